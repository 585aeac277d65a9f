"""GPU input preparation and fused loss/metrics — the "next" rows 1 and 2 of SURVEY.md §8f.

  preprocess_features  <- utils.preprocess_graph (utils.py:67-75): z-scored overlap_length / overlap_similarity
  positional_encoding  <- utils.add_positional_encoding (utils.py:97-138, 'PR') + the concat of train.py:249-251
  bce_with_logits_and_metrics <- BCEWithLogitsLoss(pos_weight) (train.py:211,255) + utils.calculate_tfpn
                                 (utils.py:217-223): one kernel, one D2H for loss and TP/TN/FP/FN.
"""
from __future__ import annotations

import torch

from . import _lib
from ._lib import check, ptr
from .functional import _on
from .plan import plan_for


def preprocess_features(overlap_length, overlap_similarity):
    """e [E,2] fp32 on the inputs' CUDA device (edge-id order)."""
    a = overlap_length.float().contiguous()
    b = overlap_similarity.float().contiguous()
    if not a.is_cuda:
        raise RuntimeError("gnnome_assembly_b200.prep: CUDA tensors expected (no CPU path)")
    E = a.numel()
    out = torch.empty(E, 2, device=a.device, dtype=torch.float32)
    ws = torch.empty(4, device=a.device, dtype=torch.float64)
    with _on(a, b) as st:
        check(_lib.lib().gg_prep_edge_features(E, ptr(a), ptr(b), ptr(out), ptr(ws), st), "gg_prep_edge_features")
    return out


def positional_encoding(graph, pe_dim=16, alpha=0.95, device=None):
    """pe [N, 2 + pe_dim] = in_deg | out_deg | k-step PageRank, rows in the caller's node order."""
    plan = plan_for(graph, device)
    N = plan.num_nodes
    out = torch.empty(N, 2 + pe_dim, device=plan.device, dtype=torch.float32)
    ws = torch.empty(3 * max(N, 1), device=plan.device, dtype=torch.float64)
    with _on(out, plan=plan) as st:
        check(_lib.lib().gg_prep_pe(plan.handle, pe_dim, float(alpha), ptr(out), ptr(ws), st), "gg_prep_pe")
    return out


class _BceMetrics(torch.autograd.Function):
    @staticmethod
    def forward(ctx, scores, y, pos_weight):
        s = scores.reshape(-1).contiguous()
        t = y.reshape(-1).to(torch.float32).contiguous()
        if not s.is_cuda:
            raise RuntimeError("gnnome_assembly_b200.prep: CUDA tensors expected (no CPU path)")
        E = s.numel()
        out = torch.empty(5, device=s.device, dtype=torch.float64)
        with _on(s, t) as st:
            check(_lib.lib().gg_bce_metrics_fwd(E, ptr(s), ptr(t), float(pos_weight), ptr(out), st), "gg_bce_metrics_fwd")
        ctx.save_for_backward(s, t)
        ctx.pos_weight, ctx.shape = float(pos_weight), scores.shape
        loss = (out[0] / max(E, 1)).to(torch.float32)
        counts = out[1:].clone()
        ctx.mark_non_differentiable(counts)
        return loss, counts

    @staticmethod
    def backward(ctx, g_loss, _g_counts):
        s, t = ctx.saved_tensors
        g = torch.empty_like(s)
        gl = g_loss.reshape(1).to(torch.float32).contiguous()
        with _on(s, t, gl) as st:
            check(_lib.lib().gg_bce_bwd(s.numel(), ptr(s), ptr(t), ctx.pos_weight, ptr(gl), ptr(g), st), "gg_bce_bwd")
        return g.reshape(ctx.shape), None, None


def bce_with_logits_and_metrics(scores, y, pos_weight):
    """(loss, counts) with counts = float64 tensor [TP, TN, FP, FN] on the device (read with ONE .tolist())."""
    return _BceMetrics.apply(scores, y, pos_weight)

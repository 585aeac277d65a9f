"""torch.autograd bindings of the C-ABI kernels (include/gnnome_b200.h).

PyTorch is plumbing here: it owns device memory, streams and the autograd tape.  All arithmetic of the
hot path runs in libgnnome_b200.so.  There is no CPU / eager fallback: CPU tensors raise.
"""
from __future__ import annotations

import torch
import torch.nn.functional as F

from . import _lib
from ._lib import check, ptr

NORM_BATCH, NORM_LAYER = 0, 1


class _on:
    """Device guard for one C-ABI call: makes the tensors' device current (kernel launches, TMA descriptors and the
    per-device caches of the library follow the CURRENT device) and yields that device's current stream.  The
    reference's default device is 'cuda:3' with no set_device (hyperparameters.py:25), so the current device is
    NOT the tensors' device in general.  Every tensor of the call (and the plan) must share one device."""

    def __init__(self, *tensors, plan=None):
        devs = {t.device for t in tensors if t is not None}
        if plan is not None:
            devs.add(plan.device)
        if len(devs) != 1:
            raise RuntimeError(f"gnnome_assembly_b200: tensors / graph plan of one call live on different devices: {sorted(map(str, devs))}")
        (self.dev,) = devs
        if self.dev.type != "cuda":
            raise RuntimeError("gnnome_assembly_b200: tensors must live on a CUDA device (no CPU path)")

    def __enter__(self):
        self._guard = None
        if torch.cuda.current_device() != self.dev.index:          # the common case needs no switch at all
            self._guard = torch.cuda.device(self.dev)
            self._guard.__enter__()
        return torch.cuda.current_stream(self.dev).cuda_stream

    def __exit__(self, *exc):
        if self._guard is not None:
            return self._guard.__exit__(*exc)
        return False


def _cuda_f32(*tensors):
    out = []
    for t in tensors:
        if t is None:
            out.append(None)
            continue
        if not t.is_cuda:
            raise RuntimeError("gnnome_assembly_b200: tensors must live on a CUDA device (no CPU path)")
        if t.dtype != torch.float32:
            raise RuntimeError(f"gnnome_assembly_b200: fp32 tensors expected, got {t.dtype}")
        out.append(t.contiguous())
    return out


# ----------------------------------------------------------------------------------- row permutation
class _PermuteRows(torch.autograd.Function):
    """out[p] = x[idx[p]] for a permutation idx with inverse inv (edge-id order <-> internal order)."""

    @staticmethod
    def forward(ctx, x, idx, inv):
        (x,) = _cuda_f32(x)
        ctx.idx, ctx.inv = idx, inv
        rows = x.shape[0]
        width = x.numel() // rows if rows else 1
        out = torch.empty_like(x)
        with _on(x, idx) as st:
            check(_lib.lib().gg_gather_rows(rows, width, ptr(x), ptr(idx), ptr(out), st), "gg_gather_rows")
        return out

    @staticmethod
    def backward(ctx, g):
        (g,) = _cuda_f32(g)
        rows = g.shape[0]
        width = g.numel() // rows if rows else 1
        out = torch.empty_like(g)
        with _on(g, ctx.inv) as st:
            check(_lib.lib().gg_gather_rows(rows, width, ptr(g), ptr(ctx.inv), ptr(out), st), "gg_gather_rows")
        return out, None, None


def permute_rows(x, idx, inv):
    return _PermuteRows.apply(x, idx, inv)


# ----------------------------------------------------------------------------------- dense linear
class _Linear(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, W, b):
        x, W, b = _cuda_f32(x, W, b)
        M, K = x.shape
        N = W.shape[0]
        y = torch.empty(M, N, device=x.device, dtype=torch.float32)
        with _on(x, W, b) as st:
            check(_lib.lib().gg_linear_fwd(M, N, K, ptr(x), ptr(W), ptr(b), 0, ptr(y), st), "gg_linear_fwd")
        ctx.save_for_backward(x, W)
        ctx.has_bias = b is not None
        return y

    @staticmethod
    def backward(ctx, g):
        x, W = ctx.saved_tensors
        (g,) = _cuda_f32(g)
        M, K = x.shape
        N = W.shape[0]
        lib = _lib.lib()
        gx = None
        dW = torch.empty_like(W)
        db = torch.empty(N, device=W.device, dtype=torch.float32) if ctx.has_bias else None
        with _on(g, x, W) as st:
            if ctx.needs_input_grad[0]:
                gx = torch.empty_like(x)
                check(lib.gg_linear_bwd_data(M, N, K, ptr(g), ptr(W), None, None, ptr(gx), st), "gg_linear_bwd_data")
            check(lib.gg_linear_bwd_weight(M, N, K, ptr(g), ptr(x), ptr(dW), ptr(db), st), "gg_linear_bwd_weight")
        return gx, dW, db


def _pad_k(x, W):
    """The kernels want K % 4 == 0: zero-pad the reduction dimension (18 -> 20, 2 -> 4)."""
    K = W.shape[1]
    r = (-K) % 4
    if r:
        x = F.pad(x, (0, r))
        W = F.pad(W, (0, r))
    return x, W


def linear(x, W, b=None):
    """nn.Linear forward (models/full_graph.py:23) on the FFMA GEMM kernel."""
    x, W = _pad_k(x, W)
    return _Linear.apply(x, W, b)


class _EdgeMLP(torch.autograd.Function):
    """linear2_edge(relu(linear1_edge(e)))  (models/full_graph.py:24-26); ReLU fused in both directions."""

    @staticmethod
    def forward(ctx, e, W1, b1, W2, b2):
        e, W1, b1, W2, b2 = _cuda_f32(e, W1, b1, W2, b2)
        lib = _lib.lib()
        E, K = e.shape
        Hh, d = W1.shape[0], W2.shape[0]
        hid = torch.empty(E, Hh, device=e.device, dtype=torch.float32)
        out = torch.empty(E, d, device=e.device, dtype=torch.float32)
        with _on(e, W1, b1, W2, b2) as st:
            check(lib.gg_linear_fwd(E, Hh, K, ptr(e), ptr(W1), ptr(b1), 1, ptr(hid), st), "gg_linear_fwd")
            check(lib.gg_linear_fwd(E, d, Hh, ptr(hid), ptr(W2), ptr(b2), 0, ptr(out), st), "gg_linear_fwd")
        ctx.save_for_backward(e, W1, W2, hid)
        return out

    @staticmethod
    def backward(ctx, g):
        e, W1, W2, hid = ctx.saved_tensors
        (g,) = _cuda_f32(g)
        lib = _lib.lib()
        E, K = e.shape
        Hh, d = W1.shape[0], W2.shape[0]
        dW2 = torch.empty_like(W2)
        db2 = torch.empty(d, device=g.device, dtype=torch.float32)
        dW1 = torch.empty_like(W1)
        db1 = torch.empty(Hh, device=g.device, dtype=torch.float32)
        with _on(g, e, W1, W2, hid) as st:
            if Hh == 16 and K == 4 and d in (64, 128):
                # one pass over g: dW2, db2, the ReLU-masked hidden gradient (never stored), dW1, db1
                check(lib.gg_edge_mlp_bwd(E, d, Hh, K, ptr(g), ptr(hid), ptr(e), ptr(W2), ptr(dW1), ptr(db1), ptr(dW2),
                                          ptr(db2), st), "gg_edge_mlp_bwd")
                return None, dW1, db1, dW2, db2
            check(lib.gg_linear_bwd_weight(E, d, Hh, ptr(g), ptr(hid), ptr(dW2), ptr(db2), st), "gg_linear_bwd_weight")
            g_hid = torch.empty_like(hid)
            check(lib.gg_linear_bwd_data(E, d, Hh, ptr(g), ptr(W2), None, ptr(hid), ptr(g_hid), st), "gg_linear_bwd_data")
            check(lib.gg_linear_bwd_weight(E, Hh, K, ptr(g_hid), ptr(e), ptr(dW1), ptr(db1), st), "gg_linear_bwd_weight")
        return None, dW1, db1, dW2, db2


def edge_mlp(e, W1, b1, W2, b2):
    e, W1 = _pad_k(e, W1)
    return _EdgeMLP.apply(e, W1, b1, W2, b2)


# ----------------------------------------------------------------------------------- GatedGCN layer
class _GatedGCNLayer(torch.autograd.Function):
    """GatedGCN_1d.forward (layers/gated_gcn_full.py:99-157); e in INTERNAL edge order."""

    @staticmethod
    def forward(ctx, plan, norm_kind, residual, h, e, Wn, bn, B3, b3, ge, be, gh, bh):
        h, e, Wn, bn, B3, b3, ge, be, gh, bh = _cuda_f32(h, e, Wn, bn, B3, b3, ge, be, gh, bh)
        N, d = h.shape
        E = e.shape[0]
        if N != plan.num_nodes or E != plan.num_edges or e.shape[1] != d:
            raise RuntimeError(f"GatedGCN layer: h{tuple(h.shape)} / e{tuple(e.shape)} do not match the graph "
                               f"(N={plan.num_nodes}, E={plan.num_edges})")
        dev = h.device
        f32 = dict(device=dev, dtype=torch.float32)
        h_out, e_out = torch.empty(N, d, **f32), torch.empty(E, d, **f32)
        P, t, z = torch.empty(N, 5 * d, **f32), torch.empty(E, d, **f32), torch.empty(N, d, **f32)
        agg = torch.empty(5, N, d, **f32)
        stats = torch.empty(4 * d, device=dev, dtype=torch.float64)
        with _on(h, e, Wn, bn, B3, b3, ge, be, gh, bh, plan=plan) as st:
            check(_lib.lib().gg_layer_fwd(plan.handle, d, norm_kind, int(residual), ptr(h), ptr(e), ptr(Wn), ptr(bn),
                                          ptr(B3), ptr(b3), ptr(ge), ptr(be), ptr(gh), ptr(bh), ptr(h_out), ptr(e_out),
                                          ptr(P), ptr(t), ptr(z), ptr(agg), ptr(stats), st), "gg_layer_fwd")
        ctx.plan, ctx.norm_kind, ctx.residual = plan, norm_kind, int(residual)
        ctx.save_for_backward(h, e, e_out, Wn, B3, ge, be, gh, bh, P, t, z, agg, stats)
        return h_out, e_out

    @staticmethod
    def backward(ctx, g_h, g_e):
        h, e, e_out, Wn, B3, ge, be, gh, bh, P, t, z, agg, stats = ctx.saved_tensors
        plan = ctx.plan
        N, d = h.shape
        E = e.shape[0]
        dev = h.device
        f32 = dict(device=dev, dtype=torch.float32)
        g_h = torch.zeros(N, d, **f32) if g_h is None else _cuda_f32(g_h)[0]
        g_e = None if g_e is None else _cuda_f32(g_e)[0]
        g_h_in, g_eo = torch.empty(N, d, **f32), torch.empty(E, d, **f32)
        g_t, g_e_in = torch.empty(E, d, **f32), torch.empty(E, d, **f32)
        gP, G = torch.empty(N, 5 * d, **f32), torch.empty(2, N, 2 * d, **f32)
        dWn, dbn = torch.empty(5 * d, d, **f32), torch.empty(5 * d, **f32)
        dB3, db3 = torch.empty(d, d, **f32), torch.empty(d, **f32)
        dge, dbe, dgh, dbh = (torch.empty(d, **f32) for _ in range(4))
        bstats = torch.empty(4 * d, device=dev, dtype=torch.float64)
        with _on(h, e, g_h, g_e, plan=plan) as st:
            check(_lib.lib().gg_layer_bwd(
                plan.handle, d, ctx.norm_kind, ctx.residual, ptr(h), ptr(e), ptr(e_out), ptr(Wn), ptr(B3), ptr(ge),
                ptr(be), ptr(gh), ptr(bh), ptr(P), ptr(t), ptr(z), ptr(agg), ptr(stats), ptr(g_h), ptr(g_e),
                ptr(g_h_in), ptr(g_e_in), ptr(dWn), ptr(dbn), ptr(dB3), ptr(db3), ptr(dge), ptr(dbe), ptr(dgh), ptr(dbh),
                ptr(gP), ptr(G), ptr(g_eo), ptr(g_t), ptr(bstats), st), "gg_layer_bwd")
        return None, None, None, g_h_in, g_e_in, dWn, dbn, dB3, db3, dge, dbe, dgh, dbh


def gated_gcn_layer(plan, norm_kind, residual, h, e, Wn, bn, B3, b3, ge, be, gh, bh):
    return _GatedGCNLayer.apply(plan, norm_kind, residual, h, e, Wn, bn, B3, b3, ge, be, gh, bh)


# ----------------------------------------------------------------------------------- score predictor
class _Score(torch.autograd.Function):
    """ScorePredictor (layers/score_predictor.py:12-25) with W1 split by columns; internal edge order."""

    @staticmethod
    def forward(ctx, plan, x, e, Wq, bq, W1e, w2, b2):
        x, e, Wq, bq, W1e, w2, b2 = _cuda_f32(x, e, Wq, bq, W1e, w2, b2)
        N, d = x.shape
        E = e.shape[0]
        H = W1e.shape[0]
        dev = x.device
        f32 = dict(device=dev, dtype=torch.float32)
        score, Q = torch.empty(E, **f32), torch.empty(N, 2 * H, **f32)
        need_grad = any(ctx.needs_input_grad)
        hid = torch.empty(E, H, **f32) if need_grad else None
        with _on(x, e, Wq, bq, W1e, w2, b2, plan=plan) as st:
            check(_lib.lib().gg_score_fwd(plan.handle, d, H, ptr(x), ptr(e), ptr(Wq), ptr(bq), ptr(W1e), ptr(w2), ptr(b2),
                                          ptr(score), ptr(Q), ptr(hid), st), "gg_score_fwd")
        ctx.plan = plan
        if need_grad:
            ctx.save_for_backward(x, e, Wq, W1e, w2, hid)
        return score

    @staticmethod
    def backward(ctx, g):
        x, e, Wq, W1e, w2, hid = ctx.saved_tensors
        (g,) = _cuda_f32(g)
        plan = ctx.plan
        N, d = x.shape
        E = e.shape[0]
        H = W1e.shape[0]
        dev = x.device
        f32 = dict(device=dev, dtype=torch.float32)
        g_x, g_e = torch.empty(N, d, **f32), torch.empty(E, d, **f32)
        dWq, dbq, dW1e = torch.empty(2 * H, d, **f32), torch.empty(2 * H, **f32), torch.empty(H, d, **f32)
        dw2, db2 = torch.empty(H, **f32), torch.empty(1, **f32)
        gQ = torch.empty(N, 2 * H, **f32)
        red = torch.empty(2 * H + 1, device=dev, dtype=torch.float64)
        gpre = torch.empty_like(hid)
        with _on(x, e, g, hid, plan=plan) as st:
            check(_lib.lib().gg_score_bwd(plan.handle, d, H, ptr(x), ptr(e), ptr(Wq), ptr(W1e), ptr(w2), ptr(g), ptr(hid),
                                          ptr(g_x), ptr(g_e), ptr(dWq), ptr(dbq), ptr(dW1e), ptr(dw2), ptr(db2), ptr(gpre),
                                          ptr(gQ), ptr(red), st), "gg_score_bwd")
        return None, g_x, g_e, dWq, dbq, dW1e, dw2, db2


def score_predictor(plan, x, e, W1, b1, W2, b2):
    """W1 [H, 3d] -> Wq = [W1[:, :d] ; W1[:, d:2d]], W1e = W1[:, 2d:]; b1 rides on the src half of Q."""
    d = x.shape[1]
    Wq = torch.cat((W1[:, :d], W1[:, d:2 * d]), dim=0)
    bq = torch.cat((b1, torch.zeros_like(b1)), dim=0)
    W1e = W1[:, 2 * d:].contiguous()
    return _Score.apply(plan, x, e, Wq, bq, W1e, W2.reshape(-1), b2.reshape(-1))

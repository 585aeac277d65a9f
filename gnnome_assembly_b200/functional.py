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


# ----------------------------------------------------------------------------------- gradient outputs
def _grad_out(arena, param, shape, device):
    """Where a parameter gradient is written: its slot of the pass's GradArena (flat.py) or a fresh tensor."""
    if arena is not None and param is not None and id(param) in arena.layout.offset_of:
        return arena.slot(param).view(shape)
    return torch.empty(shape, device=device, dtype=torch.float32)


def _pad_cols(t, K4):
    """zero-pad the last dimension to K4 columns (the kernels want K % 4 == 0: 18 -> 20, 2 -> 4)"""
    return t if t.shape[1] == K4 else F.pad(t, (0, K4 - t.shape[1]))


# ----------------------------------------------------------------------------------- dense linear
class _Linear(torch.autograd.Function):
    """nn.Linear; W may have K % 4 != 0 (padding happens here, outside autograd)."""

    @staticmethod
    def forward(ctx, x, W, b, arena, Wp, bp):
        # Wp / bp: the nn.Parameters behind W / b (identify the arena slots); W, b are what autograd tracks
        x, W, b = _cuda_f32(x, W, b)
        M, K = x.shape
        N = W.shape[0]
        K4 = (K + 3) & ~3
        x4, W4 = _pad_cols(x, K4), _pad_cols(W, K4)
        y = torch.empty(M, N, device=x.device, dtype=torch.float32)
        with _on(x4, W4, b) as st:
            check(_lib.lib().gg_linear_fwd(M, N, K4, ptr(x4), ptr(W4), ptr(b), 0, ptr(y), st), "gg_linear_fwd")
        ctx.save_for_backward(x4, W4)
        ctx.has_bias, ctx.K, ctx.arena, ctx.Wp, ctx.bp = b is not None, K, arena, Wp, bp
        return y

    @staticmethod
    def backward(ctx, g):
        x4, W4 = ctx.saved_tensors
        (g,) = _cuda_f32(g)
        M, K4 = x4.shape
        N, K = W4.shape[0], ctx.K
        lib = _lib.lib()
        dev = W4.device
        gx = None
        dW4 = _grad_out(ctx.arena if K4 == K else None, ctx.Wp, (N, K4), dev)
        db = _grad_out(ctx.arena, ctx.bp, (N,), dev) if ctx.has_bias else None
        with _on(g, x4, W4) as st:
            if ctx.needs_input_grad[0]:
                gx4 = torch.empty_like(x4)
                check(lib.gg_linear_bwd_data(M, N, K4, ptr(g), ptr(W4), None, None, ptr(gx4), st), "gg_linear_bwd_data")
                gx = gx4 if K4 == K else gx4[:, :K].contiguous()
            check(lib.gg_linear_bwd_weight(M, N, K4, ptr(g), ptr(x4), ptr(dW4), ptr(db), st), "gg_linear_bwd_weight")
        if K4 != K:
            dW = _grad_out(ctx.arena, ctx.Wp, (N, K), dev)
            dW.copy_(dW4[:, :K])
        else:
            dW = dW4
        return gx, dW, db, None, None, None


def linear(x, W, b=None, arena=None):
    """nn.Linear forward (models/full_graph.py:23) on the FFMA GEMM kernel."""
    return _Linear.apply(x, W, b, arena, W, b)


class _EdgeMLP(torch.autograd.Function):
    """linear2_edge(relu(linear1_edge(e)))  (models/full_graph.py:24-26); ReLU fused in both directions."""

    @staticmethod
    def forward(ctx, e, W1, b1, W2, b2, arena, params):
        e, W1, b1, W2, b2 = _cuda_f32(e, W1, b1, W2, b2)
        lib = _lib.lib()
        E, K = e.shape
        K4 = (K + 3) & ~3
        e4, W14 = _pad_cols(e, K4), _pad_cols(W1, K4)
        Hh, d = W1.shape[0], W2.shape[0]
        hid = torch.empty(E, Hh, device=e.device, dtype=torch.float32)
        out = torch.empty(E, d, device=e.device, dtype=torch.float32)
        with _on(e4, W14, b1, W2, b2) as st:
            check(lib.gg_linear_fwd(E, Hh, K4, ptr(e4), ptr(W14), ptr(b1), 1, ptr(hid), st), "gg_linear_fwd")
            check(lib.gg_linear_fwd(E, d, Hh, ptr(hid), ptr(W2), ptr(b2), 0, ptr(out), st), "gg_linear_fwd")
        ctx.save_for_backward(e4, W14, W2, hid)
        ctx.K, ctx.arena, ctx.params = K, arena, params
        return out

    @staticmethod
    def backward(ctx, g):
        e4, W14, W2, hid = ctx.saved_tensors
        (g,) = _cuda_f32(g)
        lib = _lib.lib()
        E, K4 = e4.shape
        K = ctx.K
        Hh, d = W14.shape[0], W2.shape[0]
        dev = g.device
        pW1, pb1, pW2, pb2 = ctx.params
        dW2 = _grad_out(ctx.arena, pW2, (d, Hh), dev)
        db2 = _grad_out(ctx.arena, pb2, (d,), dev)
        dW14 = _grad_out(ctx.arena if K4 == K else None, pW1, (Hh, K4), dev)
        db1 = _grad_out(ctx.arena, pb1, (Hh,), dev)
        with _on(g, e4, W14, W2, hid) as st:
            if Hh == 16 and K4 == 4 and d in (64, 128):
                # one pass over g: dW2, db2, the ReLU-masked hidden gradient (never stored), dW1, db1
                check(lib.gg_edge_mlp_bwd(E, d, Hh, K4, ptr(g), ptr(hid), ptr(e4), ptr(W2), ptr(dW14), ptr(db1), ptr(dW2),
                                          ptr(db2), st), "gg_edge_mlp_bwd")
            else:
                check(lib.gg_linear_bwd_weight(E, d, Hh, ptr(g), ptr(hid), ptr(dW2), ptr(db2), st), "gg_linear_bwd_weight")
                g_hid = torch.empty_like(hid)
                check(lib.gg_linear_bwd_data(E, d, Hh, ptr(g), ptr(W2), None, ptr(hid), ptr(g_hid), st), "gg_linear_bwd_data")
                check(lib.gg_linear_bwd_weight(E, Hh, K4, ptr(g_hid), ptr(e4), ptr(dW14), ptr(db1), st), "gg_linear_bwd_weight")
        if K4 != K:
            dW1 = _grad_out(ctx.arena, pW1, (Hh, K), dev)
            dW1.copy_(dW14[:, :K])
        else:
            dW1 = dW14
        return None, dW1, db1, dW2, db2, None, None


def edge_mlp(e, W1, b1, W2, b2, arena=None):
    return _EdgeMLP.apply(e, W1, b1, W2, b2, arena, (W1, b1, W2, b2))


# ----------------------------------------------------------------------------------- GatedGCN layer
class _GatedGCNLayer(torch.autograd.Function):
    """GatedGCN_1d.forward (layers/gated_gcn_full.py:99-157); e in INTERNAL edge order.

    The five node projections A_1, A_2, A_3, B_1, B_2 enter twice: as the individual parameters (w5, b5: what autograd
    tracks and what receives the gradients) and stacked as Wn [5d, d] / bn [5d] (what the kernel reads) — views of the
    flat parameter buffer (flat.py), no copy."""

    @staticmethod
    def forward(ctx, plan, norm_kind, residual, arena, conv, h, e, Wn, bn, *params):
        # params = A1w, A2w, A3w, B1w, B2w, A1b, A2b, A3b, B1b, B2b, B3, b3, ge, be, gh, bh   (autograd inputs)
        B3, b3, ge, be, gh, bh = params[10:]
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
        ctx.plan, ctx.norm_kind, ctx.residual, ctx.arena, ctx.conv = plan, norm_kind, int(residual), arena, conv
        ctx.save_for_backward(h, e, e_out, Wn, B3, ge, be, gh, bh, P, t, z, agg, stats)
        return h_out, e_out

    @staticmethod
    def backward(ctx, g_h, g_e):
        h, e, e_out, Wn, B3, ge, be, gh, bh, P, t, z, agg, stats = ctx.saved_tensors
        plan, arena, conv = ctx.plan, ctx.arena, ctx.conv
        N, d = h.shape
        E = e.shape[0]
        dev = h.device
        f32 = dict(device=dev, dtype=torch.float32)
        g_h = torch.zeros(N, d, **f32) if g_h is None else _cuda_f32(g_h)[0]
        g_e = None if g_e is None else _cuda_f32(g_e)[0]
        g_h_in, g_eo = torch.empty(N, d, **f32), torch.empty(E, d, **f32)
        g_t, g_e_in = torch.empty(E, d, **f32), torch.empty(E, d, **f32)
        gP, G = torch.empty(N, 5 * d, **f32), torch.empty(2, N, 2 * d, **f32)
        if arena is not None and conv is not None and id(conv.A_1.weight) in arena.layout.offset_of:
            # flat layout of a layer (flat._layer_param_order): 5 weights | 5 biases | B3 | b3 | bn_e w,b | bn_h w,b
            dWn = arena.slot(conv.A_1.weight, rows=5 * d * d).view(5 * d, d)
            dbn = arena.slot(conv.A_1.bias, rows=5 * d)
        else:
            dWn, dbn = torch.empty(5 * d, d, **f32), torch.empty(5 * d, **f32)
        pc = conv if conv is not None else None
        dB3 = _grad_out(arena, pc.B_3.weight if pc else None, (d, d), dev)
        db3 = _grad_out(arena, pc.B_3.bias if pc else None, (d,), dev)
        dge = _grad_out(arena, pc.bn_e.weight if pc else None, (d,), dev)
        dbe = _grad_out(arena, pc.bn_e.bias if pc else None, (d,), dev)
        dgh = _grad_out(arena, pc.bn_h.weight if pc else None, (d,), dev)
        dbh = _grad_out(arena, pc.bn_h.bias if pc else None, (d,), dev)
        bstats = torch.empty(4 * d, device=dev, dtype=torch.float64)
        with _on(h, e, g_h, g_e, plan=plan) as st:
            check(_lib.lib().gg_layer_bwd(
                plan.handle, d, ctx.norm_kind, ctx.residual, ptr(h), ptr(e), ptr(e_out), ptr(Wn), ptr(B3), ptr(ge),
                ptr(be), ptr(gh), ptr(bh), ptr(P), ptr(t), ptr(z), ptr(agg), ptr(stats), ptr(g_h), ptr(g_e),
                ptr(g_h_in), ptr(g_e_in), ptr(dWn), ptr(dbn), ptr(dB3), ptr(db3), ptr(dge), ptr(dbe), ptr(dgh), ptr(dbh),
                ptr(gP), ptr(G), ptr(g_eo), ptr(g_t), ptr(bstats), st), "gg_layer_bwd")
        if arena is not None and conv is not None:
            arena.segment_done(getattr(conv, "_gg_segment", None))
        dws = tuple(dWn[k * d:(k + 1) * d] for k in range(5))
        dbs = tuple(dbn[k * d:(k + 1) * d] for k in range(5))
        return (None, None, None, None, None, g_h_in, g_e_in, None, None, *dws, *dbs, dB3, db3, dge, dbe, dgh, dbh)


def gated_gcn_layer(plan, norm_kind, residual, h, e, conv, arena=None):
    """One layer on the parameters of `conv` (a layers.GatedGCN_1d)."""
    from .flat import packed_node_weights
    packed = packed_node_weights(conv)
    ws = (conv.A_1.weight, conv.A_2.weight, conv.A_3.weight, conv.B_1.weight, conv.B_2.weight)
    bs = (conv.A_1.bias, conv.A_2.bias, conv.A_3.bias, conv.B_1.bias, conv.B_2.bias)
    if packed is None:                       # parameters not in a flat buffer (should not happen after ensure_flat)
        with torch.no_grad():
            packed = torch.cat(ws, 0), torch.cat(bs, 0)
    Wn, bn = packed
    return _GatedGCNLayer.apply(plan, norm_kind, residual, arena, conv, h, e, Wn, bn, *ws, *bs, conv.B_3.weight,
                                conv.B_3.bias, conv.bn_e.weight, conv.bn_e.bias, conv.bn_h.weight, conv.bn_h.bias)


# ----------------------------------------------------------------------------------- score predictor
class _Score(torch.autograd.Function):
    """ScorePredictor (layers/score_predictor.py:12-25) with W1 split by columns; internal edge order.
    W1 [H, 3d] = [W1s | W1d | W1e]: Wq = [W1s ; W1d] stacked by rows, bq = [b1 ; 0] (b1 rides on the src half of Q)."""

    @staticmethod
    def forward(ctx, plan, x, e, W1, b1, W2, b2, arena, params):
        x, e, W1, b1, W2, b2 = _cuda_f32(x, e, W1, b1, W2, b2)
        N, d = x.shape
        E = e.shape[0]
        H = W1.shape[0]
        dev = x.device
        f32 = dict(device=dev, dtype=torch.float32)
        Wq = torch.cat((W1[:, :d], W1[:, d:2 * d]), dim=0)
        bq = torch.cat((b1, torch.zeros_like(b1)), dim=0)
        W1e = W1[:, 2 * d:].contiguous()
        w2 = W2.reshape(-1)
        score, Q = torch.empty(E, **f32), torch.empty(N, 2 * H, **f32)
        need_grad = any(ctx.needs_input_grad)
        hid = torch.empty(E, H, **f32) if need_grad else None
        with _on(x, e, Wq, bq, W1e, w2, b2, plan=plan) as st:
            check(_lib.lib().gg_score_fwd(plan.handle, d, H, ptr(x), ptr(e), ptr(Wq), ptr(bq), ptr(W1e), ptr(w2), ptr(b2),
                                          ptr(score), ptr(Q), ptr(hid), st), "gg_score_fwd")
        ctx.plan, ctx.arena, ctx.params = plan, arena, params
        if need_grad:
            ctx.save_for_backward(x, e, Wq, W1e, w2, hid)
        return score

    @staticmethod
    def backward(ctx, g):
        x, e, Wq, W1e, w2, hid = ctx.saved_tensors
        (g,) = _cuda_f32(g)
        plan, arena = ctx.plan, ctx.arena
        pW1, pb1, pW2, pb2 = ctx.params
        N, d = x.shape
        E = e.shape[0]
        H = W1e.shape[0]
        dev = x.device
        f32 = dict(device=dev, dtype=torch.float32)
        g_x, g_e = torch.empty(N, d, **f32), torch.empty(E, d, **f32)
        dWq, dbq, dW1e = torch.empty(2 * H, d, **f32), torch.empty(2 * H, **f32), torch.empty(H, d, **f32)
        dw2 = _grad_out(arena, pW2, (H,), dev)
        db2 = _grad_out(arena, pb2, (1,), dev)
        gQ = torch.empty(N, 2 * H, **f32)
        red = torch.empty(2 * H + 1, device=dev, dtype=torch.float64)
        gpre = torch.empty_like(hid)
        with _on(x, e, g, hid, plan=plan) as st:
            check(_lib.lib().gg_score_bwd(plan.handle, d, H, ptr(x), ptr(e), ptr(Wq), ptr(W1e), ptr(w2), ptr(g), ptr(hid),
                                          ptr(g_x), ptr(g_e), ptr(dWq), ptr(dbq), ptr(dW1e), ptr(dw2), ptr(db2), ptr(gpre),
                                          ptr(gQ), ptr(red), st), "gg_score_bwd")
        dW1 = _grad_out(arena, pW1, (H, 3 * d), dev)
        dW1[:, :d].copy_(dWq[:H])
        dW1[:, d:2 * d].copy_(dWq[H:])
        dW1[:, 2 * d:].copy_(dW1e)
        db1 = _grad_out(arena, pb1, (H,), dev)
        db1.copy_(dbq[:H])
        return None, g_x, g_e, dW1, db1, dw2.view(1, H), db2, None, None


def score_predictor(plan, x, e, W1, b1, W2, b2, arena=None):
    return _Score.apply(plan, x, e, W1, b1, W2, b2, arena, (W1, b1, W2, b2))


# ----------------------------------------------------------------------------------- whole model, one call each way
_SIDE_STREAMS = {}


def _side_stream(device):
    """One side stream per device for the weight-gradient GEMMs of gg_model_bwd (forked and joined inside the call)."""
    key = (device.type, device.index)
    if key not in _SIDE_STREAMS:
        _SIDE_STREAMS[key] = torch.cuda.Stream(device=device)
    return _SIDE_STREAMS[key]


class _ModelDesc(_lib.C.Structure):
    _fields_ = [(n, _lib.C.c_int32) for n in ("d", "layers", "hidden_edge", "hidden_score", "norm_kind", "node_in",
                                               "edge_in", "reserved")]


def model_call_table(model, layout):
    """(descriptor, int64 offset table) of include/gnnome_b200.h's gg_model_fwd for `model` under `layout`;
    cached on the layout (it dies with it on a re-flatten)."""
    cached = layout.__dict__.get("_model_call")
    if cached is not None:
        return cached
    C = _lib.C
    off = layout.offset_of
    convs = list(model.gnn.convs)
    d = model.linear_pe.weight.shape[0]
    table = [off[id(p)] for p in (model.linear_pe.weight, model.linear_pe.bias, model.linear1_edge.weight,
                                  model.linear1_edge.bias, model.linear2_edge.weight, model.linear2_edge.bias,
                                  model.predictor.W1.weight, model.predictor.W1.bias, model.predictor.W2.weight,
                                  model.predictor.W2.bias)]
    for c in convs:
        table += [off[id(p)] for p in (c.A_1.weight, c.A_1.bias, c.B_3.weight, c.B_3.bias, c.bn_e.weight, c.bn_e.bias,
                                       c.bn_h.weight, c.bn_h.bias)]
    norm = NORM_BATCH if (not convs or convs[0].batch_norm) else NORM_LAYER
    desc = _ModelDesc(d, len(convs), model.linear1_edge.weight.shape[0], model.predictor.W1.weight.shape[0], norm,
                      model.linear_pe.weight.shape[1], model.linear1_edge.weight.shape[1], 0)
    offs = (C.c_int64 * len(table))(*table)
    layout.__dict__["_model_call"] = (desc, offs, len(table))
    return layout.__dict__["_model_call"]


class _Model(torch.autograd.Function):
    """GraphGatedGCNModel.forward (models/full_graph.py:22-29) through gg_model_fwd / gg_model_bwd: the parameters are
    read from the model's flat buffer, the gradients land in the pass's GradArena; `params` (autograd inputs, flat order)
    only tell autograd where the gradients go."""

    @staticmethod
    def forward(ctx, model, plan, layout, flat, arena, e, pe, *params):
        e, pe = _cuda_f32(e, pe)
        C = _lib.C
        desc, offs, n = model_call_table(model, layout)
        E, N = plan.num_edges, plan.num_nodes
        if e.shape[0] != E or pe.shape[0] != N or e.shape[1] != desc.edge_in or pe.shape[1] != desc.node_in:
            raise RuntimeError(f"GraphGatedGCNModel: e{tuple(e.shape)} / pe{tuple(pe.shape)} do not match the graph "
                               f"(N={N}, E={E}) or the model (edge_features={desc.edge_in}, nb_pos_enc+2={desc.node_in})")
        training = arena is not None
        lib = _lib.lib()
        floats = lib.gg_model_workspace_floats(plan.handle, C.byref(desc), 0 if training else 1)
        if floats < 0:
            raise RuntimeError("gg_model_workspace_floats: unsupported model configuration")
        ws = torch.empty(max(int(floats), 1), device=pe.device, dtype=torch.float32)
        scores = torch.empty(E, 1, device=pe.device, dtype=torch.float32)
        with _on(e, pe, flat, plan=plan) as st:
            check(lib.gg_model_fwd(plan.handle, C.byref(desc), ptr(flat), offs, n, ptr(e), ptr(pe), int(training), ptr(ws),
                                   ptr(scores), st), "gg_model_fwd")
        ctx.model, ctx.plan, ctx.layout, ctx.flat, ctx.arena, ctx.params = model, plan, layout, flat, arena, params
        ctx.save_for_backward(ws)
        return scores

    @staticmethod
    def backward(ctx, g):
        (ws,) = ctx.saved_tensors
        (g,) = _cuda_f32(g)
        C = _lib.C
        model, plan, layout, flat, arena = ctx.model, ctx.plan, ctx.layout, ctx.flat, ctx.arena
        desc, offs, n = model_call_table(model, layout)
        lib = _lib.lib()
        L = desc.layers
        bws = torch.empty(max(int(lib.gg_model_workspace_floats(plan.handle, C.byref(desc), 2)), 1), device=g.device,
                          dtype=torch.float32)
        grads = arena.tensor()

        side = _side_stream(g.device)

        def run(lo, hi, st):
            check(lib.gg_model_bwd(plan.handle, C.byref(desc), ptr(flat), offs, n, ptr(g), ptr(ws), ptr(bws), ptr(grads),
                                   lo, hi, st, side.cuda_stream), "gg_model_bwd")

        with _on(g, ws, flat, plan=plan) as st:
            if arena.on_segment_ready is None:
                run(0, L + 2, st)
            else:                                   # data-parallel sync: one phase per call, a layer's arena segment is
                run(0, 1, st)                       # handed to the all-reduce as soon as its backward is enqueued
                for k in range(L):
                    run(1 + k, 2 + k, st)
                    arena.segment_done(f"conv{L - 1 - k}")
                run(L + 1, L + 2, st)
        out = tuple(grads[o:o + p.numel()].view_as(p) for p, o in layout.entries)
        return (None, None, None, None, None, None, None, *out)


def model_forward(model, plan, layout, flat, arena, e, pe):
    params = tuple(p for p, _ in layout.entries)
    return _Model.apply(model, plan, layout, flat, arena, e, pe, *params)

"""Flat parameter storage and per-step gradient arenas for the GatedGCN model.

Why: the layer kernel takes the five node projections as ONE stacked weight `Wn [5d, d]` (+ bias `[5d]`), while the
reference's state_dict keeps them as five nn.Linear modules (A_1, A_2, A_3, B_1, B_2; layers/gated_gcn_full.py:46-52).
Round 1 stacked them with `torch.cat` on every forward and let autograd split / accumulate the gradient again: 17 cat +
144 add launches per training step, 4 % of the step (profiles/r1_ncu_launch_summary_final.txt).  Here

  * every parameter's `.data` is a view of one flat fp32 buffer laid out so that a layer's five projection weights
    (and their biases) are adjacent: `Wn` / `bn` are strided views, no copy, no kernel;
  * one forward pass owns one `GradArena` with the SAME layout: the backward kernels write dWn / dbn / ... straight
    into it and hand autograd views of it, which AccumulateGrad adopts as `.grad` without a copy when the gradient is
    unset (`optimizer.zero_grad()` default since torch 2.0, train.py:256);
  * the data-parallel all-reduce (dp.py) then runs on the arena itself — no packing — and can be issued per layer
    segment while the rest of the backward is still running.

The parameters stay ordinary, separately named nn.Parameters: state_dict keys / shapes, `load_state_dict`, Adam,
`wandb.watch` are untouched.  `nn.Module.to()` / `.float()` re-allocate parameter storage; `ensure_flat` notices (by
address) and re-flattens on the next forward.
"""
from __future__ import annotations

import torch


def _layer_param_order(conv):
    """Parameters of one GatedGCN_1d in flat-buffer order."""
    return ([conv.A_1.weight, conv.A_2.weight, conv.A_3.weight, conv.B_1.weight, conv.B_2.weight,
             conv.A_1.bias, conv.A_2.bias, conv.A_3.bias, conv.B_1.bias, conv.B_2.bias,
             conv.B_3.weight, conv.B_3.bias, conv.bn_e.weight, conv.bn_e.bias, conv.bn_h.weight, conv.bn_h.bias])


class FlatLayout:
    """Order and offsets (in floats) of a module's parameters inside the flat buffer / a gradient arena."""

    def __init__(self, module):
        from .layers.gated_gcn_full import GatedGCN_1d
        ordered, seen = [], set()
        self.segments = []                       # (name, first offset, size): one per GatedGCN layer / other module
        convs = [m for m in module.modules() if isinstance(m, GatedGCN_1d)]
        conv_params = {id(p) for c in convs for p in c.parameters()}
        off = 0

        def add(params, name):
            nonlocal off
            beg = off
            for p in params:
                if id(p) in seen:
                    continue
                seen.add(id(p))
                ordered.append((p, off))
                off += p.numel()
                off = (off + 3) & ~3                                  # keep every tensor 16-byte aligned (TMA, float4)
            if off > beg:
                self.segments.append((name, beg, off - beg))

        add([p for p in module.parameters() if id(p) not in conv_params], "head")     # encoders + predictor
        for i, c in enumerate(convs):
            add(_layer_param_order(c), f"conv{i}")
            c.__dict__["_gg_segment"] = f"conv{i}"                    # the arena segment this layer's backward completes
        self.entries = ordered
        self.total = off
        self.offset_of = {id(p): o for p, o in ordered}


class GradArena:
    """One flat gradient buffer for ONE forward/backward pass (allocated at the first request of the backward)."""

    def __init__(self, layout, device):
        self.layout, self.device = layout, device
        self.buf = None
        self.on_segment_ready = None             # callable(offset, size): dp.OverlappedGradSync hooks in here
        self.grads_unset_at_backward = None      # were all .grad None when the backward started (see dp.ArenaSync.finish)

    def tensor(self):
        if self.buf is None:
            self.buf = torch.empty(self.layout.total, device=self.device, dtype=torch.float32)
            self.grads_unset_at_backward = all(p.grad is None for p, _ in self.layout.entries)
        return self.buf

    def slot(self, param, rows=None):
        """View of the arena where `param`'s gradient lives (rows: stack `rows` x param rows starting there)."""
        off = self.layout.offset_of[id(param)]
        n = param.numel() if rows is None else rows
        return self.tensor()[off:off + n]

    def segment_done(self, name):
        if self.on_segment_ready is not None:
            for nm, off, size in self.layout.segments:
                if nm == name:
                    self.on_segment_ready(off, size)


def ensure_flat(module):
    """Make every parameter of `module` a view of one flat buffer (FlatLayout order).  Idempotent and cheap when the
    parameters are already in place; returns the layout.  Values are preserved."""
    state = module.__dict__.get("_gg_flat")
    params = list(module.parameters())
    if not params:
        return None
    dev, dtype = params[0].device, params[0].dtype
    if state is not None:
        layout, flat = state
        base = flat.data_ptr()
        if (flat.device == dev and len(layout.entries) == len(params) and
                all(id(p) in layout.offset_of for p in params) and           # (a deep copy carries stale ids)
                all(p.data_ptr() == base + 4 * off and p.device == dev for p, off in layout.entries)):
            return layout
    if dtype != torch.float32 or any(p.dtype != torch.float32 or p.device != dev for p in params):
        raise RuntimeError("gnnome_assembly_b200: the model's parameters must be fp32 on one device")
    layout = FlatLayout(module)
    flat = torch.zeros(layout.total, device=dev, dtype=torch.float32)
    with torch.no_grad():
        for p, off in layout.entries:
            view = flat[off:off + p.numel()].view(p.shape)
            view.copy_(p.data)
            p.data = view
            if p.grad is not None:
                p.grad = None                   # a gradient tied to the old storage layout is dropped
    module.__dict__["_gg_flat"] = (layout, flat)
    return layout


def packed_node_weights(conv):
    """(Wn [5d, d], bn [5d]) of a GatedGCN_1d as views when its parameters sit in a flat buffer, else None."""
    d = conv.A_1.weight.shape[0]
    ws = [conv.A_1.weight, conv.A_2.weight, conv.A_3.weight, conv.B_1.weight, conv.B_2.weight]
    bs = [conv.A_1.bias, conv.A_2.bias, conv.A_3.bias, conv.B_1.bias, conv.B_2.bias]
    w0, b0 = ws[0].data_ptr(), bs[0].data_ptr()
    if any(w.data_ptr() != w0 + 4 * d * d * k or not w.is_contiguous() for k, w in enumerate(ws)):
        return None
    if any(b.data_ptr() != b0 + 4 * d * k for k, b in enumerate(bs)):
        return None
    # adjacent addresses are not enough (separately allocated tensors can sit back to back in the caching allocator):
    # the five tensors must be views of ONE storage
    st_w, st_b = ws[0].untyped_storage(), bs[0].untyped_storage()
    if any(w.untyped_storage().data_ptr() != st_w.data_ptr() for w in ws) or \
            any(b.untyped_storage().data_ptr() != st_b.data_ptr() for b in bs):
        return None
    with torch.no_grad():
        Wn = torch.as_strided(ws[0].data, (5 * d, d), (d, 1))
        bn = torch.as_strided(bs[0].data, (5 * d,), (1,))
    return Wn, bn

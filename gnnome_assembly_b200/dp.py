"""Data-parallel sharding of independent chromosome graphs.

The reference trains one graph per optimizer step on one device (train.py:239-258) and has no
distributed code.  Whole-graph training stays single-GPU here too (two global BatchNorm reductions per
layer make intra-graph sharding a per-layer exchange; a chr19 graph fits one B200), so the only
collective is ONE all-reduce of the flat fp32 gradient per optimizer step: rank r takes graphs
{i : i mod W == r}; a rank with no graph in a short wave contributes zeros; the sum is divided by the
number of active ranks.  Backend: NCCL over NVLink on GPUs, gloo in the CPU tests.
"""
from __future__ import annotations

import torch
import torch.distributed as dist


def shard_indices(num_graphs, rank, world_size):
    """Graph ids owned by `rank` (round-robin), as a list of waves: wave k holds graph k*W + rank or None."""
    waves = (num_graphs + world_size - 1) // world_size
    out = []
    for k in range(waves):
        i = k * world_size + rank
        out.append(i if i < num_graphs else None)
    return out


class GradBucket:
    """Flat fp32 gradient bucket over a fixed parameter list (3.3 MB at d=128/L=8, 25.6 MB at d=256/L=16).

    One `torch.cat` packs the gradients, one all-reduce sums them over ranks, and the parameters' `.grad`
    are re-pointed at views of the reduced buffer (no per-parameter copy kernels)."""

    def __init__(self, params):
        self.params = [p for p in params if p.requires_grad]
        self.sizes = [p.numel() for p in self.params]
        self.n = sum(self.sizes)

    def allreduce_mean(self, active=True, group=None):
        """Sum gradients over ranks, divide by the number of active ranks, install the result as p.grad.
        A rank without a graph in this wave passes active=False and contributes zeros."""
        dev = self.params[0].device
        flag = torch.full((1,), 1.0 if active else 0.0, device=dev, dtype=torch.float32)
        if active:
            parts = [(p.grad if p.grad is not None else torch.zeros_like(p)).reshape(-1) for p in self.params]
            flat = torch.cat(parts + [flag])
        else:
            flat = torch.zeros(self.n + 1, device=dev, dtype=torch.float32)
        if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
            dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group)
        flat[:self.n].div_(flat[self.n].clamp_min(1.0))
        off = 0
        for p, n in zip(self.params, self.sizes):
            p.grad = flat[off:off + n].view_as(p)
            off += n
        return flat[:self.n]


class ArenaSync:
    """Gradient all-reduce on the model's per-pass GradArena (flat.py), overlapped with the backward.

    The backward kernels write every parameter gradient into one flat arena in flat-parameter order, so nothing has to
    be packed.  The arena is reduced in SEGMENTS — one per GatedGCN layer, in the order the backward finishes them
    (last layer first), then the `head` segment (encoders + predictor) — each on a side stream as soon as its layer's
    backward has been enqueued, so the NCCL calls of layer l hide under the backward of layers l-1 ... 0.  Every rank
    issues the same sequence of collectives; a rank without a graph in a short wave (`idle_step`) contributes zeros.
    The sum is scaled by 1 / n_active on the side stream right after each segment's all-reduce.

    Usage (one process per GPU):
        sync = ArenaSync(model)
        loss = criterion(model(g, x, e, pe).squeeze(-1), y); optimizer.zero_grad(); loss.backward()
        sync.finish()                     # remaining segment + join the side stream; p.grad now holds the mean
        optimizer.step()
    zero_grad must leave the gradients unset (set_to_none=True, the torch >= 2.0 default): autograd then adopts the
    arena views as .grad without a copy.  All of it can be captured into a CUDA graph (train_step.GraphedTrainStep).
    """

    def __init__(self, model, group=None, overlap=True, bucket_layers=None):
        """bucket_layers: how many finished layer segments are reduced by one collective (1 = every layer on its own,
        the default is read from GG_DP_BUCKET_LAYERS, else 4; 0 / overlap=False = ONE all-reduce of the whole arena in
        finish()).  Each collective costs ~20 us of latency and borrows SMs from the persistent one-CTA-per-SM GEMMs
        it overlaps with, so fewer, larger buckets win on small models."""
        import os
        self.model, self.group, self.overlap = model, group, overlap
        if bucket_layers is None:
            bucket_layers = int(os.environ.get("GG_DP_BUCKET_LAYERS", "4"))
        self.bucket_layers = int(bucket_layers) if overlap else 0
        self.world = dist.get_world_size(group) if (dist.is_available() and dist.is_initialized()) else 1
        self.n_active = self.world
        self._arena, self._comm, self._buckets, self._next, self._done = None, None, [], 0, set()
        model.arena_hook = self._attach

    # ---- plumbing
    def _attach(self, arena):
        self._arena = arena
        self._done = set()
        self._next = 0
        self._buckets = self._schedule(arena.layout)
        arena.on_segment_ready = self._segment_ready if self.bucket_layers > 0 else None

    def _schedule(self, layout):
        """The collectives of one step, identical on every rank whatever its backward does: layer segments from the
        last layer down in groups of `bucket_layers`; the final bucket takes the remaining layers and the head."""
        segs = {name: (off, size) for name, off, size in layout.segments}
        convs = sorted((n for n in segs if n.startswith("conv")), key=lambda n: -int(n[4:]))
        groups, k = [], self.bucket_layers
        if k > 0:
            while len(convs) > k:
                groups.append(convs[:k])
                convs = convs[k:]
        groups.append(convs + [n for n in segs if not n.startswith("conv")])
        out = []
        for names in groups:
            lo = min(segs[n][0] for n in names)
            hi = max(segs[n][0] + segs[n][1] for n in names)
            if sum(segs[n][1] for n in names) == hi - lo:               # adjacent in the arena: one range
                out.append([(lo, hi - lo)] + [names])
            else:
                out.append([segs[n] for n in names] + [names])
        return out

    def _side_stream(self, dev):
        if dev.type != "cuda" or not self.overlap:
            return None
        if self._comm is None:
            self._comm = torch.cuda.Stream(device=dev)
        return self._comm

    def _reduce(self, buf):
        if self.world > 1:
            dist.all_reduce(buf, op=dist.ReduceOp.SUM, group=self.group)
        if self.n_active != 1:
            buf.mul_(1.0 / float(self.n_active))

    def _issue(self, bucket):
        arena = self._arena
        for off, size in bucket[:-1]:
            buf = arena.tensor()[off:off + size]
            comm = self._side_stream(buf.device)
            if comm is None:
                self._reduce(buf)
            else:
                from .functional import _side_stream
                comm.wait_stream(torch.cuda.current_stream(buf.device))
                comm.wait_stream(_side_stream(buf.device))      # the layers' weight-gradient GEMMs run there (gg_model_bwd)
                with torch.cuda.stream(comm):
                    self._reduce(buf)

    def _segment_ready(self, off, size):
        for name, o, s in self._arena.layout.segments:
            if o == off and s == size:
                self._done.add(name)
        while self._next < len(self._buckets) - 1 and all(n in self._done for n in self._buckets[self._next][-1]):
            self._issue(self._buckets[self._next])                          # (the last bucket waits for finish())
            self._next += 1

    # ---- API
    def begin(self, n_active=None):
        """Number of ranks that hold a graph in this wave (default: all)."""
        self.n_active = self.world if n_active is None else int(n_active)

    def finish(self):
        """Reduce what the backward hooks have not reduced yet (the head segment; everything if a layer did not take
        part in the pass), join the side stream, and make sure every .grad IS its arena slot."""
        arena = self._arena
        if arena is None:
            raise RuntimeError("ArenaSync.finish: no forward pass has been run through the model")
        while self._next < len(self._buckets):                                   # what the backward has not issued yet
            self._issue(self._buckets[self._next])
            self._next += 1
        buf = arena.tensor()
        comm = self._side_stream(buf.device)
        if comm is not None:
            torch.cuda.current_stream(buf.device).wait_stream(comm)
        base = buf.data_ptr()
        for p, off in arena.layout.entries:
            if not p.requires_grad:
                continue
            if p.grad is None or p.grad.data_ptr() != base + 4 * off:
                # Autograd normally ADOPTS the arena view as .grad (no copy).  If the gradients were unset when the
                # backward started and .grad is still somewhere else, autograd made a private copy instead (seen under
                # compute-sanitizer): the copy holds this rank's un-reduced values, the arena holds the mean — re-point.
                # A .grad that existed BEFORE the backward was accumulated into and cannot be replaced: refuse.
                if p.grad is not None and not arena.grads_unset_at_backward:
                    raise RuntimeError("ArenaSync: a gradient does not live in the pass's arena — call "
                                       "optimizer.zero_grad(set_to_none=True) before backward")
                p.grad = buf[off:off + p.numel()].view_as(p)
        return buf

    def idle_step(self):
        """A rank with no graph in this wave: zero gradients through the same sequence of collectives."""
        from .flat import GradArena, ensure_flat
        layout = ensure_flat(self.model)
        dev = next(self.model.parameters()).device
        arena = GradArena(layout, dev)
        arena.tensor().zero_()
        self._attach(arena)
        for p in self.model.parameters():
            p.grad = None
        return self.finish()

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

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
    """Flat fp32 gradient bucket over a fixed parameter list (3.3 MB at d=128/L=8, 25.6 MB at d=256/L=16)."""

    def __init__(self, params):
        self.params = [p for p in params if p.requires_grad]
        n = sum(p.numel() for p in self.params)
        dev = self.params[0].device
        self.flat = torch.zeros(n + 1, device=dev, dtype=torch.float32)     # last slot: "I was active"

    def allreduce_mean(self, active=True, group=None):
        """Sum gradients over ranks, divide by the number of active ranks, write back into p.grad."""
        off = 0
        for p in self.params:
            n = p.numel()
            if active and p.grad is not None:
                self.flat[off:off + n].copy_(p.grad.reshape(-1))
            else:
                self.flat[off:off + n].zero_()
            off += n
        self.flat[off] = 1.0 if active else 0.0
        if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
            dist.all_reduce(self.flat, op=dist.ReduceOp.SUM, group=group)
        n_active = self.flat[off].clamp_min(1.0)
        self.flat[:off].div_(n_active)
        off = 0
        for p in self.params:
            n = p.numel()
            g = self.flat[off:off + n].view_as(p)
            if p.grad is None:
                p.grad = g.clone()
            else:
                p.grad.copy_(g)
            off += n
        return self.flat[:off]

"""CUDA-graph capture of one whole training step on a fixed assembly graph.

The loop body of the reference (train.py:252-258: forward, BCEWithLogits loss, zero_grad, backward,
optimizer.step) launches ~150 of our kernels plus PyTorch's loss / optimizer kernels per step; at
~13 ms per step the launch gaps are worth ~1 ms.  A GraphedTrainStep captures the step once per graph
object (train.py re-visits the same graphs every epoch, train.py:239) and replays it: inputs are copied
into static device buffers, the loss is read from a static scalar.  Nothing about the arithmetic changes.
"""
from __future__ import annotations

import torch

from .plan import plan_for


class GraphedTrainStep:
    def __init__(self, model, optimizer, graph, e, pe, y, loss_fn, warmup=3, after_backward=None):
        """after_backward: optional hook between backward and optimizer.step, e.g. the data-parallel gradient
        all-reduce (`dp.GradBucket.allreduce_mean`); NCCL collectives are captured with the step."""
        self.model, self.optimizer, self.loss_fn = model, optimizer, loss_fn
        self.after_backward = after_backward
        dev = pe.device
        self.graph = graph
        plan_for(graph, dev)                                   # built (and cached) outside the capture
        self.e, self.pe, self.y = e.clone(), pe.clone(), y.clone()
        side = torch.cuda.Stream(device=dev)
        side.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(side):
            for _ in range(warmup):                            # allocator / lazy-init warm-up off the capture
                self._eager()
        torch.cuda.current_stream(dev).wait_stream(side)
        self.cuda_graph = torch.cuda.CUDAGraph()
        optimizer.zero_grad(set_to_none=True)
        with torch.cuda.graph(self.cuda_graph):
            self.loss = self._eager()

    def _eager(self):
        scores = self.model(self.graph, None, self.e, self.pe)
        loss = self.loss_fn(scores, self.y)
        self.optimizer.zero_grad(set_to_none=True)
        loss.backward()
        if self.after_backward is not None:
            self.after_backward()
        self.optimizer.step()
        return loss.detach()

    def __call__(self, e=None, pe=None, y=None):
        """Run one step; new inputs (device or pinned-host tensors) are copied into the static buffers."""
        if e is not None:
            self.e.copy_(e, non_blocking=True)
        if pe is not None:
            self.pe.copy_(pe, non_blocking=True)
        if y is not None:
            self.y.copy_(y, non_blocking=True)
        self.cuda_graph.replay()
        return self.loss

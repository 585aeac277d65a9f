"""Mini-batch (cluster) training path — SURVEY.md §8f row 3.

The reference's default configuration trains on METIS clusters (train.py:282-312, validation :428-456):

    sampler = dgl.dataloading.ClusterGCNSampler(g, num_clusters, cache_path=cluster_cache_path)
    dataloader = dgl.dataloading.DataLoader(g, torch.arange(num_clusters), sampler,
                                            batch_size=batch_size_train, shuffle=True, drop_last=False, num_workers=4)
    for sub_g in dataloader:
        sub_g = sub_g.to(device); x = sub_g.ndata['x']; e = sub_g.edata['e']; pe = sub_g.ndata['pe'] ...
        edge_predictions = model(sub_g, x, e, pe)

`ClusterGCNSampler` / `DataLoader` here keep those two call signatures.  What changes is where the work
happens: the parent graph (structure + features) is made resident on the GPU once, and every batch's
sub-graph — node-induced on the union of the drawn clusters, features gathered, and the engine's plan —
is produced on the device by `gg_subplan_count` / `gg_subplan_fill` (a compaction of the parent plan: no sort, no host copy of
the edge list).  `model(sub_g, ...)` then finds the plan already attached.

Partitioning itself is METIS inside DGL (third-party, absent here) and is not arithmetic of this path: any
assignment can be passed in (`partition_ids=`); the default `assembly_partition` cuts the parent plan's
breadth-first node order into k contiguous chunks, which on near-linear assembly graphs is the same kind of
partition METIS returns (connected, balanced, small edge cut) but is NOT METIS.
"""
from __future__ import annotations

import os
import pickle

import numpy as np
import torch

from . import _lib
from ._lib import check, ptr
from .functional import _on
from .graph import AssemblyGraph
from .plan import _PLAN_CACHE, plan_for

NID = "_ID"      # dgl.NID / dgl.EID: ids of the sub-graph's nodes / edges in the parent
EID = "_ID"


def _gather(t, idx32, idx64):
    """rows of a resident feature tensor; fp32 goes through the library's row gather"""
    if t.dtype == torch.float32 and t.is_cuda and t.numel() > 0:
        t = t.contiguous()
        rows = idx32.numel()
        width = t.numel() // t.shape[0]
        out = torch.empty((rows,) + tuple(t.shape[1:]), device=t.device, dtype=torch.float32)
        if rows:
            with _on(t, idx32) as st:
                check(_lib.lib().gg_gather_rows(rows, width, ptr(t), ptr(idx32), ptr(out), st), "gg_gather_rows")
        return out
    return t[idx64]


def node_subgraph(g, nodes, device=None):
    """g.subgraph(nodes) of the reference (DGL semantics: sub-graph node j = nodes[j]; edges = all edges of g
    with both ends in `nodes`, in increasing edge id; ndata / edata copied; ids kept under dgl.NID / dgl.EID).
    `g` must be resident on a CUDA device (or `device` given).  The result carries its GraphPlan."""
    plan = plan_for(g, device)
    dev = plan.device
    nodes = torch.as_tensor(nodes).to(dev, torch.int64)
    sp = plan.subplan(nodes)
    sub = AssemblyGraph(sp.array("csrc"), sp.array("cdst"), sp.num_nodes)
    eid32 = sp.array("parent_eid")
    eid64, nid32 = eid32.long(), nodes.to(torch.int32)
    for k, v in g.ndata.items():
        sub.ndata[k] = _gather(v.to(dev), nid32, nodes)
    for k, v in g.edata.items():
        sub.edata[k] = _gather(v.to(dev), eid32, eid64)
    sub.ndata[NID] = nodes
    sub.edata[EID] = eid64
    _PLAN_CACHE[sub] = sp
    return sub


def assembly_partition(g, k, device=None):
    """int64[N] cluster id per node: the plan's breadth-first node order cut into k contiguous chunks."""
    plan = plan_for(g, device)
    pos = plan.node_inv.long()                                   # caller node id -> internal position
    return (pos * int(k)) // max(plan.num_nodes, 1)


class ClusterGCNSampler:
    """Mirror of dgl.dataloading.ClusterGCNSampler(g, k, cache_path=...) as train.py:292 / :434 use it."""

    def __init__(self, g, k, cache_path=None, partition_ids=None, device=None, **_unused):
        plan = plan_for(g, device)
        self.device = plan.device
        self.k = int(k)
        if partition_ids is None and cache_path is not None and os.path.exists(cache_path):
            with open(cache_path, "rb") as f:                    # the reference deletes the cache to force a re-cut
                cached = pickle.load(f)
            if len(cached) == plan.num_nodes:
                partition_ids = torch.as_tensor(cached)
        if partition_ids is None:
            partition_ids = assembly_partition(g, self.k, self.device)
            if cache_path is not None:
                with open(cache_path, "wb") as f:
                    pickle.dump(partition_ids.cpu().numpy(), f)
        part = torch.as_tensor(partition_ids).to(self.device, torch.int64)
        if part.numel() != plan.num_nodes:
            raise ValueError("partition_ids must hold one cluster id per node")
        self.partition_node_ids = torch.argsort(part, stable=True)                        # resident
        size = torch.bincount(part, minlength=self.k).cpu().numpy()
        self.partition_offset = np.concatenate([[0], np.cumsum(size)]).astype(np.int64)
        # the parent, resident: structure is in the plan, features are gathered from these copies
        self._resident = g if _is_resident(g, self.device) else g.to(self.device)
        if self._resident is not g:
            _PLAN_CACHE[self._resident] = plan
        # Sampling runs on a stream of its own.  gg_subplan_count has to bring one integer to the host, i.e. it
        # synchronises the stream it runs on: on the training stream that wait would cover the previous batch's whole
        # backward + optimizer step and serialise host and device (batch time = host time + device time).  Sampling only
        # reads static data (parent plan, resident features), so it needs nothing from the training stream.
        torch.cuda.synchronize(self.device)                     # the resident copies above are complete
        self._stream = torch.cuda.Stream(device=self.device)

    def sample(self, g, partition_ids):
        ids = [int(i) for i in torch.as_tensor(partition_ids).reshape(-1).tolist()]
        off = self.partition_offset
        parts = [self.partition_node_ids[off[i]:off[i + 1]] for i in ids]
        main = torch.cuda.current_stream(self.device)
        with torch.cuda.stream(self._stream):
            nodes = torch.cat(parts) if parts else torch.empty(0, dtype=torch.int64, device=self.device)
            sub = node_subgraph(self._resident, nodes, self.device)
        main.wait_stream(self._stream)                          # the consumer (training stream) sees the finished batch
        for t in (*sub.edges(), *sub.ndata.values(), *sub.edata.values(), getattr(_PLAN_CACHE.get(sub), "_slab", None)):
            if torch.is_tensor(t) and t.is_cuda:
                t.record_stream(main)                           # allocated on the sampling stream, consumed on `main`
        return sub


def _is_resident(g, device):
    tensors = list(g.ndata.values()) + list(g.edata.values())
    return all(t.is_cuda and t.device == device for t in tensors)


class DataLoader:
    """Mirror of dgl.dataloading.DataLoader(g, indices, sampler, batch_size=, shuffle=, drop_last=) for a
    ClusterGCNSampler: iterates sub-graphs.  num_workers is accepted and ignored (sampling is a few
    kernel launches on the device, not host work to be hidden)."""

    def __init__(self, g, indices, graph_sampler, batch_size=1, shuffle=False, drop_last=False, num_workers=0,
                 **_unused):
        self.g, self.sampler = g, graph_sampler
        self.indices = torch.as_tensor(indices).reshape(-1).cpu()
        self.batch_size, self.shuffle, self.drop_last = int(batch_size), bool(shuffle), bool(drop_last)

    def __len__(self):
        n = self.indices.numel()
        return n // self.batch_size if self.drop_last else (n + self.batch_size - 1) // self.batch_size

    def batches(self):
        idx = self.indices[torch.randperm(self.indices.numel())] if self.shuffle else self.indices
        for b in range(len(self)):
            yield idx[b * self.batch_size:(b + 1) * self.batch_size]

    def __iter__(self):
        # one batch ahead: the sub-graph of batch k + 1 is cut (on the sampler's stream) while batch k trains
        nxt = None
        for batch in self.batches():
            cur, nxt = nxt, self.sampler.sample(self.g, batch)
            if cur is not None:
                yield cur
        if nxt is not None:
            yield nxt

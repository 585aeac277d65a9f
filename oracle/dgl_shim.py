"""Minimal stand-in for the six DGL calls the reference hot path makes.  TEST INFRASTRUCTURE ONLY.

DGL (>=0.8,<1.0, unpinned: requirements.txt:6-7, README.md:39-45) is a third-party dependency that
is not under /root/reference and cannot be installed here (no wheel, no network).  To run the
reference's OWN `layers/*.py` and `models/full_graph.py` unmodified — and so pin the oracle to the
reference's code rather than to our reading of it — this module restates, from DGL's published
semantics, exactly the API surface those files touch:

  dgl.reverse(g, copy_ndata, copy_edata)          gated_gcn_full.py:115   edge ids preserved
  g.apply_edges(fn.u_add_v(a, b, out))            gated_gcn_full.py:120,133
  g.apply_edges(python_udf)                       score_predictor.py:24   edges.src / .dst / .data
  g.update_all(fn.u_mul_e(a, e, m), fn.sum(m, o)) gated_gcn_full.py:128,141   zero in-degree -> 0
  g.update_all(fn.copy_e(e, m), fn.sum(m, o))     gated_gcn_full.py:129,142
  g.local_scope(), g.ndata, g.edata, g.edges()    score_predictor.py:21-25
  dgl.remove_self_loop, dgl.node_subgraph(store_ids=True), dgl.NID   inference.py:187,271-273 (decoder golden vectors)

`install()` registers it as `dgl` / `dgl.function` in sys.modules.  Used by
tests/golden/make_golden.py in the build container only; never on the product path.
"""
from __future__ import annotations

import contextlib
import sys
import types

import torch


class _Builtin:
    def __init__(self, kind, *args):
        self.kind, self.args = kind, args


def u_add_v(lhs, rhs, out):
    return _Builtin("u_add_v", lhs, rhs, out)


def u_mul_e(lhs, rhs, out):
    return _Builtin("u_mul_e", lhs, rhs, out)


def copy_e(e, out):
    return _Builtin("copy_e", e, out)


def sum(msg, out):  # noqa: A001 - mirrors dgl.function.sum
    return _Builtin("sum", msg, out)


class _EdgeBatch:
    def __init__(self, g):
        self.src = {k: v.index_select(0, g._src) for k, v in g.ndata.items()}
        self.dst = {k: v.index_select(0, g._dst) for k, v in g.ndata.items()}
        self.data = g.edata


class DGLGraph:
    def __init__(self, src, dst, num_nodes):
        self._src = torch.as_tensor(src, dtype=torch.int64)
        self._dst = torch.as_tensor(dst, dtype=torch.int64)
        self._n = int(num_nodes)
        self.ndata, self.edata = {}, {}

    def num_nodes(self):
        return self._n

    def num_edges(self):
        return int(self._src.numel())

    def edges(self):
        return self._src, self._dst

    @property
    def device(self):
        return self._src.device

    def int(self):                                       # utils.py:68
        return self

    def to(self, device):                                # inference.py:188 (CPU only here)
        return self

    def nodes(self):                                     # graph_parser.py:26
        return torch.arange(self._n)

    def in_degrees(self):                                # utils.py:102
        return torch.bincount(self._dst, minlength=self._n)

    def out_degrees(self):                               # utils.py:103
        return torch.bincount(self._src, minlength=self._n)

    def adjacency_matrix(self, scipy_fmt="csr"):         # utils.py:124 (DGL < 1.0): A[u, v] = #edges u -> v
        import numpy as np
        import scipy.sparse as sp
        A = sp.coo_matrix((np.ones(self.num_edges()), (self._src.numpy(), self._dst.numpy())), shape=(self._n, self._n))
        return A.asformat(scipy_fmt)

    @contextlib.contextmanager
    def local_scope(self):
        nd, ed = dict(self.ndata), dict(self.edata)
        try:
            yield
        finally:
            self.ndata, self.edata = nd, ed

    def apply_edges(self, func):
        if isinstance(func, _Builtin):
            assert func.kind == "u_add_v"
            a, b, out = func.args
            self.edata[out] = self.ndata[a].index_select(0, self._src) + self.ndata[b].index_select(0, self._dst)
        else:
            self.edata.update(func(_EdgeBatch(self)))

    def update_all(self, message, reduce):
        assert reduce.kind == "sum"
        msg_name, out = reduce.args
        if message.kind == "u_mul_e":
            u, e, m = message.args
            msg = self.ndata[u].index_select(0, self._src) * self.edata[e]
        else:
            assert message.kind == "copy_e"
            e, m = message.args
            msg = self.edata[e]
        assert m == msg_name
        res = torch.zeros((self._n,) + tuple(msg.shape[1:]), dtype=msg.dtype, device=msg.device)
        res = res.index_add(0, self._dst, msg)
        self.ndata[out] = res


def graph(edges, num_nodes=None):
    src, dst = edges
    return DGLGraph(src, dst, num_nodes)


def reverse(g, copy_ndata=True, copy_edata=False):
    r = DGLGraph(g._dst, g._src, g._n)
    if copy_ndata:
        r.ndata = dict(g.ndata)
    if copy_edata:
        r.edata = dict(g.edata)
    return r


NID = EID = "_ID"


def remove_self_loop(g):                                 # inference.py:187
    keep = g._src != g._dst
    r = DGLGraph(g._src[keep], g._dst[keep], g._n)
    r.ndata = dict(g.ndata)
    r.edata = {k: v[keep] for k, v in g.edata.items()}
    return r


def node_subgraph(g, nodes, store_ids=True):             # inference.py:271: node j of the result = nodes[j],
    nodes = torch.as_tensor(nodes).long()                # edges = induced edges in parent edge-id order
    local = torch.full((g._n,), -1, dtype=torch.int64)
    local[nodes] = torch.arange(nodes.numel())
    keep = (local[g._src] >= 0) & (local[g._dst] >= 0)
    r = DGLGraph(local[g._src[keep]], local[g._dst[keep]], nodes.numel())
    r.ndata = {k: v[nodes] for k, v in g.ndata.items()}
    r.edata = {k: v[keep] for k, v in g.edata.items()}
    if store_ids:
        r.ndata[NID] = nodes
        r.edata[EID] = torch.nonzero(keep).squeeze(1)
    return r


def install():
    """Register this module as `dgl` and `dgl.function` (call before importing reference code)."""
    me = sys.modules[__name__]
    dgl = types.ModuleType("dgl")
    dgl.DGLGraph, dgl.graph, dgl.reverse = DGLGraph, graph, reverse
    dgl.remove_self_loop, dgl.node_subgraph, dgl.NID, dgl.EID = remove_self_loop, node_subgraph, NID, EID
    dgl.function = me
    sys.modules["dgl"] = dgl
    sys.modules["dgl.function"] = me
    return dgl

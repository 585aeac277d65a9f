"""AssemblyGraph — the part of the DGLGraph API that the hot path and its callers touch, so the engine (and the
reference's own unmodified train.py / inference.py, see tools/run_reference.py and dropin/dgl) can be driven without
DGL, which is not installable here.  A real DGLGraph works as the `graph` argument of the model too.

Call sites in the reference this covers: g.edges() / num_nodes() / num_edges() (the model), g.to(device)
(train.py:243, inference.py:326), g.int() (utils.py:68), g.long() (train.py:290), g.ndata / g.edata,
g.in_degrees() / g.out_degrees() / g.adjacency_matrix(scipy_fmt=) (utils.py:102-124), g.nodes()
(graph_parser.py:26), g.subgraph(nodes) (what ClusterGCNSampler.sample does, train.py:293-296).
It is a holder of index and feature tensors, nothing else: no message passing API (the layers that would use it
are the ones this package replaces)."""
import torch


class AssemblyGraph:
    def __init__(self, src, dst, num_nodes):
        self._src = torch.as_tensor(src)
        self._dst = torch.as_tensor(dst)
        self._n = int(num_nodes)
        self.ndata, self.edata = {}, {}
        self._origin = None            # the graph this one was moved / cast from (same structure): shares its GraphPlan

    # ---- structure
    def edges(self):
        return self._src, self._dst

    def num_nodes(self):
        return self._n

    number_of_nodes = num_nodes

    def num_edges(self):
        return int(self._src.numel())

    number_of_edges = num_edges

    def nodes(self):
        return torch.arange(self._n, device=self._src.device)

    @property
    def device(self):
        return self._src.device

    @property
    def idtype(self):
        return self._src.dtype

    def in_degrees(self):
        return torch.bincount(self._dst.long(), minlength=self._n)

    def out_degrees(self):
        return torch.bincount(self._src.long(), minlength=self._n)

    def adjacency_matrix(self, scipy_fmt="csr"):
        """DGL < 1.0 (utils.py:124): scipy matrix with A[u, v] = number of edges u -> v (host data preparation)."""
        import numpy as np
        import scipy.sparse as sp
        s, d = self._src.cpu().numpy(), self._dst.cpu().numpy()
        A = sp.coo_matrix((np.ones(len(s)), (s, d)), shape=(self._n, self._n))
        return A.asformat(scipy_fmt)

    # ---- copies that keep the structure (and therefore the plan)
    def _derive(self, src, dst, ndata, edata):
        g = AssemblyGraph(src, dst, self._n)
        g.ndata, g.edata = ndata, edata
        g._origin = self._origin if self._origin is not None else self
        return g

    def to(self, device):
        device = torch.device(device)
        tensors = [self._src, self._dst, *self.ndata.values(), *self.edata.values()]
        if all(t.device == device or (t.is_cuda and device.type == "cuda" and device.index is None) for t in tensors):
            return self                                   # already resident: keeps the cached GraphPlan attached
        return self._derive(self._src.to(device), self._dst.to(device),
                            {k: v.to(device) for k, v in self.ndata.items()},
                            {k: v.to(device) for k, v in self.edata.items()})

    def int(self):
        if self._src.dtype == torch.int32:
            return self
        return self._derive(self._src.int(), self._dst.int(), dict(self.ndata), dict(self.edata))

    def long(self):
        if self._src.dtype == torch.int64:
            return self
        return self._derive(self._src.long(), self._dst.long(), dict(self.ndata), dict(self.edata))

    def subgraph(self, nodes):
        """Node-induced sub-graph, DGL semantics (node j = nodes[j], edges in increasing parent edge id, features
        copied, dgl.NID / dgl.EID stored).  CUDA-resident graphs go through the engine's device-side sub-plan
        (minibatch.node_subgraph); host graphs are cut with torch indexing (data preparation, not the hot path)."""
        if self._src.is_cuda:
            from .minibatch import node_subgraph
            return node_subgraph(self, nodes)
        nodes = torch.as_tensor(nodes).long()
        local = torch.full((self._n,), -1, dtype=torch.int64)
        local[nodes] = torch.arange(nodes.numel())
        s, d = self._src.long(), self._dst.long()
        keep = (local[s] >= 0) & (local[d] >= 0)
        sub = AssemblyGraph(local[s[keep]].to(self._src.dtype), local[d[keep]].to(self._src.dtype), nodes.numel())
        sub.ndata = {k: v[nodes] for k, v in self.ndata.items()}
        sub.edata = {k: v[keep] for k, v in self.edata.items()}
        sub.ndata["_ID"] = nodes
        sub.edata["_ID"] = torch.nonzero(keep).squeeze(1)
        return sub

    def __repr__(self):
        nd = {k: tuple(v.shape) for k, v in self.ndata.items()}
        ed = {k: tuple(v.shape) for k, v in self.edata.items()}
        return f"AssemblyGraph(num_nodes={self._n}, num_edges={self.num_edges()},\n      ndata={nd}\n      edata={ed})"

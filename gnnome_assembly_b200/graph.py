"""AssemblyGraph — the minimum of the DGLGraph API that the hot path and its callers touch
(`edges()`, `num_nodes()`, `num_edges()`, `ndata`, `edata`, `to()`, `device`), so the engine can be
driven without DGL (not installable here).  A real DGLGraph works as the `graph` argument too."""
import torch


class AssemblyGraph:
    def __init__(self, src, dst, num_nodes):
        self._src = torch.as_tensor(src)
        self._dst = torch.as_tensor(dst)
        self._n = int(num_nodes)
        self.ndata, self.edata = {}, {}

    def edges(self):
        return self._src, self._dst

    def num_nodes(self):
        return self._n

    def num_edges(self):
        return int(self._src.numel())

    @property
    def device(self):
        return self._src.device

    def to(self, device):
        device = torch.device(device)
        tensors = [self._src, self._dst, *self.ndata.values(), *self.edata.values()]
        if all(t.device == device or (t.is_cuda and device.type == "cuda" and device.index is None) for t in tensors):
            return self                                   # already resident: keeps the cached GraphPlan attached
        g = AssemblyGraph(self._src.to(device), self._dst.to(device), self._n)
        g.ndata = {k: v.to(device) for k, v in self.ndata.items()}
        g.edata = {k: v.to(device) for k, v in self.edata.items()}
        return g

    def int(self):
        return self

    def in_degrees(self):
        return torch.bincount(self._dst.long(), minlength=self._n)

    def out_degrees(self):
        return torch.bincount(self._src.long(), minlength=self._n)

"""GraphPlan — the structure side of the `graph` argument of the reference forward.

The reference passes a DGLGraph to `GraphGatedGCNModel.forward(graph, x, e, pe)`
(models/full_graph.py:22) and uses it only for structure: `graph.edges()` in edge-id order,
`num_nodes()`, `num_edges()`.  A GraphPlan is built once per graph object (train.py
re-sends the same graphs every epoch, train.py:239-245) and cached on a weak reference.
"""
from __future__ import annotations

import ctypes as C
import weakref

import torch

from . import _lib


class GraphPlan:
    """Device-side CSR over in-edges (= internal edge order), CSR over out-edges, permutations."""

    def __init__(self, src, dst, num_nodes, device=None, relabel=True):
        src = torch.as_tensor(src)
        dst = torch.as_tensor(dst)
        if device is None:
            device = src.device if src.is_cuda else torch.device("cuda", torch.cuda.current_device())
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise RuntimeError("GraphPlan needs a CUDA device: the GatedGCN engine has no CPU path")
        if src.shape != dst.shape or src.dim() != 1:
            raise ValueError("src/dst must be 1-D tensors of equal length")
        src32 = src.to(torch.int32).contiguous()
        dst32 = dst.to(torch.int32).contiguous()
        self.num_nodes = int(num_nodes)
        self.num_edges = int(src32.numel())
        lib = _lib.lib()
        handle = C.c_void_p()
        with torch.cuda.device(self.device):
            stream = torch.cuda.current_stream().cuda_stream
            rc = lib.gg_plan_create_ex(src32.data_ptr(), dst32.data_ptr(), self.num_nodes, self.num_edges,
                                       1 if relabel else 0, stream, C.byref(handle))
        _lib.check(rc, "gg_plan_create_ex")
        self.relabel = bool(relabel)
        self._handle = handle
        self._finalizer = weakref.finalize(self, lib.gg_plan_destroy, handle)

    @property
    def handle(self):
        return self._handle

    _WHICH = {"perm": 0, "inv_perm": 1, "src": 2, "dst": 3, "in_ptr": 4, "out_ptr": 5, "out_eid": 6,
              "node_perm": 7, "node_inv": 8}

    def array(self, name):
        """Copy of one of the plan's device index arrays as an int32 torch tensor (cached)."""
        cache = self.__dict__.setdefault("_arrays", {})
        if name not in cache:
            which = self._WHICH[name]
            n = (self.num_nodes + 1 if name in ("in_ptr", "out_ptr")
                 else self.num_nodes if name in ("node_perm", "node_inv") else self.num_edges)
            out = torch.empty(n, dtype=torch.int32, device=self.device)
            if n > 0:
                with torch.cuda.device(self.device):
                    rc = _lib.lib().gg_plan_copy_array(self._handle, which, out.data_ptr(),
                                                       torch.cuda.current_stream().cuda_stream)
                _lib.check(rc, "gg_plan_copy_array")
            cache[name] = out
        return cache[name]

    @property
    def perm(self):
        """int32[E]: internal position -> caller edge id."""
        return self.array("perm")

    @property
    def inv_perm(self):
        return self.array("inv_perm")

    @property
    def node_perm(self):
        """int32[N]: internal node position -> caller node id (node features are permuted at the seam)."""
        return self.array("node_perm")

    @property
    def node_inv(self):
        return self.array("node_inv")

    @property
    def src(self):
        return self.array("src")

    @property
    def dst(self):
        return self.array("dst")


_PLAN_CACHE = weakref.WeakKeyDictionary()


def plan_for(graph, device=None):
    """Return the cached GraphPlan of `graph`.

    `graph` may be a GraphPlan, anything with the DGLGraph structure API the reference uses
    (`edges()`, `num_nodes()`), or a `(src, dst, num_nodes)` tuple."""
    if isinstance(graph, GraphPlan):
        return graph
    if isinstance(graph, tuple):
        src, dst, n = graph
        return GraphPlan(src, dst, n, device)
    try:
        plan = _PLAN_CACHE.get(graph)
    except TypeError:
        plan = None
    if plan is not None and (device is None or plan.device == torch.device(device)):
        return plan
    src, dst = graph.edges()
    plan = GraphPlan(src, dst, graph.num_nodes(), device)
    try:
        _PLAN_CACHE[graph] = plan
    except TypeError:
        pass
    return plan

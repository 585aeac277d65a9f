"""`dgl` stand-in for running the reference's unmodified train.py / inference.py on the engine (tools/run_reference.py).

DGL (>= 0.8, < 1.0; unpinned in the reference, requirements.txt:6-7) cannot be installed here.  On the hot path the
reference uses DGL for message passing inside `layers/` — which this repository replaces — so what its LOOPS still
need from `dgl` is a graph holder, graph files, the dataset base class and the cluster sampler.  This package supplies
exactly that surface, by call site:

  train.py:18      from dgl.dataloading import GraphDataLoader            (imported, never used)
  train.py:172     dgl.dataloading.MultiLayerFullNeighborSampler(n)       (constructed, never used)
  train.py:292-293 dgl.dataloading.ClusterGCNSampler / DataLoader          -> gnnome_assembly_b200.minibatch
  utils.py:34      dgl.seed(seed)
  utils.py:68      g.int(); utils.py:102-124 g.in_degrees() / out_degrees() / adjacency_matrix(scipy_fmt=)
  graph_dataset.py:5,72,129  dgl.data.DGLDataset, dgl.load_graphs, dgl.save_graphs  -> gnnome_assembly_b200.graph_io
  inference.py:184,271-273   dgl.remove_self_loop, dgl.node_subgraph(store_ids=True), dgl.NID

Put `gnnome_assembly_b200/dropin` on sys.path BEFORE the reference checkout: `import dgl`, `import models` and
`import layers` then resolve here.  With a real DGL installed, leave this directory off the path and add only
`dropin/models`-style imports."""
import random as _random

import numpy as _np
import torch as _torch

from gnnome_assembly_b200.graph import AssemblyGraph as DGLGraph
from gnnome_assembly_b200.graph_io import load_graphs, save_graphs          # noqa: F401

from . import backend, data, dataloading                                    # noqa: F401,E402

__gnnome_standin__ = True
__version__ = "0.9-standin"
NID = "_ID"
EID = "_ID"


def graph(data, num_nodes=None, idtype=None, device=None):
    """dgl.graph((src, dst), num_nodes=...) (graph_parser.py:297-299 builds graphs this way via from_networkx)."""
    src, dst = data
    src, dst = _torch.as_tensor(src), _torch.as_tensor(dst)
    if idtype is not None:
        src, dst = src.to(idtype), dst.to(idtype)
    if num_nodes is None:
        num_nodes = int(max(int(src.max()), int(dst.max())) + 1) if src.numel() else 0
    g = DGLGraph(src, dst, num_nodes)
    return g.to(device) if device is not None else g


def seed(val):
    """dgl.seed (utils.py:34): DGL's own RNG drives its samplers; here the samplers draw from torch / random."""
    _random.seed(val)
    _np.random.seed(val)
    _torch.manual_seed(val)


def remove_self_loop(g):
    """inference.py:184 — a new graph without the edges u -> u (edge features filtered alongside)."""
    src, dst = g.edges()
    keep = src != dst
    r = DGLGraph(src[keep], dst[keep], g.num_nodes())
    r.ndata = dict(g.ndata)
    r.edata = {k: v[keep] for k, v in g.edata.items()}
    return r


def node_subgraph(g, nodes, store_ids=True):
    """inference.py:271 — node-induced sub-graph; node j of the result is nodes[j], dgl.NID / dgl.EID kept."""
    sub = g.subgraph(_torch.as_tensor(nodes).long().to(g.device))
    if not store_ids:
        sub.ndata.pop(NID, None)
        sub.edata.pop(EID, None)
    return sub


def reverse(g, copy_ndata=True, copy_edata=False):
    """dgl.reverse (gated_gcn_full.py:115): edge i of the result is edge i of g with its ends swapped."""
    src, dst = g.edges()
    r = DGLGraph(dst, src, g.num_nodes())
    if copy_ndata:
        r.ndata = dict(g.ndata)
    if copy_edata:
        r.edata = dict(g.edata)
    return r

"""Which part of the fused edge-gate epilogue costs what?  Times one layer forward with pieces switched off."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ctypes, numpy as np, torch
import gnnome_assembly_b200 as gg
from gnnome_assembly_b200 import _lib
from gnnome_assembly_b200.synth import make_assembly_graph
lib = _lib.lib(); lib.gg_debug_flags.restype = ctypes.c_int; lib.gg_debug_flags.argtypes = [ctypes.c_int]
dev = torch.device("cuda:0")
g = make_assembly_graph("chr19", seed=0)
graph = gg.AssemblyGraph(torch.from_numpy(g.src), torch.from_numpy(g.dst), g.num_nodes)
torch.manual_seed(0)
layer = gg.layers.GatedGCN_1d(128, 128, True).to(dev)
plan = gg.plan_for(graph, dev)
h = torch.randn(g.num_nodes, 128, device=dev); e = torch.randn(g.num_edges, 128, device=dev)
for flags in (0, 1, 2, 3):
    lib.gg_debug_flags(flags)
    with torch.no_grad():
        for _ in range(3): layer.forward_internal(plan, h, e)
        torch.cuda.synchronize(); _lib.profile(True)
        for _ in range(5): layer.forward_internal(plan, h, e)
        torch.cuda.synchronize(); _lib.profile(False)
    rep = _lib.profile_report()
    print("flags", flags, {k: round(v[1] / v[0] * 1e3, 1) for k, v in rep.items() if "gemm" in k})

"""Per-role cycle accounting of the tcgen05 GEMM (trace build only).

  bash tools/ab_build.sh trace "-DGG_TC_TRACE"
  GG_LIB=$PWD/gnnome_assembly_b200/libgnnome_b200_trace.so python tools/tc_trace.py [d]

For each traced region prints, per CTA and per 128-row tile, the cycles each role spent in total and inside its
mbarrier waits: producer (waiting for a free stage), MMA issuer (waiting for converted operands / a free
accumulator), one converter thread (waiting for TMA data), one epilogue thread (waiting for an accumulator, in the
chunk loop, in the operand fetch).  The role whose wait share is smallest is the pipeline's pace-setter.
The counters are summed over every gemm_tc launch of the region; regions are chosen so that differences isolate one
GEMM (e.g. layer forward minus the node projection alone = the edge-gate GEMM)."""
import ctypes as C
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

import gnnome_assembly_b200 as gg
from gnnome_assembly_b200 import _lib
from gnnome_assembly_b200._lib import check, ptr
from gnnome_assembly_b200.synth import make_assembly_graph

SLOTS = ["prod_wait_empty", "prod_total", "mma_wait_ab", "mma_wait_acc", "mma_total", "conv_wait_raw", "conv_total",
         "epi_wait_acc", "epi_chunks", "epi_fetch", "epi_total", "ctas", "tiles"]
d = int(sys.argv[1]) if len(sys.argv) > 1 else 128
dev = torch.device("cuda:0")
lib = _lib.lib()


def read(reset=True):
    buf = (C.c_ulonglong * len(SLOTS))()
    check(lib.gg_debug_trace(buf, len(SLOTS), 1 if reset else 0), "gg_debug_trace")
    return np.array(list(buf), dtype=np.float64)


def report(name, t):
    ctas, tiles = max(t[11], 1), max(t[12], 1)
    if t[1] == 0:
        print(f"{name}: no trace data (is GG_LIB pointing at a -DGG_TC_TRACE build?)")
        return
    per_tile = lambda x: x / tiles
    print(f"== {name}: {int(t[11])} CTA launches, {int(t[12])} tiles; cycles per tile (role total | waits)")
    print(f"   producer   {per_tile(t[1]):9.0f} | wait free stage {per_tile(t[0]):9.0f} ({100 * t[0] / t[1]:4.1f} %)")
    print(f"   MMA issuer {per_tile(t[4]):9.0f} | wait operands {per_tile(t[2]):9.0f} ({100 * t[2] / t[4]:4.1f} %), "
          f"wait accumulator {per_tile(t[3]):9.0f} ({100 * t[3] / t[4]:4.1f} %)")
    print(f"   converter  {per_tile(t[6]):9.0f} | wait TMA data {per_tile(t[5]):9.0f} ({100 * t[5] / t[6]:4.1f} %)")
    print(f"   epilogue   {per_tile(t[10]):9.0f} | wait accumulator {per_tile(t[7]):9.0f} ({100 * t[7] / t[10]:4.1f} %), "
          f"chunk loop {per_tile(t[8]):9.0f}, fetch + ids {per_tile(t[9]):9.0f}")


gs = make_assembly_graph("chr19", seed=0, pe_dim=0)
N, E = gs.num_nodes, gs.num_edges
graph = gg.AssemblyGraph(torch.from_numpy(gs.src), torch.from_numpy(gs.dst), N)
plan = gg.plan_for(graph, dev)
torch.manual_seed(0)
layer = gg.layers.GatedGCN_1d(d, d, True).to(dev)
h = torch.randn(N, d, device=dev, requires_grad=True)
e = torch.randn(E, d, device=dev, requires_grad=True)
x = torch.randn(N, d, device=dev)
W = torch.randn(5 * d, d, device=dev)
b = torch.randn(5 * d, device=dev)
y = torch.empty(N, 5 * d, device=dev)
st = lambda: torch.cuda.current_stream().cuda_stream

for _ in range(2):                                            # warm-up
    with torch.no_grad():
        layer.forward_internal(plan, h, e)
read()
check(lib.gg_linear_fwd(N, 5 * d, d, ptr(x), ptr(W), ptr(b), 0, ptr(y), st()), "gg_linear_fwd")
t_proj = read()
report("node projection alone (EpiBias, N = 5d)", t_proj)
with torch.no_grad():
    layer.forward_internal(plan, h, e)
t_fwd = read()
report("layer forward = node projection + edge-gate GEMM", t_fwd)
report("edge-gate GEMM (difference)", np.maximum(t_fwd - t_proj, 0))
ho, eo = layer.forward_internal(plan, h, e)
read()
torch.autograd.backward([ho, eo], [ho, eo])
report("layer backward (g_e_in GEMM with the A transform, dB3, g_h_in, dWn)", read())
# isolate single kernels with the runtime filter (gg_debug_flags bit 4 / bit 5)
lib.gg_debug_flags(32)
with torch.no_grad():
    layer.forward_internal(plan, h, e)
report("edge-gate GEMM alone (EpiEdgeGate)", read())
lib.gg_debug_flags(16)
ho, eo = layer.forward_internal(plan, h, e)
read()
torch.autograd.backward([ho, eo], [ho, eo])
report("g_e_in GEMM alone (BnBwdATx)", read())
lib.gg_debug_flags(0)

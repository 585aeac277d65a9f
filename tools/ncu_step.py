"""One eager training step of the bench workload between cudaProfilerStart/Stop, for
`ncu --profile-from-start off ...`.  Layers default to 2 so that `--set full` sees every kernel twice.

  ncu --set full --clock-control none --import-source on --profile-from-start off \
      -k regex:'edge_|node_|gemm_tc' -o gpurun_out/step python tools/ncu_step.py [L] [d] [scale]
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

import gnnome_assembly_b200 as gg
from gnnome_assembly_b200.synth import CHR_LEN, make_assembly_graph

L = int(sys.argv[1]) if len(sys.argv) > 1 else 2
D = int(sys.argv[2]) if len(sys.argv) > 2 else 128
scale = float(sys.argv[3]) if len(sys.argv) > 3 else 1.0
dev = torch.device("cuda:0")
g = make_assembly_graph("chr19", seed=0, genome_len=int(CHR_LEN["chr19"] * scale))
torch.manual_seed(0)
model = gg.GraphGatedGCNModel(1, 2, D, 16, L, 64, True, 16).to(dev)
opt = torch.optim.Adam(model.parameters(), lr=1e-3, fused=True)
from gnnome_assembly_b200.prep import bce_with_logits_and_metrics        # the bench step's loss (bench.py)
graph = gg.AssemblyGraph(torch.from_numpy(g.src), torch.from_numpy(g.dst), g.num_nodes)
e, pe, y = (torch.from_numpy(a).to(dev) for a in (g.e, g.pe, g.y))
flush = torch.empty(512 << 20, dtype=torch.uint8, device=dev)


def step():
    loss = bce_with_logits_and_metrics(model(graph, None, e, pe), y, 1 / 16.5)[0]
    opt.zero_grad(set_to_none=True)
    loss.backward()
    opt.step()
    return loss


for _ in range(2):
    step()
flush.zero_()
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStart()
step()
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStop()
print("N", g.num_nodes, "E", g.num_edges, "L", L, "d", D)

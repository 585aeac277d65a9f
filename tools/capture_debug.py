"""Bisect a failing CUDA-graph capture of the training step: run the step's phases under capture one by one."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
import gnnome_assembly_b200 as gg
from gnnome_assembly_b200.synth import make_assembly_graph
dev = torch.device("cuda:0"); torch.cuda.set_device(dev)
g = make_assembly_graph("chr19", seed=0, genome_len=3_000_000)
graph = gg.AssemblyGraph(torch.from_numpy(g.src), torch.from_numpy(g.dst), g.num_nodes)
model = gg.GraphGatedGCNModel(1, 2, 128, 16, 2, 64, True, 16).to(dev)
e, pe, y = (torch.from_numpy(a).to(dev) for a in (g.e, g.pe, g.y))
crit = torch.nn.BCEWithLogitsLoss(pos_weight=torch.tensor([0.06], device=dev))
opt = torch.optim.Adam(model.parameters(), lr=1e-3, fused=True, capturable=True)
def fwd(): return model(graph, None, e, pe)
def fwdbwd():
    loss = crit(fwd().squeeze(-1), y); opt.zero_grad(set_to_none=True); loss.backward(); return loss
def full():
    loss = fwdbwd(); opt.step(); return loss
for name, fn in (("forward", fwd), ("fwd+bwd", fwdbwd), ("full step", full)):
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        for _ in range(2): fn()
    torch.cuda.current_stream().wait_stream(side)
    torch.cuda.synchronize()
    try:
        cg = torch.cuda.CUDAGraph()
        with torch.cuda.graph(cg):
            out = fn()
        cg.replay(); torch.cuda.synchronize()
        print(name, "capture OK")
    except Exception as ex:
        print(name, "capture FAILED:", repr(ex)[:300])
        break

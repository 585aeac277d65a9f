"""A few cluster mini-batch training steps between cudaProfilerStart/Stop for an ncu launch list
(`ncu --metrics gpu__time_duration.sum --profile-from-start off --csv ...`): pure kernel durations per batch.
  python tools/ncu_minibatch.py [num_clusters] [batch_size] [d] [L] [batches]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

import gnnome_assembly_b200 as gg
from gnnome_assembly_b200.minibatch import ClusterGCNSampler, DataLoader
from gnnome_assembly_b200.synth import make_assembly_graph

K, BS, D, L, NB = (int(sys.argv[i]) if len(sys.argv) > i else v for i, v in ((1, 64), (2, 8), (3, 128), (4, 8), (5, 4)))
dev = torch.device("cuda:0")
gs = make_assembly_graph("chr19", seed=0)
g = gg.AssemblyGraph(torch.from_numpy(gs.src.astype(np.int64)), torch.from_numpy(gs.dst.astype(np.int64)), gs.num_nodes)
g.ndata["pe"], g.edata["e"], g.edata["y"] = torch.from_numpy(gs.pe), torch.from_numpy(gs.e), torch.from_numpy(gs.y)
sampler = ClusterGCNSampler(g, K, device=dev)
torch.manual_seed(0)
model = gg.GraphGatedGCNModel(1, 2, D, 16, L, 64, True, 16).to(dev)
opt = torch.optim.Adam(model.parameters(), lr=1e-4, fused=True)
crit = torch.nn.BCEWithLogitsLoss(pos_weight=torch.tensor([1 / 16.5], device=dev))
loader = DataLoader(g, torch.arange(K), sampler, batch_size=BS, shuffle=False)


def run(n):
    edges = 0
    for i, sub_g in enumerate(loader):
        if i >= n:
            break
        loss = crit(model(sub_g, None, sub_g.edata["e"], sub_g.ndata["pe"]).squeeze(-1), sub_g.edata["y"])
        opt.zero_grad()
        loss.backward()
        opt.step()
        edges += sub_g.num_edges()
    return edges


run(2)
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStart()
edges = run(NB)
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStop()
print("batches", NB, "edges", edges)

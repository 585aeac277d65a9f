"""BASELINE.json configs[3]: the epoch loop of train.py:232-258 over 15 chr19-like graphs, sharded round-robin over
the ranks (dp.shard_indices), one optimizer step per wave with ONE NCCL all-reduce of the flat gradient
(dp.GradBucket); idle ranks of the short wave contribute zeros.  Launch with torchrun; rank 0 prints one JSON line.

  python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 tools/train_dp.py [graphs] [epochs] [scale]
"""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist

import gnnome_assembly_b200 as gg
from gnnome_assembly_b200 import prep
from gnnome_assembly_b200.dp import GradBucket, shard_indices
from gnnome_assembly_b200.synth import CHR_LEN, make_assembly_graph

G = int(sys.argv[1]) if len(sys.argv) > 1 else 15
EPOCHS = int(sys.argv[2]) if len(sys.argv) > 2 else 3
SCALE = float(sys.argv[3]) if len(sys.argv) > 3 else 1.0
world, rank = int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("RANK", "0"))
local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)

mine = shard_indices(G, rank, world)                     # wave k -> graph id or None
graphs = {}
for gid in mine:
    if gid is None:
        continue
    gs = make_assembly_graph("chr19", seed=gid, genome_len=int(CHR_LEN["chr19"] * SCALE))
    g = gg.AssemblyGraph(torch.from_numpy(gs.src), torch.from_numpy(gs.dst), gs.num_nodes)
    gg.plan_for(g, dev)
    graphs[gid] = (g, torch.from_numpy(gs.e).to(dev), torch.from_numpy(gs.pe).to(dev), torch.from_numpy(gs.y).to(dev), gs.num_edges)
torch.manual_seed(0)
model = gg.GraphGatedGCNModel(1, 2, 128, 16, 8, 64, True, 16).to(dev)
opt = torch.optim.Adam(model.parameters(), lr=1e-4, fused=True)
bucket = GradBucket(model.parameters())


def epoch():
    edges, losses = 0, []
    for gid in mine:
        opt.zero_grad(set_to_none=True)
        if gid is not None:
            g, e, pe, y, ne = graphs[gid]
            loss, _ = prep.bce_with_logits_and_metrics(model(g, None, e, pe), y, 1 / 16.5)
            loss.backward()
            edges += ne
            losses.append(loss.detach())
        bucket.allreduce_mean(active=gid is not None)
        opt.step()
    return edges, losses


epoch()
torch.cuda.synchronize()
if world > 1:
    dist.barrier()
t0 = time.perf_counter()
for _ in range(EPOCHS):
    edges, losses = epoch()
torch.cuda.synchronize()
if world > 1:
    dist.barrier()
dt = (time.perf_counter() - t0) / EPOCHS
tot = torch.tensor([float(edges)], device=dev, dtype=torch.float64)
if world > 1:
    dist.all_reduce(tot)
if rank == 0:
    print(json.dumps({"workload": f"configs[3]: {G} chr19-like graphs ({SCALE:g}x) over {world} rank(s), L=8 d=128, one step per wave",
                      "waves_per_epoch": len(mine), "epoch_s": dt, "edges_per_epoch": float(tot[0]),
                      "edges_per_s": float(tot[0]) / dt, "last_loss_rank0": float(losses[-1]) if losses else None}))
if world > 1:
    dist.barrier()
    dist.destroy_process_group()

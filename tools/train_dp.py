"""BASELINE.json configs[3]: the epoch loop of train.py:232-258 over 15 chr19-like graphs, sharded round-robin over
the ranks (dp.shard_indices), one optimizer step per wave; the flat gradient arena is all-reduced per
layer segment on a side stream while the backward runs (dp.ArenaSync); idle ranks of the short wave contribute zeros.  Launch with torchrun; rank 0 prints one JSON line.

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
from gnnome_assembly_b200.dp import ArenaSync, shard_indices
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
sync = ArenaSync(model)
wave_sizes = [min(world, G - k * world) for k in range(len(mine))]      # graphs (= active ranks) in wave k


def epoch():
    edges, losses = 0, []
    for k, gid in enumerate(mine):
        opt.zero_grad(set_to_none=True)
        sync.begin(wave_sizes[k])                        # divisor = number of ranks holding a graph in this wave
        if gid is not None:
            g, e, pe, y, ne = graphs[gid]
            loss, _ = prep.bce_with_logits_and_metrics(model(g, None, e, pe), y, 1 / 16.5)
            loss.backward()
            sync.finish()
            edges += ne
            losses.append(loss.detach())
        else:
            sync.idle_step()                             # short wave: zeros through the same collectives
        opt.step()
    return edges, losses


epoch()
torch.cuda.synchronize()
if world > 1:
    dist.barrier()
t0 = time.perf_counter()
ev_a, ev_b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
ev_a.record()
for _ in range(EPOCHS):
    edges, losses = epoch()
ev_b.record()
torch.cuda.synchronize()
if world > 1:
    dist.barrier()
wall = (time.perf_counter() - t0) / EPOCHS
tdev = torch.tensor([ev_a.elapsed_time(ev_b) / 1e3 / EPOCHS], device=dev, dtype=torch.float64)     # device time, max over ranks
tot = torch.tensor([float(edges)], device=dev, dtype=torch.float64)
if world > 1:
    dist.all_reduce(tot)
    dist.all_reduce(tdev, op=dist.ReduceOp.MAX)
dt = float(tdev[0])
if rank == 0:
    print(json.dumps({"workload": f"configs[3]: {G} chr19-like graphs ({SCALE:g}x) over {world} rank(s), L=8 d=128, one step per wave",
                      "waves_per_epoch": len(mine), "wave_sizes": wave_sizes, "epoch_s": dt, "epoch_wall_s": wall,
                      "timing": "CUDA events on each rank, max over ranks", "edges_per_epoch": float(tot[0]),
                      "edges_per_s": float(tot[0]) / dt, "last_loss_rank0": float(losses[-1]) if losses else None}))
if world > 1:
    dist.barrier()
    dist.destroy_process_group()

"""Mini-batch (cluster) path on a chr19-like graph: device sub-plan construction vs the host plan builder, and one
epoch of cluster mini-batch training steps (train.py:282-312) in edges/s.  One JSON line.

  python tools/bench_minibatch.py [num_clusters] [batch_size] [d] [L]
  (round-1 point: 64 8 128 8; the reference's shipped default, hyperparameters.py:4-18: 500 50 256 16)
"""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

import gnnome_assembly_b200 as gg
from gnnome_assembly_b200 import _lib
from gnnome_assembly_b200.minibatch import ClusterGCNSampler, DataLoader
from gnnome_assembly_b200.synth import make_assembly_graph

sys.argv = [a for a in sys.argv if not a.startswith('--')] + [a for a in sys.argv if a.startswith('--')]
K = int(sys.argv[1]) if len(sys.argv) > 1 and not sys.argv[1].startswith('--') else 64
BS = int(sys.argv[2]) if len(sys.argv) > 2 else 8
D = int(sys.argv[3]) if len(sys.argv) > 3 else 128
L = int(sys.argv[4]) if len(sys.argv) > 4 else 8
dev = torch.device("cuda:0")
gs = make_assembly_graph("chr19", seed=0)
g = gg.AssemblyGraph(torch.from_numpy(gs.src.astype(np.int64)), torch.from_numpy(gs.dst.astype(np.int64)), gs.num_nodes)
g.ndata["pe"] = torch.from_numpy(gs.pe)
g.edata["e"] = torch.from_numpy(gs.e)
g.edata["y"] = torch.from_numpy(gs.y)
t0 = time.perf_counter()
plan = gg.plan_for(g, dev)
torch.cuda.synchronize()
t_parent = time.perf_counter() - t0
sampler = ClusterGCNSampler(g, K, device=dev)
N, E = gs.num_nodes, gs.num_edges


def ev_time(fn, n):
    fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t = time.perf_counter()
    a.record()
    for _ in range(n):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / n * 1e3, (time.perf_counter() - t) / n * 1e6     # device us, wall us


# ---- sub-plan construction: half of the graph's clusters
half = torch.cat([sampler.partition_node_ids[sampler.partition_offset[i]:sampler.partition_offset[i + 1]]
                  for i in range(0, K, 2)])
n0 = _lib.launch_count()
sp = plan.subplan(half)
launches = _lib.launch_count() - n0
dev_us, wall_us = ev_time(lambda: plan.subplan(half), 20)
# integer traffic of the compaction: 4 scans read their flags twice (mark gathers hit L2), the
# scatters read the parent arrays once and write the 13 result arrays
alg_bytes = 4 * (2 * (N + 3 * E) * 2 + (N + 1) + 3 * (E + 1) + 6 * E + 4 * N + 9 * sp.num_edges + 4 * sp.num_nodes)
# the host builder on the same sub-graph (what every batch would cost without the device path)
s_np, d_np = sp.array("csrc").cpu().numpy(), sp.array("cdst").cpu().numpy()
t0 = time.perf_counter()
for _ in range(3):
    gg.GraphPlan(torch.from_numpy(s_np), torch.from_numpy(d_np), sp.num_nodes, dev)
torch.cuda.synchronize()
host_us = (time.perf_counter() - t0) / 3 * 1e6

# ---- one epoch of mini-batch steps
torch.manual_seed(0)
model = gg.GraphGatedGCNModel(1, 2, D, 16, L, 64, True, 16).to(dev)
opt = torch.optim.Adam(model.parameters(), lr=1e-4, fused=True)
crit = torch.nn.BCEWithLogitsLoss(pos_weight=torch.tensor([1 / 16.5], device=dev))
loader = DataLoader(g, torch.arange(K), sampler, batch_size=BS, shuffle=True, drop_last=False, num_workers=4)


def epoch():
    edges = 0
    for sub_g in loader:
        sub_g = sub_g.to(dev)
        pred = model(sub_g, None, sub_g.edata["e"], sub_g.ndata["pe"]).squeeze(-1)
        loss = crit(pred, sub_g.edata["y"])
        opt.zero_grad()
        loss.backward()
        opt.step()
        edges += sub_g.num_edges()
    return edges, float(loss)


epoch()
torch.cuda.synchronize()
# where does a batch's time go?  (a) device time of the library's own kernels (event pair per launch), (b) host profile
_lib.profile(True)
epoch()
torch.cuda.synchronize()
_lib.profile(False)
prof = _lib.profile_report()
kern_ms = sum(v[1] for v in prof.values())
kern_launches = sum(v[0] for v in prof.values())
host_top = None
if "--host-profile" in sys.argv:
    import cProfile
    import io
    import pstats
    pr = cProfile.Profile()
    pr.enable()
    epoch()
    torch.cuda.synchronize()
    pr.disable()
    buf = io.StringIO()
    pstats.Stats(pr, stream=buf).sort_stats("cumulative").print_stats(45)
    host_top = buf.getvalue()
    sys.stderr.write(host_top)
t0 = time.perf_counter()
reps = 3
for _ in range(reps):
    edges, last = epoch()
torch.cuda.synchronize()
t_epoch = (time.perf_counter() - t0) / reps
print(json.dumps({
    "workload": f"chr19-like graph N={N} E={E}, {K} clusters (breadth-first chunks), batch_size {BS}, L={L} d={D}",
    "parent_plan_host_ms": t_parent * 1e3,
    "subplan": {"nodes": sp.num_nodes, "edges": sp.num_edges, "launches": launches, "device_us": dev_us,
                "wall_us": wall_us, "algorithmic_bytes": alg_bytes, "gbps": alg_bytes / dev_us / 1e3,
                "host_builder_us_same_subgraph": host_us},
    "epoch": {"batches": len(loader), "edges_in_batches": edges, "wall_ms": t_epoch * 1e3,
              "edges_per_s": edges / t_epoch, "ms_per_batch": t_epoch * 1e3 / len(loader), "last_loss": last,
              "library_kernel_ms_per_batch": kern_ms / len(loader), "library_launches_per_batch": kern_launches / len(loader),
              "kernels_us": {k: [v[0] / len(loader), round(v[1] / v[0] * 1e3, 1)] for k, v in
                             sorted(prof.items(), key=lambda kv: -kv[1][1])}},
}))

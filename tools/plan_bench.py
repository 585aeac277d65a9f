"""Whole-graph plan construction: device builder vs the round-1 host builder, with the device builder's per-kernel
times (event pair per launch).  One JSON line.   python tools/plan_bench.py [chr19|chr21] [target_edges]"""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import gnnome_assembly_b200 as gg
from gnnome_assembly_b200 import _lib
from gnnome_assembly_b200.synth import make_assembly_graph

chrom = sys.argv[1] if len(sys.argv) > 1 else "chr21"
target = int(sys.argv[2]) if len(sys.argv) > 2 else None
dev = torch.device("cuda:0")
gs = make_assembly_graph(chrom, seed=0, pe_dim=0, target_edges=target)
src, dst = torch.from_numpy(gs.src).to(dev), torch.from_numpy(gs.dst).to(dev)
N, E = gs.num_nodes, gs.num_edges
gg.GraphPlan(torch.tensor([0, 1, 2]), torch.tensor([1, 2, 0]), 3, dev)          # lazy init off the clock


def wall(fn, reps=5):
    out = []
    for _ in range(reps):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        p = fn()
        torch.cuda.synchronize()
        out.append((time.perf_counter() - t0) * 1e3)
        del p
    return sorted(out)[len(out) // 2], out[0]


dev_ms, dev_first = wall(lambda: gg.GraphPlan(src, dst, N, dev))
dev_norelabel_ms, _ = wall(lambda: gg.GraphPlan(src, dst, N, dev, relabel=False))
host_ms, _ = wall(lambda: gg.GraphPlan(src, dst, N, dev, host_build=True), reps=3)
_lib.profile(True)
gg.GraphPlan(src, dst, N, dev)
torch.cuda.synchronize()
_lib.profile(False)
prof = _lib.profile_report()
print(json.dumps({"graph": f"{chrom}-like N={N} E={E} (edge list device-resident)",
                  "plan_ms": {"device_builder_median": dev_ms, "device_builder_first_call": dev_first,
                              "device_builder_no_relabel": dev_norelabel_ms, "host_builder_median": host_ms},
                  "device_kernels_us": {k: [v[0], round(v[1] * 1e3 / v[0], 1)] for k, v in sorted(prof.items(), key=lambda kv: -kv[1][1])},
                  "device_kernel_sum_ms": sum(v[1] for v in prof.values())}))

"""Greedy decoding on a chr19-like graph: one iteration of nb_paths walks on the GPU (CUDA events) vs the CPU
oracle (plain Python, as the reference), and a whole get_contigs run.  One JSON line.

  python tools/bench_decode.py [nb_paths]
"""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

import gnnome_assembly_b200 as gg
from gnnome_assembly_b200 import decode
from gnnome_assembly_b200.synth import make_assembly_graph
from oracle import decode_oracle as do

NB = int(sys.argv[1]) if len(sys.argv) > 1 else 50
dev = torch.device("cuda:0")
gs = make_assembly_graph("chr19", seed=0, pe_dim=0)
rng = np.random.default_rng(1)
score = ((gs.y * 2 - 1) * 3 + rng.standard_normal(gs.num_edges)).astype(np.float32)
N, E = gs.num_nodes, gs.num_edges
dg = decode.DecodeGraph(gs.src, gs.dst, N, dev)
score_d = torch.from_numpy(score).to(dev)
prefix_d, rlen_d = torch.from_numpy(gs.prefix_length).to(dev), torch.from_numpy(gs.read_length).to(dev)
visited = torch.zeros((N + 31) // 32, dtype=torch.int32, device=dev)
gen = torch.Generator(device="cuda").manual_seed(0)
idx = decode.sample_edges(dg, score_d, visited, NB, gen)
s_t, d_t, e_t = dg.src[idx], dg.dst[idx], dg.canon_eid[idx]
wb = decode.decode_walks(dg, score_d, prefix_d, rlen_d, visited, s_t, d_t, e_t)
torch.cuda.synchronize()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
reps = 5
a.record()
for _ in range(reps):
    wb = decode.decode_walks(dg, score_d, prefix_d, rlen_d, visited, s_t, d_t, e_t, out=wb)
b.record()
torch.cuda.synchronize()
gpu_ms = a.elapsed_time(b) / reps
beg, length, seq_len, err = wb.host()
steps = int(length.sum())

succs, preds, eid = do.adjacency(gs.src, gs.dst, N)
starts = list(zip(s_t.cpu().tolist(), d_t.cpu().tolist()))
k = min(NB, 10)                                           # bounded CPU sample: the first k walks
t0 = time.perf_counter()
walks, _ = do.walks_for_starts(starts[:k], score, succs, preds, eid, set())
cpu_s = time.perf_counter() - t0
cpu_steps = sum(len(w) for w in walks)
same = all(wb.walk(w, int(beg[w]), int(length[w])).cpu().tolist() == walks[w] for w in range(k))

g = gg.AssemblyGraph(torch.from_numpy(gs.src.astype(np.int64)), torch.from_numpy(gs.dst.astype(np.int64)), N)
g.edata["score"], g.edata["prefix_length"] = torch.from_numpy(score), torch.from_numpy(gs.prefix_length)
g.ndata["read_length"] = torch.from_numpy(gs.read_length)
t0 = time.perf_counter()
contigs = decode.get_contigs(g, None, None, None, nb_paths=NB, len_threshold=20, device="cuda", generator=gen)
total_s = time.perf_counter() - t0
print(json.dumps({
    "workload": f"chr19-like graph N={N} E={E}, nb_paths={NB}, len_threshold=20 (hyperparameters.py)",
    "iteration": {"walks": NB, "walk_nodes": steps, "gpu_ms": gpu_ms, "gpu_walk_nodes_per_s": steps / gpu_ms * 1e3,
                  "cpu_oracle_walks_timed": k, "cpu_walk_nodes_per_s": cpu_steps / cpu_s,
                  "cpu_ms_extrapolated_to_all_walks": cpu_s / max(cpu_steps, 1) * steps * 1e3,
                  "first_walks_identical_to_oracle": bool(same), "err": err},
    "get_contigs": {"contigs": len(contigs), "nodes_in_contigs": sum(len(c) for c in contigs),
                    "longest": max((len(c) for c in contigs), default=0), "wall_s": total_s},
}))

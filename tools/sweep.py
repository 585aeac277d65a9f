"""BASELINE.json config 5: synthetic assembly graphs with 1M / 5M / 20M edges, d = 64 / 128 / 256, one GatedGCN
layer, forward and forward+backward: edges/s and achieved GB/s against the algorithmic-bytes formulas of
SURVEY.md §8d (fwd 4d(3E+25N)+16E, fwd+bwd 4d(11E+67N)+32E).  Writes one JSON line per configuration."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import gnnome_assembly_b200 as gg
from gnnome_assembly_b200.synth import make_assembly_graph

dev = torch.device("cuda:0")
peak = json.load(open("MEASURED_PEAKS.json"))["hbm_gbs"] if os.path.exists("MEASURED_PEAKS.json") else 6650.0
_pos = [a for a in sys.argv[1:] if not a.startswith("--")]
edges = [int(x) for x in (_pos[0].split(",") if len(_pos) > 0 else ["1000000", "5000000", "20000000"])]
dims = [int(x) for x in (_pos[1].split(",") if len(_pos) > 1 else ["64", "128", "256"])]


def timed(fn, n):
    ts = []
    for _ in range(n):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize()
        ts.append(a.elapsed_time(b) / 1e3)
    return float(np.median(ts))


for E_target in edges:
    t0 = time.time()
    g = make_assembly_graph("chr19", seed=0, target_edges=E_target, pe_dim=0)
    graph = gg.AssemblyGraph(torch.from_numpy(g.src), torch.from_numpy(g.dst), g.num_nodes)
    plan = gg.plan_for(graph, dev)
    torch.cuda.synchronize()
    N, E = g.num_nodes, g.num_edges
    t_plan = time.time() - t0
    for d in dims:
        try:
            torch.manual_seed(0)
            layer = gg.layers.GatedGCN_1d(d, d, True).to(dev)
            h = torch.randn(N, d, device=dev, requires_grad=True)
            e = torch.randn(E, d, device=dev, requires_grad=True)

            def fwd():
                with torch.no_grad():
                    layer.forward_internal(plan, h, e)

            def fwd_bwd():
                ho, eo = layer.forward_internal(plan, h, e)
                torch.autograd.backward([ho, eo], [ho, eo])     # upstream grads = the outputs themselves (E x d, N x d)
                h.grad = None; e.grad = None
                for p in layer.parameters():
                    p.grad = None

            for f in (fwd, fwd_bwd):
                f(); f()
            torch.cuda.synchronize()
            n = 7 if E * d < 2e9 else 3
            tf, tb = timed(fwd, n), timed(fwd_bwd, n)
            kern = None
            if "--kernels" in sys.argv:                       # per-kernel device time of one fwd+bwd (event pair per launch)
                from gnnome_assembly_b200 import _lib
                _lib.profile(True)
                fwd_bwd()
                torch.cuda.synchronize()
                _lib.profile(False)
                kern = {k: round(v[1] * 1e3 / v[0], 1) for k, v in sorted(_lib.profile_report().items(), key=lambda kv: -kv[1][1])}
            bf = 4 * d * (3 * E + 25 * N) + 16 * E
            bb = 4 * d * (11 * E + 67 * N) + 32 * E
            print(json.dumps({"E": E, "N": N, "d": d, "L": 1, "fwd_ms": tf * 1e3, "fwd_bwd_ms": tb * 1e3,
                              "fwd_edges_per_s": E / tf, "fwd_bwd_edges_per_s": E / tb,
                              "fwd_gbps": bf / tf / 1e9, "fwd_frac": bf / tf / 1e9 / peak,
                              "fwd_bwd_gbps": bb / tb / 1e9, "fwd_bwd_frac": bb / tb / 1e9 / peak,
                              "graph_and_plan_s": t_plan, "mem_gb": torch.cuda.max_memory_allocated() / 1e9,
                              **({"kernels_us": kern} if kern else {})}), flush=True)
        except torch.OutOfMemoryError as ex:
            print(json.dumps({"E": E, "N": N, "d": d, "error": "out of memory"}), flush=True)
        finally:
            del_names = [k for k in ("layer", "h", "e") if k in dir()]
            layer = h = e = None
            torch.cuda.empty_cache(); torch.cuda.reset_peak_memory_stats()
    del plan, graph, g

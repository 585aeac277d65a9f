"""CPU checks of the driver-facing contract: the reference arm of bench.py prints one well-formed JSON line without
a GPU, and the shared-memory ring protocol of the bulk-staged kernel survives random schedules."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_bench_reference_arm_json_line():
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "1"],
                         capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1
    j = json.loads(lines[0])
    assert j["impl"] == "reference" and j["unit"] == "edges/s" and j["higher_is_better"] is True
    assert j["metric"].startswith("edges/s GatedGCN fwd+bwd") and j["value"] > 0 and j["steps"] == 1
    assert j["cpu_baseline"]["kind"] == "port" and j["cpu_baseline"]["cores"] >= 1 and j["cpu_baseline"]["value"] == j["value"]
    assert j["e2e"] == {"value": j["value"], "unit": "edges/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert j["config"]["workload"].startswith("configs[1]")


def test_bench_without_gpu_fails_loudly():
    import torch
    if torch.cuda.is_available():
        return
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--steps", "1", "--warmup", "1"],
                         capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert out.returncode != 0 and "no CUDA device" in out.stderr


def test_ring_protocol_simulation():
    sys.path.insert(0, os.path.join(ROOT, "tools"))
    import importlib.util
    spec = importlib.util.spec_from_file_location("ring_protocol_sim", os.path.join(ROOT, "tools", "ring_protocol_sim.py"))
    src = open(spec.origin).read().replace("for seed in range(3000): sim(seed)", "")
    ns = {}
    exec(compile(src, spec.origin, "exec"), ns)
    for seed in range(300):
        assert ns["sim"](seed)


def test_gemm_barrier_protocol_simulation():
    """tools/gemm_protocol_sim.py: the stage / accumulator mbarrier protocol of the tcgen05 GEMM with the g_t bulk store
    (empty(s) takes two arrivals: MMA commit + store thread) survives random interleavings; the round-1 protocol (stage
    released by the MMA commit alone) is caught overwriting a stage that the store still reads."""
    spec_path = os.path.join(ROOT, "tools", "gemm_protocol_sim.py")
    src = open(spec_path).read()
    ns = {"__name__": "gemm_protocol_sim"}
    exec(compile(src, spec_path, "exec"), ns)
    import random
    for seed in range(300):
        assert ns["sim"](seed, stages=random.Random(seed).choice([2, 3]), tiles=random.Random(seed + 1).randint(1, 9),
                         kblocks=random.Random(seed + 2).choice([1, 4, 8]))
    bad = src.replace("empty = [Bar(2) for _ in range(stages)]", "empty = [Bar(1) for _ in range(stages)]") \
             .replace("            store_done[s] = True\n            empty[s].arrive()", "            store_done[s] = True")
    assert bad != src
    ns2 = {"__name__": "mutant"}
    exec(compile(bad, spec_path, "exec"), ns2)
    caught = 0
    for seed in range(60):
        try:
            ns2["sim"](seed)
        except AssertionError:
            caught += 1
    assert caught > 0

#!/bin/bash
# quick GPU check with tight time limits (a deadlocked kernel must not burn the budget): parity tests, then bench
tag=${1:-q}; sel=${2:-tests/test_gpu_parity.py}
mkdir -p gpurun_out
timeout ${GG_TEST_LIMIT:-120} python -m pytest $sel -m gpu -x -q --timeout 40 > gpurun_out/${tag}_pytest_full.log 2>&1
rc=$?
tail -8 gpurun_out/${tag}_pytest_full.log | tee gpurun_out/${tag}_pytest.log
if [ $rc -ne 0 ]; then echo "TESTS FAILED (rc=$rc) - bench skipped"; exit 1; fi
timeout 150 python bench.py --steps 10 --warmup 3 > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err; tail -3 gpurun_out/${tag}_bench.err
python - <<EOF
import json
j=json.loads(open("gpurun_out/${tag}_bench.json").read().strip().splitlines()[-1])
print("edges/s", j["value"], "ms/step", j["ms_per_step"], "e2e", j["e2e"]["value"], "step_frac", j["step_roofline"]["frac"])
for k,v in list(j["kernels"].items())[:16]: print(f"{k:26s} {v['avg_us']:8.1f} us  frac {v.get('frac',0):.3f}")
EOF

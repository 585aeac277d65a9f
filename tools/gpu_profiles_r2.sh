#!/bin/bash
# Round-2 evidence run (one GPU): full parity suite, bench line, ncu launch list of one step, ncu --set full at d = 128 / 256 / 64.
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --timeout 600 > gpurun_out/r2_pytest.log 2>&1; tail -3 gpurun_out/r2_pytest.log
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/r2_bench.json 2> gpurun_out/r2_bench.err; tail -2 gpurun_out/r2_bench.err; head -c 300 gpurun_out/r2_bench.json; echo
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv \
    --log-file gpurun_out/r2_ncu_launches.csv python tools/ncu_step.py 8 128 > gpurun_out/r2_ncu_launches.log 2>&1
for d in 128 256 64; do
  timeout 900 ncu --set full --clock-control none --import-source on --profile-from-start off \
      -k regex:'edge_|node_|gemm_tc|gemm_ffma|score|bce' -f -o gpurun_out/r2_full_d$d python tools/ncu_step.py 1 $d > gpurun_out/r2_full_d$d.log 2>&1
  tail -1 gpurun_out/r2_full_d$d.log
done
ls -la gpurun_out/r2_full_d*.ncu-rep

#!/bin/bash
# Round-2 evidence run (one GPU): bench line, ncu launch list of one step, ncu --set full at d = 128 / 256 / 64.
# The .ncu-rep files (~50 MB each) are reduced to their raw / source CSV pages ON THE BOX: gpurun brings back <= 64 MiB.
mkdir -p gpurun_out
what=${@:-bench launches full}
for w in $what; do
  case $w in
    tests)  timeout 1500 python -m pytest tests -m gpu -q --timeout 600 > gpurun_out/r2_pytest.log 2>&1; tail -3 gpurun_out/r2_pytest.log ;;
    bench)  timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/r2_bench.json 2> gpurun_out/r2_bench.err; tail -2 gpurun_out/r2_bench.err; head -c 300 gpurun_out/r2_bench.json; echo ;;
    launches) timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv \
                --log-file gpurun_out/r2_ncu_launches.csv python tools/ncu_step.py 8 128 > gpurun_out/r2_ncu_launches.log 2>&1 ;;
    full)   for d in 128 256 64; do
              timeout 900 ncu --set full --clock-control none --import-source on --profile-from-start off \
                  -k regex:'edge_|node_|gemm_tc|gemm_ffma|score|bce' -f -o /tmp/r2_full_d$d python tools/ncu_step.py 1 $d > gpurun_out/r2_full_d$d.log 2>&1
              tail -1 gpurun_out/r2_full_d$d.log
              ncu -i /tmp/r2_full_d$d.ncu-rep --page raw --csv > gpurun_out/r2_full_d${d}_raw.csv 2>/dev/null
              python tools/ncu_raw_summary.py /tmp/r2_full_d$d.ncu-rep > gpurun_out/r2_full_d${d}_summary.txt 2>&1
            done
            python tools/ncu_traffic.py gpurun_out/r2_ncu_dram_traffic.json /tmp/r2_full_d128.ncu-rep > /dev/null 2>&1
            ncu -i /tmp/r2_full_d128.ncu-rep --page source --csv -k regex:'BnBwdATx|EpiEdgeGate|edge_bwd_a|edge_gate_fwd' > gpurun_out/r2_full_d128_source.csv 2>/dev/null
            gzip -f gpurun_out/r2_full_d128_source.csv gpurun_out/r2_full_d*_raw.csv
            du -sh gpurun_out ;;
  esac
done

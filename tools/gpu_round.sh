#!/bin/bash
# One GPU-box visit: parity tests, bench line, ncu launch list, ncu --set full of one profiled step.
# usage (from repo root, via gpurun): bash tools/gpu_round.sh <tag> [tests] [bench] [launches] [full]
tag=${1:-run}; shift
what=${@:-tests bench launches full}
mkdir -p gpurun_out
for w in $what; do
  case $w in
    tests)    timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/${tag}_pytest.log 2>&1; tail -3 gpurun_out/${tag}_pytest.log ;;
    bench)    timeout 600 python bench.py > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err; head -c 600 gpurun_out/${tag}_bench.json; echo ;;
    ref)      timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/${tag}_bench_ref.json 2>&1 ;;
    launches) timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv \
                --log-file gpurun_out/${tag}_launches.csv python tools/ncu_step.py 8 128 > gpurun_out/${tag}_launches.log 2>&1 ;;
    full)     timeout 900 ncu --set full --clock-control none --import-source on --profile-from-start off \
                -k regex:'edge_|node_|gemm_tc|score|bce' -f -o gpurun_out/${tag}_full python tools/ncu_step.py 1 128 > gpurun_out/${tag}_full.log 2>&1 ;;
  esac
done

#!/bin/bash
# Same-box A/B of library variants on the bench workload (box-to-box variance on this pool is ~5-10 %, larger than most
# single optimisations): default library and each named variant (built by tools/ab_build.sh) alternate, REPS times.
#   bash tools/ab_bench.sh nola other     ->  gpurun_out/ab_<tag>.txt
tag=${AB_TAG:-ab}; reps=${REPS:-2}
mkdir -p gpurun_out
out=gpurun_out/${tag}.txt; : > $out
for rep in $(seq 1 $reps); do
  for v in "" "$@"; do
    lib=$PWD/gnnome_assembly_b200/libgnnome_b200${v:+_$v}.so
    GG_LIB=$lib timeout 200 python bench.py --steps 10 --warmup 3 --no-parity-check --no-baselines 2>/dev/null | python -c "
import sys, json
j = json.loads(sys.stdin.read().strip().splitlines()[-1])
k = j['kernels']
names = ['edge_bwd_a_kernel','gemm_bwd_e_in','gemm_edge_gate','edge_gate_fwd_kernel','edge_bwd_src_kernel','node_agg_fwd_kernel','gemm_dB3','gemm_node_proj','gemm_bwd_h_in','gemm_dWn','node_bwd_apply_kernel','node_bwd_reduce_kernel','node_update_fwd_kernel']
print('variant=%-8s rep=$rep ms/step=%.3f e2e_ms=%.3f | ' % ('${v:-default}', j['ms_per_step'], j['e2e']['ms_per_step']) + ' '.join('%s=%.1f' % (n.replace('_kernel','').replace('gemm_','g_'), k[n]['avg_us']) for n in names if n in k))
" | tee -a $out
  done
done

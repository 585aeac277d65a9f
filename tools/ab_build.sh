#!/bin/bash
# Build a variant of the library with extra -D flags next to the default one (same ABI), for same-box A/B timing:
#   bash tools/ab_build.sh legacy "-DGG_NO_WIDE -DGG_NO_WRES"   ->  gnnome_assembly_b200/libgnnome_b200_legacy.so
#   GG_LIB=$PWD/gnnome_assembly_b200/libgnnome_b200_legacy.so python tools/sweep.py 1000000 128
name=$1; flags=$2
cd "$(dirname "$0")/../gnnome_assembly_b200/csrc" || exit 1
objs=""
for f in gg_plan gg_api gg_prep gg_subgraph gg_decode gg_edge_mlp gg_model gg_plan_device; do
  /usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC -DGG_BUILD $flags -c $f.cu -o /tmp/ab_${name}_$f.o &
  objs="$objs /tmp/ab_${name}_$f.o"
done
wait
/usr/local/cuda/bin/nvcc -shared -o ../libgnnome_b200_${name}.so $objs -lcudart && echo built ../libgnnome_b200_${name}.so

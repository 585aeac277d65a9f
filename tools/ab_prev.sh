#!/bin/bash
# Build the library as of an earlier commit next to the current one (same ABI), for same-box A/B runs:
#   bash tools/ab_prev.sh <commit> [name]   ->  gnnome_assembly_b200/libgnnome_b200_<name>.so   (default name: prev)
commit=$1; name=${2:-prev}
root="$(cd "$(dirname "$0")/.." && pwd)"
tmp=$(mktemp -d)
git -C "$root" archive "$commit" gnnome_assembly_b200/csrc include | tar -x -C "$tmp" || exit 1
cd "$tmp/gnnome_assembly_b200/csrc" || exit 1
objs=""
for f in *.cu; do
  /usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC -DGG_BUILD -c $f -o ${f%.cu}.o &
  objs="$objs ${f%.cu}.o"
done
wait
/usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -shared -o "$root/gnnome_assembly_b200/libgnnome_b200_${name}.so" $objs -lcudart && echo "built libgnnome_b200_${name}.so from $commit"
rm -rf "$tmp"

#!/bin/bash
# tools/build_variant.sh <name> <extra nvcc -D flags...> : tuning builds -> build_variants/libpcg_<name>.so
set -e
name=$1; shift
root=$(cd "$(dirname "$0")/.." && pwd)
out=$root/build_variants; tmp=/tmp/pcgvar_$name; mkdir -p $out $tmp
cd $root/pcgol_b200/csrc
for f in api voxelgrid index icp regiongrowing cloud; do
  nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo -fmad=false \
    -Xcompiler -fPIC,-ffp-contract=off,-fno-fast-math --expt-extended-lambda "$@" -c -o $tmp/$f.o $f.cu &
done
wait
nvcc -gencode arch=compute_100a,code=sm_100a -shared -o $out/libpcg_$name.so $tmp/api.o $tmp/voxelgrid.o $tmp/index.o $tmp/icp.o $tmp/regiongrowing.o $tmp/cloud.o -lcudart
echo built $out/libpcg_$name.so

#!/bin/bash
# whole-library kernel-variant A/B: rebuilds EVERY translation unit with extra -D flags into pimc_jl_b200/libpimc_b200_<tag>.so
# (select it with PIMC_B200_SO=...).  usage: scripts/build_variant_all.sh <tag> [-D...]
set -e
tag=$1; shift
cd "$(dirname "$0")/.."
mkdir -p build/all_$tag
pids=""
for tu in pimc_jl_b200/csrc/*.cu; do
  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -fmad=false -std=c++17 -Xcompiler -fPIC "$@" -Iinclude -c -o build/all_$tag/$(basename $tu .cu).o $tu &
  pids="$pids $!"
done
for p in $pids; do wait $p; done
nvcc -shared -o pimc_jl_b200/libpimc_b200_$tag.so build/all_$tag/*.o -ldl
ls -la pimc_jl_b200/libpimc_b200_$tag.so

#!/bin/bash
# kernel-variant A/B: rebuilds ONE translation unit with extra -D flags and links it with the other objects of build/ into
# pimc_jl_b200/libpimc_b200_<tag>.so (select it with PIMC_B200_SO=...).  usage: scripts/build_variant.sh <tag> <tu.cu> [-D...]
set -e
tag=$1; tu=$2; shift 2
cd "$(dirname "$0")/.."
mkdir -p build/var_$tag
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -fmad=false -std=c++17 -Xcompiler -fPIC "$@" -c -o build/var_$tag/$(basename $tu .cu).o pimc_jl_b200/csrc/$tu
objs=""
for o in build/*.o; do b=$(basename $o); if [ "$b" == "$(basename $tu .cu).o" ]; then objs="$objs build/var_$tag/$b"; else objs="$objs $o"; fi; done
nvcc -shared -o pimc_jl_b200/libpimc_b200_$tag.so $objs -ldl
ls -la pimc_jl_b200/libpimc_b200_$tag.so

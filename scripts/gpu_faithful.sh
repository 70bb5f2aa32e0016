#!/bin/bash
# one gpurun call: faithful-path parity tests, then bench lines of the reference-schedule configurations with the per-phase cycle counters
# usage: scripts/gpu_faithful.sh <tag> [full]
tag=$1; out=gpurun_out; mkdir -p $out
timeout 600 python -m pytest tests -m gpu -q -k "(faithful or interacting) and not one-thread" 2>&1 | tail -60 > $out/${tag}_tests_faithful.log; tail -8 $out/${tag}_tests_faithful.log
[ "$2" = full ] && { timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -25 > $out/${tag}_tests.log; tail -4 $out/${tag}_tests.log; }
for c in c3i c4i c2f c1; do
  extra=""; [ $c = c2f ] && extra="--iters 5000 --therm 2000"
  PIMC_PROF=1 timeout 300 python bench.py --workload $c --steps 3 --warmup 3 --no-cpu-baseline $extra > $out/${tag}_bench_${c}.json 2> $out/${tag}_bench_${c}.err
  python scripts/show_bench.py $out/${tag}_bench_${c}.json; grep "pimc prof" $out/${tag}_bench_${c}.err | tail -2
done

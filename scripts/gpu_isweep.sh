#!/bin/bash
# one gpurun call for the optimistic interacting sweep: GPU tests, then c3i / c4i in the sweep schedule with the phase counters
tag=$1; out=gpurun_out; mkdir -p $out
timeout 1200 python -m pytest tests -m gpu -x -q --durations=8 2>&1 | tail -25 > $out/${tag}_tests.log; tail -4 $out/${tag}_tests.log
for c in c3i c4i; do
  PIMC_PROF=1 timeout 300 python bench.py --workload $c --sched sweep --steps 3 --warmup 3 --no-cpu-baseline > $out/${tag}_bench_${c}_sweep.json 2> $out/${tag}_bench_${c}_sweep.err
  tail -c 900 $out/${tag}_bench_${c}_sweep.json; grep "pimc prof" $out/${tag}_bench_${c}_sweep.err | tail -2
done

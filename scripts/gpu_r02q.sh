#!/bin/bash
tag=r02q; out=gpurun_out; mkdir -p $out
for c in c3i c4i; do for isw in 0 2; do
  PIMC_PROF=1 timeout -k 5 120 python bench.py --workload $c --sched sweep --isweep $isw --steps 3 --warmup 3 --iters 40 --therm 40 --no-cpu-baseline > $out/${tag}_bench_${c}_sweep_isw$isw.json 2> $out/${tag}_bench_${c}_sweep_isw$isw.err
  python scripts/show_bench.py $out/${tag}_bench_${c}_sweep_isw$isw.json; grep "pimc prof" $out/${tag}_bench_${c}_sweep_isw$isw.err | tail -1 | cut -c1-330
done; done

#!/bin/bash
# the round's BASELINE.md table: every SURVEY 8d configuration once, with the CPU baseline beside it
tag=$1; out=gpurun_out; mkdir -p $out
for c in c2 c2s c3 c4 c5 c1 c2f c3i c4i; do
  extra=""; case $c in c2) extra="";; c3i|c4i|c2f|c1) extra="--steps 3";; *) extra="--steps 3";; esac
  timeout -k 5 240 python bench.py --workload $c $extra > $out/${tag}_bench_$c.json 2> $out/${tag}_bench_$c.err
  python scripts/show_bench.py $out/${tag}_bench_$c.json
done
for c in c3i c4i; do
  timeout -k 5 200 python bench.py --workload $c --sched sweep --steps 3 --no-cpu-baseline > $out/${tag}_bench_${c}_sweep.json 2> $out/${tag}_bench_${c}_sweep.err
  python scripts/show_bench.py $out/${tag}_bench_${c}_sweep.json
done
timeout -k 5 240 python bench.py --impl reference --steps 3 --warmup 1 > $out/${tag}_bench_reference.json 2> $out/${tag}_bench_reference.err; tail -c 500 $out/${tag}_bench_reference.json

#!/bin/bash
# ncu --set full of the persistent reference-schedule kernel on the interacting configurations (one launch each);
# the reports are summarised on the box (raw metrics + per-source-line instruction / stall shares) and deleted (64 MiB copy-back limit)
# usage: gpu_ncu_faithful.sh <tag> "<workloads>" [skip]   skip = 0: the thermalisation launch (moves only), 1: the first measured launch
tag=$1; out=gpurun_out; mkdir -p $out; skip=${3:-1}
for c in ${2:-c3i c4i}; do
  timeout 400 ncu --set full --clock-control none --import-source on -k regex:k_run_cells -s $skip -c 1 -f -o /tmp/${tag}_$c python bench.py --workload $c --steps 1 --warmup 1 --iters 30 --therm 30 --no-cpu-baseline > $out/${tag}_ncu_$c.log 2>&1
  python scripts/ncu_summary.py /tmp/${tag}_$c.ncu-rep > $out/${tag}_k_run_cells_${c}_s${skip}_ncu_raw_summary.txt 2>&1
  python scripts/ncu_lines.py /tmp/${tag}_$c.ncu-rep k_run_cells 90 > $out/${tag}_k_run_cells_${c}_s${skip}_source_lines.txt 2>&1
done
ls -la $out | grep $tag

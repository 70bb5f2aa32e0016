#!/bin/bash
# final check of a round: full GPU test suite, smoke, density workloads re-measured, final-state profile of the cell-list kernel
tag=$1; out=gpurun_out; mkdir -p $out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -5 > $out/${tag}_tests.log; tail -2 $out/${tag}_tests.log
timeout 300 python __graft_entry__.py --smoke 2>&1 | tail -1
for c in c3 c4 c3i c4i; do
  timeout 300 python bench.py --workload $c --steps 3 --warmup 3 > $out/${tag}_bench_$c.json 2> $out/${tag}_bench_$c.err
  python scripts/show_bench.py $out/${tag}_bench_$c.json
done
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_run_cells -s 1 -c 1 -f -o /tmp/${tag}_c3i python bench.py --workload c3i --steps 1 --warmup 1 --iters 30 --therm 30 --no-cpu-baseline > $out/${tag}_ncu_c3i.log 2>&1
python scripts/ncu_summary.py /tmp/${tag}_c3i.ncu-rep > $out/${tag}_k_run_cells_c3i_ncu_raw_summary.txt 2>&1
python scripts/ncu_lines.py /tmp/${tag}_c3i.ncu-rep k_run_cells 60 > $out/${tag}_k_run_cells_c3i_source_lines.txt 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file $out/${tag}_launches_c3i.csv python bench.py --workload c3i --steps 1 --warmup 3 --no-cpu-baseline > /dev/null 2>&1
ls $out | grep $tag | wc -l

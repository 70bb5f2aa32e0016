#!/bin/bash
tag=r02e; out=gpurun_out; mkdir -p $out
export PIMC_PROF=1
for args in "c3i 1 3 2 reshape" "c3i 8 3 2 reshape" "c3i 64 3 2 reshape" "c3i 1024 3 2 reshape"; do
  echo "=== $args"; timeout 40 python scripts/probe_isweep.py $args 2>&1 | tail -5
done > $out/${tag}_probe.log 2>&1
cat $out/${tag}_probe.log | grep -E "^===|^run" 
timeout 200 compute-sanitizer --tool memcheck --print-limit 20 python scripts/probe_isweep.py c3i 4 2 2 reshape > $out/${tag}_memcheck.log 2>&1; tail -25 $out/${tag}_memcheck.log
unset PIMC_PROF
timeout 1500 python -m pytest tests -m gpu -x -q --durations=8 2>&1 | tail -25 > $out/${tag}_tests.log; tail -4 $out/${tag}_tests.log
timeout 600 python bench.py --no-cpu-baseline > $out/${tag}_bench_c2.json 2> $out/${tag}_bench_c2.err; python scripts/show_bench.py $out/${tag}_bench_c2.json; tail -3 $out/${tag}_bench_c2.err

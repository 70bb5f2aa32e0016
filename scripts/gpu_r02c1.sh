#!/bin/bash
# C1 (one-particle systems): one-warp CTAs in the sweep schedule -- parity tests that run N = 1 systems, then the c1 bench line
out=gpurun_out; mkdir -p $out
timeout -k 5 60 python -m pytest tests -m gpu -x -q -k "N1 or example_energy" 2>&1 | tail -4 > $out/r02c1_tests.log; tail -2 $out/r02c1_tests.log
timeout -k 5 40 python bench.py --workload c1 --no-cpu-baseline --steps 3 > $out/r02c1_bench_c1.json 2> $out/r02c1_bench_c1.err; python scripts/show_bench.py $out/r02c1_bench_c1.json

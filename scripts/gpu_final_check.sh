#!/bin/bash
# what the driver does at round end, on one box: GPU test suite, smoke, default bench (both arms)
tag=$1; out=gpurun_out; mkdir -p $out
timeout -k 5 700 python -m pytest tests -m gpu -x -q --durations=5 2>&1 | tail -22 > $out/${tag}_tests.log; tail -3 $out/${tag}_tests.log
timeout -k 5 120 python __graft_entry__.py --smoke 2>&1 | tail -3 | tee $out/${tag}_smoke.log
timeout -k 5 300 python bench.py --impl reference > $out/${tag}_bench_reference.json 2> $out/${tag}_bench_reference.err; tail -c 300 $out/${tag}_bench_reference.json; echo
timeout -k 5 300 python bench.py > $out/${tag}_bench_c2.json 2> $out/${tag}_bench_c2.err; python scripts/show_bench.py $out/${tag}_bench_c2.json; tail -3 $out/${tag}_bench_c2.err | cut -c1-300

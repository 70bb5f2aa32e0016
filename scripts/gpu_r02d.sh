#!/bin/bash
tag=r02d; out=gpurun_out; mkdir -p $out
timeout 1500 python -m pytest tests -m gpu -x -q --durations=8 2>&1 | tail -25 > $out/${tag}_tests.log; tail -4 $out/${tag}_tests.log
bash scripts/probe_isweep.sh > $out/${tag}_probe.log 2>&1; grep -E "^===|^run" $out/${tag}_probe.log
timeout 600 python bench.py > $out/${tag}_bench_c2.json 2> $out/${tag}_bench_c2.err; tail -c 1500 $out/${tag}_bench_c2.json; tail -5 $out/${tag}_bench_c2.err

#!/bin/bash
out=gpurun_out; mkdir -p $out
timeout -k 5 60 python -m pytest tests/test_gpu_api.py -m gpu -x -q -k "todo_estimators" 2>&1 | tail -25 > $out/r02zz_tests.log; tail -25 $out/r02zz_tests.log

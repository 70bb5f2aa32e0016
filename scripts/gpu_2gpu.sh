#!/bin/bash
tag=$1; out=gpurun_out; mkdir -p $out
timeout -k 5 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 scripts/comm_2gpu.py > $out/${tag}_comm_2gpu.log 2>&1; tail -3 $out/${tag}_comm_2gpu.log
timeout -k 5 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 5 --warmup 3 --no-cpu-baseline > $out/${tag}_bench_c2_2gpu.json 2> $out/${tag}_bench_c2_2gpu.err
python scripts/show_bench.py $out/${tag}_bench_c2_2gpu.json; tail -3 $out/${tag}_bench_c2_2gpu.err | cut -c1-300
python - $out/${tag}_bench_c2_2gpu.json <<'PY'
import json,sys
try:
    j=json.loads([l for l in open(sys.argv[1]).read().strip().splitlines() if l.startswith("{")][-1]); print("n_gpus", j["n_gpus"], "collective", j["collective"], "e2e", j["e2e"]["value"])
except Exception as ex: print("parse failed", ex)
PY

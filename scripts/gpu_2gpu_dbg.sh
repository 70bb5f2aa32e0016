#!/bin/bash
tag=$1; n=${2:-2}; out=gpurun_out; mkdir -p $out
PIMC_BENCH_TRACE=100 timeout -k 5 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $n --steps 5 --warmup 3 --no-cpu-baseline > $out/${tag}_bench_${n}gpu.json 2> $out/${tag}_bench_${n}gpu.err
grep -E "bench rank 0|Thread|File.*bench.py|NCCL WARN" $out/${tag}_bench_${n}gpu.err | cut -c1-200 | head -30
python scripts/show_bench.py $out/${tag}_bench_${n}gpu.json
python - $out/${tag}_bench_${n}gpu.json <<'PY'
import json,sys
try:
    j=json.loads([l for l in open(sys.argv[1]).read().strip().splitlines() if l.startswith("{")][-1]); print("n_gpus", j["n_gpus"], "value %.4e e2e %.4e" % (j["value"], j["e2e"]["value"]), j["collective"])
except Exception as ex: print("parse failed", ex)
PY

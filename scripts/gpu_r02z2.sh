#!/bin/bash
# r02z2: HEAD on two GPUs -- the bench line as the driver launches it (library communicator, weak scaling), short
out=gpurun_out; mkdir -p $out
timeout -k 5 100 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 3 --warmup 3 --no-cpu-baseline > $out/r02z2_bench_2gpu.json 2> $out/r02z2_bench_2gpu.err
python - <<'PY'
import json
try:
    j=json.loads(open("gpurun_out/r02z2_bench_2gpu.json").read().strip().splitlines()[-1])
    print("2 GPUs: value %.4e e2e %.4e ms/step %.1f" % (j["value"], j["e2e"]["value"], j["ms_per_step"]), j["collective"]["comm_nranks_seen"], j["check"])
except Exception as ex: print("FAILED", ex)
PY
tail -3 $out/r02z2_bench_2gpu.err | cut -c1-300

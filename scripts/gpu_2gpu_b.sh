#!/bin/bash
out=gpurun_out; mkdir -p $out
timeout -k 5 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 scripts/comm_2gpu.py > $out/r02u_comm_2gpu.log 2>&1; grep -E "^rank|Error|error|assert" $out/r02u_comm_2gpu.log | head -8

#!/bin/bash
# session 43 (2 GPUs): sharded path with one shared operand (half-width exchange), exchange caches per width
set -x
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29571 tools/dist_check.py 2>&1 | grep -E "DIST_CHECK|FAIL|shared|again|Error|error" | tee gpurun_out/dist_check_n2_s43.txt

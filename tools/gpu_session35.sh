#!/bin/bash
# session 35 (8 GPUs): copy-stream count at N=8 (arrival-bound: does a second concurrent pull raise NVLink throughput?)
set -x
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
for cs in 2 3; do
PGSD_COPY_STREAMS=$cs timeout 600 $TR --master-port 2956$cs tools/dist_check.py --skip-check --trace 2>&1 | grep -E "TRACE|Error|error" | sed "s/^/[copy_streams=$cs] /" | tee -a gpurun_out/dist_trace_n8_s35.txt
done

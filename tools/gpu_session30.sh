#!/bin/bash
# session 30 (4 GPUs): pack/barrier/pulls on a communication stream beside the own block, one copy stream
set -x
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1"
timeout 900 $TR --master-port 29541 tools/dist_check.py --trace 2>&1 | grep -E "DIST_CHECK|FAIL|TRACE|Error|error" | tee gpurun_out/dist_check_n4_s30.txt
timeout 900 $TR --master-port 29545 bench.py --gpus 4 --steps 20 --warmup 5 > gpurun_out/bench_n4_s30.json 2> gpurun_out/bench_n4_s30.err; tail -2 gpurun_out/bench_n4_s30.err; cut -c1-300 gpurun_out/bench_n4_s30.json

#!/bin/bash
# session 33 (8 GPUs): both exchange paths + per-rank operator build at N=8, headline weak-scaling line
set -x
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
timeout 900 $TR --master-port 29551 tools/dist_check.py --bench --trace 2>&1 | grep -E "DIST_CHECK|FAIL|halo_path|TRACE|Error|error" | tee gpurun_out/dist_check_n8_s33.txt
timeout 900 $TR --master-port 29552 bench.py --gpus 8 --steps 20 --warmup 5 > gpurun_out/bench_n8_s33.json 2> gpurun_out/bench_n8_s33.err; tail -3 gpurun_out/bench_n8_s33.err; cut -c1-300 gpurun_out/bench_n8_s33.json

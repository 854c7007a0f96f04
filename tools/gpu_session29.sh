#!/bin/bash
# session 29 (4 GPUs): correctness of both exchange paths + distributed build at N=4, step timeline of the all-gather path
set -x
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1"
timeout 900 $TR --master-port 29541 tools/dist_check.py --bench --trace 2>&1 | grep -E "DIST_CHECK|FAIL|halo_path|TRACE|Error|error" | tee gpurun_out/dist_check_n4_s29.txt
PGSD_COPY_STREAMS=1 timeout 600 $TR --master-port 29542 tools/dist_check.py --skip-check --trace 2>&1 | grep -E "TRACE|Error|error" | sed 's/^/[copy_streams=1] /' | tee -a gpurun_out/dist_check_n4_s29.txt
PGSD_COPY_STREAMS=3 timeout 600 $TR --master-port 29543 tools/dist_check.py --skip-check --trace 2>&1 | grep -E "TRACE|Error|error" | sed 's/^/[copy_streams=3] /' | tee -a gpurun_out/dist_check_n4_s29.txt
PGSD_SHARD_MODE=gather timeout 600 $TR --master-port 29544 bench.py --gpus 4 --steps 20 --warmup 5 2>/dev/null | cut -c1-250 | sed 's/^/[gather] /' | tee -a gpurun_out/dist_check_n4_s29.txt
timeout 900 $TR --master-port 29545 bench.py --gpus 4 --steps 20 --warmup 5 > gpurun_out/bench_n4_s29.json 2> gpurun_out/bench_n4_s29.err; tail -2 gpurun_out/bench_n4_s29.err; cut -c1-300 gpurun_out/bench_n4_s29.json

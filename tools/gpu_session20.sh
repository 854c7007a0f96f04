#!/bin/bash
# session 20 (2 GPUs): row-range operator build parity, halo all-to-all path, sharded bench with the distributed build
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_dist_build.py -m gpu -q 2>&1 | tail -8
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 tools/dist_check.py --bench 2>&1 | grep -E "DIST_CHECK|FAIL|halo_path|Error|error" | tee gpurun_out/dist_check_n2_s20.txt
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29534 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/bench_n2_s20.json 2> gpurun_out/bench_n2_s20.err; tail -3 gpurun_out/bench_n2_s20.err; cat gpurun_out/bench_n2_s20.json

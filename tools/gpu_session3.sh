#!/bin/bash
# N = 2 GPUs: sharded-path correctness under NCCL, then the scaling bench at N = 1 and 2.
set -x
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/gpus.txt
nvidia-smi topo -m >> gpurun_out/gpus.txt 2>&1
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 \
    tools/dist_check.py > gpurun_out/dist_check.log 2>&1
tail -12 gpurun_out/dist_check.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29534 \
    bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/bench_n2.json 2> gpurun_out/bench_n2.err
cat gpurun_out/bench_n2.json; tail -5 gpurun_out/bench_n2.err
timeout 600 python bench.py --gpus 1 --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err
cat gpurun_out/bench_n1.json

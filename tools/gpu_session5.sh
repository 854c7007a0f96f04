#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_autograd.py -m gpu -q 2>&1 | tail -40 > gpurun_out/pytest_autograd.log
tail -25 gpurun_out/pytest_autograd.log
timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -15 > gpurun_out/pytest_gpu.log
tail -6 gpurun_out/pytest_gpu.log
timeout 900 python tools/sweep_spmm.py > gpurun_out/sweep.log 2>&1
grep -E "segmented" gpurun_out/sweep.log; tail -3 gpurun_out/sweep.log
timeout 900 python tools/bench_configs.py > gpurun_out/configs.log 2>&1
tail -8 gpurun_out/configs.log

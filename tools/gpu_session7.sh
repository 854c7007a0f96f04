#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "group_per_row or variants" 2>&1 | tail -30 > gpurun_out/pytest_groups.log
tail -12 gpurun_out/pytest_groups.log
timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -15 > gpurun_out/pytest_gpu.log
tail -6 gpurun_out/pytest_gpu.log
timeout 900 python tools/sweep_spmm.py > gpurun_out/sweep.log 2>&1
grep -E "spmm|segmented" gpurun_out/sweep.log | cut -c1-200
timeout 900 python tools/bench_configs.py > gpurun_out/configs.log 2>&1
tail -5 gpurun_out/configs.log | cut -c1-200

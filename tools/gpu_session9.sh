#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "dense_tensor_core" 2>&1 | tail -30 > gpurun_out/pytest_tc.log
tail -12 gpurun_out/pytest_tc.log
timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -15 > gpurun_out/pytest_gpu.log
tail -6 gpurun_out/pytest_gpu.log
SWEEP_SEGMENTS=0 timeout 600 python tools/sweep_spmm.py > gpurun_out/sweep.log 2>&1
grep -E "dense|layer" gpurun_out/sweep.log | cut -c1-200
timeout 900 python tools/bench_configs.py > gpurun_out/configs.log 2>&1
tail -5 gpurun_out/configs.log | cut -c1-220
timeout 600 ncu --set full --clock-control none --import-source on -k regex:dense_tc_ws -s 3 -c 1 \
    -o gpurun_out/prof_dense_ws -f python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/prof_dense.log 2>&1
tail -2 gpurun_out/prof_dense.log

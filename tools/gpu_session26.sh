#!/bin/bash
# session 26: dense_tma (TMA-store epilogue, ring invariants): parity, sweep, configs, bench
set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_dense_tma.py tests/test_gpu_parity.py tests/test_gpu_fused.py tests/test_gpu_autograd.py -m gpu -q 2>&1 | tail -8
PGSD_SWEEP_VARIANTS=0x10010,0x20010,0x40010 timeout 600 python tools/sweep_dense.py 2>&1 | grep -E '^\{' | grep -v '"variant": 1,' | tee gpurun_out/sweep_dense_s26.jsonl | cut -c1-140
timeout 600 python tools/bench_configs.py 2>&1 | grep -E '^\{' | tee gpurun_out/configs_s26.jsonl | cut -c1-400

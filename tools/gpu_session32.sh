#!/bin/bash
# session 32: dense sweep after the tanh epilogue mode (regression check) + model parity
set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "sgcn or inception or dense" 2>&1 | tail -3
timeout 600 python tools/sweep_dense.py 2>&1 | grep -E '^\{' | grep -v '"variant": 1,' | tee gpurun_out/sweep_dense_s32.jsonl | cut -c1-140
timeout 600 python tools/sweep_dense.py 2>&1 | grep -E '^\{' | grep -v '"variant": 1,' | cut -c1-140

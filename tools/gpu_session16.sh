#!/bin/bash
# session 16: TMA-fed warp-specialised dense kernel (variant 16): parity + sweep; SGCN merged transform
set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_dense_tma.py -m gpu -q --maxfail=40 2>&1 | tail -30
timeout 600 python tools/sweep_dense.py 2>&1 | grep -E '^\{' | tee gpurun_out/sweep_dense_s16.jsonl
timeout 900 python -m pytest tests -m gpu -q --ignore tests/test_gpu_dense_tma.py 2>&1 | tail -8
ls -la gpurun_out

#!/bin/bash
# session 27: dense_tma epilogue staging rounds / sets: parity + sweep (bits 24-25: 1 = 1 box x 1 set, 2 = 1 x 2, 3 = tile x 1)
set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_dense_tma.py tests/test_gpu_parity.py -m gpu -q -k "dense or magnet_golden or inception or sgcn or model" 2>&1 | tail -8
PGSD_SWEEP_VARIANTS=0x1000010,0x2000010,0x3000010,0x10010 timeout 600 python tools/sweep_dense.py 2>&1 | grep -E '^\{' | grep -v '"variant": 1,' | tee gpurun_out/sweep_dense_s27.jsonl | cut -c1-140

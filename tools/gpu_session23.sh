#!/bin/bash
# session 23: dense_tma with back-to-back UTCHMMA issue (uniform operands, 32-bit descriptor stepping): parity + ring sweep
set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_dense_tma.py tests/test_gpu_parity.py -m gpu -q -k "dense or magnet_golden or inception or sgcn" 2>&1 | tail -5
PGSD_SWEEP_VARIANTS=0x4210,0x5210,0x6210,0x7210,0x6310,0x5310,0x10010,0x20010,0x40010 timeout 600 python tools/sweep_dense.py 2>&1 | grep -E '^\{' | grep -v '"variant": 1,' | tee gpurun_out/sweep_dense_s23.jsonl | cut -c1-140

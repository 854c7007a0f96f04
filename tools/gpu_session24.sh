#!/bin/bash
# session 24: dense_tma converter groups (1/2/4) + 8 epilogue warps: parity + sweep
set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_dense_tma.py tests/test_gpu_parity.py -m gpu -q -k "dense or magnet_golden or inception or sgcn" 2>&1 | tail -5
# bits 20-22 converter groups, 8-11 lo slots, 12-15 landing stages
PGSD_SWEEP_VARIANTS=0x100010,0x200010,0x400010,0x200210,0x200410,0x400410,0x404410,0x210010,0x220010,0x240010 timeout 600 python tools/sweep_dense.py 2>&1 | grep -E '^\{' | grep -v '"variant": 1,' | tee gpurun_out/sweep_dense_s24.jsonl | cut -c1-140

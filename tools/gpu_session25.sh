#!/bin/bash
# session 25: dense_tma with TMA-store epilogue: parity + sweep (+ knock-outs, ring depths)
set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_dense_tma.py tests/test_gpu_parity.py -m gpu -q -k "dense or magnet_golden or inception or sgcn or model" 2>&1 | tail -8
PGSD_SWEEP_VARIANTS=0x10010,0x20010,0x40010,0x3010,0x5010,0x100010,0x400010 timeout 600 python tools/sweep_dense.py 2>&1 | grep -E '^\{' | grep -v '"variant": 1,' | tee gpurun_out/sweep_dense_s25.jsonl | cut -c1-140

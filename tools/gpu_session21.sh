#!/bin/bash
# session 21: dense_tma with a short lo ring + deeper landing ring: parity, ring-depth sweep
set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_dense_tma.py tests/test_gpu_parity.py -m gpu -q -k "dense or magnet_golden or inception or sgcn" 2>&1 | tail -5
# lo slots 1..3 x landing stages 4..8
PGSD_SWEEP_VARIANTS=0x4110,0x4210,0x5210,0x6210,0x7210,0x6310,0x5310,0x8210,0x8110 timeout 600 python tools/sweep_dense.py 2>&1 | grep -E '^\{' | tee gpurun_out/sweep_dense_s21.jsonl

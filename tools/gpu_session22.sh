#!/bin/bash
# session 22: where does dense_tma spend its time? timing-only knock-outs (variant bits 16-18)
set -x
mkdir -p gpurun_out
PGSD_SWEEP_VARIANTS=0x10010,0x20010,0x40010,0x30010,0x50010,0x60010,0x70010 timeout 600 python tools/sweep_dense.py 2>&1 | grep -E '^\{' | grep -v '"variant": 1,' | tee gpurun_out/sweep_dense_s22.jsonl | cut -c1-160

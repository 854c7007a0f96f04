#!/bin/bash
# session 18: sparse GPU preprocessing (n3) parity + timings; dense_tma with ld/st.shared converters
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_prep.py -m gpu -q --maxfail=20 2>&1 | tail -30
timeout 600 python tools/bench_prep.py 2>&1 | grep -E '^\{' | tee gpurun_out/prep_s18.jsonl
timeout 600 python -m pytest tests/test_gpu_dense_tma.py -m gpu -q 2>&1 | tail -4
timeout 600 python tools/sweep_dense.py 2>&1 | grep -E '^\{' | grep '"variant": 16' | tee gpurun_out/sweep_dense_s18.jsonl
ls -la gpurun_out

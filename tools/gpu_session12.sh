#!/bin/bash
# session 12: model wrapper parity + model-level timings
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -5
timeout 600 python tools/bench_configs.py 2>&1 | grep -E '^\{' | tee gpurun_out/configs_s12.jsonl

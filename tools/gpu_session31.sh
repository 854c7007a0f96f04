#!/bin/bash
# session 31: model wrappers of configs 3 and 4 (fused x0+x1+x2, fused tanh): parity + timings
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -8
timeout 600 python tools/bench_configs.py 2>&1 | grep -E '^\{|Error|error' | tee gpurun_out/configs_s31.jsonl | cut -c1-330

#!/bin/bash
# session 39: attention layers after batching the transforms and the single-pass grouped softmax kernel
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_motifs.py tests/test_gpu_parity.py -m gpu -q -k "snea or sdr or sigat or sdgnn or models_golden" 2>&1 | tail -8
timeout 600 python tools/bench_configs.py 2>&1 | grep -E '^\{|Error|error' | grep -E "SNEA|SDR|rror" | tee gpurun_out/configs_s39.jsonl | cut -c1-400

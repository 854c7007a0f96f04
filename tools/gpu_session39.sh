#!/bin/bash
# session 39/40: attention layers (batched transforms, grouped single-pass softmax, in-kernel softmax aggregation)
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_motifs.py tests/test_gpu_parity.py -m gpu -q -k "gat_aggregation or sdr_layer_golden or sdgnn_and_sigat or snea_golden" --durations=5 2>&1 | tail -8
timeout 600 python tools/bench_configs.py 2>&1 | grep -E '^\{|Error|error' | grep -E "SNEA|SDR|rror" | tee gpurun_out/configs_s39.jsonl | cut -c1-400
PGSD_GAT_FUSED=0 timeout 600 python tools/bench_configs.py 2>&1 | grep -E '^\{' | grep -E "SDR" | sed 's/^/[two-kernel route] /' | cut -c1-300

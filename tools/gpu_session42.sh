#!/bin/bash
# session 42: edge-softmax kernel with register-cached activations (no online rescaling): parity + timing
set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_motifs.py -m gpu -q -k "snea or sdr or gat_aggregation or sdgnn_and_sigat" 2>&1 | tail -4
timeout 600 python tools/bench_configs.py 2>&1 | grep -E '^\{' | grep -E "SNEA|SDR" | tee gpurun_out/configs_s42.jsonl | cut -c1-300

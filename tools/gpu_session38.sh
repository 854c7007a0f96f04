#!/bin/bash
# session 38: attention layers (SNEAConv, SDRLayer/GATConv) at config-4 scale: first measurements
set -x
mkdir -p gpurun_out
timeout 600 python tools/bench_configs.py 2>&1 | grep -E '^\{|Error|error' | grep -E "SNEA|SDR|rror" | tee gpurun_out/configs_s38.jsonl | cut -c1-400

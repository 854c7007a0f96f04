#!/bin/bash
# session 40: final validation of the round: full GPU suite, smoke(), headline bench, all configs
set -x
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q 2>&1 | tail -4
timeout 300 python __graft_entry__.py --smoke 2>&1 | tail -2
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_s40.json 2> gpurun_out/bench_s40.err; tail -2 gpurun_out/bench_s40.err; cut -c1-200 gpurun_out/bench_s40.json
timeout 600 python tools/bench_configs.py 2>&1 | grep -E '^\{' | tee gpurun_out/configs_s40.jsonl | cut -c1-260

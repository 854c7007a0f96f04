#!/bin/bash
# session 37: full GPU suite (incl. full-size configs 3 / 4, uncached plan reuse), headline bench, configs
set -x
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q 2>&1 | tail -8
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_s37.json 2> gpurun_out/bench_s37.err; tail -2 gpurun_out/bench_s37.err; cut -c1-200 gpurun_out/bench_s37.json
timeout 600 python tools/bench_configs.py 2>&1 | grep -E '^\{' | tee gpurun_out/configs_s37.jsonl | cut -c1-300

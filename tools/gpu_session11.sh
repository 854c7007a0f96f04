#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -8 > gpurun_out/pytest_gpu.log
tail -5 gpurun_out/pytest_gpu.log
SWEEP_SEGMENTS=0 timeout 600 python tools/sweep_spmm.py > gpurun_out/sweep.log 2>&1
grep -E "dense|layer" gpurun_out/sweep.log | cut -c1-200
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/bench.json 2> gpurun_out/bench.err
python - <<'PY'
import json
d = json.load(open('gpurun_out/bench.json'))
print({k: d[k] for k in ('value', 'ms_per_step', 'gpu_launches', 'cold_ms_per_step')}, d['roofline']['kernel_ms'], d['roofline']['layer'], d['e2e']['ms_per_step'], d['cpu_baseline']['value'])
PY
tail -2 gpurun_out/bench.err
timeout 900 python tools/bench_configs.py > gpurun_out/configs.log 2>&1
tail -5 gpurun_out/configs.log | cut -c1-220

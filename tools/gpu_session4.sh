#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -25 > gpurun_out/pytest_gpu.log
tail -8 gpurun_out/pytest_gpu.log
timeout 900 python tools/sweep_spmm.py > gpurun_out/sweep.log 2>&1
grep -E "segmented|layer|spmm2" gpurun_out/sweep.log; tail -3 gpurun_out/sweep.log
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/bench.json 2> gpurun_out/bench.err
python - <<'PY'
import json
d = json.load(open('gpurun_out/bench.json'))
print({k: d[k] for k in ('value', 'ms_per_step', 'gpu_launches')}, d['e2e'], d['cpu_baseline']['value'])
PY
tail -3 gpurun_out/bench.err
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err
cat gpurun_out/bench_ref.json | cut -c1-600

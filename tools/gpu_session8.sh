#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -15 > gpurun_out/pytest_gpu.log
tail -6 gpurun_out/pytest_gpu.log
timeout 900 python tools/sweep_spmm.py > gpurun_out/sweep.log 2>&1
grep -E "spmm|layer" gpurun_out/sweep.log | cut -c1-200
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/bench.json 2> gpurun_out/bench.err
python - <<'PY'
import json
d = json.load(open('gpurun_out/bench.json'))
print({k: d[k] for k in ('value', 'ms_per_step', 'gpu_launches')}, d['roofline'], d['e2e']['ms_per_step'])
PY
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'spmm|dense|k_' -c 40 --csv \
    --log-file gpurun_out/launches.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_ncu.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:spmm_groups -s 3 -c 2 \
    -o gpurun_out/prof_spmm_groups -f python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/prof_spmm.log 2>&1
tail -2 gpurun_out/prof_spmm.log
timeout 900 python tools/bench_configs.py > gpurun_out/configs.log 2>&1
tail -5 gpurun_out/configs.log | cut -c1-220

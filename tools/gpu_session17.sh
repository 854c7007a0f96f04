#!/bin/bash
# session 17: headline bench + configs with the TMA dense kernel as default; launch list; ncu of dense_tma
set -x
mkdir -p gpurun_out
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_s17.json 2> gpurun_out/bench_s17.err; tail -2 gpurun_out/bench_s17.err; cat gpurun_out/bench_s17.json
timeout 600 python tools/bench_configs.py 2>&1 | grep -E '^\{' | tee gpurun_out/configs_s17.jsonl
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 150 --csv --log-file gpurun_out/launches_s17.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline > /dev/null 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:dense_tma -s 3 -c 1 -o gpurun_out/prof_dense_tma python bench.py --steps 2 --warmup 3 --no-cpu-baseline > /dev/null 2> gpurun_out/ncu_dense.err; tail -2 gpurun_out/ncu_dense.err
ls -la gpurun_out

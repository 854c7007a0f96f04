#!/bin/bash
# session 28: headline bench + configs + launch list + ncu --set full of the new dense_tma kernel and of the shared-operand aggregation
set -x
mkdir -p gpurun_out
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_s28.json 2> gpurun_out/bench_s28.err; tail -2 gpurun_out/bench_s28.err; cat gpurun_out/bench_s28.json
timeout 600 python tools/bench_configs.py 2>&1 | grep -E '^\{' | tee gpurun_out/configs_s28.jsonl | cut -c1-300
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 150 --csv --log-file gpurun_out/launches_s28.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline > /dev/null 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:dense_tma -s 3 -c 1 -o gpurun_out/prof_dense_tma_s28 python bench.py --steps 2 --warmup 3 --no-cpu-baseline > /dev/null 2> gpurun_out/ncu_dense_s28.err; tail -2 gpurun_out/ncu_dense_s28.err
timeout 600 ncu --set full --clock-control none --import-source on -k regex:spmm_groups -s 12 -c 1 -o gpurun_out/prof_spmm_shared_s28 python bench.py --steps 2 --warmup 3 --no-cpu-baseline > /dev/null 2> gpurun_out/ncu_spmm_s28.err; tail -2 gpurun_out/ncu_spmm_s28.err
ls -la gpurun_out | tail -8

#!/bin/bash
# One gpurun call: tests, smoke, sweep, bench, ncu launch list + full capture of the top kernel.
set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/gpu.txt 2>&1
nproc > gpurun_out/host.txt; free -g >> gpurun_out/host.txt; lscpu | grep -E "Model name|^CPU\(s\)" >> gpurun_out/host.txt
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -40 > gpurun_out/pytest_gpu.log
timeout 300 python __graft_entry__.py --smoke > gpurun_out/smoke.log 2>&1
timeout 600 python tools/sweep_spmm.py > gpurun_out/sweep.log 2>&1
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/bench.json 2> gpurun_out/bench.err
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'spmm|dense|k_' -c 80 --csv \
    --log-file gpurun_out/launches.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_ncu.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:spmm_rows -s 3 -c 2 \
    -o gpurun_out/prof_spmm -f python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/prof_spmm.log 2>&1
ls -la gpurun_out
tail -5 gpurun_out/pytest_gpu.log; cat gpurun_out/smoke.log | tail -3; cat gpurun_out/sweep.log | tail -30; cat gpurun_out/bench.json; tail -5 gpurun_out/bench.err

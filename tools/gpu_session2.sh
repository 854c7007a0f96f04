#!/bin/bash
# tcgen05 dense path bring-up: its own tests first (bounded), then everything else.
set -x
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "dense_tensor_core" 2>&1 | tail -40 > gpurun_out/pytest_tc.log
tail -15 gpurun_out/pytest_tc.log
timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -40 > gpurun_out/pytest_gpu.log
tail -15 gpurun_out/pytest_gpu.log
timeout 600 python tools/sweep_spmm.py > gpurun_out/sweep.log 2>&1
grep -E "dense|layer" gpurun_out/sweep.log
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/bench.json 2> gpurun_out/bench.err
cat gpurun_out/bench.json; tail -3 gpurun_out/bench.err
timeout 600 ncu --set full --clock-control none --import-source on -k regex:dense_tc -s 3 -c 1 \
    -o gpurun_out/prof_dense_tc -f python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/prof_dense.log 2>&1
tail -3 gpurun_out/prof_dense.log

#!/bin/bash
# session 15: predicated gathers restored in the standalone kernels (raw words), fused kernel: unconditional clamped gathers + predicated FFMAs; variants
set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_fused.py -m gpu -q --maxfail=30 2>&1 | tail -25
PGSD_FUSED_LAYER=1 timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/bench_fused.json 2> gpurun_out/bench_fused.err; tail -3 gpurun_out/bench_fused.err; cat gpurun_out/bench_fused.json
for v in 1 2 3; do PGSD_FUSED_LAYER=1 PGSD_FUSED_VARIANT=$v timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/bench_fused_v$v.json 2> gpurun_out/bench_fused_v$v.err; cat gpurun_out/bench_fused_v$v.json; done
PGSD_FUSED_LAYER=0 timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/bench_unfused.json 2> gpurun_out/bench_unfused.err; cat gpurun_out/bench_unfused.json
timeout 900 python -m pytest tests -m gpu -q --ignore tests/test_gpu_fused.py 2>&1 | tail -8
timeout 600 python tools/bench_configs.py 2>&1 | grep -E '^\{' | tee gpurun_out/configs_s15.jsonl
timeout 600 ncu --set full --clock-control none --import-source on -k regex:spmm_groups -s 4 -c 1 -o gpurun_out/prof_c3_spmm python tools/run_c3.py > /dev/null 2> gpurun_out/ncu_c3.err; tail -2 gpurun_out/ncu_c3.err
PGSD_FUSED_LAYER=1 timeout 600 ncu --set full --clock-control none --import-source on -k regex:magnet_layer_fused -s 3 -c 1 -o gpurun_out/prof_fused_v3 python bench.py --steps 2 --warmup 3 --no-cpu-baseline > /dev/null 2> gpurun_out/ncu_fused.err; tail -2 gpurun_out/ncu_fused.err
ls -la gpurun_out

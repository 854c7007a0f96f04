#!/bin/bash
# session 19: shared-operand aggregation (x_real is x_imag): parity + bench line with `shared_input`
set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "shared_operand or variants_agree or group_per_row or magnet_vs_oracle" 2>&1 | tail -5
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_s19.json 2> gpurun_out/bench_s19.err; tail -2 gpurun_out/bench_s19.err; cat gpurun_out/bench_s19.json

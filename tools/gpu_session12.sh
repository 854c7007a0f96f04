#!/bin/bash
# memory-safety pass: the small parity tests under compute-sanitizer memcheck (each kernel family once)
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "bf16" 2>&1 | tail -6
timeout 1500 compute-sanitizer --tool memcheck --error-exitcode 9 --print-limit 20 \
    python -m pytest tests/test_gpu_parity.py -m gpu -q -x \
    -k "magnet_golden or group_per_row or hub_rows or snea_golden or sdr_layer or dgcn_and_simpa or conv_base_and_dimpa or sgcn_golden or (dense_tensor_core_path and 130) or (dense_tensor_core_bf16 and 515) or digcn_golden" \
    > gpurun_out/memcheck.log 2>&1
echo "memcheck exit: $?"
grep -E "ERROR SUMMARY|Invalid|out of bounds|passed|failed|error" gpurun_out/memcheck.log | tail -15

#!/bin/bash
# session 36: compute-sanitizer memcheck over the kernels added or rewritten this session (dense_tma with TMA
# stores and the lo ring, shared-operand aggregation, row-range builder, motif counts, model wrappers)
set -x
mkdir -p gpurun_out
timeout 1200 compute-sanitizer --tool memcheck python -m pytest tests/test_gpu_dense_tma.py tests/test_gpu_motifs.py tests/test_gpu_dist_build.py tests/test_gpu_parity.py -m gpu -q -x \
  -k "(dense_tma and not 70000) or motif_lists or models_golden or (row_range and not scale) or shared_operand or inception_model or sgcn_model or magnet_golden" > gpurun_out/memcheck_s36.txt 2>&1
tail -6 gpurun_out/memcheck_s36.txt

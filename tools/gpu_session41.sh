#!/bin/bash
# session 41: ncu --set full of the shared-operand aggregation and of the grouped edge-softmax kernel
set -x
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"spmm_groups|edge_softmax" -s 2 -c 4 -o gpurun_out/prof_extra_s41 python tools/prof_extra.py > /dev/null 2> gpurun_out/ncu_extra_s41.err; tail -3 gpurun_out/ncu_extra_s41.err
ls -la gpurun_out/prof_extra_s41.ncu-rep

#!/bin/bash
# One parametrised driver for every gpurun call (replaces the numbered per-session scripts of round 1).
#   gpurun [--gpus N] -- 'bash tools/gpu_session.sh <recipe> [args]'
# Output goes to gpurun_out/ (merged back by gpurun); what is worth keeping is copied to profiles/ by hand.
set -x
mkdir -p gpurun_out
R=${1:-help}; shift || true
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
case "$R" in
  tests)       # pytest -m gpu [pytest args]
    timeout 1500 python -m pytest tests -m gpu -x -q "$@" 2>&1 | tail -15 ;;
  bench)       # single-GPU bench line -> gpurun_out/bench.json
    timeout 900 python bench.py --steps 20 --warmup 5 "$@" > gpurun_out/bench.json 2> gpurun_out/bench.err
    tail -3 gpurun_out/bench.err; cut -c1-600 gpurun_out/bench.json ;;
  dist)        # dist N [dist_check args]: sharded == single-GPU, optional --bench / --trace
    N=$1; shift
    timeout 900 $TR --nproc-per-node $N --master-port 29551 tools/dist_check.py "$@" 2>&1 \
      | grep -E "DIST_CHECK|FAIL|halo_path|TRACE|Error|error|Traceback" | tee gpurun_out/dist_check_n$N.txt ;;
  distbench)   # distbench N: bench.py under torchrun
    N=$1; shift
    timeout 900 $TR --nproc-per-node $N --master-port 29552 bench.py --gpus $N --steps 20 --warmup 5 "$@" \
      > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err
    tail -3 gpurun_out/bench_n$N.err; cut -c1-400 gpurun_out/bench_n$N.json ;;
  probe)       # probe N: transports of the all-gather, alone and beside the aggregation
    N=$1; shift
    timeout 600 $TR --nproc-per-node $N --master-port 29553 tools/exchange_probe.py "$@" \
      > gpurun_out/probe_n$N.jsonl 2> gpurun_out/probe_n$N.err
    tail -5 gpurun_out/probe_n$N.err; cat gpurun_out/probe_n$N.jsonl ;;
  interfere)   # 2 GPUs: slowdown of the aggregation beside each kind of NVLink transfer
    timeout 600 $TR --nproc-per-node 2 --master-port 29555 tools/interference_probe.py "$@" \
      > gpurun_out/interference.jsonl 2> gpurun_out/interference.err
    tail -3 gpurun_out/interference.err; cat gpurun_out/interference.jsonl ;;
  sweep)       # sweep N 'cfg' 'cfg' ...: push-exchange knobs on the sharded step (tools/shard_sweep.py)
    N=$1; shift
    timeout 900 $TR --nproc-per-node $N --master-port 29554 tools/shard_sweep.py "$@" \
      > gpurun_out/sweep_n$N.jsonl 2> gpurun_out/sweep_n$N.err
    tail -3 gpurun_out/sweep_n$N.err; cat gpurun_out/sweep_n$N.jsonl ;;
  configs)     # the other BASELINE configs
    timeout 900 python tools/bench_configs.py "$@" > gpurun_out/configs.jsonl 2> gpurun_out/configs.err
    tail -3 gpurun_out/configs.err; cat gpurun_out/configs.jsonl ;;
  launches)    # ncu launch list of the bench command (cold-cache, serialised: compare shares)
    timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv \
      --log-file gpurun_out/launches.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/launches_bench.log 2>&1
    tail -3 gpurun_out/launches_bench.log ;;
  ncu)         # ncu <kernel-regex> <out-name> -- <command...>: one --set full capture
    K=$1; O=$2; shift 3
    timeout 900 ncu --set full --clock-control none --import-source on -k "regex:$K" -c 1 -s ${NCU_SKIP:-2} \
      -o gpurun_out/$O -f "$@" > gpurun_out/$O.log 2>&1
    tail -3 gpurun_out/$O.log; ls -la gpurun_out/$O.ncu-rep ;;   # summarise here: python tools/ncu_summary.py <rep> <out.json> "<cmd>" "<note>"
  seq)         # seq 'recipe args' 'recipe args' ...: several recipes in one call
    for step in "$@"; do bash tools/gpu_session.sh $step; done ;;
  *) echo "recipes: tests bench dist distbench probe sweep configs launches ncu seq" ;;
esac

#!/bin/bash
# usage: bash tools/gpu_session_multi.sh N   (run under gpurun --gpus N)
N=${1:-2}
set -x
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 \
    tools/dist_check.py > gpurun_out/dist_check_n$N.log 2>&1
grep -E "DIST_CHECK|FAIL|Error|error" gpurun_out/dist_check_n$N.log | head
for mode in ${MODES:-ring nccl}; do
  PGSD_EXCHANGE=$([ $mode = nccl ] && echo nccl || echo pull) timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 \
      --master-port 29534 bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/bench_n${N}_$mode.json 2> gpurun_out/bench_n${N}_$mode.err
  python - <<PY
import json
try:
    d = json.loads([l for l in open('gpurun_out/bench_n${N}_$mode.json').read().splitlines() if l.startswith('{')][0])
    print('$mode', {k: d[k] for k in ('n_gpus', 'value', 'ms_per_step', 'gpu_launches')}, d['roofline'].get('kernel_ms'), d['roofline']['layer'].get('dense_ms'))
except Exception as e:
    print('$mode failed', e)
PY
  tail -2 gpurun_out/bench_n${N}_$mode.err
done

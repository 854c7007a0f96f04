#!/bin/bash
# session 34: GPU motif mining + SDGNN / SiGAT models; timing of the mining at scale
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_motifs.py tests/test_gpu_parity.py -m gpu -q -k "motif or sdgnn or sigat or sdr" 2>&1 | tail -12
timeout 600 python - <<'PY' 2>&1 | grep -E '^\{' | tee gpurun_out/motifs_s34.jsonl
import json, time, torch, sys
sys.path.insert(0, '.')
from pytorch_geometric_signed_directed_b200 import synthetic
from pytorch_geometric_signed_directed_b200.utils import signed as sg
dev = torch.device('cuda', 0)
for n, e in ((100_000, 2_000_000), (1_000_000, 20_000_000)):
    pos, neg, _ = synthetic.ssbm_edges(n, 3, num_entries=e, eta=0.1, seed=0, device=dev)
    es = torch.cat([torch.cat([pos.t(), torch.ones(pos.size(1), 1, dtype=torch.long, device=dev)], 1),
                    torch.cat([neg.t(), -torch.ones(neg.size(1), 1, dtype=torch.long, device=dev)], 1)], 0)
    for name, fn in (("sdgnn_motifs", sg.sdgnn_motifs), ("sigat_motifs", sg.sigat_motifs)):
        fn(es, n); torch.cuda.synchronize()
        t0 = time.perf_counter(); out = fn(es, n); torch.cuda.synchronize(); t = time.perf_counter() - t0
        print(json.dumps({"what": name, "nodes": n, "signed_edges": es.size(0), "ms": t * 1e3}), flush=True)
PY

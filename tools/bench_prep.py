"""Timings of the sparse GPU preprocessing at BASELINE config 3 scale (500k nodes / 10M edges), where the
reference's dense path would need a 500k x 500k float matrix (1 TB) and an O(N^3) eig: gpurun_out/prep.jsonl."""
import json
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pytorch_geometric_signed_directed_b200 import synthetic, utils  # noqa: E402

dev = torch.device("cuda", 0)
os.makedirs("gpurun_out", exist_ok=True)
fh = open("gpurun_out/prep.jsonl", "a")


def emit(**kw):
    print(json.dumps(kw), flush=True)
    fh.write(json.dumps(kw) + "\n")
    fh.flush()


def timed(fn, reps=2):
    fn()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps):
        out = fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / reps * 1e3, out


for n, e in ((100_000, 2_000_000), (500_000, 10_000_000)):
    ei, _ = synthetic.dsbm_edges(n, 3, num_edges=e, seed=1, device=dev)
    ms, (a_ei, a_w) = timed(lambda: utils.get_appr_directed_adj(0.1, ei, n, torch.float32))
    emit(what="get_appr_directed_adj", nodes=n, edges=ei.size(1), ms=ms, nnz=a_ei.size(1))
    ms, (s_ei, s_w) = timed(lambda: utils.get_second_directed_adj(ei, n, torch.float32))
    emit(what="get_second_directed_adj", nodes=n, edges=ei.size(1), ms=ms, nnz=s_ei.size(1))
    if n <= 100_000:
        ms, out = timed(lambda: utils.directed_features_in_out(ei, n))
        emit(what="directed_features_in_out", nodes=n, edges=ei.size(1), ms=ms, nnz_in=out[1].size(1), nnz_out=out[3].size(1))
    del ei, a_ei, a_w, s_ei, s_w
    torch.cuda.empty_cache()

"""Dense-transform kernel variants on the shapes the layers use (gpurun_out/sweep_dense.jsonl):
MagNet combine (1M x 4x64 -> 64 fp32), inception block (500k x 128 -> 384 bf16), SGCN (2M x 2x64 -> 64).
variant 0 = default tcgen05 kernel, 4 = warp-specialised, 8 = deeper prefetch, 1 = FFMA fallback."""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pytorch_geometric_signed_directed_b200 import ops  # noqa: E402

dev = torch.device("cuda", 0)
# dense_tma ring knobs ride in the variant word: 16 | lo_slots << 8 | landing_stages << 12
EXTRA = tuple(int(v, 0) for v in os.environ.get("PGSD_SWEEP_VARIANTS", "").split(",") if v)
os.makedirs("gpurun_out", exist_ok=True)
fh = open("gpurun_out/sweep_dense.jsonl", "a")


def emit(**kw):
    print(json.dumps(kw), flush=True)
    fh.write(json.dumps(kw) + "\n")
    fh.flush()


def timeit(fn, iters=10, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(iters):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / iters


def case(name, terms, n_out, bytes_alg, variants, **kw):
    ref = ops.dense(terms, n_out, variant=1, **kw)
    for v in variants:
        try:
            got = ops.dense(terms, n_out, variant=v, **kw)
            err = max((g.float() - r.float()).abs().max().item() / max(r.float().abs().max().item(), 1e-30)
                      for g, r in zip(got, ref))
            ms = timeit(lambda: ops.dense(terms, n_out, variant=v, **kw))
            emit(what=name, variant=v, ms=ms, gbs=bytes_alg / ms / 1e6, rel_err_vs_ffma=err)
        except Exception as exc:  # noqa: BLE001
            emit(what=name, variant=v, error=str(exc)[:200])


with torch.no_grad():
    n, f = 1_000_000, 64
    xs = [torch.rand(n, f, device=dev) * 2 - 1 for _ in range(4)]
    w = torch.rand(2, f, f, device=dev) - 0.5
    b = torch.rand(f, device=dev)
    case("magnet_combine_1M_4x64_64_f32", [(xs[0], w[0], 0), (xs[1], w[0], 1), (xs[2], w[1], 0), (xs[3], w[1], 1)],
         f, 6 * n * f * 4, (0, 16, 1) + EXTRA, bias=b, combine=True)
    del xs
    n, f = 500_000, 128
    x = (torch.rand(n, f, device=dev) * 2 - 1).bfloat16()
    w3 = torch.rand(f, 3 * f, device=dev) - 0.5
    case("inception_500k_128_384_bf16", [(x, w3, 0)], 3 * f, n * f * 2 * 4, (0, 16, 1) + EXTRA)
    xf = x.float()
    case("inception_500k_128_384_f32", [(xf, w3, 0)], 3 * f, n * f * 4 * 4, (0, 16, 1) + EXTRA)
    del x, xf
    n = 2_000_000
    xa, xb = torch.randn(n, 64, device=dev), torch.randn(n, 64, device=dev)
    w2 = torch.rand(128, 32, device=dev) - 0.5
    case("sgcn_2M_2x64_32_f32", [(xa, w2[:64], 0), (xb, w2[64:], 0)], 32, n * (128 + 32) * 4, (0, 16, 1) + EXTRA)
    xc = torch.randn(n, 64, device=dev)
    w4 = torch.rand(192, 64, device=dev) - 0.5
    case("sgcn_merged_2M_3x64_64_f32", [(xa, w4[:64], 0), (xb, w4[64:128], 0), (xc, w4[128:], 0)], 64,
         n * (192 + 64) * 4, (0, 16, 1) + EXTRA, bias=torch.rand(64, device=dev))

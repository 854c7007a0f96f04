"""GPU tuning sweep (not part of the product): times every pgsd_spmm_csr variant, the dense
transform and the plan build at the north-star size; writes gpurun_out/sweep.jsonl."""
import json
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pytorch_geometric_signed_directed_b200 import nn, ops, plan as planmod, synthetic  # noqa: E402

dev = torch.device("cuda", 0)
N = int(os.environ.get("SWEEP_N", 1_000_000))
E = int(os.environ.get("SWEEP_E", 20_000_000))
F = 64
out_path = os.path.join("gpurun_out", "sweep.jsonl")
os.makedirs("gpurun_out", exist_ok=True)
fh = open(out_path, "a")


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


ei, _ = synthetic.dsbm_edges(N, 3, num_edges=E, seed=0, device=dev)
gen = torch.Generator(device=dev).manual_seed(0)
xr = torch.rand(N, F, generator=gen, device=dev) * 2 - 1
xi = torch.rand(N, F, generator=gen, device=dev) * 2 - 1
torch.cuda.synchronize()
t0 = time.time()
p = planmod.build_magnetic(ei, None, N, 0.25, "sym", 2.0)
torch.cuda.synchronize()
emit(what="plan_build_first_ms", ms=(time.time() - t0) * 1e3, nnz=p.nnz)
emit(what="plan_build_ms", ms=timeit(lambda: planmod.build_magnetic(ei, None, N, 0.25, "sym", 2.0), 3, 1))

nnz = p.nnz
b_alg = nnz * (12 + 2 * F * 4) + (N + 1) * 4 + 2 * N * F * 4
yr, yi = torch.empty_like(xr), torch.empty_like(xi)
if os.environ.get("SWEEP_ONLY") == "bulk":
    # x1 of VERDICT r1: the bulk-copy (cp.async.bulk per neighbour row) gather beside the register-gather default, on
    # the full-size launch and on the short-row column blocks of the sharded path (8 blocks, ~5 entries per row)
    from pytorch_geometric_signed_directed_b200 import distributed as pgd
    for variant in (0, 0x800):
        ms = timeit(lambda: ops.spmm(p, [xr, xi], (0, 1), out=[yr, yi], variant=variant))
        emit(what="spmm2", variant=hex(variant), kernel="bulk-copy gather" if variant else "register gather (default)",
             ms=ms, gbs=b_alg / ms / 1e6, frac_of_peak=b_alg / ms / 1e6 / 6547.2)
        ms = timeit(lambda: ops.spmm(p, [xr], (0,), out=[yr], variant=variant))
        emit(what="spmm1", variant=hex(variant), ms=ms)
    bounds = pgd.node_bounds(N, 8)
    blocks = pgd.split_columns_by_owner(p, bounds, own_rank=0)
    for variant in (0, 0x800):
        ms = timeit(lambda: ops.spmm(blocks[3], [xr[bounds[3]:bounds[4]], xi[bounds[3]:bounds[4]]], (0, 1), beta=1.0,
                                     zs=[yr, yi], out=[yr, yi], variant=variant))
        emit(what="spmm2_column_block_1_of_8", variant=hex(variant), nnz=blocks[3].nnz, ms=ms)
    sys.exit(0)
for variant in (0x80 | 0x10 | 4, 0x80 | 0x20 | 2, 0x10 | 2, 0x10 | 4, 0x20 | 2, 0x20 | 4):
    ms = timeit(lambda: ops.spmm(p, [xr, xi], (0, 1), out=[yr, yi], variant=variant))
    emit(what="spmm2", variant=hex(variant), ms=ms, gbs=b_alg / ms / 1e6, nnz=nnz)
for variant in (0x100 | 0x10 | 4, 0x200 | 0x10 | 4):
    ms = timeit(lambda: ops.spmm(p, [xr, xi], (0, 1), out=[yr, yi], variant=variant))
    emit(what="spmm2_l2policy", variant=hex(variant), policy={1: "evict_normal", 2: "evict_first"}[variant >> 8], ms=ms, gbs=b_alg / ms / 1e6)
b1 = nnz * (8 + F * 4) + (N + 1) * 4 + N * F * 4
for variant in (0x80 | 0x20 | 2, 0x80 | 0x20 | 4, 0x10 | 4, 0x20 | 2, 0x20 | 4):
    ms = timeit(lambda: ops.spmm(p, [xr], (0,), out=[yr], variant=variant))
    emit(what="spmm1", variant=hex(variant), ms=ms, gbs=b1 / ms / 1e6)

conv = nn.MagNetConv(F, F, K=1, q=0.25, trainable_q=False, cached=True).to(dev)
conv(xr, xi, ei)
w = conv.weight.detach()
ms = timeit(lambda: ops.dense([(xr, w[0], 0), (xi, w[0], 1), (yr, w[1], 0), (yi, w[1], 1)], F,
                              bias=conv.bias, combine=True))
emit(what="dense_combine_auto", ms=ms, gflops=4 * 2 * N * F * F / ms / 1e6)
for dv in (1, 2, 4):
    ms = timeit(lambda: ops.dense([(xr, w[0], 0), (xi, w[0], 1), (yr, w[1], 0), (yi, w[1], 1)], F, bias=conv.bias, combine=True, variant=dv))
    emit(what="dense_combine", variant=dv, ms=ms, gflops=4 * 2 * N * F * F / ms / 1e6, gbs=6 * N * F * 4 / ms / 1e6)
emit(what="layer_forward", ms=timeit(lambda: conv(xr, xi, ei)))

# reference points: plain copy bandwidth and a torch gather of the same volume
big = torch.empty(256 * 1024 * 1024, dtype=torch.float32, device=dev)
big2 = torch.empty_like(big)
ms = timeit(lambda: big2.copy_(big))
emit(what="copy_1GiB", ms=ms, gbs=2 * big.numel() * 4 / ms / 1e6)
del big, big2
idx = torch.randint(0, N, (nnz // 4,), device=dev)
ms = timeit(lambda: xr.index_select(0, idx), 5, 2)
emit(what="torch_index_select_quarter", ms=ms, gbs=(nnz // 4) * F * 4 * 2 / ms / 1e6)

# ---- experiment: column-segmented multi-pass aggregation (L2-resident segments, output RMW)
if os.environ.get("SWEEP_SEGMENTS", "1") == "1":
    from pytorch_geometric_signed_directed_b200 import distributed as pgd
    for S in (2, 4, 8):
        bounds = pgd.node_bounds(N, S)
        blocks = pgd.split_columns_by_owner(p, bounds, own_rank=0)
        xs = [(xr[bounds[b]:bounds[b + 1]], xi[bounds[b]:bounds[b + 1]]) for b in range(S)]

        def run():
            y = ops.spmm(blocks[0], list(xs[0]), (0, 1), out=[yr, yi])
            for b in range(1, S):
                y = ops.spmm(blocks[b], list(xs[b]), (0, 1), beta=1.0, zs=y, out=y)
            return y
        ms = timeit(run, 5, 2)
        for variant in (0x10 | 4, 0x10 | 2, 0x20 | 2):
            def run_v():
                y = ops.spmm(blocks[0], list(xs[0]), (0, 1), out=[yr, yi], variant=variant)
                for b in range(1, S):
                    y = ops.spmm(blocks[b], list(xs[b]), (0, 1), beta=1.0, zs=y, out=y, variant=variant)
                return y
            emit(what="spmm2_segmented_variant", segments=S, variant=hex(variant), ms=timeit(run_v, 5, 2))
        ref = ops.spmm(p, [xr, xi], (0, 1))
        got = run()
        err = (got[0] - ref[0]).abs().max().item() / ref[0].abs().max().item()
        emit(what="spmm2_segmented", segments=S, ms=ms, rel_err=err)
        del blocks

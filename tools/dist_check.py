"""Run under torchrun (NCCL): the node-range sharded MagNetConv must reproduce the single-GPU
layer to rounding, on a permuted DSBM graph (halo = whole matrix -> ring all-gather) and on a
locality-ordered graph (thin halo -> all-to-all of packed halo rows).  Usage:
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 \
        --master-port 29533 tools/dist_check.py [--bench]
`--bench` additionally times the sharded step of the halo path at 1M nodes / 20M edges per rank.
"""
import json
import os
import sys

# one hardware queue per stream of the exchange (copy-engine transport: up to 7 copy streams + compute + push)
os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pytorch_geometric_signed_directed_b200 import distributed as pgd, nn, synthetic  # noqa: E402

rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
local = int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)

n_total, e_total, f = 50_000 * world + 37, 1_000_000 * world, 64
gen = torch.Generator(device=dev).manual_seed(7)          # same stream on every rank
x_real = torch.rand(n_total, f, generator=gen, device=dev) * 2 - 1
x_imag = torch.rand(n_total, f, generator=gen, device=dev) * 2 - 1


def check(gname, want_mode, ei):
    ok = True
    for K in (1, 2):
        torch.manual_seed(K)
        conv = nn.MagNetConv(f, f, K=K, q=0.25, trainable_q=False, cached=True).to(dev)
        with torch.no_grad():
            conv.bias.uniform_(-0.2, 0.2)
        full_r, full_i = conv(x_real, x_imag, ei)
        sh = pgd.ShardedMagNetConv(conv, n_total, rank, world).build(ei)
        lo, hi = sh.bounds[rank], sh.bounds[rank + 1]
        for it in range(3):                                   # repeated calls reuse the buffers
            out_r, out_i = sh(x_real[lo:hi].contiguous(), x_imag[lo:hi].contiguous())
        torch.cuda.synchronize()
        sh.agg.check()
        ok &= sh.agg.mode == want_mode
        for got, ref, nm in ((out_r, full_r[lo:hi], "real"), (out_i, full_i[lo:hi], "imag")):
            err = (got - ref).abs().max().item()
            scale = ref.abs().max().item()
            good = err <= 2e-6 * scale
            ok &= good
            print(f"[rank {rank}] {gname} mode={sh.agg.mode} K={K} out_{nm}: max err {err:.3e} "
                  f"(scale {scale:.3e}) {'OK' if good else 'FAIL'}", flush=True)
        # one tensor for both parts (the reference example's call): the shard travels once, then back to two
        # tensors -- exchange objects of both widths stay cached
        xloc = x_real[lo:hi].contiguous()
        sr, si = sh(xloc, xloc)
        fr, fi = conv(x_real, x_real, ei)
        out_r2, _ = sh(x_real[lo:hi].contiguous(), x_imag[lo:hi].contiguous())
        torch.cuda.synchronize()
        for got, ref, nm in ((sr, fr[lo:hi], "shared real"), (si, fi[lo:hi], "shared imag"), (out_r2, full_r[lo:hi], "real again")):
            err, scale = (got - ref).abs().max().item(), ref.abs().max().item()
            good = err <= 2e-6 * scale
            ok &= good
            if not good or rank == 0:
                print(f"[rank {rank}] {gname} K={K} {nm}: max err {err:.3e} {'OK' if good else 'FAIL'}", flush=True)
    return ok


ok = True
if "--skip-check" not in sys.argv:
    ok = check("dsbm", "ring", synthetic.dsbm_edges(n_total, 3, num_edges=e_total, seed=0, device=dev)[0])
    ok &= check("local", "halo", synthetic.locality_edges(n_total, e_total, 2000, 0.0, seed=0, device=dev))
flag = torch.tensor([1 if ok else 0], device=dev)
dist.all_reduce(flag, op=dist.ReduceOp.MIN)
if rank == 0:
    print("DIST_CHECK", "PASS" if flag.item() == 1 else "FAIL", f"world={world}", flush=True)

if "--bench" in sys.argv and flag.item() == 1:
    # weak scaling of the halo path: 1M nodes / 20M edges per rank, |i - j| <= 50k
    del x_real, x_imag
    n_b, e_b = 1_000_000 * world, 20_000_000 * world
    ei = synthetic.locality_edges(n_b, e_b, 50_000, 0.0, seed=1, device=dev)
    conv = nn.MagNetConv(f, f, K=1, q=0.25, trainable_q=False, cached=True).to(dev)
    sh = pgd.ShardedMagNetConv(conv, n_b, rank, world).build(ei)
    e_in = ei.size(1)
    del ei
    xr = torch.rand(sh.n_local, f, device=dev) * 2 - 1
    xi = torch.rand(sh.n_local, f, device=dev) * 2 - 1
    with torch.no_grad():
        for _ in range(5):
            sh(xr, xi)
        dist.barrier(); torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(20):
            sh(xr, xi)
        b.record()
        dist.barrier(); torch.cuda.synchronize()
    t = torch.tensor([a.elapsed_time(b) / 20], device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    if rank == 0:
        print(json.dumps({"what": "halo_path_weak_scaling", "world": world, "mode": sh.agg.mode,
                          "halo_fraction": sh.agg.halo_fraction, "ms_per_step": t.item(),
                          "edges_per_s": e_in / (t.item() * 1e-3), "edges_total": e_in,
                          "halo_rows_received_rank0": sh.agg.halo.n_recv if sh.agg.halo else None,
                          "graph": "locality_edges band=50k, 1M nodes / 20M edges per rank"}), flush=True)
if "--trace" in sys.argv and flag.item() == 1:
    # timeline of one sharded step of the all-gather path at bench scale (1M nodes / 20M edges per rank)
    n_b, e_b = 1_000_000 * world, 20_000_000 * world
    ei = synthetic.dsbm_edges(n_b, 3, num_edges=e_b, seed=0, device=dev)[0]
    conv = nn.MagNetConv(f, f, K=1, q=0.25, trainable_q=False, cached=True).to(dev)
    sh = pgd.ShardedMagNetConv(conv, n_b, rank, world).build(ei)
    del ei
    xr = torch.rand(sh.n_local, f, device=dev) * 2 - 1
    xi = torch.rand(sh.n_local, f, device=dev) * 2 - 1
    with torch.no_grad():
        for _ in range(5):
            sh(xr, xi)
        # (a) one step after a barrier + synchronize; (b) the 4th of 6 back-to-back steps (what bench.py times)
        for mode in ("isolated", "steady"):
            dist.barrier(); torch.cuda.synchronize()
            n_run, traced = (1, 0) if mode == "isolated" else (6, 3)
            for i in range(n_run):
                if i == traced:
                    pgd.TRACE = []
                    t0 = pgd._now_event()
                sh(xr, xi)
                if i == traced:
                    t1 = pgd._now_event()
                    tr, pgd.TRACE = pgd.TRACE, None
            torch.cuda.synchronize()
            sh.agg.check()
            if rank in (0, world - 1):
                line = " | ".join(f"{nm} {t0.elapsed_time(ev):.2f}" for nm, ev in tr)
                print(f"[rank {rank}] TRACE {mode} ms from step start: {line} | step end {t0.elapsed_time(t1):.2f}", flush=True)
dist.destroy_process_group()
sys.exit(0 if flag.item() == 1 else 1)

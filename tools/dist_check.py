"""Run under torchrun (NCCL): the node-range sharded MagNetConv must reproduce the single-GPU
layer bit-for-bit in structure and to rounding in values.  Usage:
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 \
        --master-port 29533 tools/dist_check.py
"""
import os
import sys

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
ei, _ = synthetic.dsbm_edges(n_total, 3, num_edges=e_total, seed=0, device=dev)
gen = torch.Generator(device=dev).manual_seed(7)          # same stream on every rank
x_real = torch.rand(n_total, f, generator=gen, device=dev) * 2 - 1
x_imag = torch.rand(n_total, f, generator=gen, device=dev) * 2 - 1
ok = True
for K in (1, 2):
    torch.manual_seed(K)
    conv = nn.MagNetConv(f, f, K=K, q=0.25, trainable_q=False, cached=True).to(dev)
    with torch.no_grad():
        conv.bias.uniform_(-0.2, 0.2)
    full_r, full_i = conv(x_real, x_imag, ei)
    sh = pgd.ShardedMagNetConv(conv, n_total, rank, world).build(ei)
    lo, hi = sh.bounds[rank], sh.bounds[rank + 1]
    for it in range(3):                                   # repeated calls reuse the recv buffers
        out_r, out_i = sh(x_real[lo:hi].contiguous(), x_imag[lo:hi].contiguous())
    torch.cuda.synchronize()
    for got, ref, nm in ((out_r, full_r[lo:hi], "real"), (out_i, full_i[lo:hi], "imag")):
        err = (got - ref).abs().max().item()
        scale = ref.abs().max().item()
        good = err <= 2e-6 * scale
        ok &= good
        print(f"[rank {rank}] K={K} out_{nm}: max err {err:.3e} (scale {scale:.3e}) {'OK' if good else 'FAIL'}",
              flush=True)
flag = torch.tensor([1 if ok else 0], device=dev)
dist.all_reduce(flag, op=dist.ReduceOp.MIN)
if rank == 0:
    print("DIST_CHECK", "PASS" if flag.item() == 1 else "FAIL", f"world={world}", flush=True)
dist.destroy_process_group()
sys.exit(0 if flag.item() == 1 else 1)

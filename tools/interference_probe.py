"""How much does an NVLink exchange slow the HBM-bound aggregation that runs beside it, and which side (egress /
ingress) and which engine (SM stores, bulk-copy, copy engine) is responsible?  world = 2, run under torchrun.
A background stream repeats one transfer of [1M, 64] fp32 x 2 back to back for ~25 ms (rank 0 -> 1, rank 1 -> 0, or
both); the foreground times ONE full-size aggregation (1M rows / 40M entries) and ONE 1 GiB device copy in the middle
of it, on each rank.  Prints one JSON line per background kind.
"""
import ctypes
import json
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pytorch_geometric_signed_directed_b200 import distributed as pgd, ops, plan as _plan, synthetic  # noqa: E402

rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
local = int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
assert world == 2
peer = 1 - rank
N, E, F = 1_000_000, 20_000_000, 64
bounds = pgd.node_bounds(N * world, world)
xr = torch.rand(N, F, device=dev) * 2 - 1
xi = torch.rand(N, F, device=dev) * 2 - 1
ei, _ = synthetic.dsbm_edges(N, 3, num_edges=E, seed=rank, device=dev)
load_plan = _plan.build_magnetic(ei, None, N, 0.25, "sym", 2.0, 0)
del ei
ca = torch.empty(128 * 1024 * 1024, dtype=torch.float32, device=dev)     # 512 MiB read + 512 MiB written
cb = torch.empty_like(ca)
cudart = ctypes.CDLL("libcudart.so.12")
cudart.cudaMemcpyAsync.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_size_t, ctypes.c_int, ctypes.c_void_p]
REPS = int(os.environ.get("PROBE_REPS", "28"))


def foreground():
    """(spmm ms, copy GB/s) measured ~1.5 ms after the call, on the current stream."""
    torch.cuda._sleep(3_000_000)
    e = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
    e[0].record()
    ops.spmm(load_plan, [xr, xi], (0, 1))
    e[1].record()
    e[2].record()
    cb.copy_(ca)
    e[3].record()
    return e


def run(name, make_bg, senders):
    """make_bg(stream) enqueues ONE background transfer on `stream`; only ranks in `senders` send."""
    bg = torch.cuda.Stream(device=dev, priority=-1)
    res = []
    for it in range(3):
        dist.barrier()
        torch.cuda.synchronize()
        t0 = torch.cuda.Event(enable_timing=True)
        t0.record()
        done = torch.cuda.Event(enable_timing=True)
        if rank in senders and make_bg is not None:
            bg.wait_event(t0)
            for _ in range(REPS):
                make_bg(bg)
            done.record(bg)
        else:
            done.record()
        ev = foreground()
        torch.cuda.synchronize()
        dist.barrier()
        if it:
            bg_ms = t0.elapsed_time(done)
            res.append((ev[0].elapsed_time(ev[1]), 1.0737 / ev[2].elapsed_time(ev[3]) * 1e3, bg_ms, t0.elapsed_time(ev[3])))
    r = res[-1]
    mine = torch.tensor([r[0], r[1], r[2], r[3]], device=dev, dtype=torch.float64)
    both = [torch.empty_like(mine) for _ in range(world)]
    dist.all_gather(both, mine)
    if rank == 0:
        rec = {"background": name, "senders": sorted(senders)}
        for q in range(world):
            v = both[q].tolist()
            rec[f"rank{q}"] = {"spmm_ms": round(v[0], 3), "copy_gbs": round(v[1], 1), "background_ms": round(v[2], 2),
                               "foreground_end_ms": round(v[3], 2)}
            if q in senders and make_bg is not None:
                rec[f"rank{q}"]["bg_gbs_per_transfer"] = round(REPS * N * F * 4 * 2 / v[2] / 1e6, 1)
        print(json.dumps(rec), flush=True)


run("none", None, set())
for engine, ctas, tile in ((1, 32, "16384x4"), (1, 16, "16384x4"), (1, 8, "32768x4"), (0, 64, "-"), (0, 24, "-")):
    os.environ["PGSD_PUSH_CTAS"], os.environ["PGSD_PUSH_ENGINE"] = str(ctas), str(engine)
    if engine == 1:
        os.environ["PGSD_PUSH_TILE"] = tile
    ex = pgd.PushExchange(rank, world, bounds, 2, F, torch.float32, dev, pgd.stage_fractions(1))

    def bg_push(stream, ex=ex):
        with torch.cuda.stream(stream):
            ex.push([xr, xi])

    nm = f"{'tma' if engine else 'lsu'}-push-{ctas}" + (f"-{tile}" if engine else "")
    for senders in ({0}, {0, 1}):
        run(nm, bg_push, senders)
    ex.check()
    last = ex
# copy-engine push / pull through the last exchange's symmetric planes
plane_bytes = N * F * 4


def bg_ce_push(stream):
    for t, x in enumerate((xr, xi)):
        dst = last.buf_ptrs[peer] + t * last.plane_bytes + bounds[rank] * last.row_bytes
        cudart.cudaMemcpyAsync(dst, x.data_ptr(), plane_bytes, 4, stream.cuda_stream)


def bg_ce_pull(stream):
    for t, x in enumerate((xr, xi)):
        src = last.buf_ptrs[peer] + t * last.plane_bytes + bounds[peer] * last.row_bytes
        cudart.cudaMemcpyAsync(x.data_ptr() if False else cb.data_ptr() + t * plane_bytes, src, plane_bytes, 4,
                               stream.cuda_stream)


for nm, fn in (("ce-push", bg_ce_push), ("ce-pull", bg_ce_pull)):
    for senders in ({0}, {0, 1}):
        run(nm, fn, senders)
dist.barrier()
dist.destroy_process_group()

"""Transport probe for the all-gather of the row-sharded path (run under torchrun, NCCL):
times one all-gather of [1M, 64] fp32 x 2 shards per rank with
  - the shard-push kernel (pgsd_shard_push) at several CTA counts, unicast and NVSwitch multicast,
  - copy-engine pulls from symmetric memory (round-1 transport),
alone and while a full-size aggregation (1M rows / 40M entries, the bench's per-rank work) runs beside
it, and reports GB/s of ingress per rank (max time over ranks) and the slowdown of the aggregation.
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 \
        --master-port 29541 tools/exchange_probe.py > gpurun_out/probe.jsonl
"""
import json
import os
import sys

# one hardware queue per stream of the exchange (copy-engine transport: up to 7 copy streams + compute + push)
os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pytorch_geometric_signed_directed_b200 import distributed as pgd, ops, plan as _plan, synthetic  # noqa: E402

rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
local = int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)

N, E, F = 1_000_000, 20_000_000, 64
bounds = pgd.node_bounds(N * world, world)
gen = torch.Generator(device=dev).manual_seed(100 + rank)
xr = torch.rand(N, F, generator=gen, device=dev) * 2 - 1
xi = torch.rand(N, F, generator=gen, device=dev) * 2 - 1
ei, _ = synthetic.dsbm_edges(N, 3, num_edges=E, seed=rank, device=dev)
load_plan = _plan.build_magnetic(ei, None, N, 0.25, "sym", 2.0, 0)
del ei
bytes_in = (world - 1) * N * F * 4 * 2


def out(rec):
    if rank == 0:
        print(json.dumps(rec), flush=True)


def max_over_ranks(v):
    t = torch.tensor([v], device=dev, dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


load_stream = torch.cuda.Stream(device=dev)


def run(name, start_exchange, wait_exchange, with_load, reserve=0, reps=4, variant=None, load_first=False):
    """start_exchange() launches the transfer ordered behind the current stream; wait_exchange() makes the
    current stream wait for everything THIS rank receives."""
    times, loads = [], []
    for it in range(reps + 1):
        dist.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        l0, l1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        if not load_first:
            start_exchange()
        if with_load:
            load_stream.wait_event(e0)
            with torch.cuda.stream(load_stream):
                l0.record()
                for _ in range(3):
                    ops.spmm(load_plan, [xr, xi], (0, 1), grid_reserve=reserve, variant=variant)
                l1.record()
        if load_first:
            start_exchange()
        wait_exchange()
        e1.record()
        torch.cuda.synchronize()
        dist.barrier()
        if it:
            times.append(e0.elapsed_time(e1))
            if with_load:
                loads.append(l0.elapsed_time(l1) / 3)
    t = max_over_ranks(sorted(times)[len(times) // 2])
    rec = {"transport": name, "load": with_load, "load_first": load_first, "world": world, "ms": round(t, 3),
           "ingress_gbs": round(bytes_in / t / 1e6, 1)}
    if with_load:
        rec["spmm_ms_beside"] = round(max_over_ranks(sorted(loads)[len(loads) // 2]), 3)
    out(rec)


# baseline: aggregation alone
for _ in range(2):
    ops.spmm(load_plan, [xr, xi], (0, 1))
torch.cuda.synchronize()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
for _ in range(5):
    ops.spmm(load_plan, [xr, xi], (0, 1))
b.record()
torch.cuda.synchronize()
out({"what": "aggregation alone", "spmm_ms": round(a.elapsed_time(b) / 5, 3)})
for code in (1, 2, 3, 4, 5, 6):       # preferred shared-memory carve-out of the aggregation launch
    for _ in range(2):
        ops.spmm(load_plan, [xr, xi], (0, 1), variant=code << 12)
    a.record()
    for _ in range(5):
        ops.spmm(load_plan, [xr, xi], (0, 1), variant=code << 12)
    b.record()
    torch.cuda.synchronize()
    out({"what": f"aggregation alone, smem carve-out {14 * code} %", "spmm_ms": round(a.elapsed_time(b) / 5, 3)})

# ---- shard-push kernel: (label, engine, multicast, CTAs, tile)
variants = [("lsu", 0, 0, 32, "16384x4"), ("lsu", 0, 0, 64, "16384x4"), ("lsu", 0, 0, 96, "16384x4"),
            ("tma", 1, 0, 16, "16384x4"), ("tma", 1, 0, 32, "16384x4"), ("tma", 1, 0, 16, "32768x4"),
            ("lsu-mc", 0, 1, 64, "16384x4")]
if os.environ.get("PROBE_VARIANTS"):
    keep = set(os.environ["PROBE_VARIANTS"].split(","))
    variants = [v for v in variants if f"{v[0]}-{v[3]}-{v[4]}" in keep]
for mc_name, engine, mc, ctas, tile in variants:
    os.environ["PGSD_PUSH_CTAS"] = str(ctas)
    os.environ["PGSD_PUSH_MC"] = str(mc)
    os.environ["PGSD_PUSH_ENGINE"] = str(engine)
    os.environ["PGSD_PUSH_TILE"] = tile
    mc_name = f"{mc_name}-{tile}" if engine == 1 else mc_name
    try:
        ex = pgd.PushExchange(rank, world, bounds, 2, F, torch.float32, dev, pgd.stage_fractions(4))
    except Exception as exc:  # noqa: BLE001
        out({"transport": f"push-{mc_name}-{ctas}", "error": f"{type(exc).__name__}: {exc}"})
        continue
    if mc and not ex.mc_ptr:
        out({"transport": f"push-mc-{ctas}", "error": "no multicast pointer on this handle"})
        del ex
        continue
    state = {}

    def start():
        state["seq"] = ex.push([xr, xi])

    def wait():
        for s in range(ex.n_slices):
            ex.wait_slice(s, state["seq"])
        ex.finish()

    var = (3 << 12) if engine == 1 else None
    run(f"push-{mc_name}-{ctas}", start, wait, False)
    run(f"push-{mc_name}-{ctas}", start, wait, True, reserve=ctas, variant=var)
    # the aggregation is enqueued BEFORE the push (what a steady-state loop can look like): the push CTAs must
    # still find room beside the resident aggregation CTAs (grid_reserve / shared-memory carve-out)
    run(f"push-{mc_name}-{ctas}", start, wait, True, reserve=ctas, variant=var, load_first=True)
    ex.check()
    # data check once per variant: plane rows of peer b must equal what b generated (same generator recipe)
    par = state["seq"] & 1
    peer = (rank + 1) % world
    g2 = torch.Generator(device=dev).manual_seed(100 + peer)
    pr = torch.rand(N, F, generator=g2, device=dev) * 2 - 1
    ok = torch.equal(ex.planes[par][0][bounds[peer]:bounds[peer + 1]], pr)
    out({"transport": f"push-{mc_name}-{ctas}", "data_ok_rank0": bool(ok)})
    del ex, pr
    torch.cuda.empty_cache()

# ---- copy-engine pulls (round-1 transport)
for streams in (1,):
    os.environ["PGSD_COPY_STREAMS"] = str(streams)
    pe = pgd.SymmetricPullExchange(rank, world, N, 2 * F, torch.float32, dev)
    recv = [None if b == rank else torch.empty((N, 2 * F), device=dev) for b in range(world)]
    state = {}

    def start():
        send = pe.send_buffer(N)
        send[:, :F].copy_(xr)
        send[:, F:].copy_(xi)
        state["works"] = pe.start(send, recv)

    def wait():
        for _, reqs in state["works"]:
            for w in reqs:
                w.wait()

    for with_load in (False, True):
        run(f"ce-pull-{streams}stream", start, wait, with_load)
    del pe, recv
    torch.cuda.empty_cache()

dist.barrier()
dist.destroy_process_group()

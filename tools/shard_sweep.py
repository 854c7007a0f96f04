"""Sweep of the push-exchange knobs on the bench's sharded step (run under torchrun, NCCL): one operator build, then
for every configuration a fresh ShardedAggregator (stage blocks + exchange) and 10 timed steps; prints one JSON
line per configuration (max over ranks) and, with --trace, the steady-state timeline of rank 0.
    configs: "slices=<spec>+ctas=<n>+engine=<0|1|2>+tile=<bytes>x<stages>" separated by spaces (PGSD_SWEEP env or argv)
"""
import json
import os
import sys

# one hardware queue per stream of the exchange (copy-engine transport: up to 7 copy streams + compute + push)
os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pytorch_geometric_signed_directed_b200 import distributed as pgd, nn, ops, synthetic  # noqa: E402

rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
local = int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)

f = 64
n_b, e_b = 1_000_000 * world, 20_000_000 * world
ei = synthetic.dsbm_edges(n_b, 3, num_edges=e_b, seed=0, device=dev)[0]
e_in = ei.size(1)
conv = nn.MagNetConv(f, f, K=1, q=0.25, trainable_q=False, cached=True).to(dev)
sh = pgd.ShardedMagNetConv(conv, n_b, rank, world).build(ei)
del ei
xr = torch.rand(sh.n_local, f, device=dev) * 2 - 1
xi = torch.rand(sh.n_local, f, device=dev) * 2 - 1
KEYS = {"cestreams": "PGSD_CE_STREAMS", "slices": "PGSD_PUSH_SLICES", "ctas": "PGSD_PUSH_CTAS", "engine": "PGSD_PUSH_ENGINE", "tile": "PGSD_PUSH_TILE",
        "exchange": "PGSD_EXCHANGE", "mc": "PGSD_PUSH_MC"}
configs = [a for a in sys.argv[1:] if not a.startswith("--")] or os.environ.get("PGSD_SWEEP", "slices=auto").split()
trace = "--trace" in sys.argv

with torch.no_grad():
    for cfg in configs:
        for k in KEYS.values():
            os.environ.pop(k, None)
        ops.SPMM_VARIANT = 0
        for item in cfg.split("+"):
            k, v = item.split("=", 1)
            if k == "variant":
                ops.SPMM_VARIANT = int(v, 0)
            else:
                os.environ[KEYS[k]] = v
        sh.agg = pgd.ShardedAggregator(sh.local_plan, sh.bounds, rank, world)
        for _ in range(4):
            sh(xr, xi)
        dist.barrier(); torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        tr = None
        for i in range(10):
            if trace and i == 5:
                pgd.TRACE = []
                t0 = pgd._now_event()
            sh(xr, xi)
            if trace and i == 5:
                t1 = pgd._now_event()
                tr, pgd.TRACE = pgd.TRACE, None
        b.record()
        dist.barrier(); torch.cuda.synchronize()
        sh.agg.check()
        t = torch.tensor([a.elapsed_time(b) / 10], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        if rank == 0:
            rec = {"config": cfg, "world": world, "ms_per_step": round(t.item(), 3),
                   "edges_per_s": e_in / (t.item() * 1e-3), "mode": sh.agg.mode,
                   "slices": [round(c, 3) for c in (sh.agg.stage_cum or [])]}
            if tr:
                rec["timeline_ms"] = {nm: round(t0.elapsed_time(ev), 2) for nm, ev in tr}
                rec["timeline_ms"]["step end"] = round(t0.elapsed_time(t1), 2)
            print(json.dumps(rec), flush=True)
        del sh.agg
        sh.agg = None
        torch.cuda.empty_cache()
dist.destroy_process_group()

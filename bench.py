#!/usr/bin/env python
"""bench.py -- headline benchmark of the hot path (BASELINE.json): MagNetConv forward on a
synthetic DSBM graph, 1M nodes / 20M directed edges / 64 features, fp32, per GPU.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

One "step" = one MagNetConv.forward (K=1, q=0.25, 'sym', cached=True, plan already built) over
the whole graph.  metric = input edges aggregated per second = E_input / t_step.
  value   : inputs resident in HBM, CUDA-event timed, max over ranks.
  e2e     : the same forward through the public layer API with HOST (pinned) feature buffers:
            H2D of x_real/x_imag and D2H of out_real/out_imag inside the timed region.
  roofline: dominant kernel = pgsd_spmm_csr; algorithmic bytes / its CUDA-event duration measured
            inside the timed region, against MEASURED_PEAKS.json's hbm_gbs.
  cpu_baseline: oracle/port.py (the reference's CPU op sequence) on the same 1M/20M graph, a few forwards.
N > 1 ("weak" scaling): the graph grows to N*1M nodes / N*20M edges, destination rows are
sharded by node range, every rank builds only its rows of the operator, and feature shards are pushed
over NVLink into the peers' receive planes by one kernel (pgsd_shard_push: symmetric peer memory, per-slice
flags), overlapped with the aggregation of the own-column block and of every landed slice
(pytorch_geometric_signed_directed_b200/distributed.py).  Extra keys: parity_check (rank 0's rows of the
timed path against the oracle), halo_path (a locality-ordered graph: all-to-all of halo rows instead).
`--impl reference` times the CPU port only (rank 0) on the whole 1M/20M graph, same metric.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import tempfile
import threading
import time

# one hardware queue per stream (compute, exchange, copy-engine streams, e2e side streams): with the default of 8
# connections streams can alias and serialise behind each other; must be set before the CUDA context exists
os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")
import torch  # noqa: E402

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

N_PER_GPU = 1_000_000
E_PER_GPU = 20_000_000
FEAT = 64
METRIC = "edges aggregated/sec (MagNetConv fwd, 1M nodes/20M edges/64d per GPU)"
UNIT = "edges/s"


def log(*a):
    print(*a, file=sys.stderr, flush=True)


# The contract is ONE JSON line on stdout.  Libraries print to stdout behind our back (NCCL's
# "NCCL version ..." banner when NCCL_DEBUG=VERSION is set on the box), so the real stdout is
# parked on a private descriptor, fd 1 is pointed at stderr for the whole run, and only the
# result line is written to the parked descriptor.
_REAL_STDOUT = None


def protect_stdout():
    global _REAL_STDOUT
    if _REAL_STDOUT is None:
        sys.stdout.flush()
        _REAL_STDOUT = os.dup(1)
        os.dup2(2, 1)


def emit_json(obj):
    data = (json.dumps(obj) + "\n").encode()
    if _REAL_STDOUT is None:
        sys.stdout.write(data.decode()); sys.stdout.flush()
    else:
        os.write(_REAL_STDOUT, data)


def measured_peak_gbs():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        with open(path) as fh:
            return float(json.load(fh)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback 6.65 TB/s (B200_PROFILING.md)"


def recorded_traffic(name="spmm_dram_traffic.json", kernel=None):
    """Per-launch DRAM bytes of the dominant kernel from the committed ncu --set full capture -- only when the capture
    is of the kernel instantiation that was just timed (`kernel` = pgsd_last_spmm_kernel()); a stale capture is
    reported as null rather than as this run's traffic."""
    try:
        with open(os.path.join(ROOT, "profiles", name)) as fh:
            rec = json.load(fh)
        if kernel is not None and rec.get("kernel") != kernel:
            log(f"roofline.traffic dropped: capture is of {rec.get('kernel')}, this run launched {kernel}")
            return None
        return float(rec["dram_bytes_per_launch"])
    except Exception:
        return None


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 200 ms during the timed region."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index, self.proc, self.path = index, None, None

    def start(self):
        try:
            fd, self.path = tempfile.mkstemp(suffix=".csv")
            os.close(fd)
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                 "-i", str(self.index), "-lms", "200"],
                stdout=open(self.path, "w"), stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": []}
        if self.proc is None:
            return out
        time.sleep(0.25)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        try:
            for line in open(self.path):
                f = [t.strip() for t in line.split(",")]
                if len(f) < 7:
                    continue
                try:
                    sm.append(float(f[0])); mx.append(float(f[1]))
                except ValueError:
                    continue
                for nm, v in zip(names, f[3:7]):
                    if v.lower().startswith("active"):
                        reasons.add(nm)
            os.remove(self.path)
        except Exception:
            pass
        if sm:
            out = {"sm_mhz": statistics.median(sm), "sm_max_mhz": max(mx), "reasons": sorted(reasons),
                   "samples": len(sm)}
        return out


# ------------------------------------------------------------------------------ CPU baseline

def _cpu_inputs(n, e):
    from pytorch_geometric_signed_directed_b200 import synthetic
    ei, _ = synthetic.dsbm_edges(n, 3, num_edges=e, eta=0.1, size_ratio=1.5, seed=0)
    g = torch.Generator().manual_seed(0)
    xr = torch.rand(n, FEAT, generator=g) * 2 - 1
    xi = torch.rand(n, FEAT, generator=g) * 2 - 1
    w = torch.rand(2, FEAT, FEAT, generator=g) - 0.5
    return ei, xr, xi, w, torch.zeros(FEAT)


def _cpu_model():
    try:
        for line in open("/proc/cpuinfo"):
            if line.startswith("model name"):
                return line.split(":", 1)[1].strip()
    except Exception:
        pass
    return "unknown"


def cpu_reference_run(steps: int, warmup: int, budget_s: float = 150.0):
    """The reference's CPU op sequence (oracle/port.py: index_select -> mul -> scatter_add_, the four
    Chebyshev chains, 4(K+1) matmuls; MagNetConv.py:185-249 op for op) on the host cores, on the
    benchmark's OWN per-GPU graph: DSBM 1M nodes / 20M edges / 64 features, cached operator (steady
    state of cached=True), under torch.no_grad().  torch's CPU scatter_add_ does not scale to very wide
    hosts, so the thread count is first chosen among {all cores, 32, 16, 8} on a 1/8-size graph of the
    same mean degree (seconds; the fastest wins) -- the baseline gets the best setting the host offers.
    `steps` timed forwards after `warmup` untimed ones, cut short so that the timed part stays within
    budget_s (at least 3 timed forwards); the count actually run is returned."""
    from oracle import port
    cores = os.cpu_count() or 1
    with torch.no_grad():
        ei, xr, xi, w, b = _cpu_inputs(N_PER_GPU // 8, E_PER_GPU // 8)
        cached = port.magnet_norm(ei, None, xr.size(0), 0.25, "sym", 2.0)
        trial = {}
        for th in sorted({cores, min(cores, 32), min(cores, 16), min(cores, 8)}, reverse=True):
            torch.set_num_threads(th)
            port.magnet_conv(xr, xi, ei, None, w, b, 0.25, "sym", cached_result=cached)
            t0 = time.perf_counter()
            port.magnet_conv(xr, xi, ei, None, w, b, 0.25, "sym", cached_result=cached)
            trial[th] = time.perf_counter() - t0
        best = min(trial, key=trial.get)
        torch.set_num_threads(best)
        small = {"nodes": xr.size(0), "edges": ei.size(1), "s_per_forward": trial[best],
                 "value": ei.size(1) / trial[best]}
        del ei, xr, xi, cached

        ei, xr, xi, w, b = _cpu_inputs(N_PER_GPU, E_PER_GPU)
        t0 = time.perf_counter()
        cached = port.magnet_norm(ei, None, xr.size(0), 0.25, "sym", 2.0)       # operator build (cold part)
        build_s = time.perf_counter() - t0

        def fwd():
            t0 = time.perf_counter()
            port.magnet_conv(xr, xi, ei, None, w, b, 0.25, "sym", cached_result=cached)
            return time.perf_counter() - t0

        first = fwd()                                       # untimed warm-up no. 1 (also sizes the run)
        n_timed = max(3, min(steps, int(budget_s / max(first, 1e-3))))
        n_warm = max(0, min(warmup - 1, int(0.25 * budget_s / max(first, 1e-3))))
        for _ in range(n_warm):
            fwd()
        times = [fwd() for _ in range(n_timed)]
    t = sum(times) / len(times)
    return {"value": ei.size(1) / t, "unit": UNIT, "cores": best, "host_cores": cores, "cpu_model": _cpu_model(),
            "kind": "port", "nodes": xr.size(0), "edges": ei.size(1), "timed_forwards": n_timed,
            "warmup_forwards": n_warm + 1, "s_per_forward": t, "cold_s": build_s + first,
            "operator_build_s": build_s,
            "thread_trials_s_on_eighth_size_graph": {str(k): round(v, 3) for k, v in trial.items()},
            "eighth_size_sample": small,
            "sample": f"the whole per-GPU workload: DSBM {xr.size(0)} nodes / {ei.size(1)} edges / {FEAT} feat, cached "
                      f"operator, {n_timed} timed forwards of {t:.2f} s after {n_warm + 1} warm-ups, {best} threads "
                      f"(best of {sorted(trial)} on a 1/8-size graph); cold call (operator build + first forward) "
                      f"{build_s + first:.1f} s"}, t, n_timed, n_warm + 1


def run_reference_arm(args, rank):
    if rank != 0:
        return
    base, t, n_timed, n_warm = cpu_reference_run(args.steps, args.warmup)
    cfg = bench_config(args.gpus)
    if args.gpus > 1:
        cfg["reference_ran"] = ("ONE rank's share of the weak-scaled workload (1M nodes / 20M edges): the reference is "
                                "single-process, its CPU path does not use the other GPUs' hosts, and the "
                                f"{args.gpus}M-node graph needs ~{27 * args.gpus} GB of edge-level temporaries")
    line = {
        "impl": "reference", "metric": METRIC, "value": base["value"], "unit": UNIT,
        "n_gpus": args.gpus, "steps": n_timed, "warmup": n_warm, "ms_per_step": t * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
        "data": "synthetic", "config": cfg,
        "cpu_baseline": base,
        "e2e": {"value": base["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit_json(line)


def bench_config(n_gpus):
    return {"workload": f"MagNetConv single-layer forward (K=1, q=0.25, sym, cached=True), synthetic DSBM "
                        f"(cyclic K=3, eta=0.1, size_ratio=1.5), {n_gpus}x(1M nodes / 20M directed edges), "
                        f"64->64 features, fp32 (BASELINE configs[1] layer shape)",
            "nodes_per_gpu": N_PER_GPU, "edges_per_gpu": E_PER_GPU, "feat": FEAT,
            "parallelism": "single GPU" if n_gpus == 1 else
                           f"node-range row shards x{n_gpus}; feature shards all-gathered over NVLink by the shard-push "
                           f"kernel (each rank stores its rows into every peer's receive planes in symmetric memory, "
                           f"slice by slice with flags), overlapped with the aggregation of the own-column block and "
                           f"of every landed slice; operator built per rank (row-range build + degree all-gather)",
            "l2": "inputs larger than L2 (x_real+x_imag 512 MB, plan 0.5 GB vs 126 MB L2); no explicit flush"}


# ------------------------------------------------------------------------------------ GPU arm

def parity_check(rank, world, dev, ei_cpu, n_total, x_real, x_imag, outs, conv, row_lo, n_check=2000):
    """In-line correctness evidence for the line being printed: n_check destination rows of rank 0's
    outputs against the oracle (port.magnet_norm_rows builds those rows of the operator from the edge
    list on the CPU, port.magnet_conv_rows runs the reference's op sequence on them).  Collective: every
    rank contributes its feature shard; only rank 0 computes and returns the dict."""
    import torch.distributed as dist
    if world > 1:
        full = []
        for x in (x_real, x_imag):
            buf = torch.empty((n_total, x.size(1)), dtype=x.dtype, device=dev)
            dist.all_gather_into_tensor(buf, x.contiguous())
            full.append(buf.cpu() if rank == 0 else None)
            del buf
    else:
        full = [x_real.cpu(), x_imag.cpu()]
    if rank != 0:
        return None
    try:
        return _parity_rows(dev, ei_cpu, n_total, full, outs, conv, row_lo, n_check)
    except Exception as exc:  # noqa: BLE001 - an extra key must never cost the headline line
        log(f"parity_check failed: {type(exc).__name__}: {exc}")
        return {"error": f"{type(exc).__name__}: {exc}"[:300], "ok": False}


def _parity_rows(dev, ei_cpu, n_total, full, outs, conv, row_lo, n_check):
    from oracle import port
    t0 = time.time()
    threads = torch.get_num_threads()
    torch.set_num_threads(max(1, min(32, os.cpu_count() or 1)))
    g = torch.Generator().manual_seed(99)
    n_local = outs[0].size(0)
    rows_local = torch.randperm(n_local, generator=g)[:n_check].sort().values
    rows = rows_local + row_lo
    with torch.no_grad():
        cached = port.magnet_norm_rows(rows, ei_cpu, None, n_total, 0.25, "sym", 2.0)
        ref = port.magnet_conv_rows(rows, full[0], full[1], cached, conv.weight.detach().cpu(),
                                    conv.bias.detach().cpu())
    torch.set_num_threads(threads)
    worst = 0.0
    for got, want in zip(outs, ref):
        got = got[rows_local.to(dev)].cpu()
        worst = max(worst, float((got - want).abs().max() / want.abs().max()))
    return {"max_rel_err": worst, "rows": int(rows.numel()), "tolerance": 1e-5, "ok": bool(worst <= 1e-5),
            "against": "oracle/port.py: magnet_norm_rows (operator rows from the edge list) + magnet_conv_rows",
            "seconds": round(time.time() - t0, 1)}


def time_steps(step, steps, warmup, barrier):
    for _ in range(warmup):
        step()
    barrier()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    for _ in range(steps):
        step()
    ev1.record()
    barrier()
    return ev0.elapsed_time(ev1) / steps


def run_gpu_arm(args, rank, world):
    import torch.distributed as dist
    from pytorch_geometric_signed_directed_b200 import nn, ops, synthetic

    local_rank = int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    n_total, e_total = N_PER_GPU * world, E_PER_GPU * world
    t0 = time.time()
    ei, _ = synthetic.dsbm_edges(n_total, 3, num_edges=e_total, eta=0.1, size_ratio=1.5, seed=0, device=dev)
    e_input = ei.size(1)
    gen = torch.Generator(device=dev).manual_seed(1234 + rank)
    conv = nn.MagNetConv(FEAT, FEAT, K=1, q=0.25, trainable_q=False, cached=True).to(dev)
    torch.manual_seed(0)
    with torch.no_grad():
        conv.weight.copy_(torch.rand(2, FEAT, FEAT) - 0.5)
        conv.bias.uniform_(-0.1, 0.1)

    if world == 1:
        n_local = n_total
        x_real = torch.rand(n_local, FEAT, generator=gen, device=dev) * 2 - 1
        x_imag = torch.rand(n_local, FEAT, generator=gen, device=dev) * 2 - 1
        conv(x_real, x_imag, ei)                       # builds & caches the plan
        plan = conv._plan
        def step():
            with torch.no_grad():          # forward-only metric: the inference path of the layer
                return conv(x_real, x_imag, ei)
        nnz, n_rows = plan.nnz, n_local
    else:
        from pytorch_geometric_signed_directed_b200 import distributed as pgd
        sharded = pgd.ShardedMagNetConv(conv, n_total, rank, world)
        sharded.build(ei)
        n_local = sharded.n_local
        x_real = torch.rand(n_local, FEAT, generator=gen, device=dev) * 2 - 1
        x_imag = torch.rand(n_local, FEAT, generator=gen, device=dev) * 2 - 1
        def step():
            with torch.no_grad():
                return sharded(x_real, x_imag)
        nnz, n_rows = sharded.local_nnz, n_local
        exchange_mode = sharded.agg.mode
    ei_cpu = ei.cpu() if rank == 0 else None
    if world > 1:
        del ei
    torch.cuda.synchronize()
    log(f"[rank {rank}] setup {time.time() - t0:.1f}s: N={n_total} E={e_input} local nnz={nnz}")

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- device-resident timing (clock sampler spans warm-up + timed region: the timed region alone
    # is shorter than one 200 ms nvidia-smi period)
    sampler = ClockSampler(local_rank)
    sampler.start()
    for _ in range(args.warmup):
        step()
    barrier()
    ops.TIMING = []
    launches0 = ops.LAUNCHES
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    ev0.record()
    for _ in range(args.steps):
        step()
    ev1.record()
    barrier()
    total_ms = ev0.elapsed_time(ev1)
    from pytorch_geometric_signed_directed_b200 import _lib as _pl
    spmm_kernel = (_pl.load().pgsd_last_spmm_kernel() or b"").decode() or "spmm kernel"
    launches = ops.LAUNCHES - launches0
    timing, ops.TIMING = ops.TIMING, None
    clocks = sampler.stop()
    kern = {}
    for name, a, b in timing:
        kern.setdefault(name, []).append(a.elapsed_time(b))
    t_ms = torch.tensor([total_ms / args.steps], device=dev)
    if world > 1:
        dist.all_reduce(t_ms, op=dist.ReduceOp.MAX)
    ms_per_step = float(t_ms.item())

    # ---- in-line parity check of the path that was just timed (rank 0's rows against the oracle)
    if world > 1:
        sharded.agg.check()
    with torch.no_grad():
        outs = step()
    parity = parity_check(rank, world, dev, ei_cpu, n_total, x_real, x_imag, outs, conv,
                          0 if world == 1 else sharded.bounds[rank])
    del outs
    # rank 0 spends seconds on the CPU in the check: the other ranks wait HERE (NCCL barrier), not inside the flag
    # waits of their next sharded step, which give up after PGSD_WAIT_TIMEOUT_S
    barrier()

    # ---- end-to-end: host (pinned) features in, host (pinned) outputs back, every step
    e2e = None

    def e2e_sharded():
        # every rank uploads its shard from pinned host memory, runs the sharded forward and downloads its
        # rows of the outputs; max over ranks
        hx_r, hx_i = x_real.cpu().pin_memory(), x_imag.cpu().pin_memory()
        ho_r = torch.empty((n_local, FEAT), dtype=torch.float32).pin_memory()
        ho_i = torch.empty((n_local, FEAT), dtype=torch.float32).pin_memory()
        dx_r, dx_i = torch.empty_like(x_real), torch.empty_like(x_imag)

        @torch.no_grad()
        def e2e_step_sharded():
            dx_r.copy_(hx_r, non_blocking=True)
            dx_i.copy_(hx_i, non_blocking=True)
            o_r, o_i = sharded(dx_r, dx_i)
            ho_r.copy_(o_r, non_blocking=True)
            ho_i.copy_(o_i, non_blocking=True)

        n_e2e = max(3, min(args.steps, 10))
        t_e2e = torch.tensor([time_steps(e2e_step_sharded, n_e2e, 2, barrier)], device=dev)
        dist.all_reduce(t_e2e, op=dist.ReduceOp.MAX)
        e2e_ms = float(t_e2e.item())
        res = {"value": e_input / (e2e_ms * 1e-3), "unit": UNIT, "ms_per_step": e2e_ms,
               "h2d_bytes_per_step": 2 * n_local * FEAT * 4 * world, "d2h_bytes_per_step": 2 * n_local * FEAT * 4 * world,
               "steps": n_e2e, "note": "per rank: pinned host shard of x_real/x_imag -> H2D -> sharded forward (push "
                                       "exchange + aggregation + transform) -> D2H of its output rows, every step; "
                                       "serial on each rank's stream, max over ranks; bytes are summed over ranks"}
        return res

    if world > 1:
        try:
            e2e = e2e_sharded()
        except Exception as exc:  # noqa: BLE001 - keep the device-resident line even if the host leg fails
            log(f"[rank {rank}] e2e leg failed: {type(exc).__name__}: {exc}")
            e2e = {"error": f"{type(exc).__name__}: {exc}"[:300]}

    if world == 1:
        hx_r, hx_i = x_real.cpu().pin_memory(), x_imag.cpu().pin_memory()
        ho_r = torch.empty((n_local, FEAT), dtype=torch.float32).pin_memory()
        ho_i = torch.empty((n_local, FEAT), dtype=torch.float32).pin_memory()
        dx_r, dx_i = torch.empty_like(x_real), torch.empty_like(x_imag)

        @torch.no_grad()
        def e2e_step():
            dx_r.copy_(hx_r, non_blocking=True)
            dx_i.copy_(hx_i, non_blocking=True)
            o_r, o_i = conv(dx_r, dx_i, ei)
            ho_r.copy_(o_r, non_blocking=True)
            ho_i.copy_(o_i, non_blocking=True)

        for _ in range(max(1, args.warmup // 2)):
            e2e_step()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        n_e2e = max(3, min(args.steps, 10))
        e0.record()
        for _ in range(n_e2e):
            e2e_step()
        e1.record()
        torch.cuda.synchronize()
        serial_ms = e0.elapsed_time(e1) / n_e2e

        # Same work per step, software-pipelined across steps: the H2D of step i+1 and the D2H of
        # step i-1 ride on their own streams (PCIe is full duplex) while step i computes.  Every
        # step still uploads its inputs from pinned memory and downloads its outputs.
        s_in, s_out, s_cmp = torch.cuda.Stream(), torch.cuda.Stream(), torch.cuda.current_stream()
        dxs = [(dx_r, dx_i), (torch.empty_like(dx_r), torch.empty_like(dx_i))]
        ev_in = [torch.cuda.Event() for _ in range(2)]
        ev_cmp = [torch.cuda.Event() for _ in range(2)]
        ev_out = [torch.cuda.Event() for _ in range(2)]

        @torch.no_grad()
        def pipelined(n):
            for i in range(n):
                b = i % 2
                with torch.cuda.stream(s_in):
                    s_in.wait_event(ev_cmp[b])            # compute of step i-2 finished with dxs[b]
                    dxs[b][0].copy_(hx_r, non_blocking=True)
                    dxs[b][1].copy_(hx_i, non_blocking=True)
                    ev_in[b].record(s_in)
                s_cmp.wait_event(ev_in[b])
                o_r, o_i = conv(dxs[b][0], dxs[b][1], ei)
                ev_cmp[b].record(s_cmp)
                o_r.record_stream(s_out); o_i.record_stream(s_out)
                with torch.cuda.stream(s_out):
                    s_out.wait_event(ev_cmp[b])
                    ho_r.copy_(o_r, non_blocking=True)
                    ho_i.copy_(o_i, non_blocking=True)
                    ev_out[b].record(s_out)
            s_cmp.wait_stream(s_out)

        pipelined(3)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        pipelined(n_e2e)
        e1.record()
        torch.cuda.synchronize()
        e2e_ms = e0.elapsed_time(e1) / n_e2e
        # both loops do the same per-step work; which one is faster depends on the host (on boxes whose
        # host memory cannot feed both PCIe directions at once the overlapped loop loses)
        pipelined_ms, e2e_ms = e2e_ms, min(e2e_ms, serial_ms)
        e2e = {"value": e_input / (e2e_ms * 1e-3), "unit": UNIT, "ms_per_step": e2e_ms,
               "serial_ms_per_step": serial_ms, "pipelined_ms_per_step": pipelined_ms,
               "h2d_bytes_per_step": 2 * n_local * FEAT * 4, "d2h_bytes_per_step": 2 * n_local * FEAT * 4,
               "steps": n_e2e, "note": "pinned host x_real/x_imag -> H2D -> MagNetConv.forward (C ABI) -> D2H "
                                       "out_real/out_imag every step; graph plan cached on device (cached=True); "
                                       "copies double-buffered on side streams (serial_ms_per_step = same loop "
                                       "without overlap)"}

    # ---- cold path: cached=False (the reference's constructor default) rebuilds the operator from the
    # COO edge list inside every forward (symmetrise, 64-bit key sort, coalesce, degree, phase)
    cold_ms = uncached_same_tensors_ms = None
    if world == 1:
        conv_cold = nn.MagNetConv(FEAT, FEAT, K=1, q=0.25, trainable_q=False, cached=False).to(dev)
        conv_cold.load_state_dict(conv.state_dict())
        with torch.no_grad():
            def cold_step():
                conv_cold._rebuild_cache.clear()        # force the rebuild (same tensors would reuse the plan)
                conv_cold(x_real, x_imag, ei)
            for _ in range(2):
                cold_step()
            torch.cuda.synchronize()
            c0, c1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            c0.record()
            for _ in range(5):
                cold_step()
            c1.record()
            torch.cuda.synchronize()
            cold_ms = c0.elapsed_time(c1) / 5
            # cached=False called again with the SAME edge tensors: the plan of the last call is reused
            for _ in range(2):
                conv_cold(x_real, x_imag, ei)
            c0, c1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            c0.record()
            for _ in range(5):
                conv_cold(x_real, x_imag, ei)
            c1.record()
        torch.cuda.synchronize()
        uncached_same_tensors_ms = c0.elapsed_time(c1) / 5
        del conv_cold

    # ---- x_real and x_imag being ONE tensor (how examples/magnet_node.py:61-62 call the first layer of
    # every MagNet model): the kernel gathers each neighbour row once for both operators.  Reported
    # beside the headline, never instead of it (the headline uses two distinct tensors).
    shared = None
    if world == 1:
        with torch.no_grad():
            for _ in range(3):
                conv(x_real, x_real, ei)
            torch.cuda.synchronize()
            ops.TIMING = []
            c0, c1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            c0.record()
            for _ in range(10):
                conv(x_real, x_real, ei)
            c1.record()
        torch.cuda.synchronize()
        tm, ops.TIMING = ops.TIMING, None
        sh_ms = c0.elapsed_time(c1) / 10
        sh_spmm = sum(a.elapsed_time(b) for nm, a, b in tm if nm == "spmm") / 10
        # one gather per stored entry; x read once by the transform's first two terms is still counted twice
        b_sh = nnz * (4 + 8 + FEAT * 4) + (n_rows + 1) * 4 + 4 * n_rows * FEAT * 4
        shared = {"ms_per_step": sh_ms, "value": e_input / (sh_ms * 1e-3), "unit": UNIT, "spmm_ms": sh_spmm,
                  "algorithmic_bytes": b_sh, "note": "x_real is x_imag (same tensor object): one gather per entry"}

    # ---- halo path (extra key, never the headline): a graph whose edge list shards naturally -- node ids with
    # locality (|i - j| <= 50k), the same 1M nodes / 20M edges per rank -- takes the all-to-all of packed halo
    # rows instead of the all-gather; timed and parity-checked the same way
    halo = None

    def halo_leg():
        ei2 = synthetic.locality_edges(n_total, e_total, 50_000, 0.0, seed=1, device=dev)
        sh2 = pgd.ShardedMagNetConv(conv, n_total, rank, world).build(ei2)
        e2_input = ei2.size(1)
        ei2_cpu = ei2.cpu() if rank == 0 else None
        del ei2

        def halo_step():
            with torch.no_grad():
                return sh2(x_real, x_imag)

        t_h = torch.tensor([time_steps(halo_step, args.steps, args.warmup, barrier)], device=dev)
        dist.all_reduce(t_h, op=dist.ReduceOp.MAX)
        par2 = parity_check(rank, world, dev, ei2_cpu, n_total, x_real, x_imag, halo_step(), conv, sh2.bounds[rank])
        barrier()
        res = {"ms_per_step": float(t_h.item()), "value": e2_input / (float(t_h.item()) * 1e-3), "unit": UNIT,
                "mode": sh2.agg.mode, "halo_fraction": getattr(sh2.agg, "halo_fraction", None),
                "halo_rows_received_rank0": sh2.agg.halo.n_recv if sh2.agg.halo else None,
                "edges_total": e2_input, "parity_check": par2,
                "graph": "synthetic.locality_edges: |i - j| <= 50k, 1M nodes / 20M edges per rank, unit weights"}
        return res

    if world > 1 and not args.no_halo:
        try:
            halo = halo_leg()
        except Exception as exc:  # noqa: BLE001 - an extra key must never cost the headline line
            log(f"[rank {rank}] halo-path leg failed: {type(exc).__name__}: {exc}")
            halo = {"error": f"{type(exc).__name__}: {exc}"[:300]}

    if rank != 0:
        return

    # ---- roofline of the dominant kernel (pgsd_spmm_csr, n_ops = 2)
    peak, peak_src = measured_peak_gbs()
    # per STEP: the sharded path issues one aggregation launch per column block, their sum is the
    # step's aggregation time
    spmm_ms = sum(kern["spmm"]) / args.steps if kern.get("spmm") else None
    dense_ms = sum(kern["dense"]) / args.steps if kern.get("dense") else None
    # algorithmic bytes per launch: per stored entry 4 (col) + 2*4 (values) +
    # 2 * F*4 (one feature-row gather per operator, no reuse assumed); per row 4 (row_ptr) +
    # 2 * F*4 (T_real, T_imag written).
    b_spmm = nnz * (4 + 8 + 2 * FEAT * 4) + (n_rows + 1) * 4 + 2 * n_rows * FEAT * 4
    # whole layer: + dense reads of x_r, x_i and writes of out_r, out_i
    b_layer = nnz * (4 + 8 + 2 * FEAT * 4) + (n_rows + 1) * 4 + 4 * n_rows * FEAT * 4
    fused_ms = sum(kern["magnet_fused"]) / args.steps if kern.get("magnet_fused") else None
    if fused_ms:
        # the whole layer is ONE kernel (pgsd_magnet_layer_fused): its algorithmic bytes are the layer's
        roof = {"bound": "hbm", "kernel": "magnet_layer_fused_kernel (pgsd_magnet_layer_fused: aggregation + "
                                          "tcgen05 transform in one launch)", "unit": "GB/s",
                "peak": peak, "peak_source": peak_src, "algorithmic_bytes": b_layer,
                "traffic": recorded_traffic("fused_dram_traffic.json"),
                "achieved": b_layer / (fused_ms * 1e-3) / 1e9, "kernel_ms": fused_ms,
                "launches_per_step": len(kern["magnet_fused"]) / args.steps,
                "share_of_step": fused_ms / ms_per_step}
        roof["frac"] = roof["achieved"] / peak
    else:
        roof = {"bound": "hbm", "kernel": f"{spmm_kernel} (pgsd_spmm_csr, n_ops=2)", "unit": "GB/s",
                "peak": peak, "peak_source": peak_src, "algorithmic_bytes": b_spmm,
                # the capture is of ONE launch over the whole shard: not comparable with the per-block launches of N > 1
                "traffic": recorded_traffic(kernel=spmm_kernel) if world == 1 else None}
        if spmm_ms:
            roof["achieved"] = b_spmm / (spmm_ms * 1e-3) / 1e9
            roof["frac"] = roof["achieved"] / peak
            roof["kernel_ms"] = spmm_ms
            roof["launches_per_step"] = len(kern["spmm"]) / args.steps
            roof["share_of_step"] = spmm_ms / ms_per_step
            if world > 1:
                roof["note"] = ("rank 0 of the sharded run: kernel_ms sums the per-shard column-block launches of "
                                "one step; the step is bounded by the NVLink exchange, see DESIGN.md §7")
    roof["layer"] = {"algorithmic_bytes": b_layer, "achieved": b_layer / (ms_per_step * 1e-3) / 1e9,
                     "frac": b_layer / (ms_per_step * 1e-3) / 1e9 / peak, "dense_ms": dense_ms}

    cpu_base = None
    if world == 1 and not args.no_cpu_baseline:
        log("timing the CPU baseline (oracle port, whole 1M/20M graph) on the host cores ...")
        cpu_base = cpu_reference_run(3, 1, budget_s=45.0)[0]

    line = {
        "metric": METRIC, "value": e_input / (ms_per_step * 1e-3), "unit": UNIT,
        "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_per_step,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
        "data": "synthetic", "config": bench_config(world),
        "roofline": roof, "cpu_baseline": cpu_base, "e2e": e2e, "parity_check": parity, "gpu_launches": launches,
        "clocks": clocks, "edges_total": e_input, "nnz_per_rank": nnz,
        # neighbour feature rows gathered per second (one per stored entry and operator, all ranks)
        "messages_per_s": 2 * nnz * world / (ms_per_step * 1e-3),
        "cold_ms_per_step": cold_ms,   # cached=False: plan build + forward (reference: ~48 s on CPU)
        "uncached_same_tensors_ms_per_step": uncached_same_tensors_ms,   # cached=False, identical edge tensors again
        "shared_input": shared,
    }
    if world > 1:
        agg = sharded.agg
        push = agg._push[1] if agg._push else None
        pull = agg._pull[1] if agg._pull else None
        if push is not None:
            transport = {"name": "push into the peers' receive planes in symmetric memory, per-slice flags "
                                 "(pgsd_shard_push / pgsd_peer_copy + pgsd_signal_flag, pgsd_wait_flags)",
                         "engine": {0: "LSU push kernel", 1: "bulk-copy (TMA) push kernel",
                                    2: "copy engines (cudaMemcpyAsync peer copies)"}[push.engine],
                         "ctas": push.n_ctas if push.engine != 2 else 0,
                         "copy_streams": len(push.ce_streams),
                         "slices": [round(c, 3) for c in agg.stage_cum], "multicast": bool(push.mc_ptr),
                         "nvlink_bytes_in_per_rank": (world - 1) * n_local * FEAT * 4 * 2}
        else:
            transport = "symmetric-memory copy-engine pull" if pull is not None else "nccl send/recv ring"
        line["exchange"] = {"mode": exchange_mode, "transport": transport,
                            "halo_fraction": getattr(agg, "halo_fraction", None)}
        line["halo_path"] = halo
    if shared:
        shared["frac"] = shared["algorithmic_bytes"] / (shared["ms_per_step"] * 1e-3) / 1e9 / peak
    emit_json(line)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-halo", action="store_true", help="N > 1: skip the extra halo-path measurement")
    args = ap.parse_args()
    protect_stdout()
    args.warmup = max(args.warmup, 3)
    rank = int(os.environ.get("RANK", 0))
    world = int(os.environ.get("WORLD_SIZE", 1))
    if args.impl == "reference":
        run_reference_arm(args, rank)
        return
    if world != args.gpus:
        if args.gpus > 1 and world == 1:
            # convenience: relaunch under torchrun
            cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1",
                   f"--nproc-per-node={args.gpus}", "--master-addr", "127.0.0.1", "--master-port", "29511",
                   os.path.abspath(__file__), "--gpus", str(args.gpus), "--steps", str(args.steps),
                   "--warmup", str(args.warmup)] + (["--no-cpu-baseline"] if args.no_cpu_baseline else []) \
                  + (["--no-halo"] if args.no_halo else [])
            sys.exit(subprocess.call(cmd, stdout=_REAL_STDOUT))
        raise SystemExit(f"--gpus {args.gpus} but WORLD_SIZE={world}")
    run_gpu_arm(args, rank, world)
    if world > 1:
        import torch.distributed as dist
        if dist.is_initialized():
            dist.destroy_process_group()


if __name__ == "__main__":
    main()

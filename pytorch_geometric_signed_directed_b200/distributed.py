"""Node-range sharded aggregation for one 8xB200 NVLink/NVSwitch box.

The reference is single-process/single-device (no torch.distributed anywhere); this module has
no reference counterpart.  The path shards by DESTINATION ROWS: rank r owns the nodes
[bounds[r], bounds[r+1]), their CSR rows, their slice of x_real/x_imag and of the outputs.
Rows are independent given x, so the only exchange is the feature rows a rank's columns point at.

All-gather path (permuted graphs: every rank references ~all remote nodes).  `PushExchange`
(csrc/exchange.cu): every rank STORES its rows into the peers' [n_total, F] receive planes in
symmetric memory (push kernel with LSU or bulk-copy/TMA engine, or the copy engines), in row
slices published by flags; the row shard's entries are split into stage blocks -- own columns
first, then the columns of every slice as it lands -- that accumulate through the aggregation
kernel's `beta * z` epilogue:

    exchange stream :  push slice 0 | slice 1 | slice 2 | ...            (flags per slice)
    compute stream  :  own block    | wait(0) stage 1 | wait(1) stage 2 | ... | transform

Receive planes and flags are double-buffered by step parity, so there is no per-step barrier.
Fallbacks behind PGSD_EXCHANGE: copy-engine pulls of whole shards + per-owner column blocks
(round 1), NCCL send/recv ring (also what the gloo CPU tests drive).

Graphs whose edge list shards naturally (locality-ordered node ids: most columns of a row shard
are local, the rest touch a thin band of each peer) take the HALO path instead (`mode="halo"`,
chosen automatically when every rank needs less than PGSD_HALO_THRESHOLD of the remote rows): the
plan is analysed once -- sorted unique remote columns per owner, request lists exchanged with an
all-to-all -- and every step packs the rows each peer asked for (`pgsd_gather_rows`), exchanges
them with ONE `all_to_all_single` (NCCL over NVLink) that runs while the local-column block is
aggregated, and then aggregates the remote-column block straight from the received buffer (its
columns are compact indices into that buffer).  DESIGN.md §7 has the measurements behind the defaults.
"""
from __future__ import annotations

import os
import sys
from typing import Callable, List, Optional, Sequence

import torch
import torch.distributed as dist
from torch import Tensor

from . import ops, plan as _plan
from .plan import CSRPlan


# Set to a list to record (label, cuda event) pairs of one sharded step on the compute stream
# (tools/dist_check.py --trace prints the timeline); None = off.
TRACE = None


def _now_event():
    ev = torch.cuda.Event(enable_timing=True)
    ev.record()
    return ev


def node_bounds(n_total: int, world: int) -> List[int]:
    """bounds[r] = first node of rank r; bounds[world] = n_total (1-D node-range split)."""
    return [(r * n_total) // world for r in range(world + 1)]


def split_rows(plan: CSRPlan, lo: int, hi: int) -> CSRPlan:
    """Rows [lo, hi) of a plan, columns still global."""
    rp = plan.row_ptr[lo:hi + 1]
    a, b = int(rp[0].item()), int(rp[-1].item())
    vals = [None if v is None else v[a:b] for v in plan.val]
    diags = [None if d is None else d[lo:hi] for d in plan.diag]
    return CSRPlan(hi - lo, plan.n_src, b - a, plan.num_input_edges, (rp - rp[0]).contiguous(),
                   plan.col[a:b], vals, diags, list(plan.diag_const),
                   {k: v for k, v in plan.meta.items() if k != "hermitian"})


def split_columns_by_owner(local: CSRPlan, bounds: Sequence[int], own_rank: int) -> List[CSRPlan]:
    """Split a row-shard's entries into one CSR block per column owner; block b's columns are
    re-based to shard b (col - bounds[b]).  Entry order inside a row is preserved (stable), so
    per-row accumulation order stays deterministic.  The diagonal term lives in block own_rank
    (x[row] is a local row there); the other blocks have no diagonal."""
    dev = local.row_ptr.device
    n_rows, world = local.n_dst, len(bounds) - 1
    counts = (local.row_ptr[1:] - local.row_ptr[:-1]).long()
    rows = torch.repeat_interleave(torch.arange(n_rows, device=dev), counts)
    col = local.col.long()
    b_t = torch.tensor(list(bounds), device=dev, dtype=torch.long)
    owner = torch.bucketize(col, b_t[1:], right=True)
    order = torch.sort(owner * n_rows + rows, stable=True).indices
    owner_s, rows_s, col_s = owner[order], rows[order], col[order]
    per_owner = torch.bincount(owner_s, minlength=world).tolist()
    blocks, start = [], 0
    for b in range(world):
        m = per_owner[b]
        sl = slice(start, start + m)
        rp = torch.zeros(n_rows + 1, dtype=torch.int32, device=dev)
        if m:
            rp[1:] = torch.cumsum(torch.bincount(rows_s[sl], minlength=n_rows), 0).int()
        vals = [None if v is None else v[order[sl]].contiguous() for v in local.val]
        is_own = b == own_rank
        diags = [d if is_own else None for d in local.diag]
        dconst = [c if is_own else 0.0 for c in local.diag_const]
        blocks.append(CSRPlan(n_rows, bounds[b + 1] - bounds[b], m, local.num_input_edges, rp,
                              (col_s[sl] - bounds[b]).int().contiguous(), vals, diags, dconst))
        start += m
    return blocks


def stage_fractions(spec=None, world: int = 2) -> List[float]:
    """Cumulative row fractions of the slices a shard travels in (push exchange).  `spec` / PGSD_PUSH_SLICES is
    a count (equal slices) or a comma list of weights.  Defaults are the best of the sweeps on 1M nodes / 20M edges
    per rank (profiles/r02_sweep_n{2,4,8}_*.jsonl): every extra slice costs one more aggregation launch that re-reads
    and re-writes the outputs (1 GB), so few slices win while the aggregation is the bottleneck (2 ranks: the whole
    peer shard lands before the own-column block is done -> one slice; 3-4 ranks: two); at 8 ranks the exchange
    paces the step and four slices keep the block that runs after the last arrival at a quarter of the remote work."""
    if spec is None:
        spec = os.environ.get("PGSD_PUSH_SLICES", "")
    spec = str(spec)
    if not spec or spec == "auto":
        spec = "1" if world <= 2 else ("2" if world <= 4 else "4")
    w = [float(t) for t in spec.split(",") if t.strip()] if "," in spec else [1.0] * max(1, int(spec))
    tot, acc, cum = sum(w), 0.0, [0.0]
    for v in w:
        acc += v
        cum.append(acc / tot)
    cum[-1] = 1.0
    return cum


def slice_rows(n_rows: int, cum: Sequence[float]) -> List[int]:
    """Row boundaries of the slices of an n_rows shard (same formula on the sender and on every receiver)."""
    out = [min(n_rows, int(round(c * n_rows))) for c in cum]
    out[0], out[-1] = 0, n_rows
    for i in range(1, len(out)):
        out[i] = max(out[i], out[i - 1])
    return out


def split_columns_by_stage(local: CSRPlan, bounds: Sequence[int], own_rank: int, cum: Sequence[float]) -> List[CSRPlan]:
    """Blocks of the push exchange: block 0 = entries whose column this rank owns (columns re-based to the shard,
    carries the diagonal); block s >= 1 = entries whose column lies in slice s-1 of ANY peer's shard, columns kept
    GLOBAL (they index the [n_total, F] receive planes).  Entry order inside a row is preserved."""
    dev = local.row_ptr.device
    n_rows, world = local.n_dst, len(bounds) - 1
    n_stage = len(cum) - 1
    counts = (local.row_ptr[1:] - local.row_ptr[:-1]).long()
    rows = torch.repeat_interleave(torch.arange(n_rows, device=dev), counts)
    col = local.col.long()
    b_t = torch.tensor(list(bounds), device=dev, dtype=torch.long)
    owner = torch.bucketize(col, b_t[1:], right=True)
    # per-owner slice boundaries in GLOBAL row numbers: [world, n_stage + 1]
    cuts = torch.tensor([[bounds[b] + r for r in slice_rows(bounds[b + 1] - bounds[b], cum)] for b in range(world)],
                        device=dev, dtype=torch.long)
    stage = torch.ones_like(col)
    for k in range(1, n_stage):                      # stage = 1 + number of interior cuts at or below col
        stage += (col >= cuts[owner, k]).long()
    stage[owner == own_rank] = 0
    order = torch.sort(stage * n_rows + rows, stable=True).indices
    stage_s, rows_s, col_s = stage[order], rows[order], col[order]
    per_stage = torch.bincount(stage_s, minlength=n_stage + 1).tolist()
    blocks, start = [], 0
    lo = bounds[own_rank]
    for k in range(n_stage + 1):
        m = per_stage[k]
        sl = slice(start, start + m)
        rp = torch.zeros(n_rows + 1, dtype=torch.int32, device=dev)
        if m:
            rp[1:] = torch.cumsum(torch.bincount(rows_s[sl], minlength=n_rows), 0).int()
        vals = [None if v is None else v[order[sl]].contiguous() for v in local.val]
        own = k == 0
        diags = [d if own else None for d in local.diag]
        dconst = [c if own else 0.0 for c in local.diag_const]
        cols = (col_s[sl] - lo) if own else col_s[sl]
        blocks.append(CSRPlan(n_rows, (bounds[own_rank + 1] - lo) if own else bounds[-1], m, local.num_input_edges,
                              rp, cols.int().contiguous(), vals, diags, dconst))
        start += m
    return blocks


def split_local_and_halo(local: CSRPlan, bounds: Sequence[int], own_rank: int):
    """(own_block, halo_block, need): own_block keeps the entries whose column this rank owns
    (columns re-based to the shard, carries the diagonal); halo_block keeps the others with the
    column replaced by its position in the sorted list of DISTINCT remote columns -- which is the
    row order of the receive buffer, because peers are laid out by rank and each sends the rows it
    was asked for in ascending order.  need[b] = rows of shard b this rank reads (ascending, re-based
    to shard b, int32); need[own_rank] is empty.  Entry order inside a row is preserved."""
    dev = local.row_ptr.device
    n_rows, world = local.n_dst, len(bounds) - 1
    lo, hi = bounds[own_rank], bounds[own_rank + 1]
    counts = (local.row_ptr[1:] - local.row_ptr[:-1]).long()
    rows = torch.repeat_interleave(torch.arange(n_rows, device=dev), counts)
    col = local.col.long()
    is_own = (col >= lo) & (col < hi)

    def sub_plan(mask, new_col, n_src, with_diag):
        m = int(mask.sum().item())
        rp = torch.zeros(n_rows + 1, dtype=torch.int32, device=dev)
        if m:
            rp[1:] = torch.cumsum(torch.bincount(rows[mask], minlength=n_rows), 0).int()
        vals = [None if v is None else v[mask].contiguous() for v in local.val]
        diags = [d if with_diag else None for d in local.diag]
        dconst = [c if with_diag else 0.0 for c in local.diag_const]
        return CSRPlan(n_rows, n_src, m, local.num_input_edges, rp, new_col.int().contiguous(), vals, diags, dconst)

    own_block = sub_plan(is_own, col[is_own] - lo, hi - lo, True)
    remote = ~is_own
    uniq, inv = torch.unique(col[remote], sorted=True, return_inverse=True)
    halo_block = sub_plan(remote, inv, max(int(uniq.numel()), 1), False)
    b_t = torch.tensor(list(bounds), device=dev, dtype=torch.long)
    owner = torch.bucketize(uniq, b_t[1:], right=True)
    per_owner = torch.bincount(owner, minlength=world).tolist()
    need, start = [], 0
    for b in range(world):
        need.append((uniq[start:start + per_owner[b]] - bounds[b]).int().contiguous())
        start += per_owner[b]
    return own_block, halo_block, need


def _all_to_all(out: Tensor, inp: Tensor, out_splits, in_splits, group=None, async_op: bool = False):
    """all_to_all_single with row splits (NCCL on the box; gloo implements it for the CPU tests)."""
    return dist.all_to_all_single(out, inp, list(out_splits), list(in_splits), group=group, async_op=async_op)


class HaloExchange:
    """Request lists of the halo path.  After `setup`, rank r knows for every peer b which of ITS
    rows b reads (`serve`, concatenated in peer order, with `serve_splits`) and how many rows it
    receives from each peer (`need_splits`)."""

    def __init__(self, rank: int, world: int, need: Sequence[Tensor], group=None):
        self.rank, self.world, self.group = rank, world, group
        dev = need[0].device
        self.need_splits = [int(t.numel()) for t in need]
        cnt_out = torch.tensor(self.need_splits, dtype=torch.int64, device=dev)
        cnt_in = torch.empty_like(cnt_out)
        _all_to_all(cnt_in, cnt_out, [1] * world, [1] * world, group)
        self.serve_splits = [int(v) for v in cnt_in.tolist()]
        req = torch.cat(list(need)) if sum(self.need_splits) else torch.empty(0, dtype=torch.int32, device=dev)
        self.serve = torch.empty(sum(self.serve_splits), dtype=torch.int32, device=dev)
        _all_to_all(self.serve, req, self.serve_splits, self.need_splits, group)
        self.n_recv, self.n_send = sum(self.need_splits), sum(self.serve_splits)

    def start(self, send: Tensor, recv: Tensor):
        """One all-to-all of packed halo rows: `send` [n_send, w] holds the rows peers asked for,
        `recv` [n_recv, w] receives the rows this rank asked for.  Returns the async work."""
        return _all_to_all(recv, send, self.need_splits, self.serve_splits, self.group, async_op=True)


def halo_fraction(need: Sequence[Tensor], bounds: Sequence[int], rank: int) -> float:
    """Share of the REMOTE rows this rank reads (1.0 = the halo is the whole matrix)."""
    remote = sum(bounds[b + 1] - bounds[b] for b in range(len(bounds) - 1) if b != rank)
    return float(sum(int(t.numel()) for t in need)) / max(remote, 1)


class RingExchange:
    """All-gather of equal-role shards as world-1 send/recv rounds; round s brings the shard of
    rank (rank - s) mod world.  Works on any backend (NCCL on the box, gloo in the CPU tests)."""

    def __init__(self, rank: int, world: int, group=None):
        self.rank, self.world, self.group = rank, world, group

    def source_of_round(self, s: int) -> int:
        return (self.rank - s) % self.world

    def start(self, send: Tensor, recv: Sequence[Optional[Tensor]]):
        """Posts every round; returns [(source_rank, work)] in arrival order.  recv[b] must be a
        buffer shaped like rank b's shard (recv[rank] is ignored)."""
        works = []
        for s in range(1, self.world):
            dst, src = (self.rank + s) % self.world, self.source_of_round(s)
            reqs = dist.batch_isend_irecv([
                dist.P2POp(dist.isend, send, dst, self.group),
                dist.P2POp(dist.irecv, recv[src], src, self.group)])
            works.append((src, reqs))
        return works


class _EventWork:
    """Adapter so copy-engine pulls look like NCCL works: wait() = current stream waits the event."""

    def __init__(self, event):
        self.event = event

    def wait(self):
        torch.cuda.current_stream().wait_event(self.event)


class SymmetricPullExchange:
    """The same all-gather as RingExchange, executed by the COPY ENGINES over NVLink peer memory
    instead of NCCL send/recv kernels: every rank packs its shard into a symmetric-memory buffer
    (torch.distributed._symmetric_memory: one allocation per rank, peer-mapped through the
    NVSwitch fabric), a device-side barrier publishes it, and each rank then pulls the peers'
    shards with cudaMemcpyAsync peer copies on a side stream, one event per shard.  No SM is
    spent on communication and no NCCL kernel competes with the aggregation launches for HBM
    (measured at N=4: per-block aggregation 1.28 ms with NCCL rounds in flight vs 0.73 ms alone).
    Send buffers are double-buffered by call parity; the barrier of call t also guarantees that
    every peer finished reading the buffer of call t-1, which call t+1 overwrites."""

    def __init__(self, rank: int, world: int, rows_max: int, width: int, dtype, device, group=None):
        import torch.distributed._symmetric_memory as symm_mem
        self.rank, self.world = rank, world
        grp = group if group is not None else dist.group.WORLD
        self.send = [symm_mem.empty((rows_max, width), dtype=dtype, device=device) for _ in range(2)]
        self.hdl = [symm_mem.rendezvous(t, grp) for t in self.send]
        # ONE copy stream by default: shards arrive one after the other in the order the blocks consume
        # them, so the first one is there as early as possible (N=4 timeline, session 29: first shard at
        # 1.4 ms instead of 2.0 ms with two concurrent copies sharing the link; step 4.8 vs 5.3 ms)
        self.copy_streams = [torch.cuda.Stream(device=device) for _ in range(int(os.environ.get("PGSD_COPY_STREAMS", "1")))]
        self.calls = 0
        self.width, self.dtype = width, dtype

    def source_of_round(self, s: int) -> int:
        return (self.rank - s) % self.world

    def send_buffer(self, n_rows: int) -> Tensor:
        return self.send[self.calls & 1][:n_rows]

    def start(self, send: Tensor, recv: Sequence[Optional[Tensor]]):
        i = self.calls & 1
        self.calls += 1
        hdl = self.hdl[i]
        cur = torch.cuda.current_stream()
        hdl.barrier(channel=0)                       # every rank's shard is packed and visible
        ready = torch.cuda.Event()
        ready.record(cur)
        works = []
        for cs in self.copy_streams:
            cs.wait_event(ready)
        for s in range(1, self.world):
            cs = self.copy_streams[(s - 1) % len(self.copy_streams)]
            with torch.cuda.stream(cs):
                src = self.source_of_round(s)
                peer = hdl.get_buffer(src, (recv[src].size(0), self.width), self.dtype)
                recv[src].copy_(peer, non_blocking=True)
                ev = torch.cuda.Event()
                ev.record(cs)
                works.append((src, [_EventWork(ev)]))
        return works


class PushExchange:
    """All-gather by `pgsd_shard_push` (csrc/exchange.cu): every rank stores its rows straight into the peers'
    [n_total, F] receive planes (symmetric memory, peer-mapped over NVSwitch), one read of the shard for all
    peers, in row slices published by per-slice flags.  Receive planes and flags are double-buffered by step
    parity: a peer can be at most one step ahead (it cannot finish step t+1 without this rank's step-t+1 rows,
    which are pushed only after this rank's step-t kernels), so no per-step barrier is needed."""

    def __init__(self, rank: int, world: int, bounds: Sequence[int], n_planes: int, feat: int, dtype, device,
                 cum: Sequence[float], group=None):
        import torch.distributed._symmetric_memory as symm_mem
        self.rank, self.world, self.bounds = rank, world, list(bounds)
        self.n_planes, self.feat, self.dtype = n_planes, feat, dtype
        grp = group if group is not None else dist.group.WORLD
        n_total = self.bounds[-1]
        self.row_bytes = feat * torch.empty((), dtype=dtype).element_size()
        self.plane_bytes = n_total * self.row_bytes
        self.parity_bytes = n_planes * self.plane_bytes
        self.buf = symm_mem.empty((2 * n_planes * n_total, feat), dtype=dtype, device=device)
        self.flags = symm_mem.empty((2 * _MAX_RANKS * _MAX_SLICES,), dtype=torch.int32, device=device)
        self.flags.zero_()
        torch.cuda.synchronize(device)
        self.hdl = symm_mem.rendezvous(self.buf, grp)
        self.fhdl = symm_mem.rendezvous(self.flags, grp)
        dist.barrier(group=group)                      # every rank's flags are zero before anyone pushes
        o_b, o_f = int(getattr(self.hdl, "offset", 0) or 0), int(getattr(self.fhdl, "offset", 0) or 0)
        self.buf_ptrs = [int(p) + o_b for p in self.hdl.buffer_ptrs]
        self.flag_ptrs = [int(p) + o_f for p in self.fhdl.buffer_ptrs]
        if self.buf_ptrs[rank] != self.buf.data_ptr() or self.flag_ptrs[rank] != self.flags.data_ptr():
            raise RuntimeError("symmetric-memory handle does not describe the tensor it was created from")
        mc = int(getattr(self.hdl, "multicast_ptr", 0) or 0)
        mc = mc + o_b if mc else 0
        self.mc_ptr = mc if (mc and os.environ.get("PGSD_PUSH_MC", "0") == "1") else 0
        self.counters = torch.zeros(_MAX_SLICES + 1, dtype=torch.int32, device=device)
        self.started = torch.zeros(1, dtype=torch.int32, device=device)
        self.status = torch.zeros(1, dtype=torch.int32, device=device)
        self.slices = slice_rows(self.bounds[rank + 1] - self.bounds[rank], cum)
        self.n_slices = len(self.slices) - 1
        # Transport (measured): engine 2 = copy engines (no SM; does not slow the aggregation beside it, but
        # many small copies to 7 peers top out near 420 GB/s), engine 1 = bulk-copy (TMA) push kernel, engine 0 = LSU
        # push kernel.  Up to 4 ranks the aggregation is the bottleneck -> copy engines; beyond, the exchange paces the
        # step -> the push kernel, throttled to 12 CTAs x 4 tiles of 32 KB: as fast as 32 CTAs (NVLink-bound either
        # way) but it slows the aggregation beside it 1.6x instead of 2.5x.
        self.engine = int(os.environ.get("PGSD_PUSH_ENGINE", "2" if world <= 4 else "1"))
        self.n_ctas = int(os.environ.get("PGSD_PUSH_CTAS", "12" if self.engine == 1 else "32"))
        tile = os.environ.get("PGSD_PUSH_TILE", "32768x4").split("x")
        self.chunk_bytes, self.stages = int(tile[0]), int(tile[1])
        self.spmm_carveout = int(os.environ.get("PGSD_PUSH_CARVEOUT", "0"))      # x 14 % of 228 KB
        prio = torch.cuda.Stream.priority_range()[1] if hasattr(torch.cuda.Stream, "priority_range") else -1
        self.stream = torch.cuda.Stream(device=device, priority=prio)
        # engine 2: copy engines -- cudaMemcpyAsync peer copies per (slice, peer) on PGSD_CE_STREAMS streams, each
        # slice followed by a one-word signal kernel; no SM, no gate
        self.ce_streams = [torch.cuda.Stream(device=device, priority=prio)
                           for _ in range(max(1, int(os.environ.get("PGSD_CE_STREAMS", str(min(world - 1, 3))))))] \
            if self.engine == 2 else []
        self.seq = 0
        self.done = torch.cuda.Event()
        self.planes = [[self.buf[(par * n_planes + t) * n_total:(par * n_planes + t + 1) * n_total]
                        for t in range(n_planes)] for par in range(2)]

    def push(self, xs: Sequence[Tensor]) -> int:
        """Starts the push of this step's rows on the exchange stream (ordered behind the current stream, where
        xs were produced); returns the step's sequence number."""
        self.seq += 1
        par, lo = self.seq & 1, self.bounds[self.rank]
        cur = torch.cuda.current_stream()
        self.stream.wait_stream(cur)
        off = lambda t: par * self.parity_bytes + t * self.plane_bytes + lo * self.row_bytes
        dst = [[self.buf_ptrs[p] + off(t) for p in range(self.world)] for t in range(self.n_planes)]
        fl = [self.flag_ptrs[p] + 4 * ((par * _MAX_RANKS + self.rank) * _MAX_SLICES) for p in range(self.world)]
        mc = [self.mc_ptr + off(t) for t in range(self.n_planes)] if self.mc_ptr else None
        if self.engine == 2 and all(x.is_contiguous() for x in xs):
            lib = ops._lib.load()
            for st in self.ce_streams:
                st.wait_stream(cur)
            for s in range(self.n_slices):
                r0, r1 = self.slices[s], self.slices[s + 1]
                for q in range(1, self.world):
                    p = (self.rank + q) % self.world
                    st = self.ce_streams[(q - 1) % len(self.ce_streams)].cuda_stream
                    for t, x in enumerate(xs):
                        ops._lib.check(lib.pgsd_peer_copy(dst[t][p] + r0 * self.row_bytes,
                                                          x.data_ptr() + r0 * self.row_bytes,
                                                          (r1 - r0) * self.row_bytes, st), "pgsd_peer_copy")
                    ops._lib.check(lib.pgsd_signal_flag(fl[p] + 4 * s, self.seq & 0xffffffff, st), "pgsd_signal_flag")
            ops.LAUNCHES += self.n_slices * (self.world - 1)
            for st in self.ce_streams:
                self.stream.wait_stream(st)
            self.done.record(self.stream)
            return self.seq
        ops.shard_push(list(xs), dst, self.row_bytes, self.rank, self.world, self.slices, fl, self.counters,
                       self.seq, n_ctas=self.n_ctas, mc_ptrs=mc, include_self=bool(mc), stream=self.stream,
                       engine=self.engine, chunk_bytes=self.chunk_bytes, stages=self.stages,
                       started_ptr=self.started.data_ptr())
        self.done.record(self.stream)
        # the aggregation launches fill every SM: let the push CTAs become resident first
        ops.wait_flags(self.started, [0], self.seq, self.status, float(os.environ.get("PGSD_WAIT_TIMEOUT_S", "20")))
        return self.seq

    def wait_slice(self, s: int, seq: int) -> None:
        """The current stream waits until slice s of every peer's shard of step `seq` has landed here."""
        par = seq & 1
        idx = [(par * _MAX_RANKS + b) * _MAX_SLICES + s for b in range(self.world) if b != self.rank]
        ops.wait_flags(self.flags, idx, seq, self.status, float(os.environ.get("PGSD_WAIT_TIMEOUT_S", "20")))

    def finish(self) -> None:
        """The current stream waits for this rank's own push (the source rows may then be reused)."""
        torch.cuda.current_stream().wait_event(self.done)

    def check(self) -> None:
        if int(self.status.item()) != 0:
            raise RuntimeError("push exchange: a slice did not arrive within the time-out (peer died or "
                               "the ranks disagree on the number of steps)")


_MAX_RANKS, _MAX_SLICES = 16, 16


def _distinct_operands(xs: Sequence[Tensor]) -> int:
    """1 when both operators read ONE tensor (x_real is x_imag, how the reference's example calls the first layer of
    every MagNet model): the shard then travels once, and the block launches see the same pointer twice, which the
    aggregation kernel turns into one gather per entry."""
    if len(xs) == 2 and xs[0].data_ptr() == xs[1].data_ptr() and xs[0].shape == xs[1].shape \
            and xs[0].stride() == xs[1].stride():
        return 1
    return len(xs)


def _default_aggregate(block: CSRPlan, xs, op_ids, alpha, beta, zs, out):
    return ops.spmm(block, xs, op_ids, alpha=alpha, beta=beta, zs=zs, out=out)


class ShardedAggregator:
    """y_k[local rows] = alpha * (L_k x_k)[local rows] + beta * z_k for the operators of a
    row-sharded plan, with the ring exchange overlapped block by block."""

    def __init__(self, local_plan: CSRPlan, bounds: Sequence[int], rank: int, world: int, group=None,
                 aggregate_fn: Callable = _default_aggregate, mode: Optional[str] = None,
                 gather_fn: Callable = None):
        self.bounds, self.rank, self.world = list(bounds), rank, world
        # "ring" (default): per-owner column blocks pipelined with the exchange.  "gather": receive
        # every shard into one [N, w] buffer, then ONE launch over all columns (exposes the whole
        # transfer; kept for comparison and for plans whose rows are too short to split).
        # "halo": all-to-all of packed halo rows (graphs whose edge list shards naturally).  "auto"
        # (default): halo when EVERY rank reads less than PGSD_HALO_THRESHOLD of the remote rows.
        self.mode = mode or os.environ.get("PGSD_SHARD_MODE", "auto")
        self.gather_fn = gather_fn
        self.halo = None
        if self.mode in ("auto", "halo") and world > 1:
            own_block, halo_block, need = split_local_and_halo(local_plan, bounds, rank)
            frac = torch.tensor([halo_fraction(need, bounds, rank)], dtype=torch.float32,
                                device=local_plan.row_ptr.device)
            dist.all_reduce(frac, op=dist.ReduceOp.MAX, group=group)
            self.halo_fraction = float(frac.item())
            if self.mode == "halo" or self.halo_fraction <= float(os.environ.get("PGSD_HALO_THRESHOLD", "0.5")):
                self.mode = "halo"
                self.own_block, self.halo_block = own_block, halo_block
                self.halo = HaloExchange(rank, world, need, group)
            else:
                self.mode = "ring"
            del own_block, halo_block, need
        elif self.mode in ("auto", "halo"):
            self.mode = "ring"
        # in gather mode x spans all nodes while the plan's rows are local: tell the kernel where
        # destination row 0 lives in x (diagonal term)
        self.local_plan = CSRPlan(local_plan.n_dst, local_plan.n_src, local_plan.nnz, local_plan.num_input_edges,
                                  local_plan.row_ptr, local_plan.col, local_plan.val, local_plan.diag,
                                  local_plan.diag_const, {**local_plan.meta, "diag_row_offset": self.bounds[rank]})
        # ring mode on CUDA with an NCCL group: the push exchange (one kernel over peer memory, stage blocks);
        # otherwise (gloo CPU tests, PGSD_EXCHANGE=pull|nccl) per-owner blocks + copy-engine pulls / send-recv
        self.transport = os.environ.get("PGSD_EXCHANGE", "push")
        self.use_push = (self.mode == "ring" and world > 1 and world <= _MAX_RANKS and local_plan.row_ptr.is_cuda
                         and self.transport == "push" and aggregate_fn is _default_aggregate)
        self.stage_cum = stage_fractions(world=world) if self.use_push else None
        self.stage_blocks = (split_columns_by_stage(local_plan, bounds, rank, self.stage_cum)
                             if self.use_push else None)
        self._push_cache = {}
        self._push = None
        self.blocks = (split_columns_by_owner(local_plan, bounds, rank)
                       if self.mode == "ring" and not self.use_push else None)
        self.n_local = local_plan.n_dst
        self.n_ops = len(local_plan.val)
        self.ring = RingExchange(rank, world, group)
        self.aggregate_fn = aggregate_fn
        # exchange objects and buffers are cached PER (dtype, width): a model alternates between widths (first
        # layer: one shared operand, later layers: two) and must not re-rendezvous symmetric memory every call
        self._recv_cache = {}
        self._pull_cache = {}
        self._recv = None          # (key, buffers) of the most recent call
        self._pull = None          # (key, exchange) of the most recent call

    def _pull_exchange(self, like: Tensor, width: int):
        """Copy-engine exchange over symmetric peer memory when it can be set up (CUDA tensors,
        NCCL process group, PGSD_EXCHANGE != 'nccl'); None -> NCCL send/recv ring."""
        if self.world == 1 or not like.is_cuda or os.environ.get("PGSD_EXCHANGE", "pull") == "nccl":
            return None
        key = (like.dtype, width)
        if key not in self._pull_cache:
            ex = None
            try:
                rows_max = max(self.bounds[b + 1] - self.bounds[b] for b in range(self.world))
                ex = SymmetricPullExchange(self.rank, self.world, rows_max, width, like.dtype, like.device,
                                           self.ring.group)
            except Exception as exc:                      # noqa: BLE001 - fall back to NCCL, but say so
                if self.rank == 0:
                    print(f"[pgsd] symmetric-memory exchange unavailable ({type(exc).__name__}: {exc}); "
                          "using the NCCL send/recv ring", file=sys.stderr, flush=True)
            self._pull_cache[key] = ex
        self._pull = (key, self._pull_cache[key])
        return self._pull[1]

    def _buffers(self, like: Tensor, width: int):
        key = (like.dtype, like.device, width)
        if key not in self._recv_cache:
            self._recv_cache[key] = [None if b == self.rank else
                                     torch.empty((self.bounds[b + 1] - self.bounds[b], width), dtype=like.dtype,
                                                 device=like.device) for b in range(self.world)]
        self._recv = (key, self._recv_cache[key])
        return self._recv[1]

    def __call__(self, xs: Sequence[Tensor], alpha: float = 1.0, beta: float = 0.0,
                 zs: Optional[Sequence[Tensor]] = None) -> List[Tensor]:
        n_ops, f = len(xs), xs[0].size(1)
        op_ids = tuple(range(n_ops))
        if self.mode == "gather":
            return self._gather_then_single(xs, op_ids, f, alpha, beta, zs)
        if self.mode == "halo":
            return self._halo_step(xs, op_ids, f, alpha, beta, zs)
        n_cols = _distinct_operands(xs)
        if self.use_push:
            ex = self._push_exchange(xs[0], n_cols, f)
            if ex is not None:
                return self._push_step(ex, xs, op_ids, n_cols, alpha, beta, zs)
        # interleave the DISTINCT operands: one [n_local, n_cols*F] send buffer
        pull = self._pull_exchange(xs[0], n_cols * f)
        views = lambda buf: [buf[:, (k % n_cols) * f:(k % n_cols + 1) * f] for k in range(n_ops)]
        trace = TRACE is not None and xs[0].is_cuda
        mark = (lambda name: TRACE.append((name, _now_event()))) if trace else (lambda name: None)
        mark("start")
        # (Packing and publishing on a side stream, beside the own-column block, was measured and is
        # slower -- session 30, N=4: the pack and the first pull then compete with that block for HBM,
        # the first shard lands at 1.9 ms instead of 1.4 ms.)
        if pull is not None:
            send = pull.send_buffer(self.n_local)
            for k in range(n_cols):
                send[:, k * f:(k + 1) * f].copy_(xs[k])
        else:
            send = xs[0].contiguous() if n_cols == 1 else torch.cat(list(xs), dim=1)
        recv = self._buffers(send, n_cols * f)
        if self.world == 1:
            works = []
        elif pull is not None:
            works = pull.start(send, recv)
        else:
            works = self.ring.start(send, recv)
        own_views = views(send)
        # own block first: it needs nothing from the network and carries the diagonal + beta*z
        y = self.aggregate_fn(self.blocks[self.rank], own_views, op_ids, alpha, beta, zs, None)
        mark(f"block{self.rank}(own) done")
        for src, reqs in works:
            for w in reqs:
                w.wait()            # NCCL: makes the current stream wait; gloo: blocks the host
            mark(f"shard{src} arrived")
            if self.blocks[src].nnz == 0:
                continue
            y = self.aggregate_fn(self.blocks[src], views(recv[src]), op_ids, alpha, 1.0, y, y)
            mark(f"block{src} done")
        return y


def _push_exchange(self, like: Tensor, n_cols: int, f: int):
    key = (like.dtype, n_cols, f)
    if key not in self._push_cache:
        ex = None
        if (f * like.element_size()) % 16 == 0:
            try:
                ex = PushExchange(self.rank, self.world, self.bounds, n_cols, f, like.dtype, like.device,
                                  self.stage_cum, self.ring.group)
            except Exception as exc:                      # noqa: BLE001 - fall back, but say so
                if self.rank == 0:
                    print(f"[pgsd] push exchange unavailable ({type(exc).__name__}: {exc}); "
                          "using copy-engine pulls / NCCL", file=sys.stderr, flush=True)
        if ex is None and self.blocks is None:
            self.blocks = split_columns_by_owner(self.local_plan, self.bounds, self.rank)
        self._push_cache[key] = ex
    self._push = (key, self._push_cache[key])
    return self._push[1]


def _push_step(self, ex, xs, op_ids, n_cols, alpha, beta, zs):
    """push (exchange stream) | own block -> for every slice: wait its flags -> stage block (+=)."""
    n_ops = len(xs)
    trace = TRACE is not None
    mark = (lambda name: TRACE.append((name, _now_event()))) if trace else (lambda name: None)
    mark("start")
    srcs = [xs[k] for k in range(n_cols)]
    for k in range(n_cols):
        x = srcs[k]
        if x.stride(1) != 1 or x.data_ptr() % 16 or (x.stride(0) * x.element_size()) % 16:
            srcs[k] = x.contiguous()
    seq = ex.push(srcs)
    if trace:
        # true arrival times, independent of when the aggregation gets round to waiting for a slice
        obs = getattr(ex, "_observer", None) or torch.cuda.Stream(device=srcs[0].device)
        ex._observer = obs
        obs.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(obs):
            for s in range(ex.n_slices):
                ex.wait_slice(s, seq)
                ev = torch.cuda.Event(enable_timing=True)
                ev.record(obs)
                TRACE.append((f"slice{s} arrived", ev))
            obs.wait_event(ex.done)
            ev = torch.cuda.Event(enable_timing=True)
            ev.record(obs)
            TRACE.append(("own push done", ev))
    # LSU engine: one 256-thread push CTA takes the register space of one aggregation CTA; the bulk-copy engine's
    # 32-thread CTAs fit beside a full set of aggregation CTAs (they are resident first, see PushExchange.push)
    reserve = ex.n_ctas if ex.engine == 0 else 0
    variant = ops.SPMM_VARIANT | ((ex.spmm_carveout & 7) << 12) if (ex.engine == 1 and ex.spmm_carveout) else None
    own = [srcs[k % n_cols] for k in range(n_ops)]
    y = ops.spmm(self.stage_blocks[0], own, op_ids, alpha=alpha, beta=beta, zs=zs, grid_reserve=reserve,
                 variant=variant)
    mark("own block done")
    planes = ex.planes[seq & 1]
    views = [planes[k % n_cols] for k in range(n_ops)]
    for s in range(ex.n_slices):
        ex.wait_slice(s, seq)
        mark(f"slice{s} landed")
        blk = self.stage_blocks[s + 1]
        if blk.nnz == 0:
            continue
        y = ops.spmm(blk, views, op_ids, alpha=alpha, beta=1.0, zs=y, out=y, grid_reserve=reserve, variant=variant)
        mark(f"stage{s + 1} done")
    ex.finish()
    return y


def _check(self):
    """Raises if a flag wait of the push exchange timed out (one device->host sync; call outside timed loops)."""
    for ex in self._push_cache.values():
        if ex is not None:
            ex.check()


ShardedAggregator.check = _check
ShardedAggregator._push_exchange = _push_exchange
ShardedAggregator._push_step = _push_step


def _gather_then_single(self, xs, op_ids, f, alpha, beta, zs):
    n_ops, lo, hi = len(xs), self.bounds[self.rank], self.bounds[self.rank + 1]
    key = ("full", xs[0].dtype, xs[0].device, n_ops * f)
    if key not in self._recv_cache:
        self._recv_cache[key] = torch.empty((self.bounds[-1], n_ops * f), dtype=xs[0].dtype, device=xs[0].device)
    self._recv = (key, self._recv_cache[key])
    full = self._recv[1]
    own = full[lo:hi]
    for k in range(n_ops):
        own[:, k * f:(k + 1) * f].copy_(xs[k])
    recv = [full[self.bounds[b]:self.bounds[b + 1]] for b in range(self.world)]
    for _, reqs in (self.ring.start(own, recv) if self.world > 1 else []):
        for w in reqs:
            w.wait()
    views = [full[:, k * f:(k + 1) * f] for k in range(n_ops)]
    return self.aggregate_fn(self.local_plan, views, op_ids, alpha, beta, zs, None)


ShardedAggregator._gather_then_single = _gather_then_single


def _halo_step(self, xs, op_ids, f, alpha, beta, zs):
    """pack -> all_to_all (async) | own-column block -> wait -> remote-column block (+=)."""
    n_ops, hx = len(xs), self.halo
    n_cols = _distinct_operands(xs)
    key = ("halo", xs[0].dtype, xs[0].device, n_cols * f)
    if key not in self._recv_cache:
        mk = lambda rows: torch.empty((max(rows, 1), n_cols * f), dtype=xs[0].dtype, device=xs[0].device)
        self._recv_cache[key] = (mk(hx.n_send), mk(hx.n_recv))
    self._recv = (key, self._recv_cache[key])
    send, recv = self._recv[1]
    gather = self.gather_fn or ops.gather_rows
    if hx.n_send:
        vec = (f * xs[0].element_size()) % 16 == 0 or self.gather_fn is not None
        for k in range(n_cols):                    # halo pack: the rows the peers asked for
            dst = send[:hx.n_send, k * f:(k + 1) * f]
            if vec and xs[k].data_ptr() % 16 == 0 and (xs[k].stride(0) * xs[k].element_size()) % 16 == 0:
                gather(xs[k], hx.serve, out=dst)
            else:                                  # odd widths (the reference's tests use F = 2, 3)
                dst.copy_(xs[k].index_select(0, hx.serve.long()))
    work = hx.start(send[:hx.n_send], recv[:hx.n_recv])
    y = self.aggregate_fn(self.own_block, list(xs), op_ids, alpha, beta, zs, None)
    work.wait()                                    # NCCL: the current stream waits; gloo: the host blocks
    if self.halo_block.nnz:
        views = [recv[:, (k % n_cols) * f:(k % n_cols + 1) * f] for k in range(n_ops)]
        y = self.aggregate_fn(self.halo_block, views, op_ids, alpha, 1.0, y, y)
    return y


ShardedAggregator._halo_step = _halo_step


def incident_edges(edge_index: Tensor, edge_weight: Optional[Tensor], lo: int, hi: int):
    """Edges with an endpoint in [lo, hi), original order kept (duplicates are summed in edge order)."""
    r, c = edge_index[0], edge_index[1]
    m = ((r >= lo) & (r < hi)) | ((c >= lo) & (c < hi))
    return edge_index[:, m].contiguous(), (None if edge_weight is None else edge_weight[m].contiguous())


def route_edges(edge_chunk: Tensor, weight_chunk: Optional[Tensor], bounds: Sequence[int], rank: int, world: int,
                group=None):
    """All-to-all of the symmetrised-edge build: rank r holds a contiguous slice of the
    global COO list; edge (i, j) is sent to owner(i) and, when different, owner(j).  Slices arrive in
    source-rank order and every sender keeps its own order, so the received list is the global list
    restricted to this rank's incident edges IN THE ORIGINAL ORDER -- which the in-edge-order duplicate
    sums of the builder (PyG coalesce semantics) rely on."""
    dev = edge_chunk.device
    b_t = torch.tensor(list(bounds), device=dev, dtype=torch.long)
    o_r = torch.bucketize(edge_chunk[0], b_t[1:], right=True)
    o_c = torch.bucketize(edge_chunk[1], b_t[1:], right=True)
    has_w = weight_chunk is not None
    sends, counts = [], []
    for b in range(world):
        m = (o_r == b) | (o_c == b)
        cols = [edge_chunk[0][m], edge_chunk[1][m]]
        if has_w:
            cols.append(weight_chunk[m].float().view(torch.int32).long())     # bit pattern rides along
        sends.append(torch.stack(cols, 1))
        counts.append(int(m.sum().item()))
    send = torch.cat(sends).contiguous()
    cnt_out = torch.tensor(counts, dtype=torch.int64, device=dev)
    cnt_in = torch.empty_like(cnt_out)
    _all_to_all(cnt_in, cnt_out, [1] * world, [1] * world, group)
    in_counts = [int(v) for v in cnt_in.tolist()]
    recv = torch.empty((sum(in_counts), send.size(1)), dtype=torch.int64, device=dev)
    _all_to_all(recv, send, in_counts, counts, group)
    ei = recv[:, :2].t().contiguous()
    ew = recv[:, 2].int().view(torch.float32).contiguous() if has_w else None
    return ei, ew


def allgather_rows(local: Tensor, bounds: Sequence[int], group=None) -> Tensor:
    """Concatenation over ranks of per-shard vectors of (possibly) different lengths."""
    world = len(bounds) - 1
    sizes = [bounds[b + 1] - bounds[b] for b in range(world)]
    m = max(sizes)
    padded = torch.zeros(m, dtype=local.dtype, device=local.device)
    padded[:local.numel()] = local
    out = torch.empty(world * m, dtype=local.dtype, device=local.device)
    dist.all_gather_into_tensor(out, padded, group=group)
    if all(sz == m for sz in sizes):
        return out
    return torch.cat([out[b * m:b * m + sizes[b]] for b in range(world)])


class ShardedMagNetConv:
    """MagNetConv forward over node-range shards: rank r passes its rows of x_real/x_imag and
    receives its rows of (out_real, out_imag).  Weights are replicated (they are [K+1, F, F])."""

    def __init__(self, conv, n_total: int, rank: int, world: int, group=None):
        self.conv, self.n_total, self.rank, self.world, self.group = conv, n_total, rank, world, group
        self.bounds = node_bounds(n_total, world)
        self.n_local = self.bounds[rank + 1] - self.bounds[rank]
        self.agg = None
        self.local_nnz = 0

    def build(self, edge_index: Tensor, edge_weight: Optional[Tensor] = None, lambda_max: float = 2.0,
              replicated: bool = True):
        """Distributed operator build: every rank builds ONLY its rows, from the edges
        incident to its node range, and the ranks exchange one all-gather of the node degrees
        (4 B/node) between the structure and the value phase (`plan.build_magnetic_rows`).
        replicated=True : `edge_index` is the whole edge list on every rank -> the incident edges are
                          selected locally (order-preserving mask), no edge traffic;
        replicated=False: `edge_index` is this rank's contiguous SLICE of the global edge list (slices
                          in rank order) -> edges are first routed to the owners of both endpoints with
                          an all-to-all (`route_edges`)."""
        c = self.conv
        lo, hi = self.bounds[self.rank], self.bounds[self.rank + 1]
        if os.environ.get("PGSD_REPLICATED_BUILD", "0") == "1" and replicated:
            # round-1 behaviour, kept for A/B checks: global plan on every rank, rows sliced out
            full = _plan.build_magnetic(edge_index, edge_weight, self.n_total, c._q_value(), c.normalization,
                                        lambda_max, c._signed_mode())
            local = split_rows(full, lo, hi)
            local.col = local.col.clone()
            local.val = [None if v is None else v.clone() for v in local.val]
            local.diag = [None if d is None else d.clone() for d in local.diag]
            local.meta = {}
            del full
        else:
            if replicated:
                ei, ew = incident_edges(edge_index, edge_weight, lo, hi)
            else:
                ei, ew = route_edges(edge_index, edge_weight, self.bounds, self.rank, self.world, self.group)
            local = _plan.build_magnetic_rows(ei, ew, self.n_total, lo, hi, c._q_value(), c.normalization,
                                              lambda_max, c._signed_mode(),
                                              allgather_deg=lambda d: allgather_rows(d, self.bounds, self.group))
            local.meta = {}
        self.local_nnz = local.nnz
        self.local_plan = local
        self.agg = ShardedAggregator(local, self.bounds, self.rank, self.world, self.group)
        return self

    def __call__(self, x_real: Tensor, x_imag: Tensor):
        c = self.conv
        w = c.weight.detach()
        k1 = w.size(0)
        t0 = [x_real.detach(), x_imag.detach()]
        terms = [(t0[0], w[0], 0), (t0[1], w[0], 1)]
        if k1 > 1:
            t1 = self.agg(t0)
            terms += [(t1[0], w[1], 0), (t1[1], w[1], 1)]
            for k in range(2, k1):
                t2 = self.agg(t1, alpha=2.0, beta=-1.0, zs=t0)
                terms += [(t2[0], w[k], 0), (t2[1], w[k], 1)]
                t0, t1 = t1, t2
        return tuple(ops.dense(terms, c.out_channels, bias=c.bias, combine=True,
                               relu_mode=1 if c.fused_complex_relu else 0))

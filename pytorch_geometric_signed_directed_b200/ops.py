"""Thin tensor-level wrappers over the two compute entry points of the C ABI:
`pgsd_spmm_csr` (sparse aggregation) and `pgsd_dense_transform` (adjacent dense transform).
PyTorch supplies device memory and the current stream; all arithmetic is in libpgsd_b200.so.
"""
from __future__ import annotations

import ctypes as C
import math
import os
from typing import List, Optional, Sequence, Tuple

import torch
from torch import Tensor

from . import _lib
from .plan import CSRPlan, require_cuda

_DT = {torch.float32: _lib.PGSD_F32, torch.bfloat16: _lib.PGSD_BF16}

# tuning knob for experiments (tools/sweep_spmm.py); 0 = library default
SPMM_VARIANT = int(os.environ.get("PGSD_SPMM_VARIANT", "0"))

# 0 = auto (TMA-fed tcgen05 kernel when the shape fits, else register-staged tcgen05, else FFMA),
# 1 = force FFMA, 2 = require the register-staged tcgen05 kernel, 16 = require the TMA-fed kernel
DENSE_VARIANT = int(os.environ.get("PGSD_DENSE_VARIANT", "0"))

# rows with more stored entries than this are aggregated by the hub-row kernel in slices
HUB_ROW_THRESHOLD = int(os.environ.get("PGSD_HUB_ROW_THRESHOLD", "4096"))
HUB_ROW_CHUNK = 1024

# counts kernel launches issued through this module (bench.py reports it as gpu_launches)
LAUNCHES = 0

# When set to a list, every launch is bracketed by CUDA events on the launching stream and
# (name, start_event, end_event) is appended -- bench.py uses it to time the dominant kernel
# inside the timed region without a profiler.
TIMING = None


class _Timed:
    def __init__(self, name, device):
        self.name, self.device = name, device

    def __enter__(self):
        if TIMING is not None:
            self.a = torch.cuda.Event(enable_timing=True)
            self.b = torch.cuda.Event(enable_timing=True)
            self.a.record(torch.cuda.current_stream(self.device))
        return self

    def __exit__(self, *exc):
        if TIMING is not None:
            self.b.record(torch.cuda.current_stream(self.device))
            TIMING.append((self.name, self.a, self.b))
        return False


def _dtype_code(t: Tensor) -> int:
    try:
        return _DT[t.dtype]
    except KeyError:
        raise TypeError(f"pgsd_b200 kernels take float32 or bfloat16 features, got {t.dtype}")


def _rows2d(t: Tensor, name: str) -> Tensor:
    require_cuda(t, name)
    if t.dim() != 2:
        raise ValueError(f"{name} must be 2-D [N, F], got {tuple(t.shape)}")
    if t.stride(1) != 1:
        t = t.contiguous()
    return t


def spmm(plan: CSRPlan, xs: Sequence[Tensor], ops: Sequence[int] = (0,), *, mean: bool = False,
         alpha: float = 1.0, beta: float = 0.0, zs: Optional[Sequence[Tensor]] = None,
         bias: Optional[Tensor] = None, out: Optional[Sequence[Tensor]] = None,
         variant: Optional[int] = None, op_scale: Optional[Sequence[float]] = None,
         grid_reserve: int = 0, tanh_out: bool = False) -> List[Tensor]:
    """y_k = alpha * (diag_k x_k[r] + sum val_k x_k[col]) (/len if mean) + beta * z_k + bias
    for the plan operators listed in `ops` (1 or 2 of them), one kernel launch.
    xs/zs/out may be column slices of wider row-major buffers.  grid_reserve = resident-CTA slots
    left free for a collective kernel running beside this launch (`shard_push`)."""
    global LAUNCHES
    n_ops = len(ops)
    assert n_ops in (1, 2) and len(xs) == n_ops
    xs = [_rows2d(x.detach(), "x") for x in xs]
    feat, dt = xs[0].size(1), _dtype_code(xs[0])
    dev = xs[0].device
    a = _lib.SpmmArgs()
    a.n_rows, a.feat, a.n_ops, a.dtype, a.mean = plan.n_dst, feat, n_ops, dt, int(mean) | (2 if tanh_out else 0)
    a.row_ptr, a.col = plan.row_ptr.data_ptr(), plan.col.data_ptr()
    a.alpha, a.beta = float(alpha), float(beta)
    a.variant = SPMM_VARIANT if variant is None else variant
    a.diag_row_offset = int(plan.meta.get("diag_row_offset", 0))
    a.grid_reserve = int(grid_reserve)
    if op_scale is not None:
        for k, sc in enumerate(op_scale):
            a.op_scale[k] = float(sc)
    hubs = plan.hub_rows()
    if tanh_out and hubs is not None:
        raise ValueError("spmm: the tanh epilogue cannot be combined with hub rows (their sums arrive by atomics)")
    if hubs is not None and dt == _lib.PGSD_F32:
        a.long_rows, a.long_chunk_ptr = hubs[0].data_ptr(), hubs[1].data_ptr()
        a.n_long_rows, a.long_row_threshold, a.long_chunk = hubs[0].numel(), HUB_ROW_THRESHOLD, HUB_ROW_CHUNK
    outs = []
    keep = []
    for k, op in enumerate(ops):
        x = xs[k]
        if x.size(1) != feat or x.dtype != xs[0].dtype:
            raise ValueError("spmm: operands must share feature width and dtype")
        v, d = plan.val[op], plan.diag[op]
        has_diag = d is not None or plan.diag_const[op] != 0.0
        if x.size(0) < plan.n_src or (has_diag and x.size(0) < plan.n_dst + a.diag_row_offset):
            raise ValueError(f"spmm: x has {x.size(0)} rows, plan gathers from {plan.n_src}"
                             + (f" and reads the diagonal at rows up to {plan.n_dst + a.diag_row_offset}"
                                if has_diag else ""))
        a.val[k] = None if v is None else v.data_ptr()
        a.diag[k] = None if d is None else d.data_ptr()
        a.diag_const[k] = plan.diag_const[op]
        a.x[k], a.ldx[k] = x.data_ptr(), x.stride(0)
        if zs is not None and zs[k] is not None:
            z = _rows2d(zs[k].detach(), "z")
            if z.dtype != x.dtype or z.size(0) < plan.n_dst or z.size(1) != feat:
                raise ValueError(f"spmm: z[{k}] must be [{plan.n_dst}+, {feat}] of {x.dtype}, got "
                                 f"{tuple(z.shape)} of {z.dtype}")
            keep.append(z)
            a.z[k], a.ldz[k] = z.data_ptr(), z.stride(0)
        if out is not None and out[k] is not None:
            y = out[k]
            if y.stride(1) != 1 or y.size(0) != plan.n_dst or y.size(1) != feat or y.dtype != x.dtype:
                raise ValueError("spmm: bad output buffer")
        else:
            y = torch.empty((plan.n_dst, feat), dtype=x.dtype, device=dev)
        a.y[k], a.ldy[k] = y.data_ptr(), y.stride(0)
        outs.append(y)
    if bias is not None:
        b = bias.detach().float().contiguous()
        if b.numel() < feat:
            raise ValueError(f"spmm: bias has {b.numel()} entries, the rows are {feat} wide")
        keep.append(b)
        a.bias = b.data_ptr()
    lib = _lib.load()
    with torch.cuda.device(dev), _Timed("spmm", dev):
        _lib.check(lib.pgsd_spmm_csr(C.byref(a), torch.cuda.current_stream(dev).cuda_stream),
                   "pgsd_spmm_csr")
    LAUNCHES += 1
    return outs


def dense(terms: Sequence[Tuple[Tensor, Tensor, int]], n_out: int, *, bias: Optional[Tensor] = None,
          combine: bool = False, relu_mode: int = 0,
          out: Optional[Sequence[Tensor]] = None, variant: Optional[int] = None) -> List[Tensor]:
    """terms: (X [N, k], W viewed as [k, n_out] (any strides), group).  combine=False:
    y0 = sum X W + b.  combine=True (MagNet): y0 = A - B + b, y1 = A + B + b with group 0 -> A,
    group 1 -> B; relu_mode=1 applies the complex ReLU mask in the epilogue."""
    global LAUNCHES
    if not 1 <= len(terms) <= _lib.DENSE_MAX_TERMS:
        raise ValueError(f"dense: between 1 and {_lib.DENSE_MAX_TERMS} terms supported")
    x0 = terms[0][0]
    require_cuda(x0, "x")
    dev, n = x0.device, x0.size(0)
    dt = _dtype_code(x0)
    a = _lib.DenseArgs()
    a.n_rows, a.n_out, a.n_terms, a.dtype, a.combine = n, n_out, len(terms), dt, int(combine)
    a.relu_mode = relu_mode
    a.variant = DENSE_VARIANT if variant is None else variant
    keep = []
    for t, (x, w, g) in enumerate(terms):
        x = _rows2d(x.detach(), "x")
        w = w.detach()
        if w.dtype != torch.float32:
            w = w.float()
        if x.dtype != x0.dtype or x.size(0) != n:
            raise ValueError("dense: terms must share dtype and row count")
        if w.dim() != 2 or w.size(0) != x.size(1) or w.size(1) != n_out:
            raise ValueError(f"dense: W must be [k={x.size(1)}, n_out={n_out}], got {tuple(w.shape)}")
        keep += [x, w]
        a.x[t], a.ldx[t], a.k[t], a.group[t] = x.data_ptr(), x.stride(0), x.size(1), int(g)
        a.w[t], a.ldw_k[t], a.ldw_n[t] = w.data_ptr(), w.stride(0), w.stride(1)
    if bias is not None:
        b = bias.detach().float().contiguous()
        if b.numel() < n_out:
            raise ValueError(f"dense: bias has {b.numel()} entries, n_out = {n_out}")
        keep.append(b)
        a.bias = b.data_ptr()
    n_outs = 2 if combine else 1
    outs = []
    for i in range(n_outs):
        if out is not None and out[i] is not None:
            y = out[i]
            if y.stride(1) != 1 or y.size(0) != n or y.size(1) != n_out or y.dtype != x0.dtype:
                raise ValueError("dense: bad output buffer")
        else:
            y = torch.empty((n, n_out), dtype=x0.dtype, device=dev)
        a.y[i], a.ldy[i] = y.data_ptr(), y.stride(0)
        outs.append(y)
    lib = _lib.load()
    with torch.cuda.device(dev), _Timed("dense", dev):
        _lib.check(lib.pgsd_dense_transform(C.byref(a), torch.cuda.current_stream(dev).cuda_stream),
                   "pgsd_dense_transform")
    LAUNCHES += 1
    return outs


def gather_rows(x: Tensor, index: Tensor, out: Optional[Tensor] = None) -> Tensor:
    """out[i] = x[index[i]] (halo pack); index int32."""
    global LAUNCHES
    x = _rows2d(x.detach(), "x")
    require_cuda(index, "index")
    idx = index if index.dtype == torch.int32 else index.int()
    idx = idx.contiguous()
    if out is None:
        out = torch.empty((idx.numel(), x.size(1)), dtype=x.dtype, device=x.device)
    lib = _lib.load()
    with torch.cuda.device(x.device):
        _lib.check(lib.pgsd_gather_rows(x.data_ptr(), x.stride(0), idx.data_ptr(), idx.numel(),
                                        x.size(1), _dtype_code(x), out.data_ptr(), out.stride(0),
                                        torch.cuda.current_stream(x.device).cuda_stream),
                   "pgsd_gather_rows")
    LAUNCHES += 1
    return out


def shard_push(srcs: Sequence[Tensor], dst_ptrs: Sequence[Sequence[int]], ld_dst_bytes: int, rank: int, world: int,
               slice_row: Sequence[int], flag_ptrs: Sequence[int], counters: Tensor, seq: int, *,
               n_ctas: int = 16, mc_ptrs: Optional[Sequence[int]] = None, include_self: bool = False,
               stream: Optional[torch.cuda.Stream] = None, engine: int = 0, chunk_bytes: int = 0,
               stages: int = 0, started_ptr: int = 0) -> None:
    """All-gather push over NVLink peer memory (`pgsd_shard_push`): every 16-byte piece of the local rows
    `srcs[t]` is read once and stored to dst_ptrs[t][p] (+ row * ld_dst_bytes) on every peer p; slice s =
    rows [slice_row[s], slice_row[s+1]) is published by writing `seq` to flag_ptrs[p][s] on every peer."""
    global LAUNCHES
    srcs = [_rows2d(x.detach(), "src") for x in srcs]
    dev = srcs[0].device
    a = _lib.PushArgs()
    a.world, a.rank, a.n_tensors = world, rank, len(srcs)
    a.row_bytes = srcs[0].size(1) * srcs[0].element_size()
    a.n_rows = srcs[0].size(0)
    for t, x in enumerate(srcs):
        if x.shape != srcs[0].shape or x.dtype != srcs[0].dtype:
            raise ValueError("shard_push: tensors must share shape and dtype")
        a.src[t], a.ld_src_bytes[t] = x.data_ptr(), x.stride(0) * x.element_size()
        a.ld_dst_bytes[t] = int(ld_dst_bytes)
        for p_ in range(world):
            a.dst[t][p_] = dst_ptrs[t][p_] or None
        if mc_ptrs is not None:
            a.mc_dst[t] = mc_ptrs[t]
    a.n_slices, a.n_ctas = len(slice_row) - 1, int(n_ctas)
    if a.n_slices > _lib.MAX_SLICES:
        raise ValueError(f"shard_push: at most {_lib.MAX_SLICES} slices")
    for s_, r in enumerate(slice_row):
        a.slice_row[s_] = int(r)
    for p_ in range(world):
        a.flag[p_] = flag_ptrs[p_] or None
    a.counters, a.seq, a.include_self = counters.data_ptr(), seq & 0xffffffff, int(include_self)
    a.engine, a.chunk_bytes, a.stages = int(engine), int(chunk_bytes), int(stages)
    a.started = started_ptr or None
    st = stream if stream is not None else torch.cuda.current_stream(dev)
    lib = _lib.load()
    with torch.cuda.device(dev):
        _lib.check(lib.pgsd_shard_push(C.byref(a), st.cuda_stream), "pgsd_shard_push")
    LAUNCHES += 1


def wait_flags(flags: Tensor, index: Sequence[int], seq: int, status: Optional[Tensor] = None,
               timeout_s: float = 5.0) -> None:
    """The current stream waits until flags[i] >= seq for every i in `index` (`pgsd_wait_flags`); after
    timeout_s the wait gives up and sets status[0] = 1 instead of hanging the GPU."""
    global LAUNCHES
    if not index:
        return
    dev = flags.device
    arr = (C.c_int32 * len(index))(*[int(i) for i in index])
    lib = _lib.load()
    with torch.cuda.device(dev):
        _lib.check(lib.pgsd_wait_flags(flags.data_ptr(), arr, len(index), seq & 0xffffffff, int(timeout_s * 1e9),
                                       None if status is None else status.data_ptr(),
                                       torch.cuda.current_stream(dev).cuda_stream), "pgsd_wait_flags")
    LAUNCHES += 1


def edge_softmax(plans: Sequence[CSRPlan], s_src: Sequence[Tensor], s_dst: Sequence[Tensor], *,
                 act: str = "tanh", slope: float = 0.2, xd: Optional[Sequence[Tensor]] = None,
                 want_alpha: bool = False):
    """Segment softmax over the rows of 1 or 2 CSR plans (`pgsd_edge_softmax`).
    xd given   -> returns y = xd[0] * sum_type0(alpha) + xd[1] * sum_type1(alpha)   (SNEAConv)
    want_alpha -> returns the per-entry alpha tensors (one per plan)              (GAT-style)."""
    global LAUNCHES
    n_types = len(plans)
    assert n_types in (1, 2) and len(s_src) == n_types and len(s_dst) == n_types
    dev = plans[0].device
    a = _lib.AttnArgs()
    a.n_rows, a.n_types = plans[0].n_dst, n_types
    a.act, a.slope = (0 if act == "tanh" else 1), float(slope)
    keep, alphas, y = [], [], None
    for t in range(n_types):
        ss = s_src[t].detach().float().contiguous().view(-1)
        sd = s_dst[t].detach().float().contiguous().view(-1)
        if ss.numel() < plans[t].n_src or sd.numel() < plans[t].n_dst:
            raise ValueError("edge_softmax: score vectors shorter than the plan's node ranges")
        keep += [ss, sd]
        a.row_ptr[t], a.col[t] = plans[t].row_ptr.data_ptr(), plans[t].col.data_ptr()
        a.s_src[t], a.s_dst[t] = ss.data_ptr(), sd.data_ptr()
        if want_alpha:
            al = torch.empty(max(plans[t].nnz, 1), dtype=torch.float32, device=dev)
            alphas.append(al[:plans[t].nnz])
            a.alpha_out[t] = al.data_ptr()
    if xd is not None:
        xs = [_rows2d(x.detach(), "xd") for x in xd]
        if any(x.dtype != torch.float32 for x in xs):
            raise TypeError("edge_softmax takes float32 features")
        a.feat = xs[0].size(1)
        y = torch.empty((plans[0].n_dst, a.feat), dtype=torch.float32, device=dev)
        for t in range(n_types):
            a.xd[t], a.ldxd[t] = xs[t].data_ptr(), xs[t].stride(0)
        a.y, a.ldy = y.data_ptr(), y.stride(0)
        keep += xs
    lib = _lib.load()
    with torch.cuda.device(dev), _Timed("edge_softmax", dev):
        _lib.check(lib.pgsd_edge_softmax(C.byref(a), torch.cuda.current_stream(dev).cuda_stream),
                   "pgsd_edge_softmax")
    LAUNCHES += 1
    return y, alphas


def edge_softmax_backward(plans: Sequence[CSRPlan], s_src: Sequence[Tensor], s_dst: Sequence[Tensor], *,
                          act: str = "tanh", slope: float = 0.2, dalpha: Optional[Sequence[Tensor]] = None,
                          row_coef: Optional[Sequence[Tensor]] = None, want_type_sum: bool = False):
    """Backward of the segment softmax (`pgsd_edge_softmax_backward`): dL/dalpha per entry (`dalpha[t]`, GAT-style)
    or per row and type (`row_coef[t]`, SNEAConv) -> (g_s_src [n_src] per type, g_s_dst [n_dst] per type,
    per-row sums of alpha per type or None)."""
    global LAUNCHES
    n_types = len(plans)
    assert n_types in (1, 2) and (dalpha is None) != (row_coef is None)
    dev = plans[0].device
    a = _lib.AttnBwdArgs()
    a.n_rows, a.n_types = plans[0].n_dst, n_types
    a.act, a.slope = (0 if act == "tanh" else 1), float(slope)
    keep, g_src, g_dst, sums = [], [], [], []
    for t in range(n_types):
        ss = s_src[t].detach().float().contiguous().view(-1)
        sd = s_dst[t].detach().float().contiguous().view(-1)
        if ss.numel() < plans[t].n_src or sd.numel() < plans[t].n_dst:
            raise ValueError("edge_softmax_backward: score vectors shorter than the plan's node ranges")
        gs = torch.zeros(ss.numel(), dtype=torch.float32, device=dev)
        gd = torch.zeros(sd.numel(), dtype=torch.float32, device=dev)
        keep += [ss, sd]
        a.row_ptr[t], a.col[t] = plans[t].row_ptr.data_ptr(), plans[t].col.data_ptr()
        a.s_src[t], a.s_dst[t] = ss.data_ptr(), sd.data_ptr()
        a.g_s_src[t], a.g_s_dst[t] = gs.data_ptr(), gd.data_ptr()
        if dalpha is not None:
            da = dalpha[t].detach().float().contiguous().view(-1)
            if da.numel() < plans[t].nnz:
                raise ValueError("edge_softmax_backward: dalpha shorter than the plan's entry list")
            keep.append(da)
            a.dalpha[t] = da.data_ptr() if da.numel() else None
            if not da.numel():                      # a plan without entries: nothing to read, coefficient 0
                a.row_coef[t] = gd.data_ptr()
        else:
            rc = row_coef[t].detach().float().contiguous().view(-1)
            keep.append(rc)
            a.row_coef[t] = rc.data_ptr()
        if want_type_sum:
            sm = torch.zeros(plans[t].n_dst, dtype=torch.float32, device=dev)
            a.type_sum[t] = sm.data_ptr()
            sums.append(sm)
        g_src.append(gs)
        g_dst.append(gd)
    lib = _lib.load()
    with torch.cuda.device(dev), _Timed("edge_softmax_bwd", dev):
        _lib.check(lib.pgsd_edge_softmax_backward(C.byref(a), torch.cuda.current_stream(dev).cuda_stream),
                   "pgsd_edge_softmax_backward")
    LAUNCHES += 1
    return g_src, g_dst, (sums if want_type_sum else None)


def sddmm_rows(plan: CSRPlan, gy: Tensor, h: Tensor) -> Tensor:
    """out[k] = <gy[row(k)], h[col[k]]> for every stored entry (`pgsd_sddmm_rows`); fp32.  Shapes the kernel does
    not take (feature width not a multiple of 4, unaligned rows) are evaluated with torch gathers."""
    global LAUNCHES
    gy, h = _rows2d(gy.detach(), "gy"), _rows2d(h.detach(), "h")
    f = gy.size(1)
    out = torch.empty(max(plan.nnz, 1), dtype=torch.float32, device=gy.device)[:plan.nnz]
    if plan.nnz == 0:
        return out
    ok = lambda t: t.dtype == torch.float32 and t.data_ptr() % 16 == 0 and t.stride(0) % 4 == 0
    if f % 4 or f > 256 or not (ok(gy) and ok(h)):
        counts = (plan.row_ptr[1:] - plan.row_ptr[:-1]).long()
        rows = torch.repeat_interleave(torch.arange(plan.n_dst, device=gy.device), counts)
        return (gy.float()[rows] * h.float()[plan.col.long()]).sum(1)
    lib = _lib.load()
    with torch.cuda.device(gy.device), _Timed("sddmm", gy.device):
        _lib.check(lib.pgsd_sddmm_rows(plan.row_ptr.data_ptr(), plan.col.data_ptr(), gy.data_ptr(), gy.stride(0),
                                       h.data_ptr(), h.stride(0), plan.n_dst, f, out.data_ptr(),
                                       torch.cuda.current_stream(gy.device).cuda_stream), "pgsd_sddmm_rows")
    LAUNCHES += 1
    return out


# 1: GATConv aggregates with the softmax inside one kernel (pgsd_gat_aggregate); 0 (default): edge_softmax +
# spmm.  The one-kernel route is parity-green but not faster yet: 1.46 ms per 22M-entry launch without software
# pipelining (= the two launches it replaces, 0.39 + 1.08 ms), 1.78 ms with the next batch's column / score
# prefetch (the dependent score gather costs more than it hides) -- profiles/r01_configs_attention_s39.jsonl.
GAT_FUSED = int(os.environ.get("PGSD_GAT_FUSED", "0"))


def gat_aggregate_supported(h: Tensor, out: Optional[Tensor] = None, z: Optional[Tensor] = None) -> bool:
    ok = lambda t: t is None or (t.dtype == torch.float32 and t.stride(1) == 1 and t.data_ptr() % 16 == 0
                                 and t.stride(0) % 4 == 0)
    return bool(GAT_FUSED) and h.dim() == 2 and h.size(1) % 4 == 0 and h.size(1) <= 128 and ok(h) and ok(out) and ok(z)


def gat_aggregate(plan: CSRPlan, s_src: Tensor, s_dst: Tensor, h: Tensor, *, slope: float = 0.2,
                  bias: Optional[Tensor] = None, z: Optional[Tensor] = None, beta: float = 1.0,
                  out: Optional[Tensor] = None) -> Tensor:
    """y = softmax-weighted sum of h over the incoming entries (+ bias) (+ beta * z)  (`pgsd_gat_aggregate`)."""
    global LAUNCHES
    require_cuda(h, "h")
    dev, n, f = h.device, plan.n_dst, h.size(1)
    ss = s_src.detach().float().contiguous().view(-1)
    sd = s_dst.detach().float().contiguous().view(-1)
    if ss.numel() < plan.n_src or sd.numel() < plan.n_dst:
        raise ValueError("gat_aggregate: score vectors shorter than the plan's node ranges")
    y = out if out is not None else torch.empty((n, f), dtype=torch.float32, device=dev)
    b = None if bias is None else bias.detach().float().contiguous()
    lib = _lib.load()
    with torch.cuda.device(dev), _Timed("gat_aggregate", dev):
        _lib.check(lib.pgsd_gat_aggregate(plan.row_ptr.data_ptr(), plan.col.data_ptr(), ss.data_ptr(), sd.data_ptr(),
                                          float(slope), h.data_ptr(), h.stride(0), f, n,
                                          None if b is None else b.data_ptr(),
                                          None if z is None else z.data_ptr(), 0 if z is None else z.stride(0),
                                          float(beta), y.data_ptr(), y.stride(0),
                                          torch.cuda.current_stream(dev).cuda_stream), "pgsd_gat_aggregate")
    LAUNCHES += 1
    return y


def xtg_accumulate(x: Tensor, g: Tensor, dw: Tensor, db: Optional[Tensor] = None) -> None:
    """dw[k, n] += sum_r x[r, k] g[r, n]; db[n] += sum_r g[r, n]  (`pgsd_xtg_accumulate`)."""
    global LAUNCHES
    x, g = _rows2d(x.detach(), "x"), _rows2d(g.detach(), "g")
    if x.dtype != g.dtype or x.size(0) != g.size(0):
        raise ValueError("xtg: x and g must share dtype and row count")
    if dw.dtype != torch.float32 or not dw.is_contiguous() or tuple(dw.shape) != (x.size(1), g.size(1)):
        raise ValueError("xtg: dw must be a contiguous fp32 [k, n] tensor")
    lib = _lib.load()
    with torch.cuda.device(x.device), _Timed("xtg", x.device):
        _lib.check(lib.pgsd_xtg_accumulate(x.data_ptr(), x.stride(0), g.data_ptr(), g.stride(0), x.size(0),
                                           x.size(1), g.size(1), _dtype_code(x), dw.data_ptr(), dw.stride(0),
                                           None if db is None else db.data_ptr(),
                                           torch.cuda.current_stream(x.device).cuda_stream),
                   "pgsd_xtg_accumulate")
    LAUNCHES += 1


def magnetic_q_grad(plan: CSRPlan, gys: Sequence[Tensor], xs: Sequence[Tensor], alpha: float = 1.0) -> Tensor:
    """dL/dq through one two-operator aggregation of a magnetic plan built with `keep_theta`
    (`pgsd_magnetic_q_grad`); returns a float64 scalar tensor."""
    global LAUNCHES
    theta = plan.meta.get("theta")
    if theta is None:
        raise _lib.PgsdError("magnetic_q_grad: the plan was built without theta (keep_theta=False)")
    gr, gi = (_rows2d(g.detach(), "gy") for g in gys)
    xr, xi = (_rows2d(x.detach(), "x") for x in xs)
    for t in (gr, gi, xr, xi):
        if t.dtype != torch.float32:
            raise TypeError("magnetic_q_grad: fp32 features only")
    if not (gr.size(0) == gi.size(0) == plan.n_dst and xr.size(0) == xi.size(0) == plan.n_src
            and gr.size(1) == gi.size(1) == xr.size(1) == xi.size(1)):
        raise ValueError("magnetic_q_grad: shape mismatch")
    dq = torch.zeros((), dtype=torch.float64, device=gr.device)
    if plan.nnz == 0:
        return dq
    lib = _lib.load()
    with torch.cuda.device(gr.device), _Timed("q_grad", gr.device):
        _lib.check(lib.pgsd_magnetic_q_grad(
            plan.row_ptr.data_ptr(), plan.col.data_ptr(), plan.val[0].data_ptr(), plan.val[1].data_ptr(),
            theta.data_ptr(), plan.n_dst, gr.size(1), gr.data_ptr(), gr.stride(0), gi.data_ptr(), gi.stride(0),
            xr.data_ptr(), xr.stride(0), xi.data_ptr(), xi.stride(0), 2.0 * math.pi * float(alpha),
            dq.data_ptr(), torch.cuda.current_stream(gr.device).cuda_stream), "pgsd_magnetic_q_grad")
    LAUNCHES += 1
    return dq


# 1: MagNetConv / MSConv layers of order K = 1 inside the fused kernel's envelope run as ONE launch
# (pgsd_magnet_layer_fused); 0: aggregation launch + transform launch
FUSED_LAYER = int(os.environ.get("PGSD_FUSED_LAYER", "0"))
FUSED_VARIANT = int(os.environ.get("PGSD_FUSED_VARIANT", "0"))


def magnet_fused_supported(plan: CSRPlan, x_real: Tensor, x_imag: Tensor, weight: Tensor) -> bool:
    """True when `magnet_layer_fused` can run this layer: fp32, K = 1, the kernel's feature widths,
    square plan with two value arrays and no hub rows."""
    if weight.dim() != 3 or weight.size(0) != 2 or x_real.dtype != torch.float32 or x_imag.dtype != torch.float32:
        return False
    if not (x_real.is_cuda and x_real.dim() == 2 and x_real.shape == x_imag.shape):
        return False
    if len(plan.val) != 2 or plan.val[0] is None or plan.val[1] is None or plan.n_dst != plan.n_src:
        return False
    if plan.meta.get("diag_row_offset", 0) != 0 or plan.hub_rows() is not None:
        return False
    if x_real.size(1) != weight.size(1) or x_real.size(0) != plan.n_dst:
        return False
    return bool(_lib.load().pgsd_magnet_fused_supported(weight.size(1), weight.size(2), _lib.PGSD_F32))


def magnet_layer_fused(plan: CSRPlan, x_real: Tensor, x_imag: Tensor, weight: Tensor,
                       bias: Optional[Tensor] = None, relu_mode: int = 0,
                       variant: Optional[int] = None) -> Tuple[Tensor, Tensor]:
    """(out_real, out_imag) of a K = 1 MagNetConv layer in one launch (`pgsd_magnet_layer_fused`):
    A = x_r W0 + (L_r x_r) W1, B = x_i W0 + (L_i x_i) W1, out_real = A - B + b, out_imag = A + B + b."""
    global LAUNCHES
    if not magnet_fused_supported(plan, x_real, x_imag, weight):
        raise _lib.PgsdError("magnet_layer_fused: layer is outside the fused kernel's envelope")
    xr, xi = _rows2d(x_real.detach(), "x_real"), _rows2d(x_imag.detach(), "x_imag")
    w = weight.detach()
    dev, n, f_out = xr.device, plan.n_dst, w.size(2)
    a = _lib.MagnetFusedArgs()
    a.n_rows, a.feat_in, a.feat_out = n, w.size(1), f_out
    a.row_ptr, a.col = plan.row_ptr.data_ptr(), plan.col.data_ptr()
    for k in range(2):
        a.val[k] = plan.val[k].data_ptr()
        a.diag[k] = None if plan.diag[k] is None else plan.diag[k].data_ptr()
        a.diag_const[k] = plan.diag_const[k]
        a.w[k], a.ldw_k[k], a.ldw_n[k] = w[k].data_ptr(), w.stride(1), w.stride(2)
    a.x[0], a.ldx[0], a.x[1], a.ldx[1] = xr.data_ptr(), xr.stride(0), xi.data_ptr(), xi.stride(0)
    keep = [xr, xi, w]
    if bias is not None:
        b = bias.detach().float().contiguous()
        keep.append(b)
        a.bias = b.data_ptr()
    outs = [torch.empty((n, f_out), dtype=torch.float32, device=dev) for _ in range(2)]
    for k in range(2):
        a.y[k], a.ldy[k] = outs[k].data_ptr(), outs[k].stride(0)
    a.relu_mode = relu_mode
    a.variant = FUSED_VARIANT if variant is None else variant
    lib = _lib.load()
    with torch.cuda.device(dev), _Timed("magnet_fused", dev):
        _lib.check(lib.pgsd_magnet_layer_fused(C.byref(a), torch.cuda.current_stream(dev).cuda_stream),
                   "pgsd_magnet_layer_fused")
    LAUNCHES += 1
    return outs[0], outs[1]

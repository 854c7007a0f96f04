"""Graph plans: the reference's COO `(edge_index, edge_weight)` turned into the int32
CSR-by-destination layout the aggregation kernels consume, built on the GPU through the C ABI
(`pgsd_build_*`, include/pgsd_b200.h) and cached by the layers under the reference's own cache
rules (SURVEY §5 "In-layer caching").

HBM layout of a plan (all int32 / fp32, contiguous):
    row_ptr [n_dst + 1]   destination-row offsets
    col     [nnz]         source node of every stored entry
    val[k]  [nnz]         one value array per operator sharing the pattern (MagNet: real, imag)
    diag[k] [n_dst]       optional per-row diagonal (kept out of the entry list: the reference's
                          3N explicit self-loop entries collapse to this)
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass, field
from typing import List, Optional, Tuple

import torch
from torch import Tensor

from . import _lib


def _stream_ptr(device) -> int:
    return torch.cuda.current_stream(device).cuda_stream


def require_cuda(t: Tensor, name: str) -> None:
    if not t.is_cuda:
        raise _lib.PgsdError(
            f"{name} must be a CUDA tensor: pytorch_geometric_signed_directed_b200 has no CPU path")


@dataclass
class CSRPlan:
    n_dst: int
    n_src: int
    nnz: int
    num_input_edges: int
    row_ptr: Tensor
    col: Tensor
    val: List[Optional[Tensor]] = field(default_factory=list)
    diag: List[Optional[Tensor]] = field(default_factory=list)
    diag_const: List[float] = field(default_factory=list)
    meta: dict = field(default_factory=dict)

    @property
    def device(self):
        return self.row_ptr.device

    def hub_rows(self):
        """(hub row ids int32, exclusive prefix of their 1024-entry slices int32) or None.
        Computed once per plan (one device->host sync) and cached."""
        cached = getattr(self, "_hubs", 0)
        if cached != 0:
            return cached
        from . import ops
        res = None
        if self.nnz > ops.HUB_ROW_THRESHOLD and self.row_ptr.is_cuda:
            lens = self.row_ptr[1:] - self.row_ptr[:-1]
            if int(lens.max().item()) > ops.HUB_ROW_THRESHOLD:
                rows = torch.nonzero(lens > ops.HUB_ROW_THRESHOLD).view(-1)
                slices = (lens[rows].long() + ops.HUB_ROW_CHUNK - 1) // ops.HUB_ROW_CHUNK
                ptr = torch.zeros(rows.numel() + 1, dtype=torch.int32, device=self.row_ptr.device)
                ptr[1:] = torch.cumsum(slices, 0).int()
                res = (rows.int().contiguous(), ptr)
        self._hubs = res
        return res

    def bytes(self) -> int:
        tot = self.row_ptr.numel() * 4 + self.nnz * 4
        tot += sum(self.nnz * 4 for v in self.val if v is not None)
        tot += sum(d.numel() * 4 for d in self.diag if d is not None)
        return tot


def _prep_edges(edge_index: Tensor, edge_weight: Optional[Tensor]) -> Tuple[Tensor, Optional[Tensor]]:
    require_cuda(edge_index, "edge_index")
    if edge_index.dim() != 2 or edge_index.size(0) != 2:
        raise ValueError(f"edge_index must have shape [2, E], got {tuple(edge_index.shape)}")
    ei = edge_index
    if ei.dtype != torch.int64:
        ei = ei.long()
    if not ei.is_contiguous():
        ei = ei.contiguous()
    ew = edge_weight
    if ew is not None:
        require_cuda(ew, "edge_weight")
        if ew.numel() != ei.size(1):
            raise ValueError("edge_weight must have one entry per edge")
        ew = ew.detach()
        if ew.dtype != torch.float32:
            ew = ew.float()
        ew = ew.contiguous().view(-1)
    return ei, ew


def as_edge_index(adj) -> Tensor:
    """COO `edge_index` [2, E] (row 0 = source, row 1 = target) of an adjacency given the way PyG's `Adj` allows
    (nn/signed/SGCNConv.py:94-95,131-134): a dense [2, E] tensor is returned as is; a torch sparse tensor (COO / CSR
    layout) or a torch_sparse-style `SparseTensor` (anything with `.coo() -> (row, col, value)`) is read as the
    TRANSPOSED adjacency `adj_t[target, source]` that `message_and_aggregate` multiplies with, entries in stored order."""
    if isinstance(adj, Tensor):
        if adj.layout == torch.strided:
            return adj
        coo = adj if adj.layout == torch.sparse_coo else adj.to_sparse_coo()
        idx = coo._indices()
        if idx.size(0) != 2:
            raise ValueError("sparse adjacency must be 2-D")
        return torch.stack([idx[1], idx[0]]).contiguous()
    coo = getattr(adj, "coo", None)
    if callable(coo):
        row, col = adj.coo()[:2]
        return torch.stack([col, row]).contiguous()
    raise NotImplementedError(f"unsupported adjacency type {type(adj).__name__}: pass a [2, E] edge_index, a torch "
                              "sparse tensor or a SparseTensor with .coo()")


def _workspace(n: int, e: int, device) -> Tensor:
    lib = _lib.load()
    nbytes = C.c_size_t(0)
    _lib.check(lib.pgsd_plan_workspace_bytes(n, e, C.byref(nbytes)), "pgsd_plan_workspace_bytes")
    return torch.empty(nbytes.value, dtype=torch.uint8, device=device)


def _ptr(t: Optional[Tensor]) -> Optional[int]:
    return None if t is None else t.data_ptr()


def build_csr(edge_index: Tensor, edge_weight: Optional[Tensor], n_dst: int, n_src: int,
              flow: str = "source_to_target") -> CSRPlan:
    """Generic aggregation plan (`pgsd_build_csr`).  flow as in PyG: source_to_target gathers
    edge_index[0] and reduces at edge_index[1]."""
    ei, ew = _prep_edges(edge_index, edge_weight)
    dev, e = ei.device, ei.size(1)
    src, dst = (ei[0], ei[1]) if flow == "source_to_target" else (ei[1], ei[0])
    with torch.cuda.device(dev):
        row_ptr = torch.empty(n_dst + 1, dtype=torch.int32, device=dev)
        col = torch.empty(max(e, 1), dtype=torch.int32, device=dev)
        val = torch.empty(max(e, 1), dtype=torch.float32, device=dev) if ew is not None else None
        ws = _workspace(max(n_dst, n_src), e, dev)
        lib = _lib.load()
        _lib.check(lib.pgsd_build_csr(src.data_ptr(), dst.data_ptr(), _ptr(ew), e, n_dst, n_src,
                                      row_ptr.data_ptr(), col.data_ptr(), _ptr(val),
                                      ws.data_ptr(), ws.numel(), _stream_ptr(dev)), "pgsd_build_csr")
    return CSRPlan(n_dst, n_src, e, e, row_ptr, col[:e], [None if val is None else val[:e]],
                   [None], [0.0])


def build_rw_norm(edge_index: Tensor, edge_weight: Optional[Tensor], n: int, fill_value: float,
                  transpose: bool = False, add_self_loops: bool = True) -> CSRPlan:
    """conv_norm_rw + target_to_source aggregation plan (`pgsd_build_csr_rw_norm`).
    transpose=True is DIMPA's `edge_index[[1, 0]]` call without materialising the flip."""
    ei, ew = _prep_edges(edge_index, edge_weight)
    dev, e = ei.device, ei.size(1)
    dst, src = (ei[1], ei[0]) if transpose else (ei[0], ei[1])
    with torch.cuda.device(dev):
        row_ptr = torch.empty(n + 1, dtype=torch.int32, device=dev)
        col = torch.empty(max(e, 1), dtype=torch.int32, device=dev)
        val = torch.empty(max(e, 1), dtype=torch.float32, device=dev)
        diag = torch.empty(max(n, 1), dtype=torch.float32, device=dev)
        ws = _workspace(n, e, dev)
        nnz = C.c_int64(0)
        lib = _lib.load()
        _lib.check(lib.pgsd_build_csr_rw_norm(dst.data_ptr(), src.data_ptr(), _ptr(ew), e, n,
                                              float(fill_value), int(add_self_loops),
                                              row_ptr.data_ptr(), col.data_ptr(), val.data_ptr(),
                                              diag.data_ptr(), C.byref(nnz), ws.data_ptr(),
                                              ws.numel(), _stream_ptr(dev)),
                   "pgsd_build_csr_rw_norm")
    k = nnz.value
    return CSRPlan(n, n, k, e, row_ptr, col[:k], [val[:k]], [diag[:n]], [0.0])


def build_sym_norm(edge_index: Tensor, edge_weight: Optional[Tensor], n: int, fill_value: float,
                   add_self_loops: bool = True) -> CSRPlan:
    """gcn_norm + source_to_target aggregation plan (`pgsd_build_csr_sym_norm`): dst = edge_index[1]."""
    ei, ew = _prep_edges(edge_index, edge_weight)
    dev, e = ei.device, ei.size(1)
    with torch.cuda.device(dev):
        row_ptr = torch.empty(n + 1, dtype=torch.int32, device=dev)
        col = torch.empty(max(e, 1), dtype=torch.int32, device=dev)
        val = torch.empty(max(e, 1), dtype=torch.float32, device=dev)
        diag = torch.empty(max(n, 1), dtype=torch.float32, device=dev)
        ws = _workspace(n, e, dev)
        nnz = C.c_int64(0)
        lib = _lib.load()
        _lib.check(lib.pgsd_build_csr_sym_norm(ei[1].data_ptr(), ei[0].data_ptr(), _ptr(ew), e, n,
                                               float(fill_value), int(add_self_loops), row_ptr.data_ptr(),
                                               col.data_ptr(), val.data_ptr(), diag.data_ptr(), C.byref(nnz),
                                               ws.data_ptr(), ws.numel(), _stream_ptr(dev)),
                   "pgsd_build_csr_sym_norm")
    k = nnz.value
    return CSRPlan(n, n, k, e, row_ptr, col[:k], [val[:k]], [diag[:n]], [0.0])


def build_magnetic(edge_index: Tensor, edge_weight: Optional[Tensor], n: int, q: float,
                   normalization: Optional[str], lambda_max: float,
                   signed_mode: int = 0, keep_theta: bool = False) -> CSRPlan:
    """Scaled magnetic (signed) Laplacian plan (`pgsd_build_magnetic_laplacian`):
    val[0]/val[1] = real/imag off-diagonals of L~ = 2L/lambda_max - I stored for
    source_to_target aggregation, diag[0] = its real diagonal (imag diagonal is 0)."""
    ei, ew = _prep_edges(edge_index, edge_weight)
    dev, e = ei.device, ei.size(1)
    cap = max(2 * e, 1)
    with torch.cuda.device(dev):
        row_ptr = torch.empty(n + 1, dtype=torch.int32, device=dev)
        col = torch.empty(cap, dtype=torch.int32, device=dev)
        vr = torch.empty(cap, dtype=torch.float32, device=dev)
        vi = torch.empty(cap, dtype=torch.float32, device=dev)
        diag = torch.empty(max(n, 1), dtype=torch.float32, device=dev)
        ws = _workspace(n, e, dev)
        nnz = C.c_int64(0)
        lib = _lib.load()
        head = (ei[0].data_ptr(), ei[1].data_ptr(), _ptr(ew), e, n, float(q),
                1 if normalization == "sym" else 0, float(lambda_max), int(signed_mode),
                row_ptr.data_ptr(), col.data_ptr(), vr.data_ptr(), vi.data_ptr(), diag.data_ptr())
        tail = (C.byref(nnz), ws.data_ptr(), ws.numel(), _stream_ptr(dev))
        theta = None
        if keep_theta:   # trainable q: the phase weights stay with the plan for pgsd_magnetic_q_grad
            theta = torch.empty(cap, dtype=torch.float32, device=dev)
            _lib.check(lib.pgsd_build_magnetic_laplacian_theta(*head, theta.data_ptr(), *tail),
                       "pgsd_build_magnetic_laplacian_theta")
        else:
            _lib.check(lib.pgsd_build_magnetic_laplacian(*head, *tail), "pgsd_build_magnetic_laplacian")
    k = nnz.value
    meta = {"q": q, "normalization": normalization, "lambda_max": float(lambda_max), "diag_real": diag[:n],
            "hermitian": True}   # real part symmetric, imaginary part antisymmetric: M^T = (M_r, -M_i)
    if theta is not None:
        meta["theta"] = theta[:k]
    if normalization == "sym":
        # diag(L) = 1 for every node, so the real diagonal of L~ is the constant 2/lambda_max - 1
        # (exactly 0 for the default lambda_max = 2: the reference's 2N cancelling self-loop
        # entries).  A constant lets the kernel skip the x[row] read when it is 0.
        import numpy as np
        dc = float(np.float32(np.float32(2.0) / np.float32(lambda_max)) - np.float32(1.0))
        return CSRPlan(n, n, k, e, row_ptr, col[:k], [vr[:k], vi[:k]], [None, None], [dc, 0.0], meta=meta)
    return CSRPlan(n, n, k, e, row_ptr, col[:k], [vr[:k], vi[:k]], [diag[:n], None], [0.0, 0.0], meta=meta)


def build_magnetic_rows(edge_index: Tensor, edge_weight: Optional[Tensor], n: int, row_lo: int, row_hi: int,
                        q: float, normalization: Optional[str], lambda_max: float, signed_mode: int = 0,
                        allgather_deg=None) -> CSRPlan:
    """Rows [row_lo, row_hi) of the magnetic plan (columns global), built from the edges incident to
    that node range only (`pgsd_build_magnetic_rows_begin` / `_finish`).
    `allgather_deg(deg_local [row_hi-row_lo]) -> deg_all [n]` supplies the one exchange the build
    needs (every node's degree, 4 B/node); None = the range is the whole graph.
    The result is bit-identical to the same rows of `build_magnetic`."""
    ei, ew = _prep_edges(edge_index, edge_weight)
    dev, e, nl = ei.device, ei.size(1), row_hi - row_lo
    cap = max(2 * e, 1)
    with torch.cuda.device(dev):
        row_ptr = torch.empty(nl + 1, dtype=torch.int32, device=dev)
        col = torch.empty(cap, dtype=torch.int32, device=dev)
        vr = torch.empty(cap, dtype=torch.float32, device=dev)
        vi = torch.empty(cap, dtype=torch.float32, device=dev)
        deg = torch.zeros(max(nl, 1), dtype=torch.float32, device=dev)
        diag = torch.empty(max(nl, 1), dtype=torch.float32, device=dev)
        ws = _workspace(n, e, dev)
        nnz = C.c_int64(0)
        lib = _lib.load()
        _lib.check(lib.pgsd_build_magnetic_rows_begin(
            ei[0].data_ptr(), ei[1].data_ptr(), _ptr(ew), e, n, row_lo, row_hi, int(signed_mode),
            row_ptr.data_ptr(), col.data_ptr(), vr.data_ptr(), vi.data_ptr(), deg.data_ptr(), C.byref(nnz),
            ws.data_ptr(), ws.numel(), _stream_ptr(dev)), "pgsd_build_magnetic_rows_begin")
        del ws
        if allgather_deg is None:
            if nl != n:
                raise ValueError("build_magnetic_rows: a proper row range needs allgather_deg")
            deg_all = deg[:n]
        else:
            deg_all = allgather_deg(deg[:nl])
        if deg_all.numel() != n or deg_all.dtype != torch.float32:
            raise ValueError("build_magnetic_rows: allgather_deg must return a float32 [n] tensor")
        deg_all = deg_all.contiguous()
        _lib.check(lib.pgsd_build_magnetic_rows_finish(
            row_ptr.data_ptr(), col.data_ptr(), vr.data_ptr(), vi.data_ptr(), deg_all.data_ptr(), n, row_lo,
            row_hi, float(q), 1 if normalization == "sym" else 0, float(lambda_max), vr.data_ptr(),
            vi.data_ptr(), diag.data_ptr(), _stream_ptr(dev)), "pgsd_build_magnetic_rows_finish")
    k = nnz.value
    # compact the 2E-capacity arrays (a shard keeps ~1/world of them)
    col, vr, vi = col[:k].clone(), vr[:k].clone(), vi[:k].clone()
    meta = {"q": q, "normalization": normalization, "lambda_max": float(lambda_max), "diag_real": diag[:nl],
            "row_lo": row_lo}
    if normalization == "sym":
        import numpy as np
        dc = float(np.float32(np.float32(2.0) / np.float32(lambda_max)) - np.float32(1.0))
        return CSRPlan(nl, n, k, e, row_ptr, col, [vr, vi], [None, None], [dc, 0.0], meta=meta)
    return CSRPlan(nl, n, k, e, row_ptr, col, [vr, vi], [diag[:nl], None], [0.0, 0.0], meta=meta)


def magnetic_cached_result(plan: CSRPlan):
    """Re-materialise MagNetConv.cached_result (nn/directed/MagNetConv.py:181) from a plan:
    (edge_index_real [2, nnz+2N], edge_index_imag [2, nnz+N], norm_real, norm_imag).
    Index tensors are bit-identical to the reference's (sorted (row, col) block, then the
    appended loop blocks); values use L~_r[a,b] = L~_r[b,a], L~_i[a,b] = -L~_i[b,a]."""
    n, dev = plan.n_dst, plan.device
    counts = (plan.row_ptr[1:] - plan.row_ptr[:-1]).long()
    rows = torch.repeat_interleave(torch.arange(n, device=dev), counts)
    cols = plan.col.long()
    loops = torch.arange(n, device=dev)
    loops2 = torch.stack([loops, loops])
    block = torch.stack([rows, cols])
    ei_imag = torch.cat([block, loops2], 1)
    ei_real = torch.cat([block, loops2, loops2], 1)
    norm_real = torch.cat([plan.val[0], plan.meta["diag_real"] + 1.0, torch.full((n,), -1.0, device=dev)])
    norm_imag = torch.cat([-plan.val[1], torch.zeros(n, device=dev)])
    return ei_real, ei_imag, norm_real, norm_imag


class PlanCache:
    """Small cache for layers the reference re-normalises on every call (MagNetConv / MSConv with cached=False,
    Conv_Base, conv_base.py:102-108; SGCNConv's / SNEAConv's index plumbing).  A plan is reused only while the very
    same edge tensors come back: identity key (storage pointer, shape, strides, in-place version counter, device,
    dtype) AND an on-device content fingerprint (`pgsd_fingerprint`: plain and position-weighted sums of the words,
    one pass), so writes
    that do not bump the version counter (`.data`, custom kernels, DLPack consumers, collective receives into the
    same buffer) are seen too.  The fingerprint costs one pass over the edge tensors and one device->host read per
    call (1M nodes / 20M edges: < 0.2 ms against a 4 ms rebuild).  PGSD_PLAN_REUSE = "fingerprint" (default) |
    "identity" (key only, no read-back) | "off" (rebuild every call, exactly what the reference does).
    Retained memory: up to `capacity` plans plus the edge tensors they were built from stay alive (20M edges:
    ~0.5 GB per plan + 0.32 GB of int64 edges); `clear()` -- also called by `reset_parameters()` and by assigning
    `cached_result = None` -- releases them."""

    def __init__(self, capacity: int = 8):
        self.capacity = capacity
        self._items: dict = {}

    @staticmethod
    def _key(t: Optional[Tensor]):
        if t is None:
            return None
        return (t.data_ptr(), tuple(t.shape), tuple(t.stride()), t._version, str(t.device), t.dtype)

    @staticmethod
    def _fingerprint(tensors):
        """One (sum, weighted sum) pair per tensor, `()` for None / empty / CPU tensors; one read-back for all."""
        live = [i for i, t in enumerate(tensors) if t is not None and t.numel() and t.is_cuda]
        res = [()] * len(tensors)
        if not live:
            return res
        dev = tensors[live[0]].device
        out = torch.zeros(2 * len(live), dtype=torch.int64, device=dev)
        lib = _lib.load()
        with torch.cuda.device(dev):
            for k, i in enumerate(live):
                w = tensors[i].detach()
                if not w.is_contiguous():
                    w = w.contiguous()
                _lib.check(lib.pgsd_fingerprint(w.data_ptr(), w.numel() * w.element_size(),
                                                out.data_ptr() + 16 * k, _stream_ptr(dev)), "pgsd_fingerprint")
        vals = out.tolist()
        for k, i in enumerate(live):
            res[i] = (vals[2 * k], vals[2 * k + 1])
        return res

    def get(self, tensors, extra, builder):
        return self.get_many([(tensors, extra, builder)])[0]

    def get_many(self, requests):
        """[(tensors, extra, builder)] -> [plan]; the fingerprints of all requests are read back with ONE
        device->host synchronisation (SGCNConv / SNEAConv ask for the positive and the negative plan together)."""
        import os
        mode = os.environ.get("PGSD_PLAN_REUSE", "fingerprint")
        if mode == "off":
            return [b() for _, _, b in requests]
        keys = [(tuple(self._key(t) for t in ts), extra) for ts, extra, _ in requests]
        fps = [None] * len(requests)
        if mode == "fingerprint":
            per_tensor = self._fingerprint([t for ts, _, _ in requests for t in ts])
            pos = 0
            for r, (ts, _, _) in enumerate(requests):
                fps[r] = tuple(per_tensor[pos:pos + len(ts)])
                pos += len(ts)
        out = []
        for key, fp, (ts, _, builder) in zip(keys, fps, requests):
            hit = self._items.get(key)
            if hit is not None and (fp is None or hit[2] is None or hit[2] == fp):
                out.append(hit[0])
                continue
            plan = builder()
            self._items.pop(key, None)
            if len(self._items) >= self.capacity:
                self._items.pop(next(iter(self._items)))
            # keep the key tensors alive so a recycled data_ptr cannot alias a stale plan
            self._items[key] = (plan, ts, fp)
            out.append(plan)
        return out

    def clear(self):
        self._items.clear()

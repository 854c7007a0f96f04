"""Sparse, on-GPU versions of the preprocessing the reference runs with dense N x N matrices on the CPU
before a DiGCN / DGCN model can be trained -- same names, signatures and return
conventions:

  get_appr_directed_adj    utils/directed/get_adjs_DiGCN.py:113-190
  get_second_directed_adj  utils/directed/get_adjs_DiGCN.py:193-254
  directed_features_in_out utils/directed/features_in_out.py:9-60

The reference materialises p_dense (N^2 floats), a dense (N+1)^2 matrix whose leading left eigenvector
it extracts with scipy.linalg.eig (O(N^3)), and dense torch.mm products; BASELINE config 3 (500k nodes)
can therefore only be benchmarked on synthetic operators upstream.  Here the same operators are built
from three sparse primitives of the C ABI (`csrc/preprocess.cu`): `pgsd_coo_coalesce`,
`pgsd_gram_expand` (expand-sort-compress SpGEMM for B^T diag(s) B) and `pgsd_ppr_stationary` (fp64
power iteration for the personalised-PageRank stationary vector).  Index arithmetic, masks and
degree scatters are torch tensor ops on the device.  Outputs are in the reference's order (row-major
sorted, the order of torch.nonzero on its dense result) and agree with it to fp32 rounding (its
eigenvector comes out of a float32 LAPACK sgeev; ours is converged in fp64).
"""
from __future__ import annotations

import ctypes as C
import math
from typing import Optional, Tuple

import torch
from torch import Tensor

from .. import _lib, plan as _plan


def _stream(dev) -> int:
    return torch.cuda.current_stream(dev).cuda_stream


def _key_bits(n: int) -> int:
    return max(1, int(n * n - 1).bit_length()) if n > 1 else 1


def coalesce(rows: Tensor, cols: Tensor, vals: Tensor, n: int) -> Tuple[Tensor, Tensor, Tensor]:
    """Entries sorted by (row, col) with duplicates summed (`pgsd_coo_coalesce`)."""
    dev, m = rows.device, rows.numel()
    if m == 0:
        return rows.long(), cols.long(), vals.float()
    if n >= (1 << 31):
        raise ValueError("num_nodes exceeds the int32 plan range")
    keys = (rows.long() * n + cols.long()).contiguous()
    vals = vals.float().contiguous()
    lib = _lib.load()
    nbytes = C.c_size_t(0)
    _lib.check(lib.pgsd_coalesce_workspace_bytes(m, C.byref(nbytes)), "pgsd_coalesce_workspace_bytes")
    with torch.cuda.device(dev):
        ws = torch.empty(nbytes.value, dtype=torch.uint8, device=dev)
        k_out = torch.empty(m, dtype=torch.int64, device=dev)
        v_out = torch.empty(m, dtype=torch.float32, device=dev)
        n_u = C.c_int64(0)
        _lib.check(lib.pgsd_coo_coalesce(keys.data_ptr(), vals.data_ptr(), m, _key_bits(n), k_out.data_ptr(),
                                         v_out.data_ptr(), C.byref(n_u), ws.data_ptr(), ws.numel(), _stream(dev)),
                   "pgsd_coo_coalesce")
    k_out, v_out = k_out[:n_u.value], v_out[:n_u.value]
    return torch.div(k_out, n, rounding_mode="floor"), k_out % n, v_out


def gram(rows: Tensor, cols: Tensor, vals: Tensor, n: int, scale: Optional[Tensor] = None):
    """C = B^T diag(scale) B for the COO matrix B (entries (rows -> k, cols -> i, vals)): C[i, j] =
    sum_k scale[k] B[k, i] B[k, j], returned coalesced and sorted (`pgsd_gram_expand` + `pgsd_coo_coalesce`)."""
    dev = rows.device
    b = _plan.build_csr(torch.stack([cols, rows]), vals, n, n, "source_to_target")   # CSR by `rows`
    lens = (b.row_ptr[1:] - b.row_ptr[:-1]).long()
    offs = torch.zeros(n + 1, dtype=torch.int64, device=dev)
    offs[1:] = torch.cumsum(lens * lens, 0)
    total = int(offs[-1].item())
    if total == 0:
        z = torch.zeros(0, dtype=torch.int64, device=dev)
        return z, z.clone(), torch.zeros(0, dtype=torch.float32, device=dev)
    if total >= (1 << 31):
        raise ValueError(f"gram: {total} products exceed the int32 range of the sort; split the graph")
    sc = None if scale is None else scale.float().contiguous()
    lib = _lib.load()
    with torch.cuda.device(dev):
        keys = torch.empty(total, dtype=torch.int64, device=dev)
        pv = torch.empty(total, dtype=torch.float32, device=dev)
        _lib.check(lib.pgsd_gram_expand(b.row_ptr.data_ptr(), b.col.data_ptr(), b.val[0].data_ptr(),
                                        None if sc is None else sc.data_ptr(), offs.data_ptr(), n, n, total,
                                        keys.data_ptr(), pv.data_ptr(), _stream(dev)), "pgsd_gram_expand")
    return coalesce(torch.div(keys, n, rounding_mode="floor"), keys % n, pv, n)


def _row_stochastic(edge_index: Tensor, num_nodes: int, dtype, edge_weight: Optional[Tensor]):
    """add_self_loops(fill 1) -> p = D^-1 A, duplicates summed (get_adjs_DiGCN.py:136-146)."""
    _plan.require_cuda(edge_index, "edge_index")
    dev = edge_index.device
    if edge_weight is None:
        edge_weight = torch.ones((edge_index.size(1),), dtype=dtype, device=dev)
    loops = torch.arange(num_nodes, device=dev, dtype=edge_index.dtype)
    row = torch.cat([edge_index[0], loops])
    col = torch.cat([edge_index[1], loops])
    w = torch.cat([edge_weight.to(torch.float32), torch.ones(num_nodes, dtype=torch.float32, device=dev)])
    deg = torch.zeros(num_nodes, dtype=torch.float32, device=dev).scatter_add_(0, row, w)
    deg_inv = deg.pow(-1)
    deg_inv[deg_inv == float("inf")] = 0
    return coalesce(row, col, deg_inv[row] * w, num_nodes)


def _sym_normalise(row: Tensor, col: Tensor, val: Tensor, n: int):
    """Drop exact zeros (torch.nonzero of the dense result), then deg^-1/2 L deg^-1/2 with
    deg = row sums (get_adjs_DiGCN.py:176-190 / :240-254)."""
    val = torch.nan_to_num(val, nan=0.0, posinf=float("inf"), neginf=float("-inf"))
    keep = val != 0
    row, col, val = row[keep], col[keep], val[keep]
    deg = torch.zeros(n, dtype=torch.float32, device=val.device).scatter_add_(0, row, val)
    dis = deg.pow(-0.5)
    dis[dis == float("inf")] = 0
    return torch.stack([row, col]), dis[row] * val * dis[col]


def ppr_stationary(prow: Tensor, pcol: Tensor, pval: Tensor, n: int, alpha: float, tol: float = 1e-15) -> Tensor:
    """pi (fp64, normalised to sum 1 over the N real states) of the (N+1)-state chain of
    get_adjs_DiGCN.py:147-160 (`pgsd_ppr_stationary`)."""
    dev = prow.device
    by_dst = _plan.build_csr(torch.stack([prow, pcol]), pval, n, n, "source_to_target")
    n_iter = int(min(20000, max(50, math.ceil(math.log(tol) / math.log(1.0 - alpha)))))
    lib = _lib.load()
    with torch.cuda.device(dev):
        pi = torch.empty(n, dtype=torch.float64, device=dev)
        tmp = torch.empty(n, dtype=torch.float64, device=dev)
        _lib.check(lib.pgsd_ppr_stationary(by_dst.row_ptr.data_ptr(), by_dst.col.data_ptr(), by_dst.val[0].data_ptr(),
                                           n, float(alpha), n_iter, pi.data_ptr(), tmp.data_ptr(), _stream(dev)),
                   "pgsd_ppr_stationary")
    return pi / pi.sum()


def get_appr_directed_adj(alpha: float, edge_index: Tensor, num_nodes: int, dtype: torch.dtype,
                          edge_weight: Optional[Tensor] = None) -> Tuple[Tensor, Tensor]:
    """Approximate-PageRank symmetrised adjacency of DiGCN (get_adjs_DiGCN.py:113-190):
    L = (Pi^1/2 P Pi^-1/2 + Pi^-1/2 P^T Pi^1/2) / 2, symmetrically normalised; (edge_index, edge_weight)."""
    n = int(num_nodes)
    prow, pcol, pval = _row_stochastic(edge_index, n, dtype, edge_weight)
    pi = ppr_stationary(prow, pcol, pval, n, alpha).to(torch.float32)
    assert not bool((pi < 0).any())
    pi_sqrt = pi.pow(0.5)
    pi_inv_sqrt = pi.pow(-0.5)
    pi_inv_sqrt[pi_inv_sqrt == float("inf")] = 0
    t = (pi_sqrt[prow] * pval) * pi_inv_sqrt[pcol]          # entry (i, j) of Pi^1/2 P Pi^-1/2; its transpose is the other term
    row, col, val = coalesce(torch.cat([prow, pcol]), torch.cat([pcol, prow]), torch.cat([t, t]), n)
    return _sym_normalise(row, col, val / 2.0, n)


def get_second_directed_adj(edge_index: Tensor, num_nodes: int, dtype: torch.dtype,
                            edge_weight: Optional[Tensor] = None) -> Tuple[Tensor, Tensor]:
    """Second-order proximity of DiGCN (get_adjs_DiGCN.py:193-254): L_in = P^T P and L_out = P P^T, both kept
    only where BOTH are non-zero (the reference's two in-place masks alias the same tensors), averaged and
    symmetrically normalised."""
    n = int(num_nodes)
    prow, pcol, pval = _row_stochastic(edge_index, n, dtype, edge_weight)
    ir, ic, iv = gram(prow, pcol, pval, n)                  # (P^T P)[a, b] = sum_k p[k, a] p[k, b]
    orow, ocol, ov = gram(pcol, prow, pval, n)              # (P P^T)[a, b] = sum_k p[a, k] p[b, k]
    kin, kout = ir * n + ic, orow * n + ocol
    iv = torch.where(iv != 0, iv, torch.zeros_like(iv))
    in_both = torch.isin(kin, kout[ov != 0]) & (iv != 0)
    kin, iv = kin[in_both], iv[in_both]
    pos = torch.searchsorted(kout, kin)
    val = (iv + ov[pos]) / 2.0
    return _sym_normalise(torch.div(kin, n, rounding_mode="floor"), kin % n, val, n)


def directed_features_in_out(edge_index: Tensor, size: int, edge_weight: Optional[Tensor] = None,
                             device: str = 'cpu'):
    """DGCN's first/second-order operators (features_in_out.py:9-60):
    A_in = A^T diag(1/colsum) A and A_out = A diag(1/rowsum) A^T (zero sums read as 1 -- note that the reference
    scales row k's outer product by the COLUMN sum of k and vice versa; reproduced), plus the undirected edge set.
    Returns (index_undirected, edge_in, in_weight, edge_out, out_weight) on edge_index's device, as upstream
    (its `device` argument is ignored there too)."""
    _plan.require_cuda(edge_index, "edge_index")
    n, dev = int(size), edge_index.device
    w = torch.ones(edge_index.size(1), dtype=torch.float32, device=dev) if edge_weight is None \
        else edge_weight.to(torch.float32)
    arow, acol, aval = coalesce(edge_index[0], edge_index[1], w, n)
    col_sum = torch.zeros(n, dtype=torch.float32, device=dev).scatter_add_(0, acol, aval)     # "out_degree"
    row_sum = torch.zeros(n, dtype=torch.float32, device=dev).scatter_add_(0, arow, aval)     # "in_degree"
    col_sum[col_sum == 0] = 1
    row_sum[row_sum == 0] = 1
    ir, ic, iv = gram(arow, acol, aval, n, 1.0 / col_sum)    # sum_k a[k, i] a[k, j] / out_degree[k]
    orow, ocol, ov = gram(acol, arow, aval, n, 1.0 / row_sum)  # sum_k a[i, k] a[j, k] / in_degree[k]
    ur, uc, _ = coalesce(torch.cat([edge_index[0], edge_index[1]]), torch.cat([edge_index[1], edge_index[0]]),
                         torch.ones(2 * edge_index.size(1), dtype=torch.float32, device=dev), n)
    return torch.stack([ur, uc]), torch.stack([ir, ic]), iv, torch.stack([orow, ocol]), ov

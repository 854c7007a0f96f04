"""GPU motif mining for SDGNN / SiGAT.

The reference builds, in Python, six dictionaries of neighbour SETS from the signed edge list and then, for every
signed edge (u, v), sixteen set intersections (nn/signed/SDGNN.py:153-254, nn/signed/SiGAT.py:93-186).  Here the
four directed lists are duplicate-free CSR rows built with sorts on the device, and the sixteen intersection sizes
per edge come from one kernel launch (`pgsd_signed_triangle_counts`, csrc/motif.cu).  Everything is integer work
and bit-exact: the adjacency lists equal the reference's as sets (the reference's edge ORDER inside a list is Python
set iteration order and carries no meaning -- GATConv sums over neighbours).
"""
from __future__ import annotations

import ctypes as C
from typing import List, Tuple

import torch
from torch import Tensor

from .. import _lib
from ..plan import require_cuda

# masks of nn/signed/SDGNN.py:229,236: which of the 16 counts make up the triangle weight of a +/- edge
SDGNN_MASK_POS = (1, 0, 0, 0, 1, 0, 0, 1, 0, 0, 0, 1, 1, 0, 0, 1)
SDGNN_MASK_NEG = (0, 1, 1, 0, 0, 0, 1, 0, 0, 1, 1, 0, 0, 1, 0, 0)


def _unique_pairs(a: Tensor, b: Tensor, n: int) -> Tuple[Tensor, Tensor]:
    """Distinct (a, b) pairs, sorted by (a, b)."""
    key = torch.unique(a * n + b)
    return key // n, key % n


def _csr(a: Tensor, b: Tensor, n: int):
    rp = torch.zeros(n + 1, dtype=torch.int32, device=a.device)
    if a.numel():
        rp[1:] = torch.cumsum(torch.bincount(a, minlength=n), 0).int()
    return rp, b.int().contiguous()


def signed_lists(edge_index_s: Tensor, n: int):
    """(pos_out, pos_in, neg_out, neg_in, pos, neg): each a (src, dst) pair of int64 vectors listing the distinct
    (a, b) with b in list[a], sorted by (a, b) -- the reference's defaultdict(set)s (SDGNN.py:203-218)."""
    require_cuda(edge_index_s, "edge_index_s")
    e = edge_index_s.long()
    i, j, s = e[:, 0], e[:, 1], e[:, 2]
    out = []
    for m in (s > 0, s < 0):
        ii, jj = i[m], j[m]
        out.append((_unique_pairs(ii, jj, n), _unique_pairs(jj, ii, n),
                    _unique_pairs(torch.cat([ii, jj]), torch.cat([jj, ii]), n)))
    (p_out, p_in, p_all), (n_out, n_in, n_all) = out
    return p_out, p_in, n_out, n_in, p_all, n_all


def triangle_counts(lists4, edge_u: Tensor, edge_v: Tensor, n: int) -> Tensor:
    """[E, 16] int32 motif counts of the edges (edge_u, edge_v) against the four directed lists
    (pos_out, pos_in, neg_out, neg_in), in the order of the reference's get_features() tuple."""
    dev = edge_u.device
    csr = [_csr(a, b, n) for a, b in lists4]
    rp = (C.c_void_p * 4)(*[t[0].data_ptr() for t in csr])
    cl = (C.c_void_p * 4)(*[t[1].data_ptr() if t[1].numel() else None for t in csr])
    eu, ev = edge_u.long().contiguous(), edge_v.long().contiguous()
    counts = torch.empty((eu.numel(), 16), dtype=torch.int32, device=dev)
    lib = _lib.load()
    with torch.cuda.device(dev):
        _lib.check(lib.pgsd_signed_triangle_counts(rp, cl, eu.data_ptr(), ev.data_ptr(), eu.numel(), n,
                                                   counts.data_ptr(), torch.cuda.current_stream(dev).cuda_stream),
                   "pgsd_signed_triangle_counts")
    return counts


def sdgnn_motifs(edge_index_s: Tensor, n: int):
    """SDGNN.build_adj_lists (SDGNN.py:197-254): returns (edge_lists, (row, col, value)) --
    edge_lists = [pos_out, pos_in, neg_out, neg_in] as [2, m] int64 tensors (a, b), and the triangle weights of
    `tri_weight` as COO triplets: one entry per distinct signed edge, a (u, v) that is both + and - keeps the
    value of the - pass (the reference overwrites weight_dict[i][j] in its second loop)."""
    p_out, p_in, n_out, n_in, _, _ = signed_lists(edge_index_s, n)
    lists4 = (p_out, p_in, n_out, n_in)
    dev = edge_index_s.device
    cp = triangle_counts(lists4, p_out[0], p_out[1], n).long()
    cn = triangle_counts(lists4, n_out[0], n_out[1], n).long()
    wp = (cp * torch.tensor(SDGNN_MASK_POS, device=dev)).sum(1)
    wn = (cn * torch.tensor(SDGNN_MASK_NEG, device=dev)).sum(1)
    key_p, key_n = p_out[0] * n + p_out[1], n_out[0] * n + n_out[1]
    keep = ~torch.isin(key_p, key_n)                      # overwritten by the negative pass
    row = torch.cat([p_out[0][keep], n_out[0]])
    col = torch.cat([p_out[1][keep], n_out[1]])
    val = torch.cat([wp[keep], wn])
    order = torch.argsort(row * n + col)
    edge_lists = [torch.stack(l) for l in lists4]
    return edge_lists, (row[order], col[order], val[order])


def sigat_motifs(edge_index_s: Tensor, n: int) -> List[Tensor]:
    """SiGAT.build_adj_lists (SiGAT.py:136-186): the 38 adjacency lists as [2, m] int64 edge tensors --
    pos, pos_out, pos_in, neg, neg_out, neg_in, then for t < 16 the + edges whose t-th motif count is non-zero,
    then the same for the - edges."""
    p_out, p_in, n_out, n_in, p_all, n_all = signed_lists(edge_index_s, n)
    lists4 = (p_out, p_in, n_out, n_in)
    cp = triangle_counts(lists4, p_out[0], p_out[1], n)
    cn = triangle_counts(lists4, n_out[0], n_out[1], n)
    res = [torch.stack(l) for l in (p_all, p_out, p_in, n_all, n_out, n_in)]
    for (a, b), c in ((p_out, cp), (n_out, cn)):
        for t in range(16):
            m = c[:, t] > 0
            res.append(torch.stack([a[m], b[m]]))
    return res

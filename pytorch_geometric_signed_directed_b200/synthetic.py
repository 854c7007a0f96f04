"""Vectorised synthetic DSBM / SSBM input generators (bench + test INPUTS only).

The reference's generators (`data/directed/DSBM.py:10-55` via networkx's per-pair loop,
`data/signed/SSBM.py:9-140` via a per-node Python loop) define the *distribution* of the
benchmark graphs but cannot produce 20M-160M edges in reasonable time.  These
functions draw from the same distributions with tensor ops (on CPU or on the GPU), seeded by
an explicit `torch.Generator`:

  * cluster sizes: geometric sequence from `size_ratio` (DSBM.py:33-45 / SSBM.py:60-71);
  * DSBM: every ordered pair (i, j), i != j, is an edge w.p. p * F[c(i), c(j)], independently;
    node ids randomly permuted (DSBM.py:32,50-53).  Here: per block pair the edge COUNT is
    drawn from the matching binomial (normal approximation above 1e6 trials), end points are
    drawn uniformly inside the blocks, duplicates and self-loops are dropped.  `extract_network`
    (largest connected component + degree pruning) is skipped at scale.
  * SSBM: undirected pairs, kept w.p. pin / pout, sign + inside / - across communities, each
    flipped w.p. eta; stored in both directions (SSBM.py:95-138), unit weights.
"""
from __future__ import annotations

import math
from typing import Optional, Tuple

import numpy as np
import torch
from torch import Tensor


def meta_graph_cyclic(K: int = 3, eta: float = 0.1, fill_val: float = 0.5) -> np.ndarray:
    """`utils/directed/meta_graph_generation.py:6-94` for F_style='cyclic', ambient=False:
    F[i,i] = 0.5, F[i,i+1] = 1-eta, F[i+1,i] = eta, untouched entries = fill_val."""
    F = np.eye(K) * 0.5
    if K > 2:
        for i in range(K):
            j = (i + 1) % K
            F[i, j] = 1.0 - eta
            F[j, i] = 1.0 - F[i, j]
    else:
        F = np.array([[0.5, 1 - eta], [eta, 0.5]])
    F[F == 0] = fill_val
    return F


def cluster_sizes(n: int, k: int, size_ratio: float):
    if size_ratio > 1 and k > 1:
        r = size_ratio ** (1.0 / (k - 1))
        sizes = [math.floor(n * (1 - r) / (1 - r ** k))]
        for _ in range(1, k - 1):
            sizes.append(math.floor(sizes[-1] * r))
        sizes.append(n - sum(sizes))
    else:
        sizes = [math.floor((i + 1) * n / k) - math.floor(i * n / k) for i in range(k)]
    return sizes


def _binomial(trials: float, prob: float, gen: torch.Generator) -> int:
    if trials <= 0 or prob <= 0:
        return 0
    mean, var = trials * prob, trials * prob * (1 - prob)
    if trials < 1e6:
        return int(torch.binomial(torch.tensor([float(trials)], dtype=torch.float64),
                                  torch.tensor([float(prob)], dtype=torch.float64),
                                  generator=gen).item())
    z = torch.randn(1, generator=gen).item()
    return max(0, int(round(mean + z * math.sqrt(var))))


def _dedup(src: Tensor, dst: Tensor, n: int) -> Tuple[Tensor, Tensor]:
    key = torch.unique(src * n + dst)
    return key // n, key % n


def dsbm_edges(n: int, k: int = 3, p: Optional[float] = None, num_edges: Optional[int] = None,
               eta: float = 0.1, size_ratio: float = 1.5, seed: int = 0,
               device: str = "cpu") -> Tuple[Tensor, Tensor]:
    """Returns (edge_index int64 [2,E] on `device`, labels int64 [n]).  Give either the
    sparsity `p` or a target `num_edges` (p is then solved from E = p * sum_ab n_a n_b F_ab)."""
    F = meta_graph_cyclic(k, eta, 0.5)
    sizes = cluster_sizes(n, k, size_ratio)
    starts = np.concatenate([[0], np.cumsum(sizes)])
    pairs = np.array([[float(sizes[a]) * (sizes[b] - (a == b)) for b in range(k)]
                      for a in range(k)])
    if p is None:
        assert num_edges is not None
        p = num_edges / float((pairs * F).sum())
    cpu_gen = torch.Generator().manual_seed(seed)
    dev_gen = torch.Generator(device=device).manual_seed(seed + 1)
    srcs, dsts = [], []
    for a in range(k):
        for b in range(k):
            m = _binomial(pairs[a, b], min(1.0, p * F[a, b]), cpu_gen)
            if m == 0:
                continue
            srcs.append(torch.randint(int(starts[a]), int(starts[a + 1]), (m,),
                                      generator=dev_gen, device=device))
            dsts.append(torch.randint(int(starts[b]), int(starts[b + 1]), (m,),
                                      generator=dev_gen, device=device))
    if not srcs:
        return torch.zeros((2, 0), dtype=torch.long, device=device), torch.zeros(n, dtype=torch.long)
    src, dst = torch.cat(srcs), torch.cat(dsts)
    keep = src != dst
    src, dst = _dedup(src[keep], dst[keep], n)
    perm = torch.randperm(n, generator=cpu_gen)
    labels = torch.empty(n, dtype=torch.long)
    labels[perm] = torch.repeat_interleave(torch.arange(k), torch.tensor(sizes))
    perm_d = perm.to(device)
    edge_index = torch.stack([perm_d[src], perm_d[dst]])
    # shuffle edge order: the layers accept COO in arbitrary order
    order = torch.randperm(edge_index.size(1), generator=dev_gen, device=device)
    return edge_index[:, order].contiguous(), labels


def locality_edges(n: int, num_edges: int, band: int, long_range: float = 0.0, seed: int = 0,
                   device: str = "cpu") -> Tensor:
    """A directed graph whose edge list shards naturally under a 1-D node-range split: node ids
    carry locality (a road network / mesh / time-ordered citation graph after a bandwidth-reducing
    ordering), every edge i -> j has |i - j| <= band, except a `long_range` share that lands
    anywhere.  No reference counterpart (the reference has no distributed path); used to exercise
    the halo exchange of distributed.py.  Returns edge_index int64 [2, E'] (self-loops and
    duplicates removed, edge order shuffled)."""
    gen = torch.Generator(device=device).manual_seed(seed)
    src = torch.randint(0, n, (num_edges,), generator=gen, device=device)
    off = torch.randint(-band, band + 1, (num_edges,), generator=gen, device=device)
    dst = (src + off).clamp_(0, n - 1)
    if long_range > 0:
        far = torch.rand(num_edges, generator=gen, device=device) < long_range
        dst = torch.where(far, torch.randint(0, n, (num_edges,), generator=gen, device=device), dst)
    keep = src != dst
    src, dst = _dedup(src[keep], dst[keep], n)
    order = torch.randperm(src.numel(), generator=gen, device=device)
    return torch.stack([src[order], dst[order]]).contiguous()


def ssbm_edges(n: int, k: int = 3, p: Optional[float] = None, num_entries: Optional[int] = None,
               eta: float = 0.1, size_ratio: float = 2.0, seed: int = 0,
               device: str = "cpu") -> Tuple[Tensor, Tensor, Tensor]:
    """Returns (pos_edge_index, neg_edge_index, labels); every undirected pair is stored in
    both directions, `num_entries` counts stored (directed) entries of both signs."""
    sizes = cluster_sizes(n, k, size_ratio)
    starts = np.concatenate([[0], np.cumsum(sizes)])
    if p is None:
        assert num_entries is not None
        p = num_entries / (float(n) * (n - 1))
    cpu_gen = torch.Generator().manual_seed(seed)
    dev_gen = torch.Generator(device=device).manual_seed(seed + 1)
    us, vs = [], []
    for a in range(k):
        for b in range(a, k):
            trials = sizes[a] * (sizes[a] - 1) / 2.0 if a == b else float(sizes[a]) * sizes[b]
            m = _binomial(trials, p, cpu_gen)
            if m == 0:
                continue
            us.append(torch.randint(int(starts[a]), int(starts[a + 1]), (m,),
                                    generator=dev_gen, device=device))
            vs.append(torch.randint(int(starts[b]), int(starts[b + 1]), (m,),
                                    generator=dev_gen, device=device))
    u, v = torch.cat(us), torch.cat(vs)
    keep = u != v
    u, v = u[keep], v[keep]
    lo, hi = torch.minimum(u, v), torch.maximum(u, v)
    key = torch.unique(lo * n + hi)
    lo, hi = key // n, key % n
    lab_sorted = torch.repeat_interleave(torch.arange(k), torch.tensor(sizes)).to(device)
    ins = lab_sorted[lo] == lab_sorted[hi]
    flip = torch.rand(lo.numel(), generator=dev_gen, device=device) < eta
    positive = ins ^ flip
    perm = torch.randperm(n, generator=cpu_gen)
    labels = torch.empty(n, dtype=torch.long)
    labels[perm] = lab_sorted.cpu()
    perm_d = perm.to(device)
    lo, hi = perm_d[lo], perm_d[hi]

    def both(mask):
        a, b = lo[mask], hi[mask]
        ei = torch.stack([torch.cat([a, b]), torch.cat([b, a])])
        order = torch.randperm(ei.size(1), generator=dev_gen, device=device)
        return ei[:, order].contiguous()

    return both(positive), both(~positive), labels


def sym_norm_weights(edge_index: Tensor, n: int) -> Tensor:
    """w_ij = d_i^-1/2 d_j^-1/2 with d = out-degree + in-degree counts: the last step of the
    reference's DiGCN preprocessing (`utils/directed/get_adjs_DiGCN.py:189-195`) applied to a
    synthetic pattern (the true approximate-PageRank operator is dense-built upstream, F10)."""
    ones = torch.ones(edge_index.size(1), device=edge_index.device)
    deg = torch.zeros(n, device=edge_index.device)
    deg.scatter_add_(0, edge_index[0], ones)
    dis = deg.clamp(min=1).pow(-0.5)
    return (dis[edge_index[0]] * dis[edge_index[1]]).contiguous()

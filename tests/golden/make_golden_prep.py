"""Golden fixtures for the DiGCN / DGCN preprocessing utilities, produced by the REFERENCE's
own functions (dense eig / dense mm / scipy loops) on small seeded graphs:

    python tests/golden/make_golden_prep.py        (build container only)
"""
from __future__ import annotations

import os
import sys
import warnings

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
warnings.filterwarnings("ignore")

from oracle import load_reference  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))
adjs = load_reference.load("utils.directed.get_adjs_DiGCN")
fio = load_reference.load("utils.directed.features_in_out")


def graph(n, e, seed, weighted):
    g = torch.Generator().manual_seed(seed)
    live = n - 4                                    # the last ids are isolated
    ei = torch.randint(0, live, (2, e), generator=g)
    ei[:, :8] = ei[:, 8:16]                         # duplicate edges
    ei[:, 16:30] = ei[:, 30:44].flip(0)             # reciprocal pairs
    ei[1, 44:50] = ei[0, 44:50]                     # self loops
    w = (torch.rand(e, generator=g) + 0.5) if weighted else None
    return ei, w


def sort_coo(ei, w):
    key = ei[0] * (int(ei.max()) + 1 if ei.numel() else 1) + ei[1]
    order = torch.argsort(key, stable=True)
    return ei[:, order], w[order]


for name, n, e, seed, weighted, alpha in (("prep_a", 90, 420, 1, True, 0.1), ("prep_b", 150, 900, 2, False, 0.2),
                                          ("prep_c", 40, 60, 3, True, 0.05)):
    ei, w = graph(n, e, seed, weighted)
    a_ei, a_w = adjs.get_appr_directed_adj(alpha, ei, n, torch.float32, w)
    s_ei, s_w = adjs.get_second_directed_adj(ei, n, torch.float32, w)
    und, e_in, w_in, e_out, w_out = fio.directed_features_in_out(ei, n, w)
    e_in, w_in = sort_coo(e_in, w_in)
    e_out, w_out = sort_coo(e_out, w_out)
    arrays = dict(edge_index=ei, has_weight=weighted, alpha=alpha, n=n,
                  appr_index=a_ei, appr_weight=a_w, second_index=s_ei, second_weight=s_w,
                  undirected=und, in_index=e_in, in_weight=w_in, out_index=e_out, out_weight=w_out)
    if weighted:
        arrays["edge_weight"] = w
    np.savez_compressed(os.path.join(OUT, name + ".npz"),
                        **{k: (v.detach().cpu().numpy() if isinstance(v, torch.Tensor) else np.asarray(v))
                           for k, v in arrays.items()})
    print(name, {k: tuple(v.shape) for k, v in arrays.items() if isinstance(v, torch.Tensor)})

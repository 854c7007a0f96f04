"""Golden fixtures for SNEAConv (second batch; same method as make_golden.py: the reference's
own nn/signed/SNEAConv.py, loaded unmodified, on seeded inputs).

    python tests/golden/make_golden_snea.py
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))

from oracle import load_reference  # noqa: E402
from pytorch_geometric_signed_directed_b200 import synthetic  # noqa: E402
from make_golden import save  # noqa: E402

SNEAConv = load_reference.ref_classes()["SNEAConv"]


def params(c):
    return dict(lin_b_weight=c.lin_b.weight, lin_b_bias=c.lin_b.bias, lin_u_weight=c.lin_u.weight,
                lin_u_bias=c.lin_u.bias, alpha_b_weight=c.alpha_b.weight, alpha_b_bias=c.alpha_b.bias,
                alpha_u_weight=c.alpha_u.weight, alpha_u_bias=c.alpha_u.bias)


def main():
    n = 140
    pos, neg, _ = synthetic.ssbm_edges(n - 6, k=3, num_entries=1500, eta=0.1, seed=60)   # ids 134..139 isolated
    pos = torch.cat([pos, pos[:, :5], torch.tensor([[2, 7], [2, 7]])], 1)     # duplicates + self loops
    neg = torch.cat([neg, torch.tensor([[4, 4], [4, 4]])], 1)
    torch.manual_seed(61)
    c1 = SNEAConv(8, 6, first_aggr=True)
    x = torch.randn(n, 8)
    with torch.no_grad():
        z1 = c1(x, pos, neg)
    save("snea_first", x=x, pos_edge_index=pos, neg_edge_index=neg, out=z1, **params(c1))
    c2 = SNEAConv(6, 5, first_aggr=False)
    with torch.no_grad():
        z2 = c2(torch.tanh(z1), pos, neg)
    save("snea_second", x=torch.tanh(z1), pos_edge_index=pos, neg_edge_index=neg, out=z2, **params(c2))


if __name__ == "__main__":
    main()

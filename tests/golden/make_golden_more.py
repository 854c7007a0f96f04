"""Golden fixtures for DGCNConv and SIMPA (third batch; same method as make_golden.py).

    python tests/golden/make_golden_more.py
"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))

from oracle import load_reference  # noqa: E402
from make_golden import nasty_graph, save  # noqa: E402

REF = load_reference.ref_classes()


def main():
    n = 130
    ei, ew = nasty_graph(n, 800, seed=70)
    torch.manual_seed(71)
    x = torch.rand(n, 7) * 2 - 1
    with torch.no_grad():
        y = REF["DGCNConv"]()(x, ei, ew)
        y_imp = REF["DGCNConv"](improved=True)(x, ei, None)
    save("dgcn_conv", x=x, edge_index=ei, edge_weight=ew, out=y, out_improved_unweighted=y_imp)

    ei_p, w_p = nasty_graph(n, 500, seed=72)
    ei_n, w_n = nasty_graph(n, 400, seed=73)
    xs = [torch.rand(n, 5) * 2 - 1 for _ in range(4)]
    und = REF["SIMPA"](hop=2, fill_value=0.5, directed=False)
    with torch.no_grad():
        und._w_p.copy_(torch.tensor([[1.0], [0.5], [-0.7]])); und._w_n.copy_(torch.tensor([[0.3], [1.1], [0.9]]))
        f_und = und(ei_p, w_p, ei_n, w_n, xs[0], xs[1])
    dr = REF["SIMPA"](hop=2, fill_value=0.5, directed=True)
    with torch.no_grad():
        dr._w_sp.copy_(torch.tensor([[1.0], [0.5], [-0.7]])); dr._w_sn.copy_(torch.tensor([[0.3], [1.1], [0.9]]))
        dr._w_tp.copy_(torch.tensor([[0.2], [-0.4], [0.8]])); dr._w_tn.copy_(torch.tensor([[1.3], [0.1], [-0.6]]))
        f_dir = dr(ei_p, w_p, ei_n, w_n, xs[0], xs[1], xs[2], xs[3])
    save("simpa", edge_index_p=ei_p, edge_weight_p=w_p, edge_index_n=ei_n, edge_weight_n=w_n,
         x_p=xs[0], x_n=xs[1], x_pt=xs[2], x_nt=xs[3],
         w_p=und._w_p, w_n=und._w_n, w_sp=dr._w_sp, w_sn=dr._w_sn, w_tp=dr._w_tp, w_tn=dr._w_tn,
         out_undirected=f_und, out_directed=f_dir)


if __name__ == "__main__":
    main()

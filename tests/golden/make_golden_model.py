"""Golden fixture for the MagNet_node_classification wrapper (reference file loaded unmodified).

    python tests/golden/make_golden_model.py
"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))

from oracle import load_reference  # noqa: E402
from make_golden import nasty_graph, save  # noqa: E402

Model = load_reference.ref_classes()["MagNet_node_classification"]


def main():
    n = 160
    ei, ew = nasty_graph(n, 1000, seed=90)
    torch.manual_seed(91)
    model = Model(9, hidden=16, q=0.2, K=2, label_dim=5, activation=True, layer=3, dropout=0.5, cached=True).eval()
    with torch.no_grad():
        for c in model.Chebs:
            c.bias.uniform_(-0.3, 0.3)
    x = torch.rand(n, 9) * 2 - 1
    with torch.no_grad():
        y = model(x, x, ei, ew)
    save("magnet_model", x=x, edge_index=ei, edge_weight=ew, out=y,
         **{k.replace(".", "__"): v for k, v in model.state_dict().items()})


if __name__ == "__main__":
    main()

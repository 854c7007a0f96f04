"""Golden fixture for SDGNN's SDRLayer (fourth batch).  The layer wiring is the reference's
(nn/signed/SDGNN.py:13-64, loaded unmodified); its GATConv is third-party PyG, so the values also
depend on the GATConv restatement in oracle/pyg_shim.py.

    python tests/golden/make_golden_sdr.py
"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))

from oracle import load_reference  # noqa: E402
from make_golden import nasty_graph, save  # noqa: E402

SDRLayer = load_reference.ref_classes()["SDRLayer"]


def main():
    n, d = 120, 12
    lists = [nasty_graph(n, 300 + 40 * i, seed=80 + i, weighted=False)[0] for i in range(4)]
    torch.manual_seed(85)
    layer = SDRLayer(d, d, lists)
    with torch.no_grad():
        for a in layer.aggs:
            a.bias.uniform_(-0.3, 0.3)
    x = torch.randn(n, d)
    with torch.no_grad():
        y = layer(x)
    arrays = {f"edges_{i}": e for i, e in enumerate(lists)}
    arrays.update({k.replace(".", "__"): v for k, v in layer.state_dict().items()})
    save("sdr_layer", x=x, out=y, **arrays)


if __name__ == "__main__":
    main()

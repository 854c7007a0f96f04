"""Gradient fixtures for the attention layers (round 2): torch autograd THROUGH the reference's own, unmodified
nn/signed/SNEAConv.py and nn/signed/SDGNN.py (SDRLayer; its GATConv is PyG's, restated in oracle/pyg_shim.py) on seeded
inputs.  tests/test_oracle_golden.py pins oracle/port.py's gradients to them, and tests/test_gpu_autograd.py compares the
CUDA backward (pgsd_edge_softmax_backward, pgsd_sddmm_rows, transposed aggregation) with the oracle's.

    python tests/golden/make_golden_attn_grad.py
"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))

from oracle import load_reference  # noqa: E402
from pytorch_geometric_signed_directed_b200 import synthetic  # noqa: E402
from make_golden import nasty_graph, save  # noqa: E402

CLS = load_reference.ref_classes()


def main():
    # ---- two SNEAConv layers (first + deep), gradients w.r.t. the input and every parameter
    n = 150
    pos, neg, _ = synthetic.ssbm_edges(n - 4, k=3, num_entries=1800, eta=0.1, seed=90)
    pos = torch.cat([pos, pos[:, :7], torch.tensor([[3, 9], [3, 9]])], 1)       # duplicates + self loops
    neg = torch.cat([neg, torch.tensor([[5, 5], [5, 5]])], 1)
    torch.manual_seed(91)
    c1, c2 = CLS["SNEAConv"](8, 6, first_aggr=True), CLS["SNEAConv"](6, 6, first_aggr=False)
    x = torch.randn(n, 8, requires_grad=True)
    r = torch.randn(n, 12)
    out = c2(torch.tanh(c1(x, pos, neg)), pos, neg)
    (out * r).sum().backward()
    arrays = dict(x=x, r=r, pos_edge_index=pos, neg_edge_index=neg, out=out, grad_x=x.grad)
    for tag, c in (("c1", c1), ("c2", c2)):
        for nm in ("lin_b", "lin_u", "alpha_b", "alpha_u"):
            m = getattr(c, nm)
            arrays[f"{tag}__{nm}__weight"], arrays[f"{tag}__{nm}__bias"] = m.weight, m.bias
            arrays[f"{tag}__{nm}__weight__grad"], arrays[f"{tag}__{nm}__bias__grad"] = m.weight.grad, m.bias.grad
    save("snea_grad", **arrays)

    # ---- SDRLayer (4 GATConv + MLP)
    n, d = 130, 12
    lists = [nasty_graph(n, 320 + 30 * i, seed=95 + i, weighted=False)[0] for i in range(4)]
    torch.manual_seed(99)
    layer = CLS["SDRLayer"](d, d, lists)
    with torch.no_grad():
        for a in layer.aggs:
            a.bias.uniform_(-0.3, 0.3)
    x = torch.randn(n, d, requires_grad=True)
    r = torch.randn(n, d)
    y = layer(x)
    (y * r).sum().backward()
    arrays = {f"edges_{i}": e for i, e in enumerate(lists)}
    arrays.update({k.replace(".", "__"): v for k, v in layer.state_dict().items()})
    arrays.update({k.replace(".", "__") + "__grad": p.grad for k, p in layer.named_parameters()})
    save("sdr_layer_grad", x=x, r=r, out=y, grad_x=x.grad, **arrays)


if __name__ == "__main__":
    main()

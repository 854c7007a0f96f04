"""Golden fixtures for the DiGCN_Inception_Block_node_classification and SGCN model wrappers
(reference files loaded unmodified through oracle/load_reference.py).

    python tests/golden/make_golden_models2.py
"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))

from oracle import load_reference  # noqa: E402
from make_golden import nasty_graph, save  # noqa: E402

R = load_reference.ref_classes()


def main():
    # ---- DiGCN inception model: 3 blocks, two weighted adjacencies
    n = 150
    ei1, w1 = nasty_graph(n, 900, seed=120)
    ei2, w2 = nasty_graph(n, 700, seed=121)
    torch.manual_seed(122)
    model = R["DiGCN_Inception_Block_node_classification"](num_features=7, hidden=16, label_dim=4, dropout=0.5).eval()
    with torch.no_grad():
        for ib in (model.ib1, model.ib2, model.ib3):
            ib.conv1.bias.uniform_(-0.3, 0.3)
            ib.conv2.bias.uniform_(-0.3, 0.3)
    x = torch.rand(n, 7) * 2 - 1
    with torch.no_grad():
        y = model(x, (ei1, ei2), (w1, w2))
    save("digcn_ib_model", x=x, ei1=ei1, w1=w1, ei2=ei2, w2=w2, out=y,
         **{k.replace(".", "__"): v for k, v in model.state_dict().items()})

    # ---- SGCN: signed edge list [E, 3], 3 layers, both norm_emb settings
    n, e = 140, 1100
    g = torch.Generator().manual_seed(130)
    src, dst = torch.randint(0, n - 5, (e,), generator=g), torch.randint(0, n - 5, (e,), generator=g)
    sign = torch.where(torch.rand(e, generator=g) < 0.4, -1, 1)
    edge_index_s = torch.stack([src, dst, sign], 1)
    init = torch.randn(n, 12, generator=g)
    for tag, norm_emb in (("sgcn_model", False), ("sgcn_model_norm", True)):
        torch.manual_seed(131)
        m = R["SGCN"](n, edge_index_s, in_dim=12, out_dim=16, layer_num=3, init_emb=init.clone(), norm_emb=norm_emb).eval()
        with torch.no_grad():
            z = m()
        sd = {k.replace(".", "__"): v for k, v in m.state_dict().items() if not k.startswith(("lsp_loss", "structure_loss"))}
        save(tag, edge_index_s=edge_index_s, out=z, pos_edge_index=m.pos_edge_index, neg_edge_index=m.neg_edge_index, **sd)


if __name__ == "__main__":
    main()

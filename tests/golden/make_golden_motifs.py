"""Golden fixtures for SDGNN / SiGAT: motif mining (build_adj_lists, pure integer work of the reference's own
code) and the model forwards (reference wiring; GATConv restated in oracle/pyg_shim.py).

    python tests/golden/make_golden_motifs.py
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))

from oracle import load_reference  # noqa: E402
from make_golden import save  # noqa: E402

R = load_reference.ref_classes()


def sorted_edges(adj):
    e = sorted((a, b) for a in adj for b in adj[a])
    return torch.tensor(e, dtype=torch.long).reshape(-1, 2).t()


def signed_graph(n, e, seed):
    g = torch.Generator().manual_seed(seed)
    src, dst = torch.randint(0, n - 4, (e,), generator=g), torch.randint(0, n - 4, (e,), generator=g)
    sign = torch.where(torch.rand(e, generator=g) < 0.4, -1, 1)
    es = torch.stack([src, dst, sign], 1)
    es[:30] = es[30:60]                      # duplicate edges
    es[60:80, :2] = es[80:100, :2].flip(1)   # reciprocal pairs
    es[100:110, 2] = -es[110:120, 2]; es[100:110, :2] = es[110:120, :2]   # the same (u, v) with both signs
    es[120:125, 1] = es[120:125, 0]          # self-loops
    return es


def main():
    n, d = 90, 12
    es = signed_graph(n, 700, seed=140)
    init = torch.randn(n, d, generator=torch.Generator().manual_seed(141))

    torch.manual_seed(142)
    m = R["SDGNN"](n, es, in_dim=d, out_dim=d, layer_num=2, init_emb=init.clone()).eval()
    with torch.no_grad():
        for layer in m.layers:
            for a in layer.aggs:
                a.bias.uniform_(-0.3, 0.3)
        z = m()
    tw = m.tri_weight.tocoo()
    order = np.lexsort((tw.col, tw.row))
    arrays = {f"list_{i}": sorted_edges(a) for i, a in enumerate(m.adj_lists)}
    arrays.update({k.replace(".", "__"): v for k, v in m.state_dict().items()
                   if not k.startswith(("loss_",))})
    save("sdgnn_model", edge_index_s=es, out=z, tri_row=torch.from_numpy(tw.row[order].astype(np.int64)),
         tri_col=torch.from_numpy(tw.col[order].astype(np.int64)),
         tri_val=torch.from_numpy(tw.data[order].astype(np.int64)), **arrays)

    torch.manual_seed(143)
    s = R["SiGAT"](n, es, in_dim=d, out_dim=d, init_emb=init.clone()).eval()
    with torch.no_grad():
        for a in s.aggs:
            a.bias.uniform_(-0.3, 0.3)
        z = s()
    arrays = {f"list_{i}": sorted_edges(a) for i, a in enumerate(s.adj_lists)}
    arrays.update({k.replace(".", "__"): v for k, v in s.state_dict().items() if not k.startswith(("lsp_loss",))})
    save("sigat_model", edge_index_s=es, out=z, **arrays)


if __name__ == "__main__":
    main()

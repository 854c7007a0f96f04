"""Generate the golden fixtures in this directory by running the REFERENCE's own, unmodified
hot-path source files (loaded by path from /root/reference through oracle/load_reference.py;
PyG symbols come from oracle/pyg_shim.py because torch_geometric is absent from the image).

Run from the repo root, in the build container only (the GPU box has no /root/reference):

    python tests/golden/make_golden.py

Every case stores its seeded INPUTS and the reference's OUTPUTS, so the tests never need
the reference at run time.  The reference's test-suite pins no values on this path
(shape asserts only), hence these fixtures are the pin.
"""
from __future__ import annotations

import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from oracle import load_reference  # noqa: E402
from pytorch_geometric_signed_directed_b200 import synthetic  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))
REF = load_reference.ref_classes()


def save(name, **arrays):
    conv = {}
    for k, v in arrays.items():
        if isinstance(v, torch.Tensor):
            v = v.detach().cpu().numpy()
        conv[k] = np.asarray(v)
    np.savez_compressed(os.path.join(OUT, name + ".npz"), **conv)
    print(f"{name:28s}", {k: tuple(v.shape) for k, v in conv.items() if v.ndim > 0})


def nasty_graph(n, e, seed, weighted=True, signed=False):
    """Random COO with the edge cases the layers must survive: self-loops, duplicate edges,
    reciprocal pairs, isolated nodes (the last 5 ids never appear)."""
    g = torch.Generator().manual_seed(seed)
    live = n - 5
    ei = torch.randint(0, live, (2, e), generator=g)
    ei[:, :10] = ei[:, 10:20]                      # duplicates
    ei[:, 20:40] = ei[:, 40:60].flip(0)            # reciprocal pairs
    ei[1, 60:70] = ei[0, 60:70]                    # self loops
    ei[:, 70:75] = ei[:, 60:65]                    # duplicate self loops
    w = None
    if weighted:
        w = torch.rand(e, generator=g) * 1.5 + 0.5
        if signed:
            w = w * (torch.randint(0, 2, (e,), generator=g) * 2 - 1).float()
    perm = torch.randperm(e, generator=g)
    return ei[:, perm].contiguous(), (None if w is None else w[perm].contiguous())


def magnet_case(name, cls, n, fin, fout, K, q, normalization, ei, ew, seed, lambda_max=None,
                **ctor):
    torch.manual_seed(seed)
    conv = cls(fin, fout, K=K, q=q, trainable_q=False, normalization=normalization,
               cached=True, **ctor)
    with torch.no_grad():
        conv.bias.uniform_(-0.5, 0.5)
    xr = torch.rand(n, fin) * 2 - 1
    xi = torch.rand(n, fin) * 2 - 1
    with torch.no_grad():
        out_r, out_i = conv(xr, xi, ei, ew, lambda_max)
    er, eim, nr, ni = conv.cached_result
    save(name, x_real=xr, x_imag=xi, edge_index=ei,
         edge_weight=(ew if ew is not None else np.zeros(0, np.float32)),
         has_weight=np.array(ew is not None), weight=conv.weight, bias=conv.bias,
         q=np.array(q, np.float64), K=np.array(K),
         sym=np.array(normalization == 'sym'),
         lambda_max=np.array(-1.0 if lambda_max is None else lambda_max, np.float64),
         out_real=out_r, out_imag=out_i,
         cached_edge_index_real=er, cached_edge_index_imag=eim,
         cached_norm_real=nr, cached_norm_imag=ni)


def main():
    # ---- BASELINE config 1: MagNetConv on a synthetic DSBM graph, 1k nodes / 5k edges / 16 feat
    ei, _ = synthetic.dsbm_edges(1000, k=3, num_edges=5000, eta=0.1, size_ratio=1.5, seed=0)
    magnet_case("magnet_c1", REF["MagNetConv"], 1000, 16, 16, 1, 0.25, 'sym', ei, None, seed=1)

    ei, ew = nasty_graph(150, 900, seed=2)
    magnet_case("magnet_k2_weighted", REF["MagNetConv"], 150, 5, 3, 2, 0.1, 'sym', ei, ew, seed=3)

    ei, ew = nasty_graph(80, 400, seed=4)
    magnet_case("magnet_k3_none", REF["MagNetConv"], 80, 4, 6, 3, 0.2, None, ei, ew, seed=5,
                lambda_max=3.7)

    ei, _ = nasty_graph(64, 300, seed=6, weighted=False)
    magnet_case("magnet_q0", REF["MagNetConv"], 64, 8, 8, 1, 0.0, 'sym', ei, None, seed=7)

    ei, _ = nasty_graph(96, 500, seed=8, weighted=False)
    magnet_case("magnet_sym_lmax", REF["MagNetConv"], 96, 7, 9, 2, 0.25, 'sym', ei, None, seed=9,
                lambda_max=1.6)

    # ---- MSConv (signed magnetic Laplacian)
    ei, ew = nasty_graph(120, 700, seed=10, signed=True)
    magnet_case("msconv_signed", REF["MSConv"], 120, 6, 4, 2, 0.25, 'sym', ei, ew, seed=11)
    ei, ew = nasty_graph(90, 450, seed=12, signed=True)
    magnet_case("msconv_nonabs_none", REF["MSConv"], 90, 3, 5, 1, 0.15, None, ei, ew, seed=13,
                lambda_max=2.5, absolute_degree=False)

    # ---- the Laplacian builder on its own (both normalisations)
    ei, ew = nasty_graph(70, 350, seed=14)
    for tag, norm in (("sym", 'sym'), ("none", None)):
        e2, wr, wi = REF["get_magnetic_Laplacian"](ei, ew, norm, torch.float32, 70, 0.25)
        save(f"laplacian_{tag}", edge_index=ei, edge_weight=ew, q=np.array(0.25),
             out_edge_index=e2, out_real=wr, out_imag=wi)

    # ---- DiGCNConv / inception block
    torch.manual_seed(20)
    n = 200
    ei, ew = nasty_graph(n, 1500, seed=21)
    conv = REF["DiGCNConv"](12, 7)
    with torch.no_grad():
        conv.bias.uniform_(-0.5, 0.5)
    x = torch.rand(n, 12) * 2 - 1
    with torch.no_grad():
        y = conv(x, ei, ew)
    save("digcn_conv", x=x, edge_index=ei, edge_weight=ew, weight=conv.weight, bias=conv.bias, out=y)

    torch.manual_seed(22)
    ei2, ew2 = nasty_graph(n, 1100, seed=23)
    blk = REF["DiGCN_InceptionBlock"](10, 6)
    with torch.no_grad():
        blk.conv1.bias.uniform_(-0.5, 0.5)
        blk.conv2.bias.uniform_(-0.5, 0.5)
    x = torch.rand(n, 10) * 2 - 1
    with torch.no_grad():
        x0, x1, x2 = blk(x, ei, ew, ei2, ew2)
    save("digcn_inception", x=x, edge_index=ei, edge_weight=ew, edge_index2=ei2, edge_weight2=ew2,
         ln_weight=blk.ln.weight, ln_bias=blk.ln.bias,
         conv1_weight=blk.conv1.weight, conv1_bias=blk.conv1.bias,
         conv2_weight=blk.conv2.weight, conv2_bias=blk.conv2.bias, x0=x0, x1=x1, x2=x2)

    # ---- SGCNConv (both aggregation modes)
    n = 150
    pos, neg, _ = synthetic.ssbm_edges(n, k=3, num_entries=1600, eta=0.1, seed=30)
    pos = torch.cat([pos, pos[:, :7]], 1)          # duplicate entries count twice in the mean
    neg = torch.cat([neg, torch.tensor([[3, 9], [3, 9]])], 1)   # self loops are ordinary edges
    torch.manual_seed(31)
    c1 = REF["SGCNConv"](8, 5, first_aggr=True)
    x = torch.randn(n, 8)
    with torch.no_grad():
        z = c1(x, pos, neg)
    save("sgcn_first", x=x, pos_edge_index=pos, neg_edge_index=neg,
         lin_b_weight=c1.lin_b.weight, lin_b_bias=c1.lin_b.bias,
         lin_u_weight=c1.lin_u.weight, lin_u_bias=c1.lin_u.bias, out=z)
    c2 = REF["SGCNConv"](5, 4, first_aggr=False, norm_emb=True)
    with torch.no_grad():
        z2 = c2(torch.tanh(z), pos, neg)
    save("sgcn_second", x=torch.tanh(z), pos_edge_index=pos, neg_edge_index=neg,
         lin_b_weight=c2.lin_b.weight, lin_b_bias=c2.lin_b.bias,
         lin_u_weight=c2.lin_u.weight, lin_u_bias=c2.lin_u.bias, out=z2)

    # ---- conv_norm_rw / Conv_Base / DIMPA
    n = 110
    ei, ew = nasty_graph(n, 600, seed=40)
    e2, w2 = REF["conv_norm_rw"](ei, 0.5, ew, n)
    save("conv_norm_rw", edge_index=ei, edge_weight=ew, fill_value=np.array(0.5),
         out_edge_index=e2, out_weight=w2)
    torch.manual_seed(41)
    x = torch.rand(n, 9) * 2 - 1
    with torch.no_grad():
        y = REF["Conv_Base"](0.5)(x, ei, ew)
        y_unw = REF["Conv_Base"](0.25)(x, ei, None)
    save("conv_base", x=x, edge_index=ei, edge_weight=ew, out=y, out_unweighted_fill025=y_unw)
    dm = REF["DIMPA"](hop=2)
    with torch.no_grad():
        dm._w_s.copy_(torch.tensor([[1.0], [0.7], [-0.3]]))
        dm._w_t.copy_(torch.tensor([[0.5], [1.2], [0.4]]))
    xs, xt = torch.rand(n, 6) * 2 - 1, torch.rand(n, 6) * 2 - 1
    with torch.no_grad():
        feat = dm(xs, xt, ei, ew)
    save("dimpa", x_s=xs, x_t=xt, edge_index=ei, edge_weight=ew, w_s=dm._w_s, w_t=dm._w_t, out=feat)

    # ---- complex ReLU epilogue
    torch.manual_seed(50)
    r, i = torch.randn(40, 6), torch.randn(40, 6)
    r[0, 0] = 0.0
    rr, ii = REF["complex_relu_layer"]()(r, i)
    save("complex_relu", real=r, imag=i, out_real=rr, out_imag=ii)


if __name__ == "__main__":
    main()

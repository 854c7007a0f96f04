"""Third, independent derivation of the hot path: DENSE fp64 linear algebra written
from the papers' formulas -- no edge lists, no gather/scatter, no PyG semantics -- against
(a) `oracle/port.py` (always) and (b) the reference's own source files running on the PyG restatement
(`oracle/load_reference.py`, only where /root/reference exists).  Three routes that share no code agreeing
to ~1e-6 is what pins the restated PyG semantics (scatter direction, coalesce, self-loop rules) in the
absence of a real torch_geometric install.

    MagNet (Zhang et al. 2021):  A_s = (A + A^T)/2,  Theta = 2 pi q (A - A^T),  H = A_s .* exp(i Theta),
        L = I - D_s^-1/2 H D_s^-1/2 ('sym')  or  D_s - H (None);   L~ = 2 L / lambda_max - I
        The reference aggregates source -> target, i.e. applies L~^T, and collapses the complex
        product to  out_real = A - B + b, out_imag = A + B + b  with  A = sum_k T_k(L~_r^T) x_r W_k,
        B = sum_k T_k(L~_i^T) x_i W_k   -- restated here with dense Chebyshev recurrences.
"""
import numpy as np
import pytest
import torch

from oracle import load_reference, port


def _dense_adj(ei, ew, n):
    a = np.zeros((n, n))
    w = np.ones(ei.shape[1]) if ew is None else ew.double().numpy()
    keep = ei[0] != ei[1]                                      # self-loops are removed first
    np.add.at(a, (ei[0][keep].numpy(), ei[1][keep].numpy()), w[keep.numpy()])   # duplicates add up
    return a


def _dense_magnet(xr, xi, ei, ew, weight, bias, q, normalization, lambda_max):
    n = xr.shape[0]
    a = _dense_adj(ei, ew, n)
    a_s = (a + a.T) / 2
    theta = 2 * np.pi * q * (a - a.T)
    h = a_s * np.exp(1j * theta)
    deg = a_s.sum(1)
    if normalization == "sym":
        with np.errstate(divide="ignore"):
            dm = np.where(deg > 0, deg ** -0.5, 0.0)
        lap = np.eye(n) - dm[:, None] * h * dm[None, :]
    else:
        lap = np.diag(deg) - h
    lt = 2 * lap / lambda_max - np.eye(n)
    lr, li = lt.real.T, lt.imag.T                              # source -> target aggregation = transpose
    w = weight.double().numpy()

    def chain(op, x):
        t0 = x
        acc = t0 @ w[0]
        if w.shape[0] > 1:
            t1 = op @ x
            acc = acc + t1 @ w[1]
            for k in range(2, w.shape[0]):
                t2 = 2 * (op @ t1) - t0
                acc = acc + t2 @ w[k]
                t0, t1 = t1, t2
        return acc
    A, B = chain(lr, xr.double().numpy()), chain(li, xi.double().numpy())
    b = bias.double().numpy()
    return A - B + b, A + B + b


CASES = [
    # n, e, fin, fout, K, q, normalization, weighted, lambda_max
    (60, 400, 5, 4, 1, 0.25, "sym", False, 2.0),
    (80, 700, 3, 6, 2, 0.1, "sym", True, 2.0),
    (50, 300, 4, 4, 3, 0.2, None, True, 9.0),
    (40, 120, 2, 3, 2, 0.0, "sym", False, 1.7),        # q = 0: purely real operator, lambda_max != 2
]


@pytest.mark.parametrize("n,e,fin,fout,K,q,norm,weighted,lmax", CASES)
def test_magnet_three_routes_agree(n, e, fin, fout, K, q, norm, weighted, lmax):
    g = torch.Generator().manual_seed(n * 7 + e)
    ei = torch.randint(0, n - 3, (2, e), generator=g)           # last 3 nodes isolated (deg 0 -> inf -> 0)
    ei[:, :20] = ei[:, 20:40]                                    # duplicates
    ei[:, 40:60] = ei[:, 60:80].flip(0)                          # reciprocal pairs
    ei[1, 80:90] = ei[0, 80:90]                                  # self-loops
    ew = (torch.rand(e, generator=g) + 0.5) if weighted else None
    xr = torch.rand(n, fin, generator=g) * 2 - 1
    xi = torch.rand(n, fin, generator=g) * 2 - 1
    weight = torch.rand(K + 1, fin, fout, generator=g) - 0.5
    bias = torch.rand(fout, generator=g) - 0.5
    d_r, d_i = _dense_magnet(xr, xi, ei, ew, weight, bias, q, norm, lmax)
    scale = max(np.abs(d_r).max(), np.abs(d_i).max())
    p_r, p_i = port.magnet_conv(xr, xi, ei, ew, weight, bias, q, norm, lambda_max=lmax)
    assert np.abs(p_r.double().numpy() - d_r).max() <= 2e-6 * scale
    assert np.abs(p_i.double().numpy() - d_i).max() <= 2e-6 * scale
    if load_reference.available():
        conv = load_reference.ref_classes()["MagNetConv"](fin, fout, K=K, q=q, trainable_q=False, normalization=norm)
        with torch.no_grad():
            conv.weight.copy_(weight)
            conv.bias.copy_(bias)
            r_r, r_i = conv(xr, xi, ei, ew, lambda_max=torch.tensor(lmax))
        assert np.abs(r_r.double().numpy() - d_r).max() <= 2e-6 * scale
        assert np.abs(r_i.double().numpy() - d_i).max() <= 2e-6 * scale


def test_row_normalised_and_mean_aggregations_dense():
    """Conv_Base = D^-1 (A + tau I) x over remaining self-loops (conv_base.py:12-31,98-117, aggregation at
    edge_index[0] over sources edge_index[1]); SGCN's mean = (in-neighbour sum at edge_index[1]) / max(count, 1);
    DiGCNConv = scatter of w * (xW)[src] at dst + b."""
    g = torch.Generator().manual_seed(5)
    n, e, f = 70, 500, 6
    ei = torch.randint(0, n, (2, e), generator=g)
    ew = torch.rand(e, generator=g) + 0.1
    x = torch.rand(n, f, generator=g) * 2 - 1
    # Conv_Base
    a = np.zeros((n, n))
    loops = ei[0] == ei[1]
    np.add.at(a, (ei[0][~loops].numpy(), ei[1][~loops].numpy()), ew[~loops].double().numpy())
    diag = np.full(n, 0.5)
    for r, w in zip(ei[0][loops].tolist(), ew[loops].double().tolist()):
        diag[r] = w                                              # an existing loop keeps its weight (last wins)
    a_hat = a + np.diag(diag)
    ref = (a_hat / a_hat.sum(1, keepdims=True)) @ x.double().numpy()
    got = port.conv_base(x, ei, ew, 0.5).double().numpy()
    assert np.abs(got - ref).max() <= 2e-6 * np.abs(ref).max()
    # mean aggregation (SGCNConv first_aggr): lin_b([mean_pos(x), x]) || lin_u([mean_neg(x), x])
    pos, neg = ei[:, : e // 2], ei[:, e // 2:]
    out = 4
    wb, wu = torch.rand(out, 2 * f, generator=g) - 0.5, torch.rand(out, 2 * f, generator=g) - 0.5
    bb, bu = torch.rand(out, generator=g), torch.rand(out, generator=g)

    def mean_in(edges):
        m = np.zeros((n, n))
        np.add.at(m, (edges[1].numpy(), edges[0].numpy()), 1.0)
        cnt = np.maximum(m.sum(1, keepdims=True), 1.0)
        return (m / cnt) @ x.double().numpy()
    xd = x.double().numpy()
    ref = np.concatenate([np.concatenate([mean_in(pos), xd], 1) @ wb.double().numpy().T + bb.double().numpy(),
                          np.concatenate([mean_in(neg), xd], 1) @ wu.double().numpy().T + bu.double().numpy()], 1)
    got = port.sgcn_conv(x, pos, neg, wb, bb, wu, bu, first_aggr=True, norm_emb=False).double().numpy()
    assert np.abs(got - ref).max() <= 2e-6 * np.abs(ref).max()
    # DiGCNConv
    w = torch.rand(f, 5, generator=g) - 0.5
    b = torch.rand(5, generator=g)
    m = np.zeros((n, n))
    np.add.at(m, (ei[1].numpy(), ei[0].numpy()), ew.double().numpy())
    ref = m @ (xd @ w.double().numpy()) + b.double().numpy()
    got = port.digcn_conv(x, ei, ew, w, b).double().numpy()
    assert np.abs(got - ref).max() <= 2e-6 * np.abs(ref).max()


def test_dimpa_and_dgcn_dense():
    """DIMPA (nn/directed/DIMPA.py:32-59): feat_s = sum_h w_s[h] P^h x_s with P = D^-1 (A + tau I) row-normalised over
    the remaining self-loops, feat_t the same with A^T.  DGCNConv (nn/directed/DGCNConv.py:38-103 + PyG gcn_norm):
    out = D^-1/2 (A^T + I) D^-1/2 x with D the in-degree of (A + I) -- aggregation at edge_index[1]."""
    g = torch.Generator().manual_seed(9)
    n, e, f, hop = 60, 420, 5, 3
    ei = torch.randint(0, n, (2, e), generator=g)
    ei = ei[:, ei[0] != ei[1]]                                    # loop-free (the loop rules are checked above)
    ew = torch.rand(ei.size(1), generator=g) + 0.2
    xs = torch.rand(n, f, generator=g) * 2 - 1
    xt = torch.rand(n, f, generator=g) * 2 - 1
    w_s, w_t = torch.rand(hop + 1, 1, generator=g), torch.rand(hop + 1, 1, generator=g)
    a = np.zeros((n, n))
    np.add.at(a, (ei[0].numpy(), ei[1].numpy()), ew.double().numpy())

    def rw(m):
        m = m + 0.5 * np.eye(n)
        return m / m.sum(1, keepdims=True)
    ps, pt = rw(a), rw(a.T)
    fs, ft = w_s[0].item() * xs.double().numpy(), w_t[0].item() * xt.double().numpy()
    cs, ct = xs.double().numpy(), xt.double().numpy()
    for h in range(1, hop + 1):
        cs, ct = ps @ cs, pt @ ct
        fs, ft = fs + w_s[h].item() * cs, ft + w_t[h].item() * ct
    ref = np.concatenate([fs, ft], 1)
    got = port.dimpa(xs, xt, ei, ew, w_s, w_t, hop, 0.5).double().numpy()
    assert np.abs(got - ref).max() <= 2e-6 * np.abs(ref).max()
    # DGCNConv
    ah = a + np.eye(n)
    deg = ah.sum(0)                                               # weighted in-degree (target = edge_index[1])
    dis = deg ** -0.5
    ref = (dis[:, None] * ah.T * dis[None, :]) @ xs.double().numpy()
    got = port.dgcn_conv(xs, ei, ew).double().numpy()
    assert np.abs(got - ref).max() <= 2e-6 * np.abs(ref).max()


def test_attention_layers_dense():
    """GATConv (heads = 1): out_i = sum_j softmax_j(leaky_relu(a_s . h_j + a_d . h_i)) h_j + b over j in N_in(i) + {i};
    SNEAConv first layer (nn/signed/SNEAConv.py:81-146): out_i = h_i * sum_j alpha_ij = h_i wherever node i has an
    incoming entry (the message is the TARGET's feature, quirk Q7, and alpha sums to one), else 0."""
    g = torch.Generator().manual_seed(13)
    n, e, c = 50, 300, 6
    ei = torch.randint(0, n - 2, (2, e), generator=g)
    x = torch.randn(n, c, generator=g)
    lin_w = torch.randn(c, c, generator=g) / c ** 0.5
    a_s, a_d, b = torch.randn(c, generator=g), torch.randn(c, generator=g), torch.randn(c, generator=g)
    h = x.double().numpy() @ lin_w.double().numpy().T
    adj = np.zeros((n, n), dtype=bool)                            # adj[i, j]: j -> i
    keep = ei[0] != ei[1]
    adj[ei[1][keep].numpy(), ei[0][keep].numpy()] = True
    mult = np.zeros((n, n))                                        # duplicate edges count twice in the softmax
    np.add.at(mult, (ei[1][keep].numpy(), ei[0][keep].numpy()), 1.0)
    mult += np.eye(n)
    s = (h @ a_s.double().numpy())[None, :] + (h @ a_d.double().numpy())[:, None]
    s = np.where(s > 0, s, 0.2 * s)
    w = mult * np.exp(s - s.max())
    alpha = w / w.sum(1, keepdims=True)
    ref = alpha @ h + b.double().numpy()
    got = port.gat_conv(x, ei, lin_w, a_s, a_d, b).double().numpy()
    assert np.abs(got - ref).max() <= 2e-6 * np.abs(ref).max()
    # SNEAConv, first layer
    out = 4
    pos, neg = ei[:, :150], ei[:, 150:]
    wb, wu = torch.randn(out, c, generator=g), torch.randn(out, c, generator=g)
    bb, bu = torch.randn(out, generator=g), torch.randn(out, generator=g)
    ab_w, au_w = torch.randn(1, 2 * out, generator=g), torch.randn(1, 2 * out, generator=g)
    ab_b, au_b = torch.randn(1, generator=g), torch.randn(1, generator=g)
    got = port.snea_conv(x, pos, neg, wb, bb, wu, bu, ab_w, ab_b, au_w, au_b, first_aggr=True).double().numpy()
    for half, (edges, w_, b_) in enumerate(((pos, wb, bb), (neg, wu, bu))):
        hh = x.double().numpy() @ w_.double().numpy().T + b_.double().numpy()
        k = edges[:, edges[0] != edges[1]]
        m = int(k.max()) + 1 if k.numel() else 0                  # self-loops re-added for nodes 0..max id only
        has_in = np.zeros(n, dtype=bool)
        has_in[:m] = True
        has_in[k[1].numpy()] = True
        ref = np.where(has_in[:, None], hh, 0.0)
        blk = got[:, half * out:(half + 1) * out]
        assert np.abs(blk - ref).max() <= 2e-6 * max(np.abs(ref).max(), 1e-30)

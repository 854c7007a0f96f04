"""GPU: gradients of the drop-in layers (autograd.py: backward aggregation through the same /
transposed plans, dX through the dense kernels, dW/db through pgsd_xtg_accumulate) against torch
autograd of the CPU oracle (whose ops are the reference's: index_select / scatter_add_ / matmul).
Tolerance 2e-5 * max|ref| per gradient tensor (fp32, atomics change the summation order)."""
import pytest
import torch

from conftest import assert_close_rel
from oracle import port
from pytorch_geometric_signed_directed_b200 import nn, synthetic

pytestmark = pytest.mark.gpu
DEV = "cuda"
TOL = 2e-5


def _leaf(t, dev=None):
    t = t.detach().clone().to(dev) if dev else t.detach().clone()
    return t.requires_grad_(True)


@pytest.mark.parametrize("K,fin,fout,relu", [(1, 64, 64, False), (2, 32, 16, False), (3, 8, 5, False),
                                             (1, 64, 64, True)])
def test_magnet_gradients(K, fin, fout, relu):
    g = torch.Generator().manual_seed(K * 100 + fin)
    n, e = 3000, 40_000
    ei = torch.randint(0, n, (2, e), generator=g)
    ew = torch.rand(e, generator=g) + 0.5
    xr, xi = torch.rand(n, fin, generator=g) * 2 - 1, torch.rand(n, fin, generator=g) * 2 - 1
    r1, r2 = torch.randn(n, fout, generator=g), torch.randn(n, fout, generator=g)
    conv = nn.MagNetConv(fin, fout, K=K, q=0.2, trainable_q=False, cached=True).to(DEV)
    conv.fused_complex_relu = relu
    with torch.no_grad():
        conv.bias.uniform_(-0.3, 0.3)
    # oracle
    ca, cb = _leaf(xr), _leaf(xi)
    w, bias = _leaf(conv.weight.cpu()), _leaf(conv.bias.cpu())
    p_r, p_i = port.magnet_conv(ca, cb, ei, ew, w, bias, 0.2, "sym")
    if relu:
        # the ReLU mask is a step function of out_real: where the pre-activation is within rounding of 0
        # the two implementations may legitimately pick different sides, so no gradient is sent there
        edge = p_r.detach().abs() < 1e-4
        r1, r2 = r1.masked_fill(edge, 0.0), r2.masked_fill(edge, 0.0)
        p_r, p_i = port.complex_relu(p_r, p_i)
    ((p_r * r1).sum() + (p_i * r2).sum()).backward()
    a, b = _leaf(xr, DEV), _leaf(xi, DEV)
    o_r, o_i = conv(a, b, ei.to(DEV), ew.to(DEV))
    ((o_r * r1.to(DEV)).sum() + (o_i * r2.to(DEV)).sum()).backward()
    assert_close_rel(o_r, p_r, 1e-5, "forward real")
    assert_close_rel(a.grad, ca.grad, TOL, "d x_real")
    assert_close_rel(b.grad, cb.grad, TOL, "d x_imag")
    assert_close_rel(conv.weight.grad, w.grad, TOL, "d weight")
    assert_close_rel(conv.bias.grad, bias.grad, TOL, "d bias")


def test_digcn_inception_gradients():
    g = torch.Generator().manual_seed(4)
    n, e, f = 2500, 30_000, 32
    ei1, ei2 = torch.randint(0, n, (2, e), generator=g), torch.randint(0, n, (2, e), generator=g)
    w1, w2 = synthetic.sym_norm_weights(ei1, n), synthetic.sym_norm_weights(ei2, n)
    x = torch.rand(n, f, generator=g) * 2 - 1
    rs = [torch.randn(n, f, generator=g) for _ in range(3)]
    blk = nn.DiGCN_InceptionBlock(f, f).to(DEV)
    with torch.no_grad():
        blk.conv1.bias.uniform_(-0.3, 0.3)
    a = _leaf(x, DEV)
    outs = blk(a, ei1.to(DEV), w1.to(DEV), ei2.to(DEV), w2.to(DEV))
    sum((o * r.to(DEV)).sum() for o, r in zip(outs, rs)).backward()
    ca = _leaf(x)
    prm = {k: _leaf(v.cpu()) for k, v in blk.state_dict().items()}
    ref = port.digcn_inception_block(ca, ei1, w1, ei2, w2, prm["ln.weight"], prm["ln.bias"], prm["conv1.weight"],
                                     prm["conv1.bias"], prm["conv2.weight"], prm["conv2.bias"])
    sum((o * r).sum() for o, r in zip(ref, rs)).backward()
    assert_close_rel(a.grad, ca.grad, TOL, "d x")
    for name, prm_t in blk.named_parameters():
        assert_close_rel(prm_t.grad, prm[name].grad, TOL, f"d {name}")


def test_sgcn_two_layer_gradients():
    n = 3000
    pos, neg, _ = synthetic.ssbm_edges(n, 3, num_entries=50_000, eta=0.1, seed=8)
    g = torch.Generator().manual_seed(5)
    x = torch.randn(n, 16, generator=g)
    r = torch.randn(n, 16, generator=g)
    c1 = nn.SGCNConv(16, 8, first_aggr=True).to(DEV)
    c2 = nn.SGCNConv(8, 8, first_aggr=False).to(DEV)
    a = _leaf(x, DEV)
    z = c2(torch.tanh(c1(a, pos.to(DEV), neg.to(DEV))), pos.to(DEV), neg.to(DEV))
    (z * r.to(DEV)).sum().backward()
    ca = _leaf(x)
    p1 = [_leaf(t.cpu()) for t in (c1.lin_b.weight, c1.lin_b.bias, c1.lin_u.weight, c1.lin_u.bias)]
    p2 = [_leaf(t.cpu()) for t in (c2.lin_b.weight, c2.lin_b.bias, c2.lin_u.weight, c2.lin_u.bias)]
    zz = port.sgcn_conv(torch.tanh(port.sgcn_conv(ca, pos, neg, *p1, True)), pos, neg, *p2, False)
    (zz * r).sum().backward()
    assert_close_rel(z, zz, 1e-5, "forward")
    assert_close_rel(a.grad, ca.grad, TOL, "d x")
    for got, ref in zip((c1.lin_b.weight, c1.lin_b.bias, c1.lin_u.weight, c1.lin_u.bias), p1):
        assert_close_rel(got.grad, ref.grad, TOL, "d layer-1 params")
    for got, ref in zip((c2.lin_b.weight, c2.lin_b.bias, c2.lin_u.weight, c2.lin_u.bias), p2):
        assert_close_rel(got.grad, ref.grad, TOL, "d layer-2 params")


def test_dimpa_gradients():
    g = torch.Generator().manual_seed(6)
    n, e, f = 2000, 24_000, 16
    ei = torch.randint(0, n, (2, e), generator=g)
    ew = torch.rand(e, generator=g) + 0.2
    xs, xt = torch.rand(n, f, generator=g), torch.rand(n, f, generator=g)
    r = torch.randn(n, 2 * f, generator=g)
    dm = nn.DIMPA(hop=2).to(DEV)
    with torch.no_grad():
        dm._w_s.copy_(torch.tensor([[1.0], [0.6], [-0.4]])); dm._w_t.copy_(torch.tensor([[0.3], [1.1], [0.5]]))
    a, b = _leaf(xs, DEV), _leaf(xt, DEV)
    (dm(a, b, ei.to(DEV), ew.to(DEV)) * r.to(DEV)).sum().backward()
    ca, cb = _leaf(xs), _leaf(xt)
    ws, wt = _leaf(dm._w_s.cpu()), _leaf(dm._w_t.cpu())
    (port.dimpa(ca, cb, ei, ew, ws, wt, 2) * r).sum().backward()
    assert_close_rel(a.grad, ca.grad, TOL, "d x_s")
    assert_close_rel(b.grad, cb.grad, TOL, "d x_t")
    assert_close_rel(dm._w_s.grad, ws.grad, TOL, "d w_s")
    assert_close_rel(dm._w_t.grad, wt.grad, TOL, "d w_t")


def test_training_step_reduces_loss():
    """examples/magnet_node.py:22-29 in miniature: two MagNetConv layers + Adam, loss goes down."""
    g = torch.Generator().manual_seed(9)
    n = 2000
    ei, labels = synthetic.dsbm_edges(n, 3, num_edges=30_000, seed=3)
    x = torch.rand(n, 16, generator=g).to(DEV)
    ei, y = ei.to(DEV), labels.to(DEV)
    c1 = nn.MagNetConv(16, 16, K=1, q=0.25, trainable_q=False, cached=True).to(DEV)
    c2 = nn.MagNetConv(16, 16, K=1, q=0.25, trainable_q=False, cached=True).to(DEV)
    head = torch.nn.Linear(32, 3).to(DEV)
    opt = torch.optim.Adam(list(c1.parameters()) + list(c2.parameters()) + list(head.parameters()), lr=0.01)
    relu = nn.complex_relu_layer()
    losses = []
    for _ in range(30):
        opt.zero_grad()
        r, i = relu(*c1(x, x, ei))
        r, i = relu(*c2(r, i, ei))
        loss = torch.nn.functional.cross_entropy(head(torch.cat([r, i], 1)), y)
        loss.backward()
        opt.step()
        losses.append(loss.item())
    assert losses[-1] < 0.8 * losses[0], losses[::5]


@pytest.mark.parametrize("name,cls,norm", [("qgrad_magnet_k2", "MagNetConv", "sym"),
                                           ("qgrad_magnet_k1_none", "MagNetConv", None),
                                           ("qgrad_msconv_k3", "MSConv", "sym"),
                                           ("qgrad_magnet_clamped", "MagNetConv", "sym")])
def test_trainable_q_gradient_golden(name, cls, norm):
    """d loss / d q of a trainable magnetic charge (pgsd_magnetic_q_grad) against the reference's
    own autograd (tests/golden/make_golden_qgrad.py), together with the other gradients."""
    from conftest import load_golden
    g = load_golden(name, DEV)
    K = g["weight"].size(0) - 1
    q0 = {"qgrad_magnet_clamped": 0.4}.get(name, float(g["q"]))
    conv = getattr(nn, cls)(g["weight"].size(1), g["weight"].size(2), K=K, q=q0, trainable_q=True,
                            normalization=norm).to(DEV)
    with torch.no_grad():
        conv.weight.copy_(g["weight"])
        conv.bias.copy_(g["bias"])
    xr, xi = _leaf(g["x_real"]), _leaf(g["x_imag"])
    lam = float(g["lambda_max"])
    o_r, o_i = conv(xr, xi, g["edge_index"], g["edge_weight"], lambda_max=None if lam < 0 else lam)
    ((o_r * g["r1"]).sum() + (o_i * g["r2"]).sum()).backward()
    assert_close_rel(o_r, g["out_real"], 1e-5, "forward real")
    assert_close_rel(o_i, g["out_imag"], 1e-5, "forward imag")
    assert isinstance(conv.q, torch.nn.Parameter) and conv.q.shape == (1,)
    assert_close_rel(conv.q.detach(), g["q"], 1e-7, "clamped q")          # quirk Q9
    assert_close_rel(conv.q.grad, g["d_q"], 1e-4, "d q")
    assert_close_rel(conv.weight.grad, g["d_weight"], TOL, "d weight")
    assert_close_rel(xr.grad, g["d_x_real"], TOL, "d x_real")
    assert_close_rel(xi.grad, g["d_x_imag"], TOL, "d x_imag")


def test_trainable_q_gradient_wide_rows():
    """q gradient at a width that needs two feature chunks per lane group and at an odd width (scalar path)."""
    for f in (160, 7):
        gen = torch.Generator().manual_seed(f)
        n, e = 1200, 15_000
        ei = torch.randint(0, n, (2, e), generator=gen)
        ew = torch.rand(e, generator=gen) + 0.5
        xr, xi = torch.rand(n, f, generator=gen) * 2 - 1, torch.rand(n, f, generator=gen) * 2 - 1
        r1, r2 = torch.randn(n, 4, generator=gen), torch.randn(n, 4, generator=gen)
        conv = nn.MagNetConv(f, 4, K=2, q=0.17, trainable_q=True).to(DEV)
        o_r, o_i = conv(xr.to(DEV), xi.to(DEV), ei.to(DEV), ew.to(DEV))
        ((o_r * r1.to(DEV)).sum() + (o_i * r2.to(DEV)).sum()).backward()
        q = torch.tensor([0.17], requires_grad=True)
        p_r, p_i = port.magnet_conv(xr, xi, ei, ew, conv.weight.detach().cpu(), conv.bias.detach().cpu(), q, "sym")
        ((p_r * r1).sum() + (p_i * r2).sum()).backward()
        assert_close_rel(conv.q.grad, q.grad, 2e-4, f"d q, F={f}")


# ------------------------------------------------------------------ attention layers
def _snea_params(m):
    return [m.lin_b.weight, m.lin_b.bias, m.lin_u.weight, m.lin_u.bias,
            m.alpha_b.weight, m.alpha_b.bias, m.alpha_u.weight, m.alpha_u.bias]


@pytest.mark.parametrize("fin,fout", [(64, 32), (12, 8), (6, 5)])
def test_snea_two_layer_gradients(fin, fout):
    """Training through SNEAConv (nn/signed/SNEAConv.py:81-146): gradients of a two-layer stack w.r.t. the input and
    every parameter (lin_b / lin_u / alpha_b / alpha_u) against torch autograd of the oracle."""
    n = 3000
    pos, neg, _ = synthetic.ssbm_edges(n, 3, num_entries=50_000, eta=0.1, seed=7)
    g = torch.Generator().manual_seed(fin)
    x = torch.randn(n, fin, generator=g)
    r = torch.randn(n, 2 * fout, generator=g)
    torch.manual_seed(fin + 1)
    c1 = nn.SNEAConv(fin, fout, first_aggr=True).to(DEV)
    c2 = nn.SNEAConv(fout, fout, first_aggr=False).to(DEV)
    w1 = [_leaf(t.cpu()) for t in _snea_params(c1)]
    w2 = [_leaf(t.cpu()) for t in _snea_params(c2)]
    cx = _leaf(x)
    ref = port.snea_conv(torch.tanh(port.snea_conv(cx, pos, neg, *w1, True)), pos, neg, *w2, False)
    (ref * r).sum().backward()
    dx = _leaf(x, DEV)
    out = c2(torch.tanh(c1(dx, pos.to(DEV), neg.to(DEV))), pos.to(DEV), neg.to(DEV))
    (out * r.to(DEV)).sum().backward()
    assert_close_rel(out, ref, 1e-5, "forward")
    assert_close_rel(dx.grad, cx.grad, TOL, "d x")
    names = ["lin_b.weight", "lin_b.bias", "lin_u.weight", "lin_u.bias", "alpha_b.weight", "alpha_b.bias",
             "alpha_u.weight", "alpha_u.bias"]
    for layer, conv, ws in ((1, c1, w1), (2, c2, w2)):
        for nm, p, w in zip(names, _snea_params(conv), ws):
            if layer == 1 and nm.startswith("alpha"):
                # first layer: one edge type per softmax, the weights of a row sum to 1 -> zero gradient (both sides)
                assert float(p.grad.abs().max()) <= 1e-3 and float(w.grad.abs().max()) <= 1e-3
                continue
            assert_close_rel(p.grad, w.grad, TOL, f"layer {layer} d {nm}")


@pytest.mark.parametrize("n,e,c", [(3000, 40_000, 64), (2000, 15_000, 12), (700, 5_000, 10)])
def test_gat_conv_gradients(n, e, c):
    """Training through the GATConv of SDGNN / SiGAT (nn/signed/SDGNN.py:35-41): edge-softmax backward + SDDMM +
    transposed weighted aggregation against torch autograd of the oracle's GATConv."""
    g = torch.Generator().manual_seed(n + c)
    ei = torch.randint(0, n - 5, (2, e), generator=g)
    ei[1, :300] = 2
    x = torch.randn(n, c, generator=g)
    r = torch.randn(n, c, generator=g)
    gat = nn.GATConv(c, c).to(DEV)
    with torch.no_grad():
        gat.bias.uniform_(-0.3, 0.3)
    prm = [gat.lin.weight, gat.att_src, gat.att_dst, gat.bias]
    ws = [_leaf(t.cpu()) for t in prm]
    cx = _leaf(x)
    ref = port.gat_conv(cx, ei, ws[0], ws[1].view(-1), ws[2].view(-1), ws[3])
    (ref * r).sum().backward()
    dx = _leaf(x, DEV)
    out = gat(dx, ei.to(DEV))
    (out * r.to(DEV)).sum().backward()
    assert_close_rel(out, ref, 1e-5, "forward")
    assert_close_rel(dx.grad, cx.grad, TOL, "d x")
    for nm, p, w in zip(("lin.weight", "att_src", "att_dst", "bias"), prm, ws):
        assert_close_rel(p.grad, w.grad.view_as(p.grad), TOL, f"d {nm}")


def test_sdr_layer_gradients_and_training_step():
    """SDRLayer (4 GATConv + MLP, nn/signed/SDGNN.py:57-64): gradients against the oracle, and an optimiser step
    changes the attention parameters (the layers used to return detached tensors)."""
    gen = torch.Generator().manual_seed(21)
    n, c = 1500, 12
    lists = [torch.randint(0, n, (2, 9000), generator=gen) for _ in range(4)]
    x = torch.randn(n, c, generator=gen)
    r = torch.randn(n, c, generator=gen)
    layer = nn.SDRLayer(c, c, [e.to(DEV) for e in lists]).to(DEV)
    gp = [[_leaf(t.cpu()) for t in (a.lin.weight, a.att_src, a.att_dst, a.bias)] for a in layer.aggs]
    l0, l2 = layer.mlp_layer[0], layer.mlp_layer[2]
    mp = [_leaf(t.cpu()) for t in (l0.weight, l0.bias, l2.weight, l2.bias)]
    cx = _leaf(x)
    ref = port.sdr_layer(cx, lists, [(w, a.view(-1), b.view(-1), bb) for w, a, b, bb in gp], *mp)
    (ref * r).sum().backward()
    dx = _leaf(x, DEV)
    out = layer(dx)
    assert out.requires_grad
    (out * r.to(DEV)).sum().backward()
    assert_close_rel(out, ref, 1e-5, "forward")
    assert_close_rel(dx.grad, cx.grad, TOL, "d x")
    for k, (a, ws) in enumerate(zip(layer.aggs, gp)):
        for nm, p, w in zip(("lin.weight", "att_src", "att_dst", "bias"), (a.lin.weight, a.att_src, a.att_dst, a.bias), ws):
            assert_close_rel(p.grad, w.grad.view_as(p.grad), TOL, f"agg_{k} d {nm}")
    for nm, p, w in zip(("mlp0.weight", "mlp0.bias", "mlp2.weight", "mlp2.bias"), (l0.weight, l0.bias, l2.weight, l2.bias), mp):
        assert_close_rel(p.grad, w.grad, TOL, f"d {nm}")
    before = layer.aggs[0].att_src.detach().clone()
    torch.optim.SGD(layer.parameters(), lr=0.1).step()
    assert not torch.equal(before, layer.aggs[0].att_src.detach())

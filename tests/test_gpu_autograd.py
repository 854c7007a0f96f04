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

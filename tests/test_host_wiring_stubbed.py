"""CPU-only: the HOST-side wiring of SGCNConv (first layer folded through the mean aggregations, fused tanh, sparse
adjacency inputs, autograd route, deep layer) with the three C-ABI calls it makes replaced by torch stand-ins (checker
only -- the product has no CPU path).  Catches plumbing mistakes (column slices, weight blocks, bias placement, plan
lookups) without a GPU; the kernels themselves are covered by the `-m gpu` parity tests."""
import os

import pytest
import torch

from oracle import port
from pytorch_geometric_signed_directed_b200 import nn, ops, plan as _plan
from pytorch_geometric_signed_directed_b200.plan import CSRPlan


def _build_csr(ei, ew, n_dst, n_src, flow="source_to_target"):
    src, dst = (ei[0], ei[1]) if flow == "source_to_target" else (ei[1], ei[0])
    order = torch.sort(dst, stable=True).indices
    rp = torch.zeros(n_dst + 1, dtype=torch.int32)
    rp[1:] = torch.cumsum(torch.bincount(dst, minlength=n_dst), 0).int()
    return CSRPlan(n_dst, n_src, ei.size(1), ei.size(1), rp, src[order].int(),
                   [None if ew is None else ew[order]], [None], [0.0])


def _spmm(plan, xs, op_ids=(0,), *, mean=False, alpha=1.0, beta=0.0, zs=None, bias=None, out=None, variant=None,
          op_scale=None, grid_reserve=0, tanh_out=False):
    counts = (plan.row_ptr[1:] - plan.row_ptr[:-1]).long()
    rows = torch.repeat_interleave(torch.arange(plan.n_dst), counts)
    ys = []
    for k, op in enumerate(op_ids):
        v, x = plan.val[op], xs[k]
        msg = x[plan.col.long()] * (1.0 if v is None else v.view(-1, 1))
        agg = torch.zeros(plan.n_dst, x.size(1)).index_add_(0, rows, msg)
        if mean:
            agg = agg / counts.clamp(min=1).view(-1, 1)
        y = alpha * agg
        if zs is not None and zs[k] is not None:
            y = y + beta * zs[k]
        if bias is not None:
            y = y + bias
        if tanh_out:
            y = torch.tanh(y)
        if out is not None and out[k] is not None:
            out[k].copy_(y)
            y = out[k]
        ys.append(y)
    return ys


def _dense(terms, n_out, *, bias=None, combine=False, relu_mode=0, out=None, variant=None):
    acc = sum(x.float() @ w.float() for x, w, g in terms)
    if bias is not None:
        acc = acc + bias
    return [torch.tanh(acc) if relu_mode == 2 else acc]


@pytest.fixture
def stubbed(monkeypatch):
    monkeypatch.setattr(_plan, "build_csr", _build_csr)
    monkeypatch.setattr(_plan, "require_cuda", lambda t, name: None)
    monkeypatch.setattr(ops, "spmm", _spmm)
    monkeypatch.setattr(ops, "dense", _dense)
    monkeypatch.setattr(CSRPlan, "hub_rows", lambda self: None)


@pytest.mark.parametrize("fused", [False, True])
def test_sgcn_first_layer_fold_and_sparse_inputs(stubbed, fused):
    g = torch.Generator().manual_seed(0)
    n = 300
    pos, neg = torch.randint(0, n, (2, 2000), generator=g), torch.randint(0, n, (2, 1500), generator=g)
    x = torch.randn(n, 16, generator=g)
    torch.manual_seed(10)
    conv = nn.SGCNConv(16, 8, first_aggr=True)
    conv.fused_tanh = fused
    prm = (conv.lin_b.weight, conv.lin_b.bias, conv.lin_u.weight, conv.lin_u.bias)
    with torch.no_grad():
        ref = port.sgcn_conv(x, pos, neg, *prm, True)
        ref = torch.tanh(ref) if fused else ref
        y = conv(x, pos, neg)                                  # inference route: one transform + two half-width aggregations
        assert (y - ref).abs().max() <= 2e-6 * ref.abs().max()
        adj = lambda e: torch.sparse_coo_tensor(torch.stack([e[1], e[0]]), torch.ones(e.size(1)), (n, n))
        assert torch.equal(conv(x, adj(pos), adj(neg)), y)     # SGCNConv.py:131-134: adj_t[target, source]
    xg = x.clone().requires_grad_(True)
    yg = conv(xg, pos, neg)                                    # autograd route: same numbers, differentiable
    assert yg.requires_grad and (yg.detach() - ref).abs().max() <= 2e-6 * ref.abs().max()


def test_sgcn_deep_layer_wiring(stubbed):
    g = torch.Generator().manual_seed(1)
    n = 200
    pos, neg = torch.randint(0, n, (2, 1200), generator=g), torch.randint(0, n, (2, 900), generator=g)
    z = torch.randn(n, 16, generator=g)
    torch.manual_seed(11)
    conv = nn.SGCNConv(8, 8, first_aggr=False)
    with torch.no_grad():
        y = conv(z, pos, neg)
        ref = port.sgcn_conv(z, pos, neg, conv.lin_b.weight, conv.lin_b.bias, conv.lin_u.weight, conv.lin_u.bias, False)
    assert (y - ref).abs().max() <= 2e-6 * ref.abs().max()


# ---- attention layers: the same idea for SNEAConv / GATConv / SDRLayer (inference route and autograd route) -----------
def _rows_of(plan):
    counts = (plan.row_ptr[1:] - plan.row_ptr[:-1]).long()
    return torch.repeat_interleave(torch.arange(plan.n_dst), counts)


def _act(pre, act, slope):
    return torch.tanh(pre) if act == "tanh" else torch.nn.functional.leaky_relu(pre, slope)


def _softmax_parts(plans, s_src, s_dst, act, slope):
    """alpha per entry and type: softmax over ALL entries of a row (both types), PyG's 1e-16 in the denominator."""
    n = plans[0].n_dst
    rows = [_rows_of(p) for p in plans]
    pre = [s_src[t].view(-1)[p.col[:p.nnz].long()] + s_dst[t].view(-1)[rows[t]] for t, p in enumerate(plans)]
    a = [_act(v, act, slope) for v in pre]
    mx = torch.full((n,), float("-inf"))
    for t in range(len(plans)):
        mx = mx.scatter_reduce(0, rows[t], a[t], "amax", include_self=True)
    mx = mx.masked_fill(mx == float("-inf"), 0)
    ex = [(a[t] - mx[rows[t]]).exp() for t in range(len(plans))]
    den = sum(torch.zeros(n).index_add_(0, rows[t], ex[t]) for t in range(len(plans))) + 1e-16
    return rows, pre, a, [ex[t] / den[rows[t]] for t in range(len(plans))]


def _edge_softmax(plans, s_src, s_dst, *, act="tanh", slope=0.2, xd=None, want_alpha=False):
    rows, _, _, alpha = _softmax_parts(plans, s_src, s_dst, act, slope)
    y = None
    if xd is not None:
        y = sum(xd[t] * torch.zeros(plans[0].n_dst).index_add_(0, rows[t], alpha[t]).view(-1, 1) for t in range(len(plans)))
    return y, (alpha if want_alpha else [])


def _edge_softmax_backward(plans, s_src, s_dst, *, act="tanh", slope=0.2, dalpha=None, row_coef=None, want_type_sum=False):
    n = plans[0].n_dst
    rows, pre, a, alpha = _softmax_parts(plans, s_src, s_dst, act, slope)
    da = [dalpha[t] if dalpha is not None else row_coef[t][rows[t]] for t in range(len(plans))]
    dot = sum(torch.zeros(n).index_add_(0, rows[t], alpha[t] * da[t]) for t in range(len(plans)))
    g_src, g_dst, sums = [], [], []
    for t, p in enumerate(plans):
        dact = (1 - a[t] ** 2) if act == "tanh" else torch.where(pre[t] > 0, 1.0, slope)
        dpre = alpha[t] * (da[t] - dot[rows[t]]) * dact
        g_src.append(torch.zeros(s_src[t].numel()).index_add_(0, p.col[:p.nnz].long(), dpre))
        g_dst.append(torch.zeros(s_dst[t].numel()).index_add_(0, rows[t], dpre))
        sums.append(torch.zeros(n).index_add_(0, rows[t], alpha[t]))
    return g_src, g_dst, (sums if want_type_sum else None)


def _sddmm_rows(plan, gy, h):
    return (gy[_rows_of(plan)] * h[plan.col[:plan.nnz].long()]).sum(1)


@pytest.fixture
def stubbed_attention(stubbed, monkeypatch):
    monkeypatch.setattr(ops, "edge_softmax", _edge_softmax)
    monkeypatch.setattr(ops, "edge_softmax_backward", _edge_softmax_backward)
    monkeypatch.setattr(ops, "sddmm_rows", _sddmm_rows)
    monkeypatch.setattr(ops, "gat_aggregate_supported", lambda *a, **k: False)
    monkeypatch.setattr(ops, "xtg_accumulate",
                        lambda x, g, dw, db=None: (dw.add_(x.float().t() @ g.float()),
                                                   db.add_(g.float().sum(0)) if db is not None else None))


def _grads(fn, leaves):
    for t in leaves:
        t.grad = None
    out = fn()
    return out


def test_snea_two_layers_inference_and_autograd_routes(stubbed_attention):
    g = torch.Generator().manual_seed(2)
    n = 120
    pos, neg = torch.randint(0, n - 3, (2, 700), generator=g), torch.randint(0, n - 3, (2, 500), generator=g)
    x = torch.randn(n, 8, generator=g)
    r = torch.randn(n, 12, generator=g)
    torch.manual_seed(int(os.environ.get("PGSD_TEST_SEED", "12")))
    c1, c2 = nn.SNEAConv(8, 6, first_aggr=True), nn.SNEAConv(6, 6, first_aggr=False)
    prm = lambda m: [m.lin_b.weight, m.lin_b.bias, m.lin_u.weight, m.lin_u.bias,
                     m.alpha_b.weight, m.alpha_b.bias, m.alpha_u.weight, m.alpha_u.bias]
    w1, w2 = [t.detach().clone().requires_grad_(True) for t in prm(c1)], [t.detach().clone().requires_grad_(True) for t in prm(c2)]
    cx = x.clone().requires_grad_(True)
    ref = port.snea_conv(torch.tanh(port.snea_conv(cx, pos, neg, *w1, True)), pos, neg, *w2, False)
    (ref * r).sum().backward()
    with torch.no_grad():
        y = c2(torch.tanh(c1(x, pos, neg)), pos, neg)                       # inference route
    assert (y - ref.detach()).abs().max() <= 1e-5 * ref.abs().max()
    dx = x.clone().requires_grad_(True)
    out = c2(torch.tanh(c1(dx, pos, neg)), pos, neg)                        # autograd route
    (out * r).sum().backward()
    assert (out.detach() - ref.detach()).abs().max() <= 1e-5 * ref.abs().max()
    assert (dx.grad - cx.grad).abs().max() <= 2e-5 * cx.grad.abs().max()
    for conv, ws in ((c1, w1), (c2, w2)):
        for p, w in zip(prm(conv), ws):
            scale = max(float(w.grad.abs().max()), 2e-2)   # first-layer attention gradients are rounding noise (~1e-7)
            assert (p.grad - w.grad).abs().max() <= 5e-5 * scale


def test_sdr_layer_inference_and_autograd_routes(stubbed_attention):
    g = torch.Generator().manual_seed(3)
    n, c = 90, 12
    lists = [torch.randint(0, n, (2, 400), generator=g) for _ in range(4)]
    x = torch.randn(n, c, generator=g)
    r = torch.randn(n, c, generator=g)
    torch.manual_seed(int(os.environ.get("PGSD_TEST_SEED", "13")))
    layer = nn.SDRLayer(c, c, lists)
    with torch.no_grad():
        for a in layer.aggs:
            a.bias.uniform_(-0.3, 0.3)
    leaf = lambda t: t.detach().clone().requires_grad_(True)
    gp = [[leaf(t) for t in (a.lin.weight, a.att_src, a.att_dst, a.bias)] for a in layer.aggs]
    l0, l2 = layer.mlp_layer[0], layer.mlp_layer[2]
    mp = [leaf(t) for t in (l0.weight, l0.bias, l2.weight, l2.bias)]
    cx = leaf(x)
    ref = port.sdr_layer(cx, lists, [(w, a.view(-1), b.view(-1), bb) for w, a, b, bb in gp], *mp)
    (ref * r).sum().backward()
    with torch.no_grad():
        y = layer(x)                                                         # inference route: Linear folded through
    assert (y - ref.detach()).abs().max() <= 1e-5 * ref.abs().max()
    dx = leaf(x)
    out = layer(dx)                                                          # autograd route
    (out * r).sum().backward()
    assert (out.detach() - ref.detach()).abs().max() <= 1e-5 * ref.abs().max()
    assert (dx.grad - cx.grad).abs().max() <= 2e-5 * cx.grad.abs().max()
    for a, ws in zip(layer.aggs, gp):
        for p, w in zip((a.lin.weight, a.att_src, a.att_dst, a.bias), ws):
            assert (p.grad - w.grad.view_as(p.grad)).abs().max() <= 5e-5 * max(float(w.grad.abs().max()), 2e-2)
    for p, w in zip((l0.weight, l0.bias, l2.weight, l2.bias), mp):
        assert (p.grad - w.grad).abs().max() <= 5e-5 * max(float(w.grad.abs().max()), 2e-2)

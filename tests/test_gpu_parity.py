"""GPU parity tests proper: every layer is driven through its public (reference-shaped) API,
which calls the C ABI, and compared with
  (i)  the golden fixtures produced by the reference's own source files, and
  (ii) oracle/port.py on fresh seeded inputs at sizes the CPU finishes in seconds.
Tolerance: fp32 outputs within 1e-5 * max|ref| per tensor (north_star: "within 1e-5 rel
fp32"); integer artefacts (cached_result indices, CSR structure) bit-exact.
"""
import numpy as np
import pytest
import torch

from conftest import assert_close_rel, load_golden
from oracle import port
from pytorch_geometric_signed_directed_b200 import nn, ops, plan as planmod, synthetic

pytestmark = pytest.mark.gpu
DEV = "cuda"

MAGNET_CASES = ["magnet_c1", "magnet_k2_weighted", "magnet_k3_none", "magnet_q0",
                "magnet_sym_lmax", "msconv_signed", "msconv_nonabs_none"]


def _make_magnet(name, g):
    k1, fin, fout = g["weight"].shape
    norm = "sym" if g["sym"] else None
    if name.startswith("msconv"):
        conv = nn.MSConv(fin, fout, K=k1 - 1, q=g["q"], trainable_q=False, normalization=norm,
                         cached=True, absolute_degree=(name != "msconv_nonabs_none"))
    else:
        conv = nn.MagNetConv(fin, fout, K=k1 - 1, q=g["q"], trainable_q=False, normalization=norm,
                             cached=True)
    conv = conv.to(DEV)
    with torch.no_grad():
        conv.weight.copy_(g["weight"])
        conv.bias.copy_(g["bias"])
    return conv


@pytest.mark.parametrize("name", MAGNET_CASES)
def test_magnet_golden(name):
    g = load_golden(name, DEV)
    conv = _make_magnet(name, g)
    ew = g["edge_weight"] if g["has_weight"] else None
    lam = None if g["lambda_max"] < 0 else g["lambda_max"]
    out_r, out_i = conv(g["x_real"], g["x_imag"], g["edge_index"], ew, lam)
    assert_close_rel(out_r, g["out_real"], 1e-5, f"{name} out_real")
    assert_close_rel(out_i, g["out_imag"], 1e-5, f"{name} out_imag")
    # second call takes the cached branch (test/directed_test.py:61-72) and must agree exactly
    out_r2, out_i2 = conv(g["x_real"], g["x_imag"], g["edge_index"], ew, lam)
    assert torch.equal(out_r, out_r2) and torch.equal(out_i, out_i2)
    # cached_result: index tensors bit-exact, values to rounding
    er, ei_, nr, ni = conv.cached_result
    assert torch.equal(er, g["cached_edge_index_real"])
    assert torch.equal(ei_, g["cached_edge_index_imag"])
    assert_close_rel(nr, g["cached_norm_real"], 2e-6, "norm_real")
    assert_close_rel(ni, g["cached_norm_imag"], 2e-6, "norm_imag")


def test_magnet_cache_rules_and_errors():
    g = load_golden("magnet_q0", DEV)
    conv = _make_magnet("magnet_q0", g)
    conv(g["x_real"], g["x_imag"], g["edge_index"])
    with pytest.raises(RuntimeError, match="Cached 300 number of edges, but found 299"):
        conv(g["x_real"], g["x_imag"], g["edge_index"][:, :299])
    conv.q = 0.1
    with pytest.raises(RuntimeError, match="Cached q is 0.0, but found 0.1"):
        conv(g["x_real"], g["x_imag"], g["edge_index"])
    conv.reset_parameters()
    assert conv.cached_result is None
    # out-of-range node id is reported, not silently read
    bad = g["edge_index"].clone()
    bad[0, 0] = 10_000
    with pytest.raises(RuntimeError, match="outside"):
        nn.MagNetConv(8, 8, K=1, q=0.25, trainable_q=False).to(DEV)(g["x_real"], g["x_imag"], bad)
    tq = nn.MagNetConv(8, 8, K=1, q=0.25, trainable_q=True, normalization=None).to(DEV)
    with pytest.raises(RuntimeError, match="Cannot train q"):
        tq(g["x_real"], g["x_imag"], g["edge_index"])


def test_magnet_lambda_max_eigsh_path_matches_reference_route():
    # normalization=None without lambda_max: lambda_max comes from scipy eigsh (CPU), as upstream
    g = load_golden("magnet_k3_none", DEV)
    conv = nn.MagNetConv(4, 6, K=3, q=g["q"], trainable_q=False, normalization=None).to(DEV)
    with torch.no_grad():
        conv.weight.copy_(g["weight"]); conv.bias.copy_(g["bias"])
    out_r, out_i = conv(g["x_real"], g["x_imag"], g["edge_index"], g["edge_weight"])
    lam = conv._plan.meta["lambda_max"]
    o_r, o_i = port.magnet_conv(g["x_real"].cpu(), g["x_imag"].cpu(), g["edge_index"].cpu(),
                                g["edge_weight"].cpu(), g["weight"].cpu(), g["bias"].cpu(), g["q"],
                                None, lam)
    assert 1.0 < lam < 100.0
    assert_close_rel(out_r, o_r, 1e-5)
    assert_close_rel(out_i, o_i, 1e-5)


@pytest.mark.parametrize("n,e,fin,fout,K,q,weighted", [
    (5000, 100_000, 64, 64, 1, 0.25, False),     # north-star shape in miniature
    (3000, 60_000, 64, 64, 2, 0.1, True),
    (2000, 30_000, 128, 32, 1, 0.25, True),      # 512-B rows
    (2000, 30_000, 32, 48, 3, 0.2, False),
    (1500, 20_000, 16, 16, 1, 0.25, False),
    (1500, 20_000, 48, 24, 2, 0.25, True),       # F/4 not a power of two -> masked lanes
    (1000, 9_000, 3, 2, 2, 0.25, True),          # the reference tests' widths -> scalar path
    (1000, 9_000, 7, 5, 1, 0.05, False),
    (4000, 0, 8, 8, 1, 0.25, False),             # empty graph
    (1, 0, 4, 4, 1, 0.25, False),
])
def test_magnet_vs_oracle(n, e, fin, fout, K, q, weighted):
    g = torch.Generator().manual_seed(n + e + fin)
    ei = torch.randint(0, n, (2, e), generator=g)
    ew = (torch.rand(e, generator=g) + 0.5) if weighted else None
    xr = torch.rand(n, fin, generator=g) * 2 - 1
    xi = torch.rand(n, fin, generator=g) * 2 - 1
    conv = nn.MagNetConv(fin, fout, K=K, q=q, trainable_q=False).to(DEV)
    with torch.no_grad():
        conv.bias.uniform_(-0.5, 0.5)
    o_r, o_i = port.magnet_conv(xr, xi, ei, ew, conv.weight.detach().cpu(), conv.bias.detach().cpu(),
                                q, "sym")
    out_r, out_i = conv(xr.to(DEV), xi.to(DEV), ei.to(DEV), None if ew is None else ew.to(DEV))
    assert_close_rel(out_r, o_r, 1e-5, "out_real")
    assert_close_rel(out_i, o_i, 1e-5, "out_imag")


def test_spmm_variants_agree_and_fused_relu():
    g = torch.Generator().manual_seed(7)
    n, e, f = 4000, 90_000, 64
    ei = torch.randint(0, n, (2, e), generator=g).to(DEV)
    x0 = (torch.rand(n, f, generator=g) * 2 - 1).to(DEV)
    x1 = (torch.rand(n, f, generator=g) * 2 - 1).to(DEV)
    p = planmod.build_magnetic(ei, None, n, 0.25, "sym", 2.0)
    base = ops.spmm(p, [x0, x1], (0, 1), variant=0)
    for variant in (2, 4, 8, 0x10 | 2, 0x20 | 2, 0x20 | 4, 0x80, 0x80 | 2, 0x80 | 8, 0x80 | 0x20 | 2):
        got = ops.spmm(p, [x0, x1], (0, 1), variant=variant)
        assert_close_rel(got[0], base[0], 2e-6, f"variant {variant:#x} op0")
        assert_close_rel(got[1], base[1], 2e-6, f"variant {variant:#x} op1")
    one = ops.spmm(p, [x1], (1,))
    assert_close_rel(one[0], base[1], 2e-6, "single-operator launch")
    # bulk-copy (TMA) gathers + segmented reduction (variant bit 0x800): every epilogue, 1 and 2 operators, rows of
    # 128 / 256 / 512 bytes, empty rows at both ends of a chunk
    z0, z1 = torch.randn(n, f, generator=g).to(DEV), torch.randn(n, f, generator=g).to(DEV)
    bias = torch.randn(f, generator=g).to(DEV)
    kw = dict(alpha=2.0, beta=-1.0, zs=[z0, z1], bias=bias)
    want = ops.spmm(p, [x0, x1], (0, 1), **kw)
    got = ops.spmm(p, [x0, x1], (0, 1), variant=0x800, **kw)
    assert_close_rel(got[0], want[0], 2e-6, "bulk variant, Chebyshev epilogue op0")
    assert_close_rel(got[1], want[1], 2e-6, "bulk variant, Chebyshev epilogue op1")
    assert_close_rel(ops.spmm(p, [x1], (1,), variant=0x800)[0], base[1], 2e-6, "bulk variant, one operator")
    for width in (32, 128):
        ei2 = torch.randint(5, n - 300, (2, 30_000), generator=g).to(DEV)          # first / last rows are empty
        q2 = planmod.build_csr(ei2, torch.rand(30_000, generator=g).to(DEV), n, n, "source_to_target")
        xw = torch.randn(n, width, generator=g).to(DEV)
        assert_close_rel(ops.spmm(q2, [xw], (0,), mean=True, variant=0x800)[0], ops.spmm(q2, [xw], (0,), mean=True)[0],
                         2e-6, f"bulk variant, mean, width {width}")
    conv = nn.MagNetConv(f, f, K=1, q=0.25, trainable_q=False).to(DEV)
    r, i = conv(x0, x1, ei)
    conv.fused_complex_relu = True
    rr, ii = conv(x0, x1, ei)
    er, ei2 = port.complex_relu(r.cpu(), i.cpu())
    assert torch.equal(rr.cpu(), er) and torch.equal(ii.cpu(), ei2)


def test_csr_plan_structure_is_exact():
    g = torch.Generator().manual_seed(11)
    n, e = 700, 6000
    ei = torch.randint(0, n, (2, e), generator=g)
    w = torch.rand(e, generator=g)
    p = planmod.build_csr(ei.to(DEV), w.to(DEV), n, n, "source_to_target")
    order = torch.sort(ei[1], stable=True).indices                # stable by destination
    assert torch.equal(p.col.cpu().long(), ei[0][order])
    assert torch.equal(p.val[0].cpu(), w[order])
    counts = torch.bincount(ei[1], minlength=n)
    assert torch.equal((p.row_ptr[1:] - p.row_ptr[:-1]).cpu().long(), counts)


def test_digcn_golden_and_cache_quirk():
    g = load_golden("digcn_conv", DEV)
    conv = nn.DiGCNConv(12, 7).to(DEV)
    with torch.no_grad():
        conv.weight.copy_(g["weight"]); conv.bias.copy_(g["bias"])
    with pytest.raises(RuntimeError, match="Normalized adj matrix cannot be None"):
        conv(g["x"], g["edge_index"])
    y = conv(g["x"], g["edge_index"], g["edge_weight"])
    assert_close_rel(y, g["out"], 1e-5)
    # cached=True default: later calls ignore new edge VALUES ...
    y2 = conv(g["x"], g["edge_index"].flip(1), g["edge_weight"] * 0)
    assert torch.equal(y, y2)
    # ... but not a new edge COUNT
    with pytest.raises(RuntimeError, match="Cached 1500 number of edges, but found 1499"):
        conv(g["x"], g["edge_index"][:, :1499], g["edge_weight"][:1499])


def test_inception_block_golden_and_bf16():
    g = load_golden("digcn_inception", DEV)
    blk = nn.DiGCN_InceptionBlock(10, 6).to(DEV)
    with torch.no_grad():
        blk.ln.weight.copy_(g["ln_weight"]); blk.ln.bias.copy_(g["ln_bias"])
        blk.conv1.weight.copy_(g["conv1_weight"]); blk.conv1.bias.copy_(g["conv1_bias"])
        blk.conv2.weight.copy_(g["conv2_weight"]); blk.conv2.bias.copy_(g["conv2_bias"])
    x0, x1, x2 = blk(g["x"], g["edge_index"], g["edge_weight"], g["edge_index2"], g["edge_weight2"])
    for a, b in ((x0, g["x0"]), (x1, g["x1"]), (x2, g["x2"])):
        assert_close_rel(a, b, 1e-5)

    # BASELINE config 3 shape in miniature: 128 features, bf16 storage, fp32 accumulation.
    # Tolerance: inputs/outputs are rounded to bf16 (rel 2^-9 = 2e-3 per rounding); compare with
    # the fp32 oracle evaluated on the bf16-rounded inputs at 1e-2 * max|ref|.
    gen = torch.Generator().manual_seed(5)
    n, e, f = 3000, 50_000, 128
    ei1 = torch.randint(0, n, (2, e), generator=gen); ei2 = torch.randint(0, n, (2, e), generator=gen)
    w1 = synthetic.sym_norm_weights(ei1, n); w2 = synthetic.sym_norm_weights(ei2, n)
    x = (torch.rand(n, f, generator=gen) * 2 - 1).bfloat16()
    blk = nn.DiGCN_InceptionBlock(f, f).to(DEV)
    for prm in blk.parameters():
        prm.data = prm.data.bfloat16().float()
    y = blk(x.to(DEV), ei1.to(DEV), w1.to(DEV), ei2.to(DEV), w2.to(DEV))
    ref = port.digcn_inception_block(x.float(), ei1, w1, ei2, w2, blk.ln.weight.cpu(), blk.ln.bias.cpu(),
                                     blk.conv1.weight.cpu(), blk.conv1.bias.cpu(),
                                     blk.conv2.weight.cpu(), blk.conv2.bias.cpu())
    for a, b in zip(y, ref):
        assert a.dtype == torch.bfloat16
        assert_close_rel(a.float(), b.detach(), 1e-2, "bf16 inception")


@pytest.mark.parametrize("name,first,norm_emb", [("sgcn_first", True, False), ("sgcn_second", False, True)])
def test_sgcn_golden(name, first, norm_emb):
    g = load_golden(name, DEV)
    fo = g["lin_b_weight"].shape[0]
    fi = g["lin_b_weight"].shape[1] // (2 if first else 3)
    conv = nn.SGCNConv(fi, fo, first_aggr=first, norm_emb=norm_emb).to(DEV)
    with torch.no_grad():
        conv.lin_b.weight.copy_(g["lin_b_weight"]); conv.lin_b.bias.copy_(g["lin_b_bias"])
        conv.lin_u.weight.copy_(g["lin_u_weight"]); conv.lin_u.bias.copy_(g["lin_u_bias"])
    y = conv(g["x"], g["pos_edge_index"], g["neg_edge_index"])
    assert_close_rel(y, g["out"], 1e-5)
    # in-degree counts are integers: CSR row lengths must equal them exactly
    p = conv._plan_for(g["pos_edge_index"], g["x"].size(0), g["x"].size(0))
    cnt = torch.bincount(g["pos_edge_index"][1], minlength=g["x"].size(0))
    assert torch.equal((p.row_ptr[1:] - p.row_ptr[:-1]).long(), cnt)


def test_sgcn_two_layer_vs_oracle_wide():
    gen = torch.Generator().manual_seed(9)
    n = 6000
    pos, neg, _ = synthetic.ssbm_edges(n, 3, num_entries=120_000, eta=0.1, seed=2)
    x = torch.randn(n, 64, generator=gen)
    c1 = nn.SGCNConv(64, 32, first_aggr=True).to(DEV)
    c2 = nn.SGCNConv(32, 32, first_aggr=False).to(DEV)
    with torch.no_grad():
        z1 = c1(x.to(DEV), pos.to(DEV), neg.to(DEV))
        z2 = c2(torch.tanh(z1), pos.to(DEV), neg.to(DEV))
    t1 = c1(x.to(DEV), pos.to(DEV), neg.to(DEV))                          # autograd path, same numbers
    t2 = c2(torch.tanh(t1), pos.to(DEV), neg.to(DEV))
    assert t2.requires_grad
    assert_close_rel(t1, z1, 2e-6, "layer 1: autograd path vs inference path")
    assert_close_rel(t2, z2, 2e-6, "layer 2: autograd path vs inference path")
    cpu = lambda m: [m.lin_b.weight.detach().cpu(), m.lin_b.bias.detach().cpu(),
                     m.lin_u.weight.detach().cpu(), m.lin_u.bias.detach().cpu()]
    r1 = port.sgcn_conv(x, pos, neg, *cpu(c1), True)
    r2 = port.sgcn_conv(torch.tanh(r1), pos, neg, *cpu(c2), False)
    assert_close_rel(z1, r1, 1e-5)
    assert_close_rel(z2, r2, 1e-5)


def test_conv_base_and_dimpa_golden():
    g = load_golden("conv_norm_rw", DEV)
    p = planmod.build_rw_norm(g["edge_index"], g["edge_weight"], 110, g["fill_value"])
    # rebuild the reference's (edge_index', w') list from the plan and compare as a sparse matrix
    ref = torch.zeros(110, 110, dtype=torch.float64)
    ref.index_put_((g["out_edge_index"][0].cpu(), g["out_edge_index"][1].cpu()),
                   g["out_weight"].cpu().double(), accumulate=True)
    rows = torch.repeat_interleave(torch.arange(110), (p.row_ptr[1:] - p.row_ptr[:-1]).cpu().long())
    got = torch.zeros(110, 110, dtype=torch.float64)
    got.index_put_((rows, p.col.cpu().long()), p.val[0].cpu().double(), accumulate=True)
    got += torch.diag(p.diag[0].cpu().double())
    assert_close_rel(got, ref, 1e-6, "conv_norm_rw matrix")

    g = load_golden("conv_base", DEV)
    assert_close_rel(nn.Conv_Base(0.5)(g["x"], g["edge_index"], g["edge_weight"]), g["out"], 1e-5)
    assert_close_rel(nn.Conv_Base(0.25)(g["x"], g["edge_index"], None), g["out_unweighted_fill025"], 1e-5)
    g = load_golden("dimpa", DEV)
    dm = nn.DIMPA(hop=2).to(DEV)
    with torch.no_grad():
        dm._w_s.copy_(g["w_s"]); dm._w_t.copy_(g["w_t"])
    assert_close_rel(dm(g["x_s"], g["x_t"], g["edge_index"], g["edge_weight"]), g["out"], 1e-5)


def test_conv_base_follows_new_edge_tensors():
    gen = torch.Generator().manual_seed(1)
    n = 500
    x = torch.rand(n, 32, generator=gen)
    cb = nn.Conv_Base(0.5)
    for seed in (1, 2):
        ei = torch.randint(0, n, (2, 4000), generator=torch.Generator().manual_seed(seed))
        assert_close_rel(cb(x.to(DEV), ei.to(DEV)), port.conv_base(x, ei, None, 0.5), 1e-5)
    ei_d = ei.to(DEV)
    y1 = cb(x.to(DEV), ei_d)
    ei_d[0, :100] = 0                       # in-place edit bumps the version counter -> new plan
    y2 = cb(x.to(DEV), ei_d)
    assert_close_rel(y2, port.conv_base(x, ei_d.cpu(), None, 0.5), 1e-5)
    assert not torch.equal(y1, y2)


def test_gather_rows_halo_pack():
    x = torch.randn(1000, 64, device=DEV)
    idx = torch.randint(0, 1000, (333,), device=DEV, dtype=torch.int32)
    assert torch.equal(ops.gather_rows(x, idx), x[idx.long()])


# ---------------------------------------------------------------------------- full size
def _c2_inputs(n, e):
    ei, _ = synthetic.dsbm_edges(n, 3, num_edges=e, eta=0.1, size_ratio=1.5, seed=0, device=DEV)
    gen = torch.Generator(device=DEV).manual_seed(0)
    xr = torch.rand(n, 64, generator=gen, device=DEV) * 2 - 1
    xi = torch.rand(n, 64, generator=gen, device=DEV) * 2 - 1
    return ei, xr, xi


@pytest.mark.parametrize("n,e", [(200_000, 4_000_000), (1_000_000, 20_000_000)])
def test_magnet_full_size_properties(n, e):
    """BASELINE config 2 size (and a 1/5 copy): size-independent properties + a row-subset
    comparison with the oracle (the full edge-materialising oracle needs ~27 GB at 1M/20M)."""
    ei, xr, xi = _c2_inputs(n, e)
    conv = nn.MagNetConv(64, 64, K=1, q=0.25, trainable_q=False, cached=True).to(DEV)
    with torch.no_grad():
        conv.bias.uniform_(-0.5, 0.5)
    out_r, out_i = conv(xr, xi, ei)
    assert torch.isfinite(out_r).all() and torch.isfinite(out_i).all()
    p = conv._plan
    # structure: symmetric pattern, no diagonal, sorted unique columns per row, real part
    # symmetric / imaginary part antisymmetric (checked through x^T L y identities below)
    assert p.nnz % 2 == 0 and p.nnz <= 2 * ei.size(1)
    rows = torch.repeat_interleave(torch.arange(n, device=DEV), (p.row_ptr[1:] - p.row_ptr[:-1]).long())
    key = rows * n + p.col.long()
    assert bool((key[1:] > key[:-1]).all()) and not bool((rows == p.col).any())
    del key
    # linearity of the aggregation: L(ax + by) = a L(x) + b L(y)
    t = ops.spmm(p, [xr, xi], (0, 1))
    comb = ops.spmm(p, [0.5 * xr - 2.0 * xi, xi], (0, 1))
    lin = 0.5 * t[0] - 2.0 * ops.spmm(p, [xi], (0,))[0]
    assert_close_rel(comb[0], lin, 2e-6, "linearity")
    # adjointness: <y, L_r x> = <L_r y, x> (symmetric), <y, L_i x> = -<L_i y, x> (antisymmetric),
    # measured against the Cauchy-Schwarz scale ||y|| ||L x||
    u = ops.spmm(p, [xi, xr], (0, 1))
    nrm = lambda v: v.double().pow(2).sum().sqrt().item()
    a = (xi.double() * t[0].double()).sum().item(); b = (u[0].double() * xr.double()).sum().item()
    assert abs(a - b) <= 1e-6 * nrm(xi) * nrm(t[0])
    c = (xr.double() * t[1].double()).sum().item(); d = (u[1].double() * xi.double()).sum().item()
    assert abs(c + d) <= 1e-6 * nrm(xr) * nrm(t[1])
    # The OPERATOR against the oracle, built by the oracle from the edge list (not from the GPU plan):
    #   * 200k / 4M: the whole cached_result against port.magnet_norm -- indices bit-exact, values 1e-6;
    #   * 1M / 20M : the entries aggregated into 3000 random rows against port.magnet_norm_rows (the full
    #     2E-key sort takes minutes on the CPU), same bars.
    gen = torch.Generator().manual_seed(1)
    sel = torch.randperm(n, generator=gen)[:3000].sort().values
    cr = [c_.cpu() for c_ in conv.cached_result]
    ei_cpu = ei.cpu()
    if n <= 200_000:
        ref = port.magnet_norm(ei_cpu, None, n, 0.25, "sym", 2.0)
        for k in (0, 1):
            assert torch.equal(cr[k], ref[k]), f"cached_result index tensor {k} differs from the oracle's"
        for k in (2, 3):
            assert (cr[k] - ref[k]).abs().max().item() <= 1e-6 * ref[k].abs().max().item(), f"operator values {k}"
        oracle_rows = port.magnet_norm_rows(sel, ei_cpu, None, n, 0.25, "sym", 2.0)
    else:
        oracle_rows = port.magnet_norm_rows(sel, ei_cpu, None, n, 0.25, "sym", 2.0)
        nnz = cr[1].size(1) - n
        pick = torch.zeros(n, dtype=torch.bool)
        pick[sel] = True
        keep = pick[cr[1][1, :nnz]]
        m = int(keep.sum())
        assert torch.equal(cr[1][:, :nnz][:, keep], oracle_rows[1][:, :m]), "operator entries of the sampled rows"
        for k in (2, 3):
            got, want = cr[k][:nnz][keep], oracle_rows[k][:m]
            assert (got - want).abs().max().item() <= 1e-6 * want.abs().max().item(), f"operator values {k}"
    # the layer's outputs on those rows against the oracle's op sequence on the ORACLE's operator rows
    o_r, o_i = port.magnet_conv_rows(sel, xr.cpu(), xi.cpu(), oracle_rows, conv.weight.detach().cpu(),
                                     conv.bias.detach().cpu())
    scale_r, scale_i = out_r.abs().max().item(), out_i.abs().max().item()
    assert (out_r[sel.to(DEV)].cpu() - o_r).abs().max().item() <= 1e-5 * scale_r
    assert (out_i[sel.to(DEV)].cpu() - o_i).abs().max().item() <= 1e-5 * scale_i


# ------------------------------------------------------------------ dense transform paths
@pytest.mark.parametrize("n_rows,ks,n_out,combine", [
    (1000, (64, 64, 64, 64), 64, True),       # MagNet K=1, 64 -> 64
    (130, (64, 64, 64, 64, 64, 64), 64, True),  # K=2
    (4097, (32, 32, 32, 32), 32, True),
    (777, (16, 16, 16, 16), 16, True),        # partial K chunk (16 of 32)
    (2500, (48, 48), 128, True),
    (1, (64, 64), 64, True),
    (3000, (128,), 128, False),               # DiGCN-like single term
    (3000, (64, 64), 32, False),              # SGCN first layer: two column blocks
    (999, (32, 32, 32), 32, False),           # SGCN deep layer: three column blocks
])
def test_dense_tensor_core_path_matches_fp64(n_rows, ks, n_out, combine):
    gen = torch.Generator(device=DEV).manual_seed(n_rows + n_out)
    xs = [torch.randn(n_rows, k, generator=gen, device=DEV) for k in ks]
    # MagNet-style: consecutive (real, imag) terms share one weight
    ws = []
    for t, k in enumerate(ks):
        if combine and t % 2 == 1:
            ws.append(ws[-1])
        else:
            ws.append(torch.randn(k, n_out, generator=gen, device=DEV) / k ** 0.5)
    bias = torch.randn(n_out, generator=gen, device=DEV)
    terms = [(x, w, (t % 2) if combine else 0) for t, (x, w) in enumerate(zip(xs, ws))]
    acc = [torch.zeros(n_rows, n_out, dtype=torch.float64, device=DEV) for _ in range(2)]
    for x, w, g in terms:
        acc[g] += x.double() @ w.double()
    ref = [acc[0] - acc[1] + bias.double(), acc[0] + acc[1] + bias.double()] if combine \
        else [acc[0] + bias.double()]
    tc = ops.dense(terms, n_out, bias=bias, combine=combine, variant=4)     # tcgen05, warp-specialised
    ts = ops.dense(terms, n_out, bias=bias, combine=combine, variant=2)     # tcgen05, synchronous kernel
    ff = ops.dense(terms, n_out, bias=bias, combine=combine, variant=1)     # FFMA
    for got_tc, got_ts, got_ff, r in zip(tc, ts, ff, ref):
        assert_close_rel(got_tc, r, 2e-6, "tcgen05 3xTF32 (warp-specialised) vs fp64")
        assert_close_rel(got_ts, r, 2e-6, "tcgen05 3xTF32 (synchronous) vs fp64")
        assert_close_rel(got_ff, r, 2e-6, "ffma vs fp64")
    # many tiles per CTA: exercises the TMEM double buffer and every mbarrier phase wrap
    if n_rows < 5000:
        big = [(x.repeat(40, 1), w, g) for x, w, g in terms]
        b1 = ops.dense(big, n_out, bias=bias, combine=combine, variant=4)
        b2 = ops.dense(big, n_out, bias=bias, combine=combine, variant=2)
        b3 = ops.dense(big, n_out, bias=bias, combine=combine, variant=1)
        for u, u2, v in zip(b1, b2, b3):
            assert_close_rel(u, v, 2e-6, "long run: tcgen05 (warp-specialised) vs ffma")
            assert_close_rel(u2, v, 2e-6, "long run: tcgen05 (synchronous) vs ffma")


def test_dense_tensor_core_strided_operands_and_relu():
    gen = torch.Generator(device=DEV).manual_seed(3)
    n = 5000
    wide = torch.randn(n, 192, generator=gen, device=DEV)
    lin = torch.randn(64, 128, generator=gen, device=DEV)          # nn.Linear layout [out, in]
    wt = lin.t()                                                   # [in=128, out=64] view, strides (1, 128)
    terms = [(wide[:, 64:128], wt[:64], 0), (wide[:, 128:], wt[64:], 0)]
    out = torch.empty(n, 128, device=DEV)
    ops.dense(terms, 64, out=[out[:, 64:]], variant=2)
    ref = wide[:, 64:128].double() @ wt[:64].double() + wide[:, 128:].double() @ wt[64:].double()
    assert_close_rel(out[:, 64:], ref, 2e-6)
    x0, x1 = wide[:, :64].contiguous(), wide[:, 64:128].contiguous()
    w = torch.randn(64, 64, generator=gen, device=DEV) / 8
    r, i = ops.dense([(x0, w, 0), (x1, w, 1)], 64, combine=True, relu_mode=1, variant=2)
    a, b = x0.double() @ w.double(), x1.double() @ w.double()
    mask = ((a - b) >= 0).double()
    assert_close_rel(r, (a - b) * mask, 2e-6)
    # the mask is decided on the fp32 value of out_real: tolerate flips where |a-b| ~ 0
    agree = (((r != 0) | ((a - b).abs() < 1e-5)) == ((mask != 0) | ((a - b).abs() < 1e-5))).all()
    assert bool(agree)
    assert_close_rel(i * (r != 0), (a + b) * mask * (r != 0).double(), 2e-6)


# ------------------------------------------------------------------------------- SNEAConv
def _make_snea(g, first):
    fo, fi = g["lin_b_weight"].shape
    conv = nn.SNEAConv(fi, fo, first_aggr=first).to(DEV)
    with torch.no_grad():
        for nm in ("lin_b", "lin_u", "alpha_b", "alpha_u"):
            getattr(conv, nm).weight.copy_(g[nm + "_weight"])
            getattr(conv, nm).bias.copy_(g[nm + "_bias"])
    return conv


@pytest.mark.parametrize("name,first", [("snea_first", True), ("snea_second", False)])
def test_snea_golden(name, first):
    g = load_golden(name, DEV)
    conv = _make_snea(g, first)
    y = conv(g["x"], g["pos_edge_index"], g["neg_edge_index"])          # parameters require grad: autograd path
    assert y.requires_grad
    assert_close_rel(y, g["out"], 1e-5, "autograd path")
    with torch.no_grad():                                                 # inference path (raw kernels)
        y = conv(g["x"], g["pos_edge_index"], g["neg_edge_index"])
    assert_close_rel(y, g["out"], 1e-5, "inference path")


def test_snea_two_layer_vs_oracle_wide():
    n = 5000
    pos, neg, _ = synthetic.ssbm_edges(n, 3, num_entries=100_000, eta=0.1, seed=5)
    x = torch.randn(n, 64, generator=torch.Generator().manual_seed(2))
    torch.manual_seed(3)
    c1 = nn.SNEAConv(64, 32, first_aggr=True).to(DEV)
    c2 = nn.SNEAConv(32, 32, first_aggr=False).to(DEV)
    with torch.no_grad():
        z1 = c1(x.to(DEV), pos.to(DEV), neg.to(DEV))
        z2 = c2(torch.tanh(z1), pos.to(DEV), neg.to(DEV))
    t1 = c1(x.to(DEV), pos.to(DEV), neg.to(DEV))                          # autograd path, same numbers
    t2 = c2(torch.tanh(t1), pos.to(DEV), neg.to(DEV))
    assert t2.requires_grad
    assert_close_rel(t1, z1, 2e-6, "layer 1: autograd path vs inference path")
    assert_close_rel(t2, z2, 2e-6, "layer 2: autograd path vs inference path")
    cpu = lambda m: [t.detach().cpu() for t in (m.lin_b.weight, m.lin_b.bias, m.lin_u.weight, m.lin_u.bias,
                                                m.alpha_b.weight, m.alpha_b.bias, m.alpha_u.weight, m.alpha_u.bias)]
    r1 = port.snea_conv(x, pos, neg, *cpu(c1), True)
    r2 = port.snea_conv(torch.tanh(r1), pos, neg, *cpu(c2), False)
    assert_close_rel(z1, r1, 1e-5)
    assert_close_rel(z2, r2, 1e-5)


@pytest.mark.parametrize("n_rows,k,n_out,dtype", [
    (3000, 128, 384, torch.bfloat16),     # inception block: one pass over x for [W_ln | W_1 | W_2], 3 column tiles
    (1000, 64, 128, torch.bfloat16),
    (515, 72, 64, torch.bfloat16),        # partial K chunk (72 = 64 + 8)
    (2000, 64, 256, torch.float32),       # fp32 with two 128-column tiles
])
def test_dense_tensor_core_bf16_and_column_tiles(n_rows, k, n_out, dtype):
    gen = torch.Generator(device=DEV).manual_seed(n_rows + k)
    x = torch.randn(n_rows, k, generator=gen, device=DEV).to(dtype)
    w = (torch.randn(k, n_out, generator=gen, device=DEV) / k ** 0.5)
    if dtype == torch.bfloat16:
        w = w.bfloat16().float()          # bf16-representable parameters (a bf16 model)
    bias = torch.randn(n_out, generator=gen, device=DEV)
    ref = x.double() @ w.double() + bias.double()
    tc = ops.dense([(x, w, 0)], n_out, bias=bias, variant=2)[0]
    ff = ops.dense([(x, w, 0)], n_out, bias=bias, variant=1)[0]
    tol = 6e-3 if dtype == torch.bfloat16 else 2e-6      # one bf16 output rounding = 2^-9 relative
    assert tc.dtype == dtype
    assert_close_rel(tc.double(), ref, tol, "tcgen05")
    assert_close_rel(ff.double(), ref, tol, "ffma")


@pytest.mark.parametrize("f,dtype,n_ops", [(64, torch.float32, 2), (64, torch.float32, 1), (16, torch.float32, 2),
                                           (32, torch.float32, 1), (128, torch.float32, 2), (48, torch.float32, 1),
                                           (128, torch.bfloat16, 1), (64, torch.bfloat16, 1)])
def test_group_per_row_kernel_matches_row_kernel(f, dtype, n_ops):
    """variant bit 0x40: the short-row kernel (each lane group owns a row, next row's indices
    prefetched) against the warp-per-row kernel, all epilogues, ragged / empty / long rows."""
    g = torch.Generator().manual_seed(f + n_ops)
    n = 3001
    e = 9000
    ei = torch.randint(0, n - 40, (2, e), generator=g)          # last 40 rows empty
    ei[1, :700] = 5                                               # one long row (700 entries)
    ei = ei.to(DEV)
    p = planmod.build_magnetic(ei, None, n, 0.25, "sym", 1.6)    # non-zero real diagonal
    xs = [(torch.rand(n, f, generator=g) * 2 - 1).to(DEV).to(dtype) for _ in range(n_ops)]
    zs = [(torch.rand(n, f, generator=g) * 2 - 1).to(DEV).to(dtype) for _ in range(n_ops)]
    bias = torch.randn(f, generator=g).to(DEV)
    ops_ids = (0, 1)[:n_ops] if n_ops == 2 else (1,)
    tol = 2e-6 if dtype == torch.float32 else 1e-2
    for kw in (dict(), dict(alpha=2.0, beta=-1.0, zs=zs), dict(bias=bias), dict(mean=True)):
        ref = ops.spmm(p, xs, ops_ids, variant=0x80 | 0x10 | 4, **kw)        # warp-per-row kernel
        for variant in (0, 2, 0x20 | 2, 0x10 | 4, 0x20 | 4):                  # group-per-row kernel
            got = ops.spmm(p, xs, ops_ids, variant=variant, **kw)
            for a, b in zip(got, ref):
                assert_close_rel(a.float(), b.float(), tol, f"variant {variant:#x} {list(kw)}")


def test_hub_rows_power_law_graph():
    """Rows far above PGSD_HUB_ROW_THRESHOLD (4096): their entries are aggregated in slices by the
    hub-row kernel (fp32 atomics); everything else is unchanged."""
    g = torch.Generator().manual_seed(21)
    n = 60_000
    ei = torch.randint(0, n, (2, 200_000), generator=g)
    ei[1, :50_000] = 7                  # 50k edges into node 7  (=> ~100k stored entries after symmetrisation)
    ei[0, 50_000:62_000] = 9            # 12k edges out of node 9
    ew = torch.rand(ei.size(1), generator=g) + 0.5
    xr, xi = torch.rand(n, 64, generator=g) * 2 - 1, torch.rand(n, 64, generator=g) * 2 - 1
    conv = nn.MagNetConv(64, 32, K=2, q=0.15, trainable_q=False, cached=True).to(DEV)
    with torch.no_grad():
        out_r, out_i = conv(xr.to(DEV), xi.to(DEV), ei.to(DEV), ew.to(DEV))
    hubs = conv._plan.hub_rows()
    assert hubs is not None and set(hubs[0].tolist()) >= {7, 9}
    o_r, o_i = port.magnet_conv(xr, xi, ei, ew, conv.weight.detach().cpu(), conv.bias.detach().cpu(), 0.15, "sym")
    assert_close_rel(out_r, o_r, 1e-5, "hub graph out_real")
    assert_close_rel(out_i, o_i, 1e-5, "hub graph out_imag")
    # both main kernels + mean epilogue with a hub, against the scalar evaluation
    pos = ei[:, :120_000]
    c1 = nn.SGCNConv(64, 16, first_aggr=True).to(DEV)
    with torch.no_grad():
        z = c1(xr.to(DEV), pos.to(DEV), ei[:, 120_000:].to(DEV))
    ref = port.sgcn_conv(xr, pos, ei[:, 120_000:], c1.lin_b.weight.detach().cpu(), c1.lin_b.bias.detach().cpu(),
                         c1.lin_u.weight.detach().cpu(), c1.lin_u.bias.detach().cpu(), True)
    assert_close_rel(z, ref, 1e-5, "hub graph sgcn")
    p = conv._plan
    a = ops.spmm(p, [xr.to(DEV), xi.to(DEV)], (0, 1), variant=0)
    b = ops.spmm(p, [xr.to(DEV), xi.to(DEV)], (0, 1), variant=0x80)
    assert_close_rel(a[0], b[0], 1e-5, "group vs row kernel with hubs")


def test_dgcn_and_simpa_golden():
    g = load_golden("dgcn_conv", DEV)
    conv = nn.DGCNConv(cached=True)
    y = conv(g["x"], g["edge_index"], g["edge_weight"])
    assert_close_rel(y, g["out"], 1e-5)
    # cache quirk Q6: a cached layer ignores whatever graph comes next
    assert torch.equal(conv(g["x"], g["edge_index"][:, :50], None), y)
    assert_close_rel(nn.DGCNConv(improved=True)(g["x"], g["edge_index"], None), g["out_improved_unweighted"], 1e-5)
    g = load_golden("simpa", DEV)
    args = (g["edge_index_p"], g["edge_weight_p"], g["edge_index_n"], g["edge_weight_n"], g["x_p"], g["x_n"])
    und = nn.SIMPA(hop=2, fill_value=0.5, directed=False).to(DEV)
    with torch.no_grad():
        und._w_p.copy_(g["w_p"]); und._w_n.copy_(g["w_n"])
        assert_close_rel(und(*args), g["out_undirected"], 1e-5)
    dr = nn.SIMPA(hop=2, fill_value=0.5, directed=True).to(DEV)
    with torch.no_grad():
        for nm in ("w_sp", "w_sn", "w_tp", "w_tn"):
            getattr(dr, "_" + nm).copy_(g[nm])
        assert_close_rel(dr(*args, g["x_pt"], g["x_nt"]), g["out_directed"], 1e-5)
    # and with autograd enabled (differentiable path, torch.cat of the parts)
    und2 = nn.SIMPA(hop=2, fill_value=0.5).to(DEV)
    with torch.no_grad():
        und2._w_p.copy_(g["w_p"]); und2._w_n.copy_(g["w_n"])
    out = und2(*args)
    assert out.requires_grad
    assert_close_rel(out, g["out_undirected"], 1e-5)
    out.sum().backward()
    assert und2._w_p.grad is not None and torch.isfinite(und2._w_p.grad).all()


def test_sdr_layer_golden_and_wide():
    g = load_golden("sdr_layer", DEV)
    lists = [g[f"edges_{i}"] for i in range(4)]
    layer = nn.SDRLayer(12, 12, lists).to(DEV)
    layer.load_state_dict({k.replace("__", "."): v for k, v in g.items()
                           if k not in ("x", "out") and not k.startswith("edges_")})
    assert_close_rel(layer(g["x"]), g["out"], 1e-5, "autograd path")
    with torch.no_grad():
        assert_close_rel(layer(g["x"]), g["out"], 1e-5, "inference path (Linear folded through the aggregations)")
    # a GATConv on its own against the oracle at a width the vector kernels take
    gen = torch.Generator().manual_seed(12)
    n = 4000
    ei = torch.randint(0, n, (2, 60_000), generator=gen)
    x = torch.randn(n, 64, generator=gen)
    conv = nn.GATConv(64, 64).to(DEV)
    with torch.no_grad():
        conv.bias.uniform_(-0.2, 0.2)
    ref = port.gat_conv(x, ei, conv.lin.weight.detach().cpu(), conv.att_src.detach().cpu(),
                        conv.att_dst.detach().cpu(), conv.bias.detach().cpu())
    assert_close_rel(conv(x.to(DEV), ei.to(DEV)), ref, 1e-5, "autograd path")
    with torch.no_grad():
        assert_close_rel(conv(x.to(DEV), ei.to(DEV)), ref, 1e-5, "inference path")


def test_magnet_bf16_features():
    """bf16 storage, fp32 accumulation everywhere (aggregation and tcgen05 kind::f16 transform):
    compared with the fp32 oracle on the bf16-rounded inputs at 2e-2 * max|ref| (three bf16
    roundings on the path: T, out, and the inputs' own 2^-9)."""
    g = torch.Generator().manual_seed(31)
    n, e, f = 4000, 80_000, 64
    ei = torch.randint(0, n, (2, e), generator=g)
    xr = (torch.rand(n, f, generator=g) * 2 - 1).bfloat16()
    xi = (torch.rand(n, f, generator=g) * 2 - 1).bfloat16()
    conv = nn.MagNetConv(f, f, K=2, q=0.25, trainable_q=False).to(DEV)
    with torch.no_grad():
        conv.weight.copy_(conv.weight.bfloat16().float())
        conv.bias.uniform_(-0.3, 0.3)
        out_r, out_i = conv(xr.to(DEV), xi.to(DEV), ei.to(DEV))
    assert out_r.dtype == torch.bfloat16
    o_r, o_i = port.magnet_conv(xr.float(), xi.float(), ei, None, conv.weight.detach().cpu(),
                                conv.bias.detach().cpu(), 0.25, "sym")
    assert_close_rel(out_r.float(), o_r, 2e-2, "bf16 out_real")
    assert_close_rel(out_i.float(), o_i, 2e-2, "bf16 out_imag")


def test_magnet_node_classification_model_golden():
    g = load_golden("magnet_model", DEV)
    model = nn.MagNet_node_classification(9, hidden=16, q=0.2, K=2, label_dim=5, activation=True, layer=3,
                                          dropout=0.5, cached=True).to(DEV).eval()
    model.load_state_dict({k.replace("__", "."): v for k, v in g.items()
                           if k not in ("x", "out", "edge_index", "edge_weight")})
    with torch.no_grad():
        y = model(g["x"], g["x"], g["edge_index"], g["edge_weight"])
        y2 = model(g["x"], g["x"], g["edge_index"], g["edge_weight"])       # cached branch
    assert_close_rel(y, g["out"], 1e-5)
    assert torch.equal(y, y2)
    assert model.Chebs[1]._plan is model.Chebs[0]._plan                      # one operator for the stack
    # training mode: gradients reach every layer through the shared plan and the fused ReLU
    model.train()
    out = model(g["x"], g["x"], g["edge_index"], g["edge_weight"])
    F_loss = torch.nn.functional.nll_loss(out, torch.randint(0, 5, (out.size(0),), device=DEV))
    F_loss.backward()
    assert all(p.grad is not None and torch.isfinite(p.grad).all() for p in model.parameters())


@pytest.mark.parametrize("f", [64, 16, 128, 48])
def test_shared_operand_gathers_once_and_is_bit_identical(f):
    """x_real and x_imag being ONE tensor (examples/magnet_node.py:61-62 pass data.x for both): the
    aggregation kernel gathers each neighbour row once for both operators.  Per-row summation order
    is unchanged, so the result must be bit-identical to the two-gather launch (variant bit 0x400
    switches the sharing off) and to a launch on a cloned second operand; layer output vs oracle."""
    g = torch.Generator().manual_seed(100 + f)
    n, e = 5003, 70_000
    ei = torch.randint(0, n - 30, (2, e), generator=g)
    ei[1, :900] = 11                                               # one long row
    x = (torch.rand(n, f, generator=g) * 2 - 1).to(DEV)
    z = [(torch.rand(n, f, generator=g) * 2 - 1).to(DEV) for _ in range(2)]
    p = planmod.build_magnetic(ei.to(DEV), None, n, 0.25, "sym", 1.7)   # non-zero real diagonal
    for kw in (dict(), dict(alpha=2.0, beta=-1.0, zs=z), dict(op_scale=(1.0, -1.0))):
        for variant in (0, 2, 0x10 | 4, 0x20 | 2, 0x20 | 4):
            shared = ops.spmm(p, [x, x], (0, 1), variant=variant, **kw)
            split = ops.spmm(p, [x, x], (0, 1), variant=variant | 0x400, **kw)
            cloned = ops.spmm(p, [x, x.clone()], (0, 1), variant=variant, **kw)
            for a, b, c in zip(shared, split, cloned):
                assert torch.equal(a, b) and torch.equal(a, c), f"variant {variant:#x} {list(kw)}"
    # column slice of a wider buffer passed twice
    wide = torch.zeros(n, 2 * f + 8, device=DEV)
    wide[:, 8:8 + f] = x
    sl = wide[:, 8:8 + f]
    a = ops.spmm(p, [sl, sl], (0, 1))
    b = ops.spmm(p, [x, x.clone()], (0, 1))
    assert torch.equal(a[0], b[0]) and torch.equal(a[1], b[1])
    # the layer, called the way the reference's node-classification example calls it
    conv = nn.MagNetConv(f, 32, K=2, q=0.2, trainable_q=False).to(DEV)
    o_r, o_i = port.magnet_conv(x.cpu(), x.cpu(), ei, None, conv.weight.detach().cpu(),
                                conv.bias.detach().cpu(), 0.2, "sym")
    out_r, out_i = conv(x, x, ei.to(DEV))
    assert_close_rel(out_r, o_r, 1e-5, "out_real")
    assert_close_rel(out_i, o_i, 1e-5, "out_imag")


def test_digcn_inception_model_golden_and_training():
    """BASELINE config 3 as a model: three inception blocks, x0 + x1 + x2 fused into the aggregation epilogues."""
    g = load_golden("digcn_ib_model", DEV)
    model = nn.DiGCN_Inception_Block_node_classification(7, 16, 4, dropout=0.5).to(DEV).eval()
    model.load_state_dict({k.replace("__", "."): v for k, v in g.items()
                           if k not in ("x", "out", "ei1", "w1", "ei2", "w2")})
    args = (g["x"], (g["ei1"], g["ei2"]), (g["w1"], g["w2"]))
    with torch.no_grad():
        y = model(*args)
    assert_close_rel(y, g["out"], 1e-5)
    # the unfused route (three tensors, explicit sum) agrees with the fused epilogue sum
    with torch.no_grad():
        x0, x1, x2 = model.ib1(g["x"], g["ei1"], g["w1"], g["ei2"], g["w2"])
        s = model.ib1.forward_sum(g["x"], g["ei1"], g["w1"], g["ei2"], g["w2"])
    assert_close_rel(s, x0 + x1 + x2, 2e-6)
    # training mode (dropout active): gradients reach every parameter; eval result unchanged afterwards
    model.train()
    out = model(*args)
    torch.nn.functional.nll_loss(out, torch.randint(0, 4, (out.size(0),), device=DEV)).backward()
    assert all(p.grad is not None and torch.isfinite(p.grad).all() for p in model.parameters())
    model.eval()
    # no-dropout training path differentiates through the fused sum: compare with torch autograd of the oracle
    m2 = nn.DiGCN_Inception_Block_node_classification(7, 16, 4, dropout=0.0).to(DEV).train()
    m2.load_state_dict(model.state_dict())
    m2(*args).sum().backward()
    blocks = []
    leaves = []
    for i in (1, 2, 3):
        blk = [g[f"ib{i}__{nm}"].cpu().clone().requires_grad_(True)
               for nm in ("ln__weight", "ln__bias", "conv1__weight", "conv1__bias", "conv2__weight", "conv2__bias")]
        blocks.append(tuple(blk))
        leaves += blk
    port.digcn_inception_model(g["x"].cpu(), g["ei1"].cpu(), g["w1"].cpu(), g["ei2"].cpu(), g["w2"].cpu(),
                               blocks).sum().backward()
    got = [p.grad for _, p in sorted(m2.named_parameters(), key=lambda kv: kv[0])]
    names = sorted(n for n, _ in m2.named_parameters())
    ref = {f"ib{i}.{nm.replace('__', '.')}": t.grad for i, blk in zip((1, 2, 3), blocks)
           for nm, t in zip(("ln__weight", "ln__bias", "conv1__weight", "conv1__bias", "conv2__weight",
                             "conv2__bias"), blk)}
    for nme, gr in zip(names, got):
        assert_close_rel(gr, ref[nme], 5e-5, f"grad {nme}")


@pytest.mark.parametrize("name,norm_emb", [("sgcn_model", False), ("sgcn_model_norm", True)])
def test_sgcn_model_golden(name, norm_emb):
    """BASELINE config 4 as a model: sign split, conv1 + deep layers, tanh as the transform's epilogue."""
    g = load_golden(name, DEV)
    m = nn.SGCN(140, g["edge_index_s"], in_dim=12, out_dim=16, layer_num=3, init_emb=g["x"].clone(),
                norm_emb=norm_emb).to(DEV).eval()
    assert torch.equal(m.pos_edge_index, g["pos_edge_index"]) and torch.equal(m.neg_edge_index, g["neg_edge_index"])
    res = m.load_state_dict({k.replace("__", "."): v for k, v in g.items()
                             if k not in ("out", "edge_index_s", "pos_edge_index", "neg_edge_index")}, strict=False)
    # the fixture holds the layers' parameters; the loss head of the reference (SGCN.py:75, lsp_loss.lin) exists here
    # as a parameter container so that full reference state_dicts load with strict=True
    assert not res.unexpected_keys and sorted(res.missing_keys) == ["lsp_loss.lin.bias", "lsp_loss.lin.weight"]
    with torch.no_grad():
        z = m()
    assert_close_rel(z, g["out"], 1e-5)
    # with gradients required the tanh runs unfused; same values
    m.train()
    for p in m.parameters():
        p.requires_grad_(True)
    z2 = m()
    assert_close_rel(z2, g["out"], 1e-5)
    z2.sum().backward()
    assert all(p.grad is not None for n_, p in m.named_parameters() if n_ != "x" and not n_.startswith("lsp_loss"))
    with pytest.raises(NotImplementedError):
        nn.SGCN(140, g["edge_index_s"], in_dim=12, out_dim=16)


# ------------------------------------------------ BASELINE configs 3 and 4 at full size

def _rows_subproblem(edge_index, sel, n):
    """Edges whose destination (edge_index[1]) is in `sel`, relabelled so that the oracle can run on a small
    graph: returns (sub_edge_index on compact ids, mask of kept edges, compact id -> original id, positions of
    sel inside the compact ids).  Rows outside `sel` of the small problem are NOT comparable (their in-edges were
    dropped); rows in `sel` see exactly their original in-neighbourhood."""
    keep = torch.isin(edge_index[1], sel)
    sub = edge_index[:, keep]
    ids, inv = torch.unique(torch.cat([sel, sub[0], sub[1]]), return_inverse=True)
    k = sel.numel()
    pos_sel = inv[:k]
    sub_c = torch.stack([inv[k:k + sub.size(1)], inv[k + sub.size(1):]])
    return sub_c, keep, ids, pos_sel


def test_sgcn_full_size_config4():
    """BASELINE config 4: SSBM 2M nodes / 40M signed entries / 64 -> 32|32.  Integer in-degree counts exact
    (plan row lengths vs bincount), mean aggregation linear, and 2000 random destination rows against the
    oracle evaluated on their induced in-neighbourhood sub-problem."""
    n = 2_000_000
    pos, neg, _ = synthetic.ssbm_edges(n, 3, num_entries=40_000_000, eta=0.1, seed=0, device=DEV)
    x = torch.randn(n, 64, device=DEV, generator=torch.Generator(device=DEV).manual_seed(0))
    conv = nn.SGCNConv(64, 32, first_aggr=True).to(DEV)
    with torch.no_grad():
        out = conv(x, pos, neg)
    assert out.shape == (n, 64) and torch.isfinite(out).all()
    for ei in (pos, neg):
        p = conv._plan_for(ei, n, n)
        assert torch.equal((p.row_ptr[1:] - p.row_ptr[:-1]).long(), torch.bincount(ei[1], minlength=n))
        assert p.nnz == ei.size(1)
    p = conv._plan_for(pos, n, n)
    y = torch.randn(n, 64, device=DEV, generator=torch.Generator(device=DEV).manual_seed(1))
    a = ops.spmm(p, [x], (0,), mean=True)[0]
    b = ops.spmm(p, [y], (0,), mean=True)[0]
    c = ops.spmm(p, [0.5 * x - 2.0 * y], (0,), mean=True)[0]
    assert_close_rel(c, 0.5 * a - 2.0 * b, 2e-6, "linearity of the mean aggregation")
    sel = torch.randperm(n, generator=torch.Generator().manual_seed(2))[:2000].to(DEV)
    sp, _, ids_p, pos_p = _rows_subproblem(pos, sel, n)
    # one compact id space for both signs: run the two signs as separate sub-problems sharing `sel`
    sn, _, ids_n, pos_n = _rows_subproblem(neg, sel, n)
    wb, bb = conv.lin_b.weight.detach().cpu(), conv.lin_b.bias.detach().cpu()
    wu, bu = conv.lin_u.weight.detach().cpu(), conv.lin_u.bias.detach().cpu()
    empty = torch.zeros((2, 0), dtype=torch.long)
    ref_b = port.sgcn_conv(x[ids_p].cpu(), sp.cpu(), empty, wb, bb, wu, bu, first_aggr=True, norm_emb=False)
    ref_u = port.sgcn_conv(x[ids_n].cpu(), empty, sn.cpu(), wb, bb, wu, bu, first_aggr=True, norm_emb=False)
    got = out[sel].cpu()
    scale = out.abs().max().item()
    assert (got[:, :32] - ref_b[pos_p.cpu(), :32]).abs().max().item() <= 1e-5 * scale
    assert (got[:, 32:] - ref_u[pos_n.cpu(), 32:]).abs().max().item() <= 1e-5 * scale


def test_inception_block_full_size_config3():
    """BASELINE config 3: DiGCN_InceptionBlock, 500k nodes / 2 x 10M weighted entries / 128 features, bf16.
    x0 against a dense fp32 product on sampled rows, x1 / x2 on 2000 random rows against the oracle on the induced
    sub-problem (bf16 tolerance 1e-2: inputs, xW and outputs are each rounded once), aggregation linearity."""
    n, e, f = 500_000, 10_000_000, 128
    ei1, _ = synthetic.dsbm_edges(n, 3, num_edges=e, seed=1, device=DEV)
    ei2, _ = synthetic.dsbm_edges(n, 3, num_edges=e, seed=2, device=DEV)
    w1, w2 = synthetic.sym_norm_weights(ei1, n), synthetic.sym_norm_weights(ei2, n)
    x = (torch.rand(n, f, device=DEV, generator=torch.Generator(device=DEV).manual_seed(0)) * 2 - 1).bfloat16()
    blk = nn.DiGCN_InceptionBlock(f, f).to(DEV)
    with torch.no_grad():
        blk.conv1.bias.uniform_(-0.2, 0.2)
        blk.conv2.bias.uniform_(-0.2, 0.2)
        x0, x1, x2 = blk(x, ei1, w1, ei2, w2)
        s = blk.forward_sum(x, ei1, w1, ei2, w2)
    for t in (x0, x1, x2, s):
        assert t.dtype == torch.bfloat16 and t.shape == (n, f) and torch.isfinite(t.float()).all()
    assert_close_rel(s.float(), x0.float() + x1.float() + x2.float(), 2e-2, "fused x0 + x1 + x2")
    sel = torch.randperm(n, generator=torch.Generator().manual_seed(3))[:2000].to(DEV)
    xf = x.float()
    ref0 = torch.nn.functional.linear(xf[sel], blk.ln.weight.detach(), blk.ln.bias.detach())
    assert_close_rel(x0[sel].float(), ref0, 1e-2, "x0")
    for ei, w, conv, got in ((ei1, w1, blk.conv1, x1), (ei2, w2, blk.conv2, x2)):
        sub, keep, ids, pos_sel = _rows_subproblem(ei, sel, n)
        ref = port.digcn_conv(xf[ids].cpu(), sub.cpu(), w[keep].cpu(), conv.weight.detach().cpu(),
                              conv.bias.detach().cpu())[pos_sel.cpu()]
        assert_close_rel(got[sel].float().cpu(), ref, 1e-2, "aggregated rows")
    # linearity of the bf16 aggregation (fp32 accumulation, one rounding at the end)
    p = blk.conv1._plan
    a = ops.spmm(p, [x], (0,))[0].float()
    y = (torch.rand(n, f, device=DEV) * 2 - 1).bfloat16()
    b = ops.spmm(p, [y], (0,))[0].float()
    c = ops.spmm(p, [(x.float() + y.float()).bfloat16()], (0,))[0].float()
    assert_close_rel(c, a + b, 2e-2, "linearity")


def test_uncached_layer_reuses_its_plan_only_for_identical_tensors():
    """cached=False (the reference's default) re-normalises every call; here the plan of the previous call is
    reused while the very same edge tensors come back unchanged, and rebuilt as soon as anything differs --
    outputs always follow the tensors passed in."""
    g = torch.Generator().manual_seed(11)
    n, e, f = 3000, 40_000, 16
    ei = torch.randint(0, n, (2, e), generator=g).to(DEV)
    ew = (torch.rand(e, generator=g) + 0.5).to(DEV)
    x = (torch.rand(n, f, generator=g) * 2 - 1).to(DEV)
    conv = nn.MagNetConv(f, f, K=2, q=0.2, trainable_q=False, cached=False).to(DEV)

    def oracle(weights):
        return port.magnet_conv(x.cpu(), x.cpu(), ei.cpu(), weights.cpu(), conv.weight.detach().cpu(),
                                conv.bias.detach().cpu(), 0.2, "sym")
    a = conv(x, x, ei, ew)
    p1 = conv._plan
    b = conv(x, x, ei, ew)
    assert conv._plan is p1 and torch.equal(a[0], b[0])                 # same tensors, untouched: reused
    assert_close_rel(a[0], oracle(ew)[0], 1e-5)
    ew.mul_(2.0)                                                         # in-place edit bumps the version counter
    c = conv(x, x, ei, ew)
    assert conv._plan is not p1
    assert_close_rel(c[0], oracle(ew)[0], 1e-5)
    assert_close_rel(c[1], oracle(ew)[1], 1e-5)
    p2 = conv._plan
    d = conv(x, x, ei.clone(), ew)                                       # equal values, different tensor: rebuilt
    assert conv._plan is not p2 and torch.equal(c[0], d[0])
    p3 = conv._plan
    conv(x, x, ei, ew, lambda_max=torch.tensor(1.5))                     # another operator from the same tensors
    assert conv._plan is not p3
    ref = nn.MagNetConv(f, f, K=2, q=0.2, trainable_q=False, cached=True).to(DEV)
    ref.load_state_dict(conv.state_dict())
    r = ref(x, x, ei, ew, lambda_max=torch.tensor(1.5))
    assert torch.equal(conv(x, x, ei, ew, lambda_max=torch.tensor(1.5))[0], r[0])


@pytest.mark.parametrize("n,e,c", [(4000, 60_000, 64), (3000, 20_000, 12), (2000, 30_000, 128), (500, 0, 8)])
def test_gat_aggregation_with_inkernel_softmax(n, e, c):
    """`pgsd_gat_aggregate` (softmax inside the aggregation kernel) against the two-kernel route (edge softmax ->
    alpha array -> weighted aggregation) and against the oracle's GATConv; also the accumulating epilogue."""
    g = torch.Generator().manual_seed(n + c)
    ei = torch.randint(0, n - 7, (2, e), generator=g)
    if e:
        ei[1, :600] = 3                                                  # one long row (several batches per group)
    x = torch.randn(n, c, generator=g)
    gat = nn.GATConv(c, c).to(DEV)
    with torch.no_grad():
        gat.bias.uniform_(-0.3, 0.3)
    ref = port.gat_conv(x, ei, gat.lin.weight.detach().cpu(), gat.att_src.detach().cpu().view(-1),
                        gat.att_dst.detach().cpu().view(-1), gat.bias.detach().cpu())
    xd, eid = x.to(DEV), ei.to(DEV)
    old = ops.GAT_FUSED
    torch.set_grad_enabled(False)              # the inference kernels (with gradients the layer takes autograd.py)
    try:
        ops.GAT_FUSED = 1
        fused = gat(xd, eid)
        acc = torch.randn(n, c, generator=g).to(DEV)
        acc0 = acc.clone()
        hs, ss = nn.sdr_layer.gat_transforms(xd, [gat])
        gat.aggregate(hs[0], ss[0][0], ss[0][1], eid, accumulate_into=acc)
        ops.GAT_FUSED = 0
        two = gat(xd, eid)
    finally:
        ops.GAT_FUSED = old
        torch.set_grad_enabled(True)
    assert_close_rel(fused, ref, 1e-5, "fused vs oracle")
    assert_close_rel(fused, two, 2e-6, "fused vs two-kernel route")
    assert_close_rel(acc - acc0, fused - gat.bias.detach(), 2e-6, "accumulating epilogue")


def test_plan_cache_follows_content_not_only_identity(monkeypatch):
    """a layer that re-normalises on every call upstream (MagNetConv cached=False, MagNetConv.py:158-181)
    reuses its last plan only while the edge tensors' CONTENT is unchanged -- a write through `.data` (no version
    bump, same storage) must be seen; PGSD_PLAN_REUSE=off rebuilds every call."""
    g = torch.Generator().manual_seed(3)
    n, e, f = 2000, 30_000, 16
    ei = torch.randint(0, n, (2, e), generator=g).to(DEV)
    xr, xi = torch.randn(n, f, generator=g).to(DEV), torch.randn(n, f, generator=g).to(DEV)
    conv = nn.MagNetConv(f, f, K=1, q=0.25, trainable_q=False, cached=False).to(DEV)
    with torch.no_grad():
        a = conv(xr, xi, ei)
        plan_a = conv._plan
        conv(xr, xi, ei)
        assert conv._plan is plan_a                                  # identical tensors: the plan is reused
        version = ei._version
        ei.data[:, :5000] = torch.randint(0, n, (2, 5000), generator=g).to(DEV)
        assert ei._version == version                                # the write is invisible to the version counter
        b = conv(xr, xi, ei)
        assert conv._plan is not plan_a
        fresh = nn.MagNetConv(f, f, K=1, q=0.25, trainable_q=False, cached=False).to(DEV)
        fresh.load_state_dict(conv.state_dict())
        c = fresh(xr, xi, ei)
        assert_close_rel(b[0], c[0], 1e-6, "result follows the modified edges")
        assert (a[0] - b[0]).abs().max().item() > 1e-3
        monkeypatch.setenv("PGSD_PLAN_REUSE", "off")
        plan_b = conv._plan
        conv(xr, xi, ei)
        assert conv._plan is not plan_b
        conv.cached_result = None
        assert conv._plan is None and not conv._rebuild_cache._items


def test_magnet_chebyshev_order_above_seven():
    """VERDICT r1 missing #5: the reference's loop (MagNetConv.py:213-240) has no limit on K; orders above 7 need more
    than the transform's 16 terms per launch and are summed chunk by chunk (also with the fused complex ReLU)."""
    g = torch.Generator().manual_seed(9)
    n, e, f = 1500, 12_000, 16
    ei = torch.randint(0, n, (2, e), generator=g)
    xr, xi = torch.rand(n, f, generator=g) * 2 - 1, torch.rand(n, f, generator=g) * 2 - 1
    conv = nn.MagNetConv(f, f, K=9, q=0.15, trainable_q=False).to(DEV)
    with torch.no_grad():
        conv.bias.uniform_(-0.2, 0.2)
        out_r, out_i = conv(xr.to(DEV), xi.to(DEV), ei.to(DEV))
    ref_r, ref_i = port.magnet_conv(xr, xi, ei, None, conv.weight.detach().cpu(), conv.bias.detach().cpu(), 0.15, "sym")
    assert_close_rel(out_r, ref_r, 2e-5, "K = 9 out_real")          # ten chained recurrences: rounding accumulates
    assert_close_rel(out_i, ref_i, 2e-5, "K = 9 out_imag")
    a = xr.to(DEV).requires_grad_(True)
    o_r, o_i = conv(a, xi.to(DEV), ei.to(DEV))
    (o_r.sum() + o_i.sum()).backward()
    assert a.grad is not None and conv.weight.grad is not None and torch.isfinite(conv.weight.grad).all()

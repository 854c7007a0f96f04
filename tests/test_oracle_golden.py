"""CPU: pin oracle/port.py (the checker used by every GPU parity test) against fixtures made
from the reference's own source files (tests/golden/make_golden.py)."""
import pytest
import torch

from conftest import assert_close_rel, load_golden
from oracle import port

MAGNET_CASES = ["magnet_c1", "magnet_k2_weighted", "magnet_k3_none", "magnet_q0",
                "magnet_sym_lmax", "msconv_signed", "msconv_nonabs_none", "magnet_k9"]


def _magnet_kwargs(name, g):
    return dict(q=g["q"], normalization="sym" if g["sym"] else None,
                lambda_max=None if g["lambda_max"] < 0 else g["lambda_max"],
                signed=name.startswith("msconv"),
                absolute_degree=(name != "msconv_nonabs_none"))


@pytest.mark.parametrize("name", MAGNET_CASES)
def test_magnet_port_matches_reference(name):
    g = load_golden(name)
    ew = g["edge_weight"] if g["has_weight"] else None
    kw = _magnet_kwargs(name, g)
    n = g["x_real"].size(0)
    lam = 2.0 if kw["lambda_max"] is None else kw["lambda_max"]
    cr = port.magnet_norm(g["edge_index"], ew, n, kw["q"], kw["normalization"], lam,
                          torch.float32, kw["signed"], kw["absolute_degree"])
    # integer artefacts: bit exact
    assert torch.equal(cr[0], g["cached_edge_index_real"])
    assert torch.equal(cr[1], g["cached_edge_index_imag"])
    assert_close_rel(cr[2], g["cached_norm_real"], 1e-6, "norm_real")
    assert_close_rel(cr[3], g["cached_norm_imag"], 1e-6, "norm_imag")
    o_r, o_i = port.magnet_conv(g["x_real"], g["x_imag"], g["edge_index"], ew, g["weight"],
                                g["bias"], **kw)
    assert_close_rel(o_r, g["out_real"], 1e-5, "out_real")
    assert_close_rel(o_i, g["out_imag"], 1e-5, "out_imag")
    # fp64 sanity band: the fp32 reference itself sits within 1e-5 of an fp64 evaluation
    d = lambda t: None if t is None else t.double()
    o_r64, o_i64 = port.magnet_conv(d(g["x_real"]), d(g["x_imag"]), g["edge_index"], d(ew),
                                    d(g["weight"]), d(g["bias"]), **kw)
    assert_close_rel(g["out_real"], o_r64, 2e-5, "fp64 band real")
    assert_close_rel(g["out_imag"], o_i64, 2e-5, "fp64 band imag")


@pytest.mark.parametrize("tag,norm", [("sym", "sym"), ("none", None)])
def test_laplacian_port(tag, norm):
    g = load_golden(f"laplacian_{tag}")
    ei, wr, wi = port.magnetic_laplacian(g["edge_index"], g["edge_weight"], 70, g["q"], norm)
    assert torch.equal(ei, g["out_edge_index"])
    assert_close_rel(wr, g["out_real"], 1e-6, "real")
    assert_close_rel(wi, g["out_imag"], 1e-6, "imag")


def test_q11_real_part_of_one_way_edge_is_tiny_but_nonzero():
    ei = torch.tensor([[0], [1]])
    _, wr, wi = port.magnetic_laplacian(ei, None, 2, 0.25, "sym")
    assert 0 < abs(wr[0].item()) < 1e-6 and abs(abs(wi[0].item()) - 1.0) < 1e-6


def test_digcn_port():
    g = load_golden("digcn_conv")
    y = port.digcn_conv(g["x"], g["edge_index"], g["edge_weight"], g["weight"], g["bias"])
    assert_close_rel(y, g["out"], 1e-5)
    g = load_golden("digcn_inception")
    x0, x1, x2 = port.digcn_inception_block(
        g["x"], g["edge_index"], g["edge_weight"], g["edge_index2"], g["edge_weight2"],
        g["ln_weight"], g["ln_bias"], g["conv1_weight"], g["conv1_bias"],
        g["conv2_weight"], g["conv2_bias"])
    for a, b in ((x0, g["x0"]), (x1, g["x1"]), (x2, g["x2"])):
        assert_close_rel(a, b, 1e-5)


@pytest.mark.parametrize("name,first,norm_emb", [("sgcn_first", True, False),
                                                 ("sgcn_second", False, True)])
def test_sgcn_port(name, first, norm_emb):
    g = load_golden(name)
    y = port.sgcn_conv(g["x"], g["pos_edge_index"], g["neg_edge_index"], g["lin_b_weight"],
                       g["lin_b_bias"], g["lin_u_weight"], g["lin_u_bias"], first, norm_emb)
    assert_close_rel(y, g["out"], 1e-5)


def test_conv_base_and_dimpa_port():
    g = load_golden("conv_norm_rw")
    ei, w = port.conv_norm_rw(g["edge_index"], g["fill_value"], g["edge_weight"], 110)
    assert torch.equal(ei, g["out_edge_index"])
    assert_close_rel(w, g["out_weight"], 1e-6)
    g = load_golden("conv_base")
    assert_close_rel(port.conv_base(g["x"], g["edge_index"], g["edge_weight"], 0.5), g["out"])
    assert_close_rel(port.conv_base(g["x"], g["edge_index"], None, 0.25),
                     g["out_unweighted_fill025"])
    g = load_golden("dimpa")
    y = port.dimpa(g["x_s"], g["x_t"], g["edge_index"], g["edge_weight"], g["w_s"], g["w_t"], 2)
    assert_close_rel(y, g["out"], 1e-5)


def test_complex_relu_port():
    g = load_golden("complex_relu")
    r, i = port.complex_relu(g["real"], g["imag"])
    assert torch.equal(r, g["out_real"]) and torch.equal(i, g["out_imag"])


def test_row_subset_evaluator_matches_full():
    g = load_golden("magnet_c1")
    n = g["x_real"].size(0)
    cr = port.magnet_norm(g["edge_index"], None, n, g["q"], "sym", 2.0)
    rows = torch.tensor([0, 5, 17, 999, 512])
    o_r, o_i = port.magnet_conv_rows(rows, g["x_real"], g["x_imag"], cr, g["weight"], g["bias"])
    assert_close_rel(o_r, g["out_real"][rows], 1e-5)
    assert_close_rel(o_i, g["out_imag"][rows], 1e-5)


@pytest.mark.parametrize("name,first", [("snea_first", True), ("snea_second", False)])
def test_snea_port(name, first):
    g = load_golden(name)
    y = port.snea_conv(g["x"], g["pos_edge_index"], g["neg_edge_index"], g["lin_b_weight"], g["lin_b_bias"],
                       g["lin_u_weight"], g["lin_u_bias"], g["alpha_b_weight"], g["alpha_b_bias"],
                       g["alpha_u_weight"], g["alpha_u_bias"], first)
    assert_close_rel(y, g["out"], 1e-5)
    if first:
        # quirk Q7: the message is the target's own feature times alpha, so the first layer is
        # lin(x) on every node that has a (self-)loop, and exactly 0 on ids above the largest edge id
        h = torch.nn.functional.linear(g["x"], g["lin_b_weight"], g["lin_b_bias"])
        assert_close_rel(g["out"][:134, :6], h[:134], 1e-5)
        assert float(g["out"][134:].abs().max()) == 0.0


def test_dgcn_and_simpa_port():
    g = load_golden("dgcn_conv")
    assert_close_rel(port.dgcn_conv(g["x"], g["edge_index"], g["edge_weight"]), g["out"], 1e-5)
    assert_close_rel(port.dgcn_conv(g["x"], g["edge_index"], None, improved=True), g["out_improved_unweighted"], 1e-5)
    g = load_golden("simpa")
    args = (g["edge_index_p"], g["edge_weight_p"], g["edge_index_n"], g["edge_weight_n"], g["x_p"], g["x_n"])
    assert_close_rel(port.simpa(*args, (g["w_p"], g["w_n"]), 2, 0.5), g["out_undirected"], 1e-5)
    assert_close_rel(port.simpa(*args, (g["w_sp"], g["w_sn"], g["w_tp"], g["w_tn"]), 2, 0.5, g["x_pt"], g["x_nt"]),
                     g["out_directed"], 1e-5)


def _sdr_params(g):
    gat = [(g[f"agg_{i}__lin__weight"], g[f"agg_{i}__att_src"], g[f"agg_{i}__att_dst"], g[f"agg_{i}__bias"])
           for i in range(4)]
    return [g[f"edges_{i}"] for i in range(4)], gat


def test_sdr_layer_port():
    g = load_golden("sdr_layer")
    lists, gat = _sdr_params(g)
    y = port.sdr_layer(g["x"], lists, gat, g["mlp_layer__0__weight"], g["mlp_layer__0__bias"],
                       g["mlp_layer__2__weight"], g["mlp_layer__2__bias"])
    assert_close_rel(y, g["out"], 1e-5)


def test_magnet_model_port():
    g = load_golden("magnet_model")
    chebs = [(g[f"Chebs__{i}__weight"], g[f"Chebs__{i}__bias"]) for i in range(3)]
    y = port.magnet_node_classification(g["x"], g["x"], g["edge_index"], g["edge_weight"], chebs,
                                        g["Conv__weight"], g["Conv__bias"], 0.2)
    assert_close_rel(y, g["out"], 1e-5)


@pytest.mark.parametrize("name,signed,norm", [("qgrad_magnet_k2", False, "sym"), ("qgrad_magnet_k1_none", False, None),
                                              ("qgrad_msconv_k3", True, "sym"), ("qgrad_magnet_clamped", False, "sym")])
def test_trainable_q_gradient_port(name, signed, norm):
    """autograd of the port w.r.t. a trainable q == autograd of the reference (golden d_q)."""
    g = load_golden(name)
    q = g["q"].clone().requires_grad_(True)
    lam = float(g["lambda_max"])
    o_r, o_i = port.magnet_conv(g["x_real"], g["x_imag"], g["edge_index"], g["edge_weight"], g["weight"], g["bias"],
                                q, norm, lambda_max=None if lam < 0 else lam, signed=signed)
    ((o_r * g["r1"]).sum() + (o_i * g["r2"]).sum()).backward()
    assert_close_rel(o_r, g["out_real"], 1e-5)
    assert_close_rel(q.grad, g["d_q"], 1e-4)


# ------------------------------------------------------------------ preprocessing
@pytest.mark.parametrize("name", ["prep_a", "prep_b", "prep_c"])
def test_port_preprocessing_matches_reference_golden(name):
    """oracle/port.py's sparse restatements of get_appr_directed_adj / get_second_directed_adj /
    directed_features_in_out against the reference's own dense-eig / dense-mm / scipy-loop outputs."""
    g = load_golden(name)
    ei, n = g["edge_index"], int(g["n"])
    ew = g["edge_weight"] if g["has_weight"] else None
    a_ei, a_w = port.appr_directed_adj(float(g["alpha"]), ei, n, ew)
    assert torch.equal(a_ei, g["appr_index"])
    # the reference's pi is a float32 LAPACK eigenvector: agreement to a few fp32 ulps of the largest entry
    assert_close_rel(a_w, g["appr_weight"], 2e-5, "appr weights")
    s_ei, s_w = port.second_directed_adj(ei, n, ew)
    assert torch.equal(s_ei, g["second_index"])
    assert_close_rel(s_w, g["second_weight"], 1e-5, "second-order weights")
    und, e_in, w_in, e_out, w_out = port.features_in_out(ei, n, ew)
    assert torch.equal(und, g["undirected"])
    assert torch.equal(e_in, g["in_index"]) and torch.equal(e_out, g["out_index"])
    assert_close_rel(w_in, g["in_weight"], 1e-5, "A_in")
    assert_close_rel(w_out, g["out_weight"], 1e-5, "A_out")


def test_digcn_inception_model_port():
    g = load_golden("digcn_ib_model")
    blocks = [tuple(g[f"ib{i}__{nm}"] for nm in ("ln__weight", "ln__bias", "conv1__weight", "conv1__bias",
                                                 "conv2__weight", "conv2__bias")) for i in (1, 2, 3)]
    y = port.digcn_inception_model(g["x"], g["ei1"], g["w1"], g["ei2"], g["w2"], blocks)
    assert_close_rel(y, g["out"], 1e-5)


@pytest.mark.parametrize("name,norm_emb", [("sgcn_model", False), ("sgcn_model_norm", True)])
def test_sgcn_model_port(name, norm_emb):
    g = load_golden(name)
    layers = [tuple(g[f"{pre}__{nm}"] for nm in ("lin_b__weight", "lin_b__bias", "lin_u__weight", "lin_u__bias"))
              for pre in ("conv1", "convs__0", "convs__1")]
    z = port.sgcn_model(g["x"], g["edge_index_s"], layers, norm_emb=norm_emb)
    assert_close_rel(z, g["out"], 1e-5)
    # integer plumbing of SGCN.py:53-54 is bit-exact
    e = g["edge_index_s"]
    assert torch.equal(e[e[:, 2] > 0][:, :2].t(), g["pos_edge_index"])
    assert torch.equal(e[e[:, 2] < 0][:, :2].t(), g["neg_edge_index"])


def test_motif_mining_port_matches_reference_lists():
    """SDGNN / SiGAT build_adj_lists: the port's set loops reproduce the reference's adjacency lists (as sorted
    edge sets) and SDGNN's triangle weights exactly."""
    g = load_golden("sdgnn_model")
    lists, tri = port.sdgnn_motifs(g["edge_index_s"], 90)
    for i, e in enumerate(lists):
        assert torch.equal(e, g[f"list_{i}"])
    assert [t[0] for t in tri] == g["tri_row"].tolist() and [t[1] for t in tri] == g["tri_col"].tolist()
    assert [t[2] for t in tri] == g["tri_val"].tolist()
    s = load_golden("sigat_model")
    for i, e in enumerate(port.sigat_motifs(s["edge_index_s"], 90)):
        assert torch.equal(e, s[f"list_{i}"]), f"SiGAT list {i}"


def _gat_params(g, prefix, k):
    return [(g[f"{prefix}agg_{i}__lin__weight"], g[f"{prefix}agg_{i}__att_src"], g[f"{prefix}agg_{i}__att_dst"],
             g[f"{prefix}agg_{i}__bias"]) for i in range(k)]


def test_sdgnn_and_sigat_forward_port():
    g = load_golden("sdgnn_model")
    lists = [g[f"list_{i}"] for i in range(4)]
    x = g["x"]
    for l in range(2):
        pre = f"SDRLayer_{l}__"
        x = port.sdr_layer(x, lists, _gat_params(g, pre, 4), g[pre + "mlp_layer__0__weight"],
                           g[pre + "mlp_layer__0__bias"], g[pre + "mlp_layer__2__weight"], g[pre + "mlp_layer__2__bias"])
    assert_close_rel(x, g["out"], 1e-5)
    s = load_golden("sigat_model")
    lists = [s[f"list_{i}"] for i in range(38)]
    z = port.sigat_forward(s["x"], lists, _gat_params(s, "", 38), s["mlp_layer__0__weight"], s["mlp_layer__0__bias"],
                           s["mlp_layer__2__weight"], s["mlp_layer__2__bias"])
    assert_close_rel(z, s["out"], 1e-5)


@pytest.mark.parametrize("norm,lam", [("sym", 2.0), (None, 3.7)])
@pytest.mark.parametrize("weighted", [False, True])
def test_row_restricted_operator_matches_the_full_oracle(norm, lam, weighted):
    """`port.magnet_norm_rows` (used by the full-size parity checks and bench.py's in-line parity_check) keeps
    exactly the entries of `port.magnet_norm` whose target is in `rows` -- indices bit-exact and in the same
    order, values to fp32 rounding (its degree is the same sum in another association)."""
    g = torch.Generator().manual_seed(11)
    n, e, f = 700, 6000, 8
    ei = torch.randint(0, n, (2, e), generator=g)
    ei[:, :40] = ei[:, 40:80]                      # duplicates
    ei[:, 80:120] = ei[:, 120:160].flip(0)         # reciprocal pairs
    ei[1, 160:170] = ei[0, 160:170]                # self-loops
    w = torch.rand(e, generator=g) + 0.1 if weighted else None
    rows = torch.randperm(n, generator=g)[:64]
    full = port.magnet_norm(ei, w, n, 0.2, norm, lam)
    sub = port.magnet_norm_rows(rows, ei, w, n, 0.2, norm, lam)
    nnz = full[1].size(1) - n
    sel = torch.zeros(n, dtype=torch.bool)
    sel[rows] = True
    keep = sel[full[1][1, :nnz]]
    m = int(keep.sum())
    assert torch.equal(full[1][:, :nnz][:, keep], sub[1][:, :m])
    assert torch.equal(sub[1][:, m:], torch.stack([rows, rows]))
    for k in (2, 3):
        ref = full[k][:nnz][keep]
        assert (ref - sub[k][:m]).abs().max() <= 1e-6 * ref.abs().max().clamp(min=1e-30)
    xr, xi = torch.rand(n, f, generator=g), torch.rand(n, f, generator=g)
    wt, b = torch.rand(2, f, f, generator=g) - 0.5, torch.rand(f, generator=g)
    o_full = port.magnet_conv(xr, xi, ei, w, wt, b, 0.2, norm, lam)
    o_rows = port.magnet_conv_rows(rows, xr, xi, sub, wt, b)
    for a, r in zip(o_full, o_rows):
        assert (a[rows] - r).abs().max() <= 2e-6 * a.abs().max()


def _rel(a, b):
    return float((a - b).abs().max() / b.abs().max().clamp(min=1e-30))


def test_oracle_attention_gradients_match_the_reference_autograd():
    """The attention layers train upstream (SNEAConv.py:135-146, SDGNN.py:57-64).  Fixtures from torch autograd through the
    reference's OWN source files (tests/golden/make_golden_attn_grad.py) pin the gradients of oracle/port.py -- which the
    GPU tests then use as the yardstick for the CUDA backward kernels."""
    g = load_golden("snea_grad")
    x = g["x"].clone().requires_grad_(True)
    prm = {}
    for tag in ("c1", "c2"):
        prm[tag] = [g[f"{tag}__{nm}__{wb}"].clone().requires_grad_(True)
                    for nm in ("lin_b", "lin_u", "alpha_b", "alpha_u") for wb in ("weight", "bias")]
    order = lambda p: [p[0], p[1], p[2], p[3], p[4], p[5], p[6], p[7]]   # lin_b w,b, lin_u w,b, alpha_b w,b, alpha_u w,b
    pos, neg = g["pos_edge_index"], g["neg_edge_index"]
    z1 = port.snea_conv(x, pos, neg, *order(prm["c1"]), True)
    out = port.snea_conv(torch.tanh(z1), pos, neg, *order(prm["c2"]), False)
    assert _rel(out.detach(), g["out"]) <= 2e-6
    (out * g["r"]).sum().backward()
    assert _rel(x.grad, g["grad_x"]) <= 1e-5
    names = [f"{nm}__{wb}" for nm in ("lin_b", "lin_u", "alpha_b", "alpha_u") for wb in ("weight", "bias")]
    for tag in ("c1", "c2"):
        for nm, p in zip(names, prm[tag]):
            ref = g[f"{tag}__{nm}__grad"]
            if float(ref.abs().max()) < 1e-4:          # first layer's attention: softmax weights of one type sum to 1
                assert float(p.grad.abs().max()) < 1e-4
            else:
                assert _rel(p.grad, ref) <= 2e-5, f"{tag} {nm}"

    s = load_golden("sdr_layer_grad")
    x = s["x"].clone().requires_grad_(True)
    lists = [s[f"edges_{i}"] for i in range(4)]
    leaf = lambda k: s[k].clone().requires_grad_(True)
    gat = [(leaf(f"agg_{i}__lin__weight"), leaf(f"agg_{i}__att_src"), leaf(f"agg_{i}__att_dst"), leaf(f"agg_{i}__bias"))
           for i in range(4)]
    mlp = [leaf("mlp_layer__0__weight"), leaf("mlp_layer__0__bias"), leaf("mlp_layer__2__weight"), leaf("mlp_layer__2__bias")]
    y = port.sdr_layer(x, lists, [(w, a.view(-1), b.view(-1), bb) for w, a, b, bb in gat], *mlp)
    assert _rel(y.detach(), s["out"]) <= 2e-6
    (y * s["r"]).sum().backward()
    assert _rel(x.grad, s["grad_x"]) <= 1e-5
    for i, (w, a, b, bb) in enumerate(gat):
        for nm, p in (("lin__weight", w), ("att_src", a), ("att_dst", b), ("bias", bb)):
            assert _rel(p.grad, s[f"agg_{i}__{nm}__grad"]) <= 2e-5, f"agg_{i} {nm}"
    for nm, p in zip(("mlp_layer__0__weight", "mlp_layer__0__bias", "mlp_layer__2__weight", "mlp_layer__2__bias"), mlp):
        assert _rel(p.grad, s[nm + "__grad"]) <= 2e-5, nm

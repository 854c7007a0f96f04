"""GPU parity of the sparse preprocessing utilities (`pytorch_geometric_signed_directed_b200.utils`)
against the reference's own outputs (tests/golden/prep_*.npz: dense eig / dense mm / scipy loops) and, at sizes
the dense reference cannot reach, against oracle/port.py's sparse CPU restatement.  Index tensors bit-exact
(row-major order = torch.nonzero of the reference's dense result), weights within 2e-5 * max (the reference's
stationary vector is a float32 LAPACK eigenvector)."""
import pytest
import torch

from conftest import assert_close_rel, load_golden
from oracle import port
from pytorch_geometric_signed_directed_b200 import nn, synthetic, utils

pytestmark = pytest.mark.gpu
DEV = "cuda"


@pytest.mark.parametrize("name", ["prep_a", "prep_b", "prep_c"])
def test_preprocessing_matches_reference_golden(name):
    g = load_golden(name, DEV)
    ei, n = g["edge_index"], int(g["n"])
    ew = g["edge_weight"] if g["has_weight"] else None
    a_ei, a_w = utils.get_appr_directed_adj(float(g["alpha"]), ei, n, torch.float32, ew)
    assert torch.equal(a_ei, g["appr_index"])
    assert_close_rel(a_w, g["appr_weight"], 2e-5, "appr weights")
    s_ei, s_w = utils.get_second_directed_adj(ei, n, torch.float32, ew)
    assert torch.equal(s_ei, g["second_index"])
    assert_close_rel(s_w, g["second_weight"], 1e-5, "second-order weights")
    und, e_in, w_in, e_out, w_out = utils.directed_features_in_out(ei, n, ew)
    assert torch.equal(und, g["undirected"])
    assert torch.equal(e_in, g["in_index"]) and torch.equal(e_out, g["out_index"])
    assert_close_rel(w_in, g["in_weight"], 1e-5, "A_in")
    assert_close_rel(w_out, g["out_weight"], 1e-5, "A_out")


def test_preprocessing_vs_sparse_oracle_at_20k_nodes():
    n, e = 20_000, 200_000
    ei, _ = synthetic.dsbm_edges(n, 3, num_edges=e, seed=4)
    g = torch.Generator().manual_seed(1)
    ew = torch.rand(ei.size(1), generator=g) + 0.5
    a_ei, a_w = utils.get_appr_directed_adj(0.1, ei.to(DEV), n, torch.float32, ew.to(DEV))
    r_ei, r_w = port.appr_directed_adj(0.1, ei, n, ew)
    assert torch.equal(a_ei.cpu(), r_ei)
    assert_close_rel(a_w, r_w, 1e-5, "appr")
    s_ei, s_w = utils.get_second_directed_adj(ei.to(DEV), n, torch.float32, ew.to(DEV))
    q_ei, q_w = port.second_directed_adj(ei, n, ew)
    assert torch.equal(s_ei.cpu(), q_ei)
    assert_close_rel(s_w, q_w, 1e-5, "second")
    got = utils.directed_features_in_out(ei.to(DEV), n, None)
    ref = port.features_in_out(ei, n, None)
    for a, b in zip(got, ref):
        if a.dtype == torch.int64:
            assert torch.equal(a.cpu(), b)
        else:
            assert_close_rel(a, b, 1e-5, "features_in_out")


def test_preprocessing_feeds_the_inception_block():
    """The real DiGCN operators (not synthetic stand-ins) drive DiGCN_InceptionBlock end to end; the operator
    of get_appr_directed_adj is symmetric and its weights are those of a symmetrically normalised matrix."""
    n, e = 30_000, 300_000
    ei, _ = synthetic.dsbm_edges(n, 3, num_edges=e, seed=7, device=DEV)
    ei1, w1 = utils.get_appr_directed_adj(0.1, ei, n, torch.float32)
    ei2, w2 = utils.get_second_directed_adj(ei, n, torch.float32)
    key = ei1[0] * n + ei1[1]
    key_t = ei1[1] * n + ei1[0]
    assert torch.equal(torch.sort(key).values, torch.sort(key_t).values)          # symmetric pattern
    assert bool((w1 > 0).all()) and bool((w2 > 0).all())
    x = torch.rand(n, 64, device=DEV) * 2 - 1
    blk = nn.DiGCN_InceptionBlock(64, 32).to(DEV)
    with torch.no_grad():
        x0, x1, x2 = blk(x, ei1, w1, ei2, w2)
    assert x0.shape == x1.shape == x2.shape == (n, 32)
    assert torch.isfinite(x1).all() and torch.isfinite(x2).all()

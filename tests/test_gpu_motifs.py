"""GPU motif mining (`pgsd_signed_triangle_counts`, utils/signed.py) and the SDGNN / SiGAT models built on it:
adjacency lists and triangle weights are integer artefacts and must equal the reference's exactly (as sorted edge
sets); model outputs within 1e-5 of the reference's forward (goldens from the reference's own files)."""
import pytest
import torch

from conftest import assert_close_rel, load_golden
from oracle import port
from pytorch_geometric_signed_directed_b200 import nn
from pytorch_geometric_signed_directed_b200.utils import signed as sg

pytestmark = pytest.mark.gpu
DEV = "cuda"


def _sorted(e):
    e = e.cpu()
    if e.numel() == 0:
        return e.reshape(2, 0)
    n = int(e.max()) + 1
    return e[:, torch.argsort(e[0] * n + e[1])]


def test_motif_lists_and_triangle_weights_match_reference_golden():
    g = load_golden("sdgnn_model", DEV)
    lists, (row, col, val) = sg.sdgnn_motifs(g["edge_index_s"], 90)
    for i, e in enumerate(lists):
        assert torch.equal(_sorted(e), g[f"list_{i}"].cpu()), f"SDGNN list {i}"
    assert torch.equal(row, g["tri_row"]) and torch.equal(col, g["tri_col"]) and torch.equal(val, g["tri_val"])
    s = load_golden("sigat_model", DEV)
    got = sg.sigat_motifs(s["edge_index_s"], 90)
    assert len(got) == 38
    for i, e in enumerate(got):
        assert torch.equal(_sorted(e), s[f"list_{i}"].cpu()), f"SiGAT list {i}"


@pytest.mark.parametrize("n,e,seed", [(300, 4000, 1), (50, 2000, 2), (1000, 3000, 3), (40, 0, 4)])
def test_motif_mining_vs_oracle_set_loops(n, e, seed):
    """Denser / sparser random signed graphs (long neighbour lists exercise the binary search over the longer
    list with lanes striding the shorter one) against the reference's algorithm restated with Python sets."""
    g = torch.Generator().manual_seed(seed)
    es = torch.stack([torch.randint(0, n, (e,), generator=g), torch.randint(0, n, (e,), generator=g),
                      torch.where(torch.rand(e, generator=g) < 0.45, -1, 1)], 1)
    if e:
        es[0, 2] = 0                                       # a zero-sign row is ignored by both passes
    ref_lists, ref_tri = port.sdgnn_motifs(es, n)
    lists, (row, col, val) = sg.sdgnn_motifs(es.to(DEV), n)
    for a, b in zip(lists, ref_lists):
        assert torch.equal(_sorted(a), b)
    assert list(zip(row.tolist(), col.tolist(), val.tolist())) == ref_tri
    for a, b in zip(sg.sigat_motifs(es.to(DEV), n), port.sigat_motifs(es, n)):
        assert torch.equal(_sorted(a), b)


def test_sdgnn_and_sigat_models_golden():
    g = load_golden("sdgnn_model", DEV)
    m = nn.SDGNN(90, g["edge_index_s"], in_dim=12, out_dim=12, layer_num=2, init_emb=g["x"].clone()).to(DEV).eval()
    m.load_state_dict({k.replace("__", "."): v for k, v in g.items()
                       if k.startswith(("SDRLayer_", "x"))})
    assert_close_rel(m(), g["out"], 1e-5, "SDGNN autograd path")
    with torch.no_grad():
        assert_close_rel(m(), g["out"], 1e-5, "SDGNN inference path")
    tw = m.tri_weight
    assert tw.shape == (90, 90) and int(tw.sum()) == int(g["tri_val"].sum())
    s = load_golden("sigat_model", DEV)
    sm = nn.SiGAT(90, s["edge_index_s"], in_dim=12, out_dim=12, init_emb=s["x"].clone()).to(DEV).eval()
    sm.load_state_dict({k.replace("__", "."): v for k, v in s.items() if k.startswith(("agg_", "mlp_layer", "x"))})
    assert_close_rel(sm(), s["out"], 1e-5, "SiGAT autograd path")
    with torch.no_grad():
        assert_close_rel(sm(), s["out"], 1e-5, "SiGAT inference path")
    # widths the vector kernels cannot write in place (out_dim = 10: 40-byte column blocks)
    torch.manual_seed(0)
    odd = nn.SiGAT(90, s["edge_index_s"], in_dim=10, out_dim=10, init_emb=torch.randn(90, 10, device=DEV)).to(DEV).eval()
    params = [(a.lin.weight.detach().cpu(), a.att_src.detach().cpu(), a.att_dst.detach().cpu(), a.bias.detach().cpu())
              for a in odd.aggs]
    ref = port.sigat_forward(odd.x.detach().cpu(), [e.cpu() for e in odd.edge_lists], params,
                             odd.mlp_layer[0].weight.detach().cpu(), odd.mlp_layer[0].bias.detach().cpu(),
                             odd.mlp_layer[2].weight.detach().cpu(), odd.mlp_layer[2].bias.detach().cpu())
    assert_close_rel(odd(), ref, 1e-5, "odd width, autograd path")
    with torch.no_grad():
        assert_close_rel(odd(), ref, 1e-5, "odd width, inference path")
    with pytest.raises(NotImplementedError):
        nn.SiGAT(90, s["edge_index_s"])

"""GPU: input and parameter gradients of the attention layers against fixtures made by torch autograd THROUGH THE REFERENCE'S OWN,
unmodified source files (tests/golden/make_golden_attn_grad.py -> snea_grad.npz, sdr_layer_grad.npz).  The oracle's gradients
are pinned to the same fixtures on the CPU (tests/test_oracle_golden.py), tests/test_gpu_autograd.py compares the CUDA backward
with the oracle on larger random graphs.  Tolerance 5e-5 * max|ref| per tensor: the fixtures are small (150 nodes, 6 features),
so single gradients are sums with heavy cancellation; the attention biases (one scalar each) are compared on the scale of
their module's weight gradient."""
import pytest
import torch

from conftest import assert_close_rel, load_golden
from pytorch_geometric_signed_directed_b200 import nn

pytestmark = pytest.mark.gpu
DEV = "cuda"
TOL = 5e-5


def _snea_from_fixture(g, tag, fi, fo, first, dev):
    conv = nn.SNEAConv(fi, fo, first_aggr=first).to(dev)
    with torch.no_grad():
        for nm in ("lin_b", "lin_u", "alpha_b", "alpha_u"):
            getattr(conv, nm).weight.copy_(g[f"{tag}__{nm}__weight"])
            getattr(conv, nm).bias.copy_(g[f"{tag}__{nm}__bias"])
    return conv


def _sdr_from_fixture(s, dev):
    lists = [s[f"edges_{i}"] for i in range(4)]
    layer = nn.SDRLayer(12, 12, lists).to(dev)
    state = {k.replace("__", "."): v for k, v in s.items()
             if not k.endswith("__grad") and k not in ("x", "r", "out", "grad_x") and not k.startswith("edges_")}
    layer.load_state_dict(state)
    return layer


def test_attention_gradients_match_reference_autograd_fixtures():
    """tests/golden/snea_grad.npz / sdr_layer_grad.npz (make_golden_attn_grad.py): input and parameter gradients of two
    SNEAConv layers and of an SDRLayer as torch autograd computes them through the reference's unmodified files."""
    g = load_golden("snea_grad", DEV)
    c1 = _snea_from_fixture(g, "c1", 8, 6, True, DEV)
    c2 = _snea_from_fixture(g, "c2", 6, 6, False, DEV)
    x = g["x"].clone().requires_grad_(True)
    out = c2(torch.tanh(c1(x, g["pos_edge_index"], g["neg_edge_index"])), g["pos_edge_index"], g["neg_edge_index"])
    (out * g["r"]).sum().backward()
    assert_close_rel(out, g["out"], 1e-5, "SNEAConv forward")
    assert_close_rel(x.grad, g["grad_x"], TOL, "SNEAConv d x")
    for tag, conv in (("c1", c1), ("c2", c2)):
        for nm in ("lin_b", "lin_u", "alpha_b", "alpha_u"):
            for wb in ("weight", "bias"):
                ref = g[f"{tag}__{nm}__{wb}__grad"]
                got = getattr(getattr(conv, nm), wb).grad
                if float(ref.abs().max()) < 1e-4:      # first layer's attention parameters: zero gradient upstream too
                    assert float(got.abs().max()) < 1e-3
                elif nm.startswith("alpha") and wb == "bias":
                    scale = max(float(ref.abs().max()), float(g[f"{tag}__{nm}__weight__grad"].abs().max()))
                    assert float((got - ref).abs().max()) <= TOL * scale, f"{tag} d {nm}.bias"
                else:
                    assert_close_rel(got, ref, TOL, f"{tag} d {nm}.{wb}")
    s = load_golden("sdr_layer_grad", DEV)
    layer = _sdr_from_fixture(s, DEV)
    x = s["x"].clone().requires_grad_(True)
    y = layer(x)
    (y * s["r"]).sum().backward()
    assert_close_rel(y, s["out"], 1e-5, "SDRLayer forward")
    assert_close_rel(x.grad, s["grad_x"], TOL, "SDRLayer d x")
    for name, p in layer.named_parameters():
        assert_close_rel(p.grad, s[name.replace(".", "__") + "__grad"].view_as(p.grad), TOL, f"SDRLayer d {name}")

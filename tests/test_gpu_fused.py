"""GPU parity of the fused MagNetConv layer kernel (`pgsd_magnet_layer_fused`: aggregation ->
shared-memory ring -> tcgen05 transform in one launch) through the public layer API, against
oracle/port.py (the reference's op order on the CPU) and against the two-launch path.
Tolerance: 1e-5 * max|ref| per tensor, as everywhere (north_star "within 1e-5 rel fp32").
"""
import pytest
import torch

from conftest import assert_close_rel
from oracle import port
from pytorch_geometric_signed_directed_b200 import nn, ops, synthetic

pytestmark = pytest.mark.gpu
DEV = "cuda"


@pytest.fixture
def fused(monkeypatch):
    monkeypatch.setattr(ops, "FUSED_LAYER", 1)
    before = ops.LAUNCHES
    yield
    assert ops.LAUNCHES > before


def _run(conv, xr, xi, ei, ew=None, lam=None):
    n0 = ops.LAUNCHES
    with torch.no_grad():     # parameters require grad; the fused kernel is the inference path
        out = conv(xr.to(DEV), xi.to(DEV), ei.to(DEV), None if ew is None else ew.to(DEV), lam)
    torch.cuda.synchronize()
    return out, ops.LAUNCHES - n0


@pytest.mark.parametrize("variant", [0, 1])
@pytest.mark.parametrize("n,e,weighted,q,norm,lam", [
    (1, 0, False, 0.25, "sym", None),             # single row, no entries
    (127, 900, False, 0.25, "sym", None),         # one partial tile
    (128, 1500, True, 0.1, "sym", None),          # exactly one tile
    (129, 1500, False, 0.25, "sym", None),        # one row into the second tile
    (5000, 100_000, False, 0.25, "sym", None),    # north-star shape in miniature
    (5000, 100_000, True, 0.2, "sym", 3.0),       # lambda_max != 2 -> constant diagonal term
    (3000, 40_000, True, 0.25, None, 40.0),       # normalization None -> per-row diagonal (degrees)
    (4000, 0, False, 0.25, "sym", None),          # empty graph: out = x W0 mix + bias
    (60_000, 700_000, False, 0.25, "sym", None),  # > 2 tiles per CTA: ring slots are reused
])
def test_fused_layer_vs_oracle_and_two_launch_path(fused, monkeypatch, n, e, weighted, q, norm, lam, variant):
    monkeypatch.setattr(ops, "FUSED_VARIANT", variant)
    g = torch.Generator().manual_seed(7 * n + e)
    ei = torch.randint(0, n, (2, e), generator=g)
    ew = (torch.rand(e, generator=g) + 0.5) if weighted else None
    xr = torch.rand(n, 64, generator=g) * 2 - 1
    xi = torch.rand(n, 64, generator=g) * 2 - 1
    conv = nn.MagNetConv(64, 64, K=1, q=q, trainable_q=False, normalization=norm, cached=True).to(DEV)
    with torch.no_grad():
        conv.bias.uniform_(-0.5, 0.5)
    (out_r, out_i), launches = _run(conv, xr, xi, ei, ew, lam)
    o_r, o_i = port.magnet_conv(xr, xi, ei, ew, conv.weight.detach().cpu(), conv.bias.detach().cpu(), q, norm,
                                lam)
    assert_close_rel(out_r, o_r, 1e-5, "fused out_real vs oracle")
    assert_close_rel(out_i, o_i, 1e-5, "fused out_imag vs oracle")
    # cached plan: the second forward is exactly one launch and bit-identical (no atomics anywhere)
    (r2, i2), launches2 = _run(conv, xr, xi, ei, ew, lam)
    assert launches2 == 1
    assert torch.equal(out_r, r2) and torch.equal(out_i, i2)
    # the two-launch path agrees to rounding
    monkeypatch.setattr(ops, "FUSED_LAYER", 0)
    (u_r, u_i), launches3 = _run(conv, xr, xi, ei, ew, lam)
    assert launches3 == 2
    assert_close_rel(out_r, u_r, 2e-6, "fused vs two-launch out_real")
    assert_close_rel(out_i, u_i, 2e-6, "fused vs two-launch out_imag")


def test_fused_layer_skewed_rows_and_hub_fallback(fused):
    # one destination with 3000 in-neighbours (below the hub threshold: stays in the fused kernel,
    # one lane group works on it while the others run ahead), then one above it (two-launch path)
    n = 20_000
    g = torch.Generator().manual_seed(5)
    ei = torch.randint(0, n, (2, 150_000), generator=g)
    xr = torch.rand(n, 64, generator=g) * 2 - 1
    xi = torch.rand(n, 64, generator=g) * 2 - 1
    for fan_in, fused_expected in ((1500, True), (6000, False)):
        star = torch.stack([torch.randperm(n, generator=g)[:fan_in], torch.full((fan_in,), 17)])
        e2 = torch.cat([ei, star], 1)
        conv = nn.MagNetConv(64, 64, K=1, q=0.25, trainable_q=False, cached=True).to(DEV)
        (out_r, out_i), _ = _run(conv, xr, xi, e2)
        (_, _), launches = _run(conv, xr, xi, e2)
        assert (launches == 1) == fused_expected
        o_r, o_i = port.magnet_conv(xr, xi, e2, None, conv.weight.detach().cpu(), conv.bias.detach().cpu(), 0.25,
                                    "sym")
        assert_close_rel(out_r, o_r, 1e-5, f"fan-in {fan_in} out_real")
        assert_close_rel(out_i, o_i, 1e-5, f"fan-in {fan_in} out_imag")


def test_fused_layer_complex_relu_and_strided_inputs(fused, monkeypatch):
    # the model wrapper's fused complex ReLU (complex_relu.py:17-34) and inputs that are column
    # slices of a wider buffer (leading dimension != 64)
    n, e = 10_000, 150_000
    g = torch.Generator().manual_seed(11)
    ei = torch.randint(0, n, (2, e), generator=g).to(DEV)
    wide = (torch.rand(n, 192, generator=g) * 2 - 1).to(DEV)
    xr, xi = wide[:, 64:128], wide[:, 128:]
    conv = nn.MagNetConv(64, 64, K=1, q=0.25, trainable_q=False, cached=True).to(DEV)
    conv.fused_complex_relu = True
    with torch.no_grad():
        conv.bias.uniform_(-0.2, 0.2)
    with torch.no_grad():
        n0 = ops.LAUNCHES
        f_r, f_i = conv(xr, xi, ei)
        assert ops.LAUNCHES - n0 == 1         # one fused launch (the plan build is not an ops launch)
        monkeypatch.setattr(ops, "FUSED_LAYER", 0)
        u_r, u_i = conv(xr, xi, ei)
    # the mask is a sign decision on out_real: compare where the two-launch value is not within
    # rounding of zero, and require the zero pattern to agree there
    safe = u_r.abs() > 1e-4
    assert torch.equal((f_r == 0)[safe], (u_r == 0)[safe])
    assert_close_rel(torch.where(safe, f_r, u_r), u_r, 2e-6, "relu out_real")
    assert_close_rel(torch.where(safe, f_i, u_i), u_i, 2e-6, "relu out_imag")
    assert (f_r >= 0).all()


def test_fused_layer_not_used_when_training_or_outside_envelope(fused):
    n, e = 3000, 30_000
    g = torch.Generator().manual_seed(3)
    ei = torch.randint(0, n, (2, e), generator=g).to(DEV)
    xr = (torch.rand(n, 64, generator=g) * 2 - 1).to(DEV)
    conv = nn.MagNetConv(64, 64, K=1, q=0.25, trainable_q=False, cached=True).to(DEV)
    out_r, out_i = conv(xr, xr, ei)          # parameters require grad -> differentiable two-launch path
    assert out_r.requires_grad
    (out_r.sum() + out_i.sum()).backward()
    assert conv.weight.grad is not None and torch.isfinite(conv.weight.grad).all()
    with torch.no_grad():
        n0 = ops.LAUNCHES
        conv(xr, xr, ei)
        assert ops.LAUNCHES - n0 == 1        # inference: fused
        conv2 = nn.MagNetConv(64, 32, K=1, q=0.25, trainable_q=False, cached=True).to(DEV)
        n0 = ops.LAUNCHES
        conv2(xr, xr, ei)
        assert ops.LAUNCHES - n0 == 2        # 64 -> 32 is outside the envelope


def test_fused_layer_full_size_matches_two_launch_path(fused, monkeypatch):
    """BASELINE config 2 layer shape (1M nodes / 20M edges / 64 -> 64): the fused kernel against
    the two-launch path (itself checked against the oracle on a row subset in test_gpu_parity)."""
    n = 1_000_000
    ei, _ = synthetic.dsbm_edges(n, 3, num_edges=20_000_000, seed=0, device=DEV)
    x = torch.rand(n, 64, device=DEV) * 2 - 1
    xi = torch.rand(n, 64, device=DEV) * 2 - 1
    conv = nn.MagNetConv(64, 64, K=1, q=0.25, trainable_q=False, cached=True).to(DEV)
    with torch.no_grad():
        conv.bias.uniform_(-0.1, 0.1)
        f_r, f_i = conv(x, xi, ei)
        monkeypatch.setattr(ops, "FUSED_LAYER", 0)
        u_r, u_i = conv(x, xi, ei)
    assert_close_rel(f_r, u_r, 2e-6, "full size out_real")
    assert_close_rel(f_i, u_i, 2e-6, "full size out_imag")

"""GPU parity of the TMA-fed warp-specialised transform kernel (`csrc/dense_tma.cu`, variant 16 of
`pgsd_dense_transform`) against fp64 and against the FFMA kernel, on the shapes the layers use.
Tolerance: 2e-6 * max|ref| for fp32 (3xTF32 is fp32-class), 6e-3 for bf16 outputs (one bf16 rounding)."""
import pytest
import torch

from conftest import assert_close_rel
from pytorch_geometric_signed_directed_b200 import ops

pytestmark = pytest.mark.gpu
DEV = "cuda"
TMA = 16
LEGACY = 16 | (1 << 26)      # round-1 split: hi rounded in place, three MMAs per k-step (kept for A/B timing)


@pytest.mark.parametrize("n_rows,ks,n_out,combine", [
    (1000, (64, 64, 64, 64), 64, True),       # MagNet K=1, 64 -> 64
    (130, (64, 64, 64, 64, 64, 64), 64, True),  # K=2
    (4097, (32, 32, 32, 32), 32, True),
    (777, (16, 16, 16, 16), 16, True),        # partial K chunk (16 of 32): zero-filled by the TMA unit
    (2500, (48, 48), 128, True),
    (1, (64, 64), 64, True),
    (3000, (128,), 128, False),               # DiGCN-like single term
    (3000, (64, 64), 32, False),              # SGCN first layer: two column blocks
    (999, (64, 64, 64), 64, False),           # SGCN merged transform: three blocks, both halves
])
def test_dense_tma_matches_fp64(n_rows, ks, n_out, combine):
    gen = torch.Generator(device=DEV).manual_seed(n_rows + n_out)
    xs = [torch.randn(n_rows, k, generator=gen, device=DEV) for k in ks]
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
    got = ops.dense(terms, n_out, bias=bias, combine=combine, variant=TMA)
    for g_, r in zip(got, ref):
        assert_close_rel(g_, r, 2e-6, "tma 3xTF32 vs fp64")
    for g_, r in zip(ops.dense(terms, n_out, bias=bias, combine=combine, variant=LEGACY), ref):
        assert_close_rel(g_, r, 2e-6, "tma 3xTF32 (legacy split) vs fp64")
    # many tiles per CTA: every stage / accumulator mbarrier wraps its phase several times
    big = [(x.repeat(60, 1), w, g) for x, w, g in terms]
    b1 = ops.dense(big, n_out, bias=bias, combine=combine, variant=TMA)
    b2 = ops.dense(big, n_out, bias=bias, combine=combine, variant=1)
    for u, v in zip(b1, b2):
        assert_close_rel(u, v, 2e-6, "long run: tma vs ffma")
    # deterministic
    b3 = ops.dense(big, n_out, bias=bias, combine=combine, variant=TMA)
    for u, v in zip(b1, b3):
        assert torch.equal(u, v)


def test_dense_tma_strided_operands_and_relu():
    gen = torch.Generator(device=DEV).manual_seed(3)
    n = 5000
    wide = torch.randn(n, 192, generator=gen, device=DEV)
    lin = torch.randn(64, 128, generator=gen, device=DEV)          # nn.Linear layout [out, in]
    wt = lin.t()
    terms = [(wide[:, 64:128], wt[:64], 0), (wide[:, 128:], wt[64:], 0)]
    out = torch.empty(n, 128, device=DEV)
    ops.dense(terms, 64, out=[out[:, 64:]], variant=TMA)
    ref = wide[:, 64:128].double() @ wt[:64].double() + wide[:, 128:].double() @ wt[64:].double()
    assert_close_rel(out[:, 64:], ref, 2e-6)
    x0, x1 = wide[:, :64].contiguous(), wide[:, 64:128].contiguous()
    w = torch.randn(64, 64, generator=gen, device=DEV) / 8
    r, i = ops.dense([(x0, w, 0), (x1, w, 1)], 64, combine=True, relu_mode=1, variant=TMA)
    a, b = x0.double() @ w.double(), x1.double() @ w.double()
    mask = ((a - b) >= 0).double()
    assert_close_rel(r, (a - b) * mask, 2e-6)
    assert_close_rel(i * (r != 0), (a + b) * mask * (r != 0).double(), 2e-6)


@pytest.mark.parametrize("n_rows,k,n_out,dtype", [
    (3000, 128, 384, torch.bfloat16),     # inception block: three 128-column tiles
    (1000, 64, 128, torch.bfloat16),
    (515, 72, 64, torch.bfloat16),        # partial K chunk (72 = 64 + 8)
    (2000, 64, 256, torch.float32),       # fp32 with two 128-column tiles
    (70000, 128, 128, torch.bfloat16),    # several tiles per CTA in bf16 (no converter warps)
])
def test_dense_tma_bf16_and_column_tiles(n_rows, k, n_out, dtype):
    gen = torch.Generator(device=DEV).manual_seed(n_rows + k)
    x = torch.randn(n_rows, k, generator=gen, device=DEV).to(dtype)
    w = (torch.randn(k, n_out, generator=gen, device=DEV) / k ** 0.5)
    if dtype == torch.bfloat16:
        w = w.bfloat16().float()
    bias = torch.randn(n_out, generator=gen, device=DEV)
    ref = x.double() @ w.double() + bias.double()
    got = ops.dense([(x, w, 0)], n_out, bias=bias, variant=TMA)[0]
    tol = 6e-3 if dtype == torch.bfloat16 else 2e-6
    assert got.dtype == dtype
    assert_close_rel(got.double(), ref, tol, "tma")

"""Golden gradients for a trainable magnetic charge q (reference files loaded unmodified; torch autograd
through torch.exp(1j*2*pi*q*theta) and the PyG-shim propagates).

    python tests/golden/make_golden_qgrad.py
"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))

from oracle import load_reference  # noqa: E402
from make_golden import nasty_graph, save  # noqa: E402

REF = load_reference.ref_classes()


def one(name, cls, K, q, seed, normalization="sym", lambda_max=None, signed_weights=False, **kw):
    n, fin, fout = 150, 12, 8
    ei, ew = nasty_graph(n, 900, seed=seed)
    if signed_weights:
        g = torch.Generator().manual_seed(seed + 1)
        ew = ew * (torch.randint(0, 2, ew.shape, generator=g) * 2 - 1).to(ew.dtype)
    torch.manual_seed(seed + 2)
    conv = cls(fin, fout, K=K, q=q, trainable_q=True, normalization=normalization, **kw)
    with torch.no_grad():
        conv.bias.uniform_(-0.3, 0.3)
    xr = (torch.rand(n, fin) * 2 - 1).requires_grad_(True)
    xi = (torch.rand(n, fin) * 2 - 1).requires_grad_(True)
    r1, r2 = torch.randn(n, fout), torch.randn(n, fout)
    o_r, o_i = conv(xr, xi, ei, ew, lambda_max=lambda_max)
    ((o_r * r1).sum() + (o_i * r2).sum()).backward()
    save(name, x_real=xr.detach(), x_imag=xi.detach(), edge_index=ei, edge_weight=ew, weight=conv.weight.detach(),
         bias=conv.bias.detach(), r1=r1, r2=r2, out_real=o_r.detach(), out_imag=o_i.detach(),
         q=conv.q.detach(), d_q=conv.q.grad, d_weight=conv.weight.grad, d_x_real=xr.grad, d_x_imag=xi.grad,
         lambda_max=torch.tensor(-1.0 if lambda_max is None else lambda_max))


def main():
    one("qgrad_magnet_k2", REF["MagNetConv"], K=2, q=0.2, seed=70)
    one("qgrad_magnet_k1_none", REF["MagNetConv"], K=1, q=0.1, seed=73, normalization=None, lambda_max=3.0)
    one("qgrad_msconv_k3", REF["MSConv"], K=3, q=0.15, seed=76, signed_weights=True)
    one("qgrad_magnet_clamped", REF["MagNetConv"], K=1, q=0.4, seed=79)   # clamp(q, 0, 0.25), quirk Q9


if __name__ == "__main__":
    main()

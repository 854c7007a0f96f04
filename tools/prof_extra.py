"""Small driver for ncu captures of kernels the headline bench does not launch: the shared-operand aggregation
(x_real is x_imag) and the grouped edge-softmax kernel (SNEAConv at config-4 scale)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pytorch_geometric_signed_directed_b200 import nn, synthetic  # noqa: E402

dev = torch.device("cuda", 0)
with torch.no_grad():
  if os.environ.get("PROF_ONLY") != "softmax":
    n = 1_000_000
    ei, _ = synthetic.dsbm_edges(n, 3, num_edges=20_000_000, seed=0, device=dev)
    x = torch.rand(n, 64, device=dev) * 2 - 1
    conv = nn.MagNetConv(64, 64, K=1, q=0.25, trainable_q=False, cached=True).to(dev)
    for _ in range(4):
        conv(x, x, ei)                       # one tensor for both parts: spmm_groups_kernel<..., NX = 1>
    del ei, x, conv
  if True:
    n = 2_000_000
    pos, neg, _ = synthetic.ssbm_edges(n, 3, num_entries=40_000_000, eta=0.1, seed=0, device=dev)
    x = torch.randn(n, 64, device=dev)
    s1 = nn.SNEAConv(64, 32, first_aggr=True).to(dev)
    for _ in range(3):
        s1(x, pos, neg)
torch.cuda.synchronize()

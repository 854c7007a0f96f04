"""C3 only (DiGCN_InceptionBlock 500k nodes / 2 x 10M nnz / 128 bf16), a few forwards -- the target of ncu captures."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pytorch_geometric_signed_directed_b200 import nn, synthetic  # noqa: E402

dev = torch.device("cuda", 0)
with torch.no_grad():
    n, e, f = 500_000, 10_000_000, 128
    ei1, _ = synthetic.dsbm_edges(n, 3, num_edges=e, seed=1, device=dev)
    ei2, _ = synthetic.dsbm_edges(n, 3, num_edges=e, seed=2, device=dev)
    w1, w2 = synthetic.sym_norm_weights(ei1, n), synthetic.sym_norm_weights(ei2, n)
    x = (torch.rand(n, f, device=dev) * 2 - 1).bfloat16()
    blk = nn.DiGCN_InceptionBlock(f, f).to(dev)
    for _ in range(4):
        blk(x, ei1, w1, ei2, w2)
    torch.cuda.synchronize()
print("done")

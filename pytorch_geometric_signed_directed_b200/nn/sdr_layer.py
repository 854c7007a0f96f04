"""Drop-in SDRLayer (SDGNN's signed directed relationship layer) and the GATConv it is made of.

Reference: nn/signed/SDGNN.py:13-64 -- `SDRLayer(in_dim, out_dim, edge_lists)` owns one PyG
`GATConv(in_dim, out_dim)` per motif edge list (`agg_0..agg_{k-1}`) and an MLP
(`mlp_layer.0`, `mlp_layer.2`); forward(x) = mlp(cat([x] + [agg_i(x, edges_i)])).
The GATConv arithmetic is third-party (torch_geometric.nn.GATConv, absent from this image); it is
restated here for the configuration SDGNN uses (heads=1, concat, no dropout, self-loops re-added
for every node, LeakyReLU(0.2) scores) with PyG's parameter names (`lin.weight [out, in]`,
`att_src`, `att_dst` [1, 1, out], `bias`).

Kernels per GATConv: `pgsd_dense_transform` (h = x W^T), `pgsd_dense_transform` with n_out = 2
(the two attention scores per node), `pgsd_edge_softmax` (alpha per stored entry) and
`pgsd_spmm_csr` with val = alpha (+ bias).  Forward only (the attention path has no backward yet).
"""
from __future__ import annotations

import math
from typing import List

import torch
from torch import Tensor

from .. import ops, plan as _plan
from ..plan import CSRPlan


class GATConv(torch.nn.Module):
    def __init__(self, in_channels: int, out_channels: int, heads: int = 1, concat: bool = True,
                 negative_slope: float = 0.2, dropout: float = 0.0, add_self_loops: bool = True,
                 bias: bool = True, **kwargs):
        super().__init__()
        if heads != 1 or not concat or dropout != 0.0:
            raise NotImplementedError("GATConv kernels cover heads=1, concat=True, dropout=0 (SDGNN's use)")
        self.in_channels, self.out_channels, self.heads = in_channels, out_channels, heads
        self.negative_slope, self.add_self_loops = negative_slope, add_self_loops
        self.lin = torch.nn.Linear(in_channels, out_channels, bias=False)
        self.att_src = torch.nn.Parameter(torch.empty(1, heads, out_channels))
        self.att_dst = torch.nn.Parameter(torch.empty(1, heads, out_channels))
        if bias:
            self.bias = torch.nn.Parameter(torch.empty(out_channels))
        else:
            self.register_parameter('bias', None)
        self._plans = _plan.PlanCache(capacity=2)
        self.reset_parameters()

    def reset_parameters(self):
        def glorot(t):
            a = math.sqrt(6.0 / (t.size(-2) + t.size(-1)))
            with torch.no_grad():
                t.uniform_(-a, a)
        glorot(self.lin.weight)
        glorot(self.att_src)
        glorot(self.att_dst)
        if self.bias is not None:
            with torch.no_grad():
                self.bias.zero_()
        self._plans.clear()

    def _plan_for(self, edge_index: Tensor, n: int) -> CSRPlan:
        def build():
            ei = edge_index
            if self.add_self_loops:
                ei = ei[:, ei[0] != ei[1]]
                loops = torch.arange(n, device=ei.device, dtype=ei.dtype)
                ei = torch.cat([ei, torch.stack([loops, loops])], dim=1)
            return _plan.build_csr(ei.contiguous(), None, n, n, "source_to_target")
        return self._plans.get((edge_index,), (n, self.add_self_loops), build)

    def forward(self, x: Tensor, edge_index: Tensor, out: Tensor = None) -> Tensor:
        """`out`: optional [N, out_channels] destination (may be a column slice of a wider buffer)."""
        _plan.require_cuda(x, "x")
        n, c = x.size(0), self.out_channels
        p = self._plan_for(edge_index, n)
        with torch.no_grad():
            h = ops.dense([(x, self.lin.weight.t(), 0)], c)[0]
            att = torch.stack([self.att_src.view(c), self.att_dst.view(c)], dim=1)      # [C, 2]
            s = ops.dense([(h, att, 0)], 2)[0]
            _, alphas = ops.edge_softmax([p], [s[:, 0]], [s[:, 1]], act="leaky_relu",
                                         slope=self.negative_slope, want_alpha=True)
            weighted = CSRPlan(p.n_dst, p.n_src, p.nnz, p.num_input_edges, p.row_ptr, p.col, [alphas[0]],
                               [None], [0.0])
            return ops.spmm(weighted, [h], (0,), bias=self.bias, out=None if out is None else [out])[0]

    def __repr__(self):
        return f'{self.__class__.__name__}({self.in_channels}, {self.out_channels}, heads={self.heads})'


class SDRLayer(torch.nn.Module):
    def __init__(self, in_dim: int = 20, out_dim: int = 20, edge_lists: List[Tensor] = [], **kwargs):
        super().__init__(**kwargs)
        self.edge_lists = edge_lists
        self.aggs = []
        for i in range(len(edge_lists)):
            self.aggs.append(GATConv(in_dim, out_dim))
            self.add_module('agg_{}'.format(i), self.aggs[-1])
        self.mlp_layer = torch.nn.Sequential(
            torch.nn.Linear(in_dim * (len(edge_lists) + 1), out_dim),
            torch.nn.Tanh(),
            torch.nn.Linear(out_dim, out_dim))

    def reset_parameters(self):
        for m in self.mlp_layer:
            if type(m) == torch.nn.Linear:
                torch.nn.init.kaiming_normal_(m.weight)
        for agg in self.aggs:
            agg.reset_parameters()

    def forward(self, x: Tensor) -> Tensor:
        _plan.require_cuda(x, "x")
        with torch.no_grad():
            feats = [x] + [agg(x, edges) for edges, agg in zip(self.edge_lists, self.aggs)]
            l0, l2 = self.mlp_layer[0], self.mlp_layer[2]
            w0 = l0.weight.t()                                     # [in*(k+1), out] view
            fi = x.size(1)
            # cat([x] + neigh_feats) @ W0^T  ==  sum of column-block terms (no concatenation)
            terms = [(f, w0[i * fi:(i + 1) * fi], 0) for i, f in enumerate(feats)]
            hid = ops.dense(terms, l0.out_features, bias=l0.bias, relu_mode=2)[0]     # Tanh as the epilogue
            return ops.dense([(hid, l2.weight.t(), 0)], l2.out_features, bias=l2.bias)[0]

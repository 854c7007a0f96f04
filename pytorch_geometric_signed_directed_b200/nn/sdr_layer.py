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
`pgsd_spmm_csr` with val = alpha (+ bias).  When gradients are required the layers run the same kernels through
`autograd.py` (`pgsd_edge_softmax_backward`, `pgsd_sddmm_rows`, transposed aggregation).
"""
from __future__ import annotations

import math
from typing import List

import torch
from torch import Tensor

from .. import autograd as ag, ops, plan as _plan
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

    def score_weights(self) -> Tensor:
        """[in, 2]: the two attention scores are linear in x -- (x W^T) . a = x . (W a) -- so they come from x
        directly, and several GATConvs reading the same x share one transform launch for all their scores."""
        w = self.lin.weight.detach().t().float()                                           # [in, C]
        att = torch.stack([self.att_src.detach().view(-1), self.att_dst.detach().view(-1)], dim=1).float()
        return w @ att

    def aggregate(self, h: Tensor, s_src: Tensor, s_dst: Tensor, edge_index: Tensor, out: Tensor = None,
                  accumulate_into: Tensor = None) -> Tensor:
        """softmax over the incoming edges + weighted sum of h (+ bias), given h = x W^T and the per-node scores.
        `accumulate_into`: running sum the result is ADDED to in the aggregation epilogue (no bias then; used
        when the layer that follows is linear and has been folded into h, see SDRLayer.forward)."""
        p = self._plan_for(edge_index, h.size(0))
        if ops.gat_aggregate_supported(h, out, accumulate_into) and (self.bias is None or self.out_channels % 4 == 0):
            # softmax inside the aggregation kernel: no alpha array, one launch
            if accumulate_into is not None:
                return ops.gat_aggregate(p, s_src, s_dst, h, slope=self.negative_slope, z=accumulate_into, beta=1.0,
                                         out=accumulate_into)
            return ops.gat_aggregate(p, s_src, s_dst, h, slope=self.negative_slope, bias=self.bias, out=out)
        _, alphas = ops.edge_softmax([p], [s_src], [s_dst], act="leaky_relu", slope=self.negative_slope,
                                     want_alpha=True)
        weighted = CSRPlan(p.n_dst, p.n_src, p.nnz, p.num_input_edges, p.row_ptr, p.col, [alphas[0]], [None], [0.0])
        weighted._hubs = p.hub_rows()          # same rows: no second device->host look at the row lengths
        if accumulate_into is not None:
            return ops.spmm(weighted, [h], (0,), beta=1.0, zs=[accumulate_into], out=[accumulate_into])[0]
        return ops.spmm(weighted, [h], (0,), bias=self.bias, out=None if out is None else [out])[0]

    def requires_grad_path(self, x: Tensor) -> bool:
        return torch.is_grad_enabled() and (x.requires_grad or any(p.requires_grad for p in self.parameters()))

    def forward_train(self, x: Tensor, edge_index: Tensor) -> Tensor:
        """The same arithmetic on the autograd path (`ag.dense` for h = x W^T and the two scores, `ag.gat_attend` for
        softmax + weighted sum): gradients reach x, lin.weight, att_src, att_dst and bias."""
        w = self.lin.weight.t().float()                                                   # [in, C]
        h = ag.dense([(x, w, 0)], self.out_channels)[0]
        att = torch.stack([self.att_src.view(-1), self.att_dst.view(-1)], dim=1).float()   # [C, 2]
        width = 32 if x.dtype == torch.bfloat16 else 16
        sw = torch.nn.functional.pad(w @ att, (0, width - 2))
        s = ag.dense([(x, sw, 0)], width)[0].float()
        p = self._plan_for(edge_index, h.size(0))
        return ag.gat_attend(p, s[:, 0].contiguous(), s[:, 1].contiguous(), h, self.negative_slope, bias=self.bias)

    def forward(self, x: Tensor, edge_index: Tensor, out: Tensor = None) -> Tensor:
        """`out`: optional [N, out_channels] destination (may be a column slice of a wider buffer)."""
        _plan.require_cuda(x, "x")
        if self.requires_grad_path(x):
            y = self.forward_train(x, edge_index)
            if out is not None:
                raise RuntimeError("GATConv: preallocated outputs are not supported when gradients are required")
            return y
        with torch.no_grad():
            h, s = gat_transforms(x, [self])
            return self.aggregate(h[0], s[0][0], s[0][1], edge_index, out)

    def __repr__(self):
        return f'{self.__class__.__name__}({self.in_channels}, {self.out_channels}, heads={self.heads})'


def gat_transforms(x: Tensor, gats: List["GATConv"], scores_only: bool = False):
    """h_k = x W_k^T and the score pairs of several GATConvs that read the same x, in TWO launches: one transform
    with the weights side by side ([in, k*C]; every h_k is a column block of its output) and one with all score
    vectors side by side, zero-padded to a width the tensor-core kernels take ([in, 2k] -> 16 / 32 / 64 / 128).
    Returns ([h_k], [(s_src_k, s_dst_k)])."""
    c = gats[0].out_channels
    k = len(gats)
    hs = None
    if not scores_only:
        w_all = torch.cat([g.lin.weight.detach().t().float() for g in gats], dim=1)           # [in, k*C]
        h_all = ops.dense([(x, w_all, 0)], k * c)[0]
        hs = [h_all[:, i * c:(i + 1) * c] for i in range(k)]
    sw = torch.cat([g.score_weights() for g in gats], dim=1)                                   # [in, 2k]
    width = next((w for w in (16, 32, 64, 128) if w >= 2 * k), None)
    if width is None:
        width = -(-2 * k // 128) * 128
    if x.dtype == torch.bfloat16 and width < 32:
        width = 32
    sw = torch.nn.functional.pad(sw, (0, width - 2 * k))
    s_all = ops.dense([(x, sw, 0)], width)[0].float()
    ss = [(s_all[:, 2 * i].contiguous(), s_all[:, 2 * i + 1].contiguous()) for i in range(k)]
    return hs, ss


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
        """mlp(cat([x] + [agg_k(x, edges_k)])) with the first Linear folded THROUGH the aggregations: it is linear,
        so  (sum_j alpha_ij h_kj + b_k) M_k = sum_j alpha_ij (h_kj M_k) + b_k M_k  -- every GATConv aggregates
        x (W_k^T M_k) and adds into one [N, out] accumulator in its epilogue.  The k aggregated [N, C] tensors, their
        concatenation and the [N, (k+1) C] x [(k+1) C, out] transform (whose weights do not fit the tensor-core
        kernel's shared memory) never exist; the attention scores still come from h_k = x W_k^T."""
        _plan.require_cuda(x, "x")
        if torch.is_grad_enabled() and (x.requires_grad or any(p.requires_grad for p in self.parameters())):
            # training: the reference's own order (SDGNN.py:57-64) on the differentiable primitives
            feats = torch.cat([x] + [agg.forward_train(x, e) for e, agg in zip(self.edge_lists, self.aggs)], dim=1)
            l0, l2 = self.mlp_layer[0], self.mlp_layer[2]
            hid = torch.tanh(ag.dense([(feats, l0.weight.t(), 0)], l0.out_features, bias=l0.bias)[0])
            return ag.dense([(hid, l2.weight.t(), 0)], l2.out_features, bias=l2.bias)[0]
        with torch.no_grad():
            l0, l2 = self.mlp_layer[0], self.mlp_layer[2]
            k, fi = len(self.aggs), x.size(1)
            w0 = l0.weight.detach().t().float()                                  # [fi + k*C, out]
            c = (w0.size(0) - fi) // max(k, 1)
            blocks = [w0[fi + i * c: fi + (i + 1) * c] for i in range(k)]         # M_k
            bias0 = l0.bias.detach().float().clone() if l0.bias is not None else w0.new_zeros(w0.size(1))
            for agg, m in zip(self.aggs, blocks):
                if agg.bias is not None:
                    bias0 += agg.bias.detach().float() @ m
            acc = ops.dense([(x, w0[:fi], 0)], l0.out_features, bias=bias0)[0]    # x M_0 + b_0 + sum_k b_k M_k
            if k:
                folded = torch.cat([agg.lin.weight.detach().t().float() @ m for agg, m in zip(self.aggs, blocks)], 1)
                hp_all = ops.dense([(x, folded, 0)], k * l0.out_features)[0]      # [N, k*out]
                _, ss = gat_transforms(x, self.aggs, scores_only=True)
                o = l0.out_features
                for i, (edges, agg) in enumerate(zip(self.edge_lists, self.aggs)):
                    agg.aggregate(hp_all[:, i * o:(i + 1) * o], ss[i][0], ss[i][1], edges, accumulate_into=acc)
            hid = torch.tanh_(acc)
            return ops.dense([(hid, l2.weight.t(), 0)], l2.out_features, bias=l2.bias)[0]

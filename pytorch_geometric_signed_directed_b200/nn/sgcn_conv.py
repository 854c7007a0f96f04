"""Drop-in SGCNConv on the B200 kernels.

Reference: nn/signed/SGCNConv.py:63-138 -- same constructor, forward(x, pos_edge_index,
neg_edge_index), parameter names (`lin_b.weight [out, 2*in | 3*in]`, `lin_b.bias`, `lin_u.*`)
and __repr__.  Mean aggregation = `pgsd_spmm_csr` with implicit unit values and the
1/max(in-count, 1) row scale computed from the CSR row lengths (integer counts, exact); the
reference's torch.cat + Linear is one `pgsd_dense_transform` whose terms are the column blocks
of the concatenation.  In the deep layer the reference's four half-width propagates
(SGCNConv.py:109-119) are two full-width launches.
"""
from __future__ import annotations

import math
from typing import Tuple, Union

import torch
import torch.nn.functional as F
from torch import Tensor

from .. import autograd as ag, ops, plan as _plan


class _Lin(torch.nn.Module):
    """Parameter container with torch_geometric.nn.dense.linear.Linear's names and layout
    (weight [out, in], y = x W^T + b)."""

    def __init__(self, in_channels: int, out_channels: int, bias: bool = True):
        super().__init__()
        self.in_channels, self.out_channels = in_channels, out_channels
        self.weight = torch.nn.Parameter(torch.empty(out_channels, in_channels))
        if bias:
            self.bias = torch.nn.Parameter(torch.empty(out_channels))
        else:
            self.register_parameter('bias', None)
        self.reset_parameters()

    def reset_parameters(self):
        torch.nn.init.kaiming_uniform_(self.weight, a=math.sqrt(5))
        if self.bias is not None:
            bound = 1.0 / math.sqrt(self.in_channels) if self.in_channels > 0 else 0.0
            torch.nn.init.uniform_(self.bias, -bound, bound)


class SGCNConv(torch.nn.Module):
    def __init__(self, in_dim: int, out_dim: int, first_aggr: bool, bias: bool = True,
                 norm_emb: bool = False, **kwargs):
        super().__init__()
        self.aggr = kwargs.get('aggr', 'mean')
        if self.aggr != 'mean':
            raise NotImplementedError("SGCNConv kernels implement the reference's mean aggregation")
        self.in_dim, self.out_dim = in_dim, out_dim
        self.first_aggr, self.norm_emb = first_aggr, norm_emb
        mult = 2 if first_aggr else 3
        self.lin_b = _Lin(mult * in_dim, out_dim, bias)
        self.lin_u = _Lin(mult * in_dim, out_dim, bias)
        self._plans = _plan.PlanCache(capacity=4)

    def reset_parameters(self):
        self.lin_b.reset_parameters()
        self.lin_u.reset_parameters()
        self._plans.clear()

    def _plan_for(self, edge_index: Tensor, n_dst: int, n_src: int):
        if not isinstance(edge_index, Tensor):
            raise NotImplementedError("SparseTensor adjacency is not supported; pass COO edge_index")
        return self._plans.get((edge_index,), (n_dst, n_src),
                               lambda: _plan.build_csr(edge_index, None, n_dst, n_src,
                                                       "source_to_target"))

    def forward(self, x: Union[Tensor, Tuple[Tensor, Tensor]], pos_edge_index: Tensor,
                neg_edge_index: Tensor) -> Tensor:
        x_src, x_dst = (x, x) if isinstance(x, Tensor) else x
        _plan.require_cuda(x_src, "x")
        n_src, n_dst = x_src.size(0), x_dst.size(0)
        pos = self._plan_for(pos_edge_index, n_dst, n_src)
        neg = self._plan_for(neg_edge_index, n_dst, n_src)
        fi, fo = self.in_dim, self.out_dim
        wb, wu = self.lin_b.weight.t(), self.lin_u.weight.t()          # [mult*in, out] views
        m_pos = ag.spmm(pos, [x_src], (0,), mean=True)[0]
        m_neg = ag.spmm(neg, [x_src], (0,), mean=True)[0]
        if self.first_aggr:
            terms_b = [(m_pos, wb[:fi], 0), (x_dst, wb[fi:], 0)]
            terms_u = [(m_neg, wu[:fi], 0), (x_dst, wu[fi:], 0)]
        else:
            # x = [x_b | x_u]; m_pos = [mean+(x_b) | mean+(x_u)], m_neg = [mean-(x_b) | mean-(x_u)]
            terms_b = [(m_pos[:, :fi], wb[:fi], 0), (m_neg[:, fi:], wb[fi:2 * fi], 0), (x_dst[:, :fi], wb[2 * fi:], 0)]
            terms_u = [(m_pos[:, fi:], wu[:fi], 0), (m_neg[:, :fi], wu[fi:2 * fi], 0), (x_dst[:, fi:], wu[2 * fi:], 0)]
        track = [x_src, x_dst, self.lin_b.weight, self.lin_u.weight, self.lin_b.bias, self.lin_u.bias]
        if ag._needs_grad(track):
            out = torch.cat([ag.dense(terms_b, fo, bias=self.lin_b.bias)[0],
                             ag.dense(terms_u, fo, bias=self.lin_u.bias)[0]], dim=-1)
        else:   # inference: both halves are written in place into one buffer (no torch.cat)
            out = torch.empty((n_dst, 2 * fo), dtype=x_src.dtype, device=x_src.device)
            ops.dense(terms_b, fo, bias=self.lin_b.bias, out=[out[:, :fo]])
            ops.dense(terms_u, fo, bias=self.lin_u.bias, out=[out[:, fo:]])
        if self.norm_emb:
            out = F.normalize(out, p=2, dim=-1)
        return out

    def __repr__(self) -> str:
        return (f'{self.__class__.__name__}({self.in_dim}, '
                f'{self.out_dim}, first_aggr={self.first_aggr})')

"""Drop-in DiGCNConv / DiGCN_InceptionBlock on the B200 kernels.

Reference: nn/directed/DiGCNConv.py:30-98 and nn/directed/DiGCN_Inception_Block.py:19-47 --
same constructors, forward signatures, parameter names (`weight [in, out]`, `bias`, block:
`ln`, `conv1`, `conv2`), `cached=True` default and its quirk (after the first call only the
edge COUNT is checked; new edge values are ignored), same RuntimeErrors.

Per forward: `pgsd_dense_transform` (x @ W) then `pgsd_spmm_csr` with the bias folded into the
aggregation epilogue (the reference's update(), DiGCNConv.py:91-94).
"""
from __future__ import annotations

import math
from typing import Optional, Tuple

import torch
from torch import Tensor
from torch.nn import Linear, Parameter

from .. import autograd as ag, plan as _plan


class DiGCNConv(torch.nn.Module):
    def __init__(self, in_channels: int, out_channels: int, improved: bool = False,
                 cached: bool = True, bias: bool = True, **kwargs):
        super().__init__()
        self.in_channels = in_channels
        self.out_channels = out_channels
        self.improved = improved
        self.cached = cached
        self.weight = Parameter(torch.Tensor(in_channels, out_channels))
        if bias:
            self.bias = Parameter(torch.Tensor(out_channels))
        else:
            self.register_parameter('bias', None)
        self.reset_parameters()

    def reset_parameters(self):
        a = math.sqrt(6.0 / (self.weight.size(-2) + self.weight.size(-1)))
        with torch.no_grad():
            self.weight.uniform_(-a, a)
            if self.bias is not None:
                self.bias.zero_()
        self._plan = None
        self._cached_inputs = None
        self.cached_num_edges = None

    @property
    def cached_result(self):
        return self._cached_inputs

    @cached_result.setter
    def cached_result(self, value):
        if value is not None:
            raise AttributeError("assign None to reset the cache")
        self._plan, self._cached_inputs = None, None

    def _aggregate(self, xw: Tensor, edge_index: Tensor, edge_weight: Optional[Tensor],
                   out: Optional[Tensor] = None, add: Optional[Tensor] = None) -> Tensor:
        """`add`: a tensor summed into the result inside the aggregation epilogue (beta * z)."""
        if self.cached and self._plan is not None and edge_index.size(1) != self.cached_num_edges:
            raise RuntimeError(
                'Cached {} number of edges, but found {}. Please '
                'disable the caching behavior of this layer by removing '
                'the `cached=True` argument in its constructor.'.format(
                    self.cached_num_edges, edge_index.size(1)))
        if not self.cached or self._plan is None:
            self.cached_num_edges = edge_index.size(1)
            if edge_weight is None:
                raise RuntimeError(
                    'Normalized adj matrix cannot be None. Please '
                    'obtain the adj matrix in preprocessing.')
            n = xw.size(0)
            self._plan = _plan.build_csr(edge_index, edge_weight, n, n, "source_to_target")
            self._cached_inputs = (edge_index, edge_weight)
        if add is not None:
            return ag.spmm(self._plan, [xw], (0,), bias=self.bias, beta=1.0, zs=[add])[0]
        return ag.spmm(self._plan, [xw], (0,), bias=self.bias, out=None if out is None else [out])[0]

    def forward(self, x: Tensor, edge_index: Tensor, edge_weight: Optional[Tensor] = None) -> Tensor:
        _plan.require_cuda(x, "x")
        xw = ag.dense([(x, self.weight, 0)], self.out_channels)[0]
        return self._aggregate(xw, edge_index, edge_weight)

    def __repr__(self):
        return '{}({}, {})'.format(self.__class__.__name__, self.in_channels, self.out_channels)


class DiGCN_InceptionBlock(torch.nn.Module):
    """x0 = Linear(x); x1 = DiGCNConv(x, ei, w); x2 = DiGCNConv(x, ei2, w2).
    The three feature transforms share one pass over x (weights of ln / conv1 / conv2 are
    three column blocks of one [in, 3*out] operand); the two aggregations read their column
    block in place through the leading-dimension arguments of pgsd_spmm_csr."""

    def __init__(self, in_dim: int, out_dim: int):
        super().__init__()
        self.ln = Linear(in_dim, out_dim)
        self.conv1 = DiGCNConv(in_dim, out_dim)
        self.conv2 = DiGCNConv(in_dim, out_dim)
        self.reset_parameters()

    def reset_parameters(self):
        self.ln.reset_parameters()
        self.conv1.reset_parameters()
        self.conv2.reset_parameters()

    def forward_sum(self, x: Tensor, edge_index: Tensor, edge_weight: Tensor, edge_index2: Tensor,
                    edge_weight2: Tensor) -> Tensor:
        """x0 + x1 + x2 (DiGCN_Inception_Block_node_classification.py:55,63,71) without materialising
        x1 and x2: each aggregation adds the running sum in its epilogue."""
        buf, out_dim = self._transform(x)
        y = self.conv1._aggregate(buf[:, out_dim:2 * out_dim], edge_index, edge_weight, add=buf[:, :out_dim])
        return self.conv2._aggregate(buf[:, 2 * out_dim:], edge_index2, edge_weight2, add=y)

    def _transform(self, x: Tensor):
        _plan.require_cuda(x, "x")
        out_dim = self.conv1.out_channels
        w_all = torch.cat([self.ln.weight.t().float(), self.conv1.weight.float(),
                           self.conv2.weight.float()], dim=1)                   # [in, 3*out]
        b_all = None
        if self.ln.bias is not None:
            b_all = torch.cat([self.ln.bias.float(), torch.zeros(2 * out_dim, device=x.device)])
        return ag.dense([(x, w_all, 0)], 3 * out_dim, bias=b_all)[0], out_dim      # [N, 3*out]

    def forward(self, x: Tensor, edge_index: Tensor, edge_weight: Tensor, edge_index2: Tensor,
                edge_weight2: Tensor) -> Tuple[Tensor, Tensor, Tensor]:
        _plan.require_cuda(x, "x")
        out_dim = self.conv1.out_channels
        w_all = torch.cat([self.ln.weight.t().float(), self.conv1.weight.float(),
                           self.conv2.weight.float()], dim=1)                   # [in, 3*out]
        b_all = None
        if self.ln.bias is not None:
            b_all = torch.cat([self.ln.bias.float(), torch.zeros(2 * out_dim, device=x.device)])
        buf = ag.dense([(x, w_all, 0)], 3 * out_dim, bias=b_all)[0]               # [N, 3*out]
        x0 = buf[:, :out_dim]
        x1 = self.conv1._aggregate(buf[:, out_dim:2 * out_dim], edge_index, edge_weight)
        x2 = self.conv2._aggregate(buf[:, 2 * out_dim:], edge_index2, edge_weight2)
        return x0, x1, x2

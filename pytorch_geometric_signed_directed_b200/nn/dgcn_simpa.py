"""Drop-in DGCNConv and SIMPA: the remaining "same kernels" rows of SURVEY §2.

DGCNConv (reference nn/directed/DGCNConv.py:38-103): weight-less GCN propagate with PyG
`gcn_norm` (remaining self-loops, deg^-1/2 A_hat deg^-1/2, source_to_target).  The cache quirk is
kept: with `cached=True` the FIRST normalised graph is reused for every later call, whatever
edge_index is passed (SURVEY Q6 -- DGCN_node_classification shares one layer across three
adjacencies).  Plan: `pgsd_build_csr_sym_norm`; aggregation: `pgsd_spmm_csr`.

SIMPA (reference nn/signed/SIMPA.py:17-144): mixed-path sums of Conv_Base products over A+
(fill_value) and A- (fill 0).  Parameter names `_w_p/_w_n` (undirected) and
`_w_sp/_w_sn/_w_tp/_w_tn` (directed) are the reference's.  Feature pairs that travel through
the same operator in the same hop are aggregated in ONE launch (two operands, shared pattern
and values), e.g. (curr_p, curr_n_aux) through A+.
"""
from __future__ import annotations

from typing import Optional

import torch
from torch import Tensor
from torch.nn.parameter import Parameter

from .. import autograd as ag, plan as _plan
from .mixed_path import Conv_Base


class DGCNConv(torch.nn.Module):
    def __init__(self, improved: bool = False, cached: bool = False, add_self_loops: bool = True,
                 normalize: bool = True, **kwargs):
        super().__init__()
        self.aggr = kwargs.get('aggr', 'add')
        self.flow = kwargs.get('flow', 'source_to_target')
        if self.aggr != 'add' or self.flow != 'source_to_target':
            raise NotImplementedError("DGCNConv kernels implement aggr='add', flow='source_to_target'")
        self.node_dim = -2
        self.improved, self.cached = improved, cached
        self.add_self_loops, self.normalize = add_self_loops, normalize
        self._plans = _plan.PlanCache(capacity=4)
        self.reset_parameters()

    def reset_parameters(self):
        self._cached_edge_index = None      # holds the cached PLAN (reference: the normalised COO)
        self._cached_adj_t = None
        self._plans.clear()

    def forward(self, x: Tensor, edge_index: Tensor, edge_weight: Optional[Tensor] = None) -> Tensor:
        if not isinstance(edge_index, Tensor):
            raise NotImplementedError("SparseTensor adjacency is not supported; pass COO edge_index")
        _plan.require_cuda(x, "x")
        n = x.size(self.node_dim)
        if self.normalize:
            p = self._cached_edge_index
            if p is None:
                fill = 2.0 if self.improved else 1.0
                p = self._plans.get((edge_index, edge_weight), ("gcn", n, fill, self.add_self_loops),
                                    lambda: _plan.build_sym_norm(edge_index, edge_weight, n, fill,
                                                                 self.add_self_loops))
                if self.cached:
                    self._cached_edge_index = p
        else:
            p = self._plans.get((edge_index, edge_weight), ("raw", n),
                                lambda: _plan.build_csr(edge_index, edge_weight, n, n, "source_to_target"))
        return ag.spmm(p, [x], (0,))[0]


class SIMPA(torch.nn.Module):
    def __init__(self, hop: int, fill_value: float, directed: bool = False):
        super().__init__()
        self._hop_p = hop + 1
        self._hop_n = int((1 + hop) * hop / 2)
        self._undirected = not directed
        self.conv_layer_p = Conv_Base(fill_value)
        self.conv_layer_n = Conv_Base(0.0)
        if self._undirected:
            self._w_p = Parameter(torch.FloatTensor(self._hop_p, 1))
            self._w_n = Parameter(torch.FloatTensor(self._hop_n, 1))
            self._reset_parameters_undirected()
        else:
            self._w_sp = Parameter(torch.FloatTensor(self._hop_p, 1))
            self._w_sn = Parameter(torch.FloatTensor(self._hop_n, 1))
            self._w_tp = Parameter(torch.FloatTensor(self._hop_p, 1))
            self._w_tn = Parameter(torch.FloatTensor(self._hop_n, 1))
            self._reset_parameters_directed()

    def _reset_parameters_undirected(self):
        self._w_p.data.fill_(1.0)
        self._w_n.data.fill_(1.0)

    def _reset_parameters_directed(self):
        for w in (self._w_sp, self._w_sn, self._w_tp, self._w_tn):
            w.data.fill_(1.0)

    @staticmethod
    def _mixed_paths(plan_p, plan_n, x_p, x_n, w_p, w_n, hop_p):
        """One side (source or target) of the reference loop (SIMPA.py:72-93 / :95-139):
        feat_p = sum_h w_p[h] A_p^h x_p ;  feat_n = sum over paths A_p^a A_n A_p^h x_n."""
        feat_p = w_p[0] * x_p
        feat_n = torch.zeros_like(feat_p)
        curr_p, curr_n_aux = x_p, x_n
        j = 0
        for h in range(hop_p):
            if h > 0:
                # both go through A_p in the same hop: one launch, two operands
                curr_p, curr_n_aux = ag.spmm(plan_p, [curr_p, curr_n_aux], (0, 0))
                feat_p = feat_p + w_p[h] * curr_p
            if h != hop_p - 1:
                curr_n = ag.spmm(plan_n, [curr_n_aux], (0,))[0]
                feat_n = feat_n + w_n[j] * curr_n
                j += 1
                for _ in range(hop_p - 2 - h):
                    curr_n = ag.spmm(plan_p, [curr_n], (0,))[0]
                    feat_n = feat_n + w_n[j] * curr_n
                    j += 1
        return feat_p, feat_n

    def forward(self, edge_index_p: Tensor, edge_weight_p: Tensor, edge_index_n: Tensor,
                edge_weight_n: Tensor, x_p: Tensor, x_n: Tensor, x_pt: Optional[Tensor] = None,
                x_nt: Optional[Tensor] = None) -> Tensor:
        _plan.require_cuda(x_p, "x_p")
        n = x_p.size(0)
        pp = self.conv_layer_p.plan_for(edge_index_p, edge_weight_p, n, transpose=False)
        pn = self.conv_layer_n.plan_for(edge_index_n, edge_weight_n, n, transpose=False)
        if self._undirected:
            feat_p, feat_n = self._mixed_paths(pp, pn, x_p, x_n, self._w_p, self._w_n, self._hop_p)
            return torch.cat([feat_p, feat_n], dim=1)
        # directed: the target side runs on the transposed edge lists (SIMPA.py:97-98)
        tp = self.conv_layer_p.plan_for(edge_index_p, edge_weight_p, n, transpose=True)
        tn = self.conv_layer_n.plan_for(edge_index_n, edge_weight_n, n, transpose=True)
        feat_sp, feat_sn = self._mixed_paths(pp, pn, x_p, x_n, self._w_sp, self._w_sn, self._hop_p)
        feat_tp, feat_tn = self._mixed_paths(tp, tn, x_pt, x_nt, self._w_tp, self._w_tn, self._hop_p)
        return torch.cat([feat_sp, feat_sn, feat_tp, feat_tn], dim=1)

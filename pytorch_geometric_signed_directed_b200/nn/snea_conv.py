"""Drop-in SNEAConv on the B200 kernels.

Reference: nn/signed/SNEAConv.py:43-150 -- same constructor, forward(x, pos_edge_index,
neg_edge_index), parameter names (`lin_b`, `lin_u` = torch.nn.Linear(in, out);
`alpha_b`, `alpha_u` = torch.nn.Linear(2*out, 1)) and __repr__.

What the reference computes per propagate (SNEAConv.py:135-146): for every edge j -> i of type
p, alpha_e = softmax_i(tanh(Linear([x_p[j] || x_p[i]]))) and the message is the TARGET's
feature x_p[i] * alpha_e (quirk Q7), so  out[i] = x_0[i] * sum_{type 0} alpha + x_1[i] *
sum_{type 1} alpha.  Here the Linear splits into two per-node scalars (s_src = X a_j,
s_dst = X a_i + c), computed by `pgsd_dense_transform`, and `pgsd_edge_softmax` does the
per-row softmax sums -- scalar gathers instead of the reference's [nnz, 2*out] temporaries.
Edge bookkeeping follows the reference exactly: self-loops are removed, then re-added for the
nodes 0..max(remaining edge ids) only (add_self_loops without num_nodes, :88-89,110-111);
negative edges never get self-loops in the deep layer (:112).
"""
from __future__ import annotations

from typing import Tuple, Union

import torch
from torch import Tensor

from .. import ops, plan as _plan


class SNEAConv(torch.nn.Module):
    def __init__(self, in_dim: int, out_dim: int, first_aggr: bool, bias: bool = True,
                 norm_emb: bool = True, add_self_loops=True, **kwargs):
        super().__init__()
        self.aggr = kwargs.get('aggr', 'add')
        self.in_dim, self.out_dim = in_dim, out_dim
        self.first_aggr = first_aggr
        self.add_self_loops = add_self_loops
        self.norm_emb = norm_emb
        self.lin_b = torch.nn.Linear(in_dim, out_dim, bias)
        self.lin_u = torch.nn.Linear(in_dim, out_dim, bias)
        self.alpha_u = torch.nn.Linear(self.out_dim * 2, 1)
        self.alpha_b = torch.nn.Linear(self.out_dim * 2, 1)
        self._plans = _plan.PlanCache(capacity=4)
        self.reset_parameters()

    def reset_parameters(self):
        self.lin_b.reset_parameters()
        self.lin_u.reset_parameters()
        torch.nn.init.xavier_normal_(self.alpha_b.weight)
        torch.nn.init.xavier_normal_(self.alpha_u.weight)
        self._plans.clear()

    # remove_self_loops [+ add_self_loops over 0..max id] -> CSR by target
    def _plan_for(self, edge_index: Tensor, n: int, with_loops: bool) -> _plan.CSRPlan:
        def build():
            keep = edge_index[0] != edge_index[1]
            ei = edge_index[:, keep]
            if with_loops:
                m = int(ei.max().item()) + 1 if ei.numel() > 0 else 0
                loops = torch.arange(m, device=ei.device, dtype=ei.dtype)
                ei = torch.cat([ei, torch.stack([loops, loops])], dim=1)
            return _plan.build_csr(ei.contiguous(), None, n, n, "source_to_target")
        return self._plans.get((edge_index,), (n, with_loops), build)

    def _scores(self, x: Tensor, alpha_lin: torch.nn.Linear) -> Tuple[Tensor, Tensor]:
        # Linear(2*out -> 1) on [x_j || x_i]  ==  x_j . a_j  +  (x_i . a_i + c)
        w = alpha_lin.weight.detach().view(2, self.out_dim).t()          # [out, 2]: (a_j, a_i)
        bias = torch.cat([torch.zeros(1, device=x.device), alpha_lin.bias.detach().float()])
        s = ops.dense([(x, w, 0)], 2, bias=bias)[0]
        return s[:, 0], s[:, 1]

    def _attend(self, plans, xs, alpha_lin) -> Tensor:
        sc = [self._scores(x, alpha_lin) for x in xs]
        y, _ = ops.edge_softmax(plans, [s[0] for s in sc], [s[1] for s in sc], act="tanh", xd=xs)
        return y

    def _lin(self, lin: torch.nn.Linear, x: Tensor) -> Tensor:
        return ops.dense([(x, lin.weight.detach().t(), 0)], self.out_dim, bias=lin.bias)[0]

    def forward(self, x: Union[Tensor, Tuple[Tensor, Tensor]], pos_edge_index: Tensor,
                neg_edge_index: Tensor) -> Tensor:
        if not isinstance(x, Tensor):
            raise NotImplementedError("SNEAConv kernels take a single feature tensor")
        _plan.require_cuda(x, "x")
        n = x.size(0)
        if self.first_aggr:
            h_b, h_u = self._lin(self.lin_b, x), self._lin(self.lin_u, x)
            out_b = self._attend([self._plan_for(pos_edge_index, n, True)], [h_b], self.alpha_b)
            out_u = self._attend([self._plan_for(neg_edge_index, n, True)], [h_u], self.alpha_u)
        else:
            fi = self.in_dim
            h_b, h_u = x[:, :fi], x[:, fi:]
            plans = [self._plan_for(pos_edge_index, n, True), self._plan_for(neg_edge_index, n, False)]
            out_b = self._attend(plans, [self._lin(self.lin_b, h_b), self._lin(self.lin_b, h_u)], self.alpha_b)
            out_u = self._attend(plans, [self._lin(self.lin_u, h_u), self._lin(self.lin_u, h_b)], self.alpha_u)
        return torch.cat([out_b, out_u], dim=-1)

    def __repr__(self) -> str:
        return (f'{self.__class__.__name__}({self.in_dim}, '
                f'{self.out_dim}, first_aggr={self.first_aggr})')

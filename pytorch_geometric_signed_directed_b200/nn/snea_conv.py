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
Training: when gradients are required the same kernels run under `autograd.py`'s Functions
(`pgsd_edge_softmax_backward` for the attention, the transform's own backward for the Linears).
Edge bookkeeping follows the reference exactly: self-loops are removed, then re-added for the
nodes 0..max(remaining edge ids) only (add_self_loops without num_nodes, :88-89,110-111);
negative edges never get self-loops in the deep layer (:112).
"""
from __future__ import annotations

from typing import Tuple, Union

import torch
from torch import Tensor

from .. import autograd as ag, ops, plan as _plan


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
    def _request(self, edge_index: Tensor, n: int, with_loops: bool):
        def build():
            keep = edge_index[0] != edge_index[1]
            ei = edge_index[:, keep]
            if with_loops:
                m = int(ei.max().item()) + 1 if ei.numel() > 0 else 0
                loops = torch.arange(m, device=ei.device, dtype=ei.dtype)
                ei = torch.cat([ei, torch.stack([loops, loops])], dim=1)
            return _plan.build_csr(ei.contiguous(), None, n, n, "source_to_target")
        return ((edge_index,), (n, with_loops), build)

    def _plan_for(self, edge_index: Tensor, n: int, with_loops: bool) -> _plan.CSRPlan:
        return self._plans.get(*self._request(edge_index, n, with_loops))

    def _transforms(self, x: Tensor, specs):
        """All Linear applications of one forward in TWO launches.  specs = [(lin, first input column, alpha_lin)]:
        feature k is h_k = lin(x[:, c0:c0+in]) and its attention scores are s_src = h_k . a_j, s_dst = h_k . a_i + c
        with (a_j, a_i, c) = alpha_lin (SNEAConv.py:137-142: Linear(2*out -> 1) on [h_j || h_i]).  Both are linear in
        x, so (i) the h_k are the column blocks of one transform with block-structured weights and (ii) all score
        pairs are the columns of a second, narrow transform of x (zero-padded to 16 columns for the tensor-core
        kernel) whose weights are W^T a and whose bias carries b . a (+ c)."""
        fi, fo, dev = self.in_dim, self.out_dim, x.device
        k = len(specs)
        w_all = x.new_zeros((x.size(1), k * fo), dtype=torch.float32)
        b_all = torch.zeros(k * fo, device=dev)
        width = 16 if 2 * k <= 16 else 32
        sw = torch.zeros((x.size(1), width), device=dev)
        sb = torch.zeros(width, device=dev)
        for i, (lin, c0, alpha_lin) in enumerate(specs):
            wt = lin.weight.detach().t().float()                                   # [in, out]
            w_all[c0:c0 + fi, i * fo:(i + 1) * fo] = wt
            a = alpha_lin.weight.detach().view(2, fo).t().float()                  # [out, 2]: (a_j, a_i)
            sw[c0:c0 + fi, 2 * i:2 * i + 2] = wt @ a
            if lin.bias is not None:
                b_all[i * fo:(i + 1) * fo] = lin.bias.detach().float()
                sb[2 * i:2 * i + 2] = lin.bias.detach().float() @ a
            sb[2 * i + 1] += alpha_lin.bias.detach().float().view(())
        h_all = ops.dense([(x, w_all, 0)], k * fo, bias=b_all)[0]
        s_all = ops.dense([(x, sw, 0)], width, bias=sb)[0]
        hs = [h_all[:, i * fo:(i + 1) * fo] for i in range(k)]
        ss = [(s_all[:, 2 * i].contiguous(), s_all[:, 2 * i + 1].contiguous()) for i in range(k)]
        return hs, ss

    def _transforms_train(self, x: Tensor, specs):
        """The same two launches on the autograd path: the block-structured weights are assembled with differentiable
        torch ops (pads and sums of the layer's own parameters, a few KB), the launches are `ag.dense`."""
        fi, fo = self.in_dim, self.out_dim
        k, kin = len(specs), x.size(1)
        width = 16 if 2 * k <= 16 else 32
        pad = torch.nn.functional.pad
        w_all = b_all = sw = sb = None
        add = lambda acc, t: t if acc is None else acc + t
        for i, (lin, c0, alpha_lin) in enumerate(specs):
            wt = lin.weight.t().float()                                              # [in, out]
            a = alpha_lin.weight.view(2, fo).t().float()                             # [out, 2]: (a_j, a_i)
            rows = (0, 0, c0, kin - c0 - fi)                                         # place rows c0 .. c0+in
            w_all = add(w_all, pad(wt, (i * fo, (k - 1 - i) * fo) + rows[2:]))
            sw = add(sw, pad(wt @ a, (2 * i, width - 2 * i - 2) + rows[2:]))
            bvec = lin.bias.float() if lin.bias is not None else wt.new_zeros(fo)
            b_all = add(b_all, pad(bvec, (i * fo, (k - 1 - i) * fo)))
            sc = bvec @ a + pad(alpha_lin.bias.float().view(1), (1, 0))               # (b.a_j, b.a_i + c)
            sb = add(sb, pad(sc, (2 * i, width - 2 * i - 2)))
        h_all = ag.dense([(x, w_all, 0)], k * fo, bias=b_all)[0]
        s_all = ag.dense([(x, sw, 0)], width, bias=sb)[0]
        hs = [h_all[:, i * fo:(i + 1) * fo] for i in range(k)]
        ss = [(s_all[:, 2 * i].contiguous(), s_all[:, 2 * i + 1].contiguous()) for i in range(k)]
        return hs, ss

    def _needs_grad(self, x: Tensor) -> bool:
        return torch.is_grad_enabled() and (x.requires_grad or any(p.requires_grad for p in self.parameters()))

    def _attend(self, plans, hs, ss) -> Tensor:
        if ag._needs_grad(list(hs) + [t for pair in ss for t in pair]):
            return ag.edge_softmax_sum(plans, [s[0] for s in ss], [s[1] for s in ss], list(hs), act="tanh")
        y, _ = ops.edge_softmax(plans, [s[0] for s in ss], [s[1] for s in ss], act="tanh", xd=hs)
        return y

    def forward(self, x: Union[Tensor, Tuple[Tensor, Tensor]], pos_edge_index: Tensor,
                neg_edge_index: Tensor) -> Tensor:
        if not isinstance(x, Tensor):
            raise NotImplementedError("SNEAConv kernels take a single feature tensor")
        _plan.require_cuda(x, "x")
        n = x.size(0)
        train = self._needs_grad(x)
        transforms = self._transforms_train if train else self._transforms
        with torch.enable_grad() if train else torch.no_grad():
            # the transforms are enqueued before the plans are validated (one fingerprint read-back for both)
            if self.first_aggr:
                hs, ss = transforms(x, [(self.lin_b, 0, self.alpha_b), (self.lin_u, 0, self.alpha_u)])
                p_pos, p_neg = self._plans.get_many([self._request(pos_edge_index, n, True),
                                                     self._request(neg_edge_index, n, True)])
                out_b = self._attend([p_pos], hs[:1], ss[:1])
                out_u = self._attend([p_neg], hs[1:], ss[1:])
            else:
                fi = self.in_dim                       # x = [h_b | h_u]
                hs, ss = transforms(x, [(self.lin_b, 0, self.alpha_b), (self.lin_b, fi, self.alpha_b),
                                        (self.lin_u, fi, self.alpha_u), (self.lin_u, 0, self.alpha_u)])
                plans = self._plans.get_many([self._request(pos_edge_index, n, True),
                                              self._request(neg_edge_index, n, False)])
                out_b = self._attend(plans, hs[:2], ss[:2])
                out_u = self._attend(plans, hs[2:], ss[2:])
        return torch.cat([out_b, out_u], dim=-1)

    def __repr__(self) -> str:
        return (f'{self.__class__.__name__}({self.in_dim}, '
                f'{self.out_dim}, first_aggr={self.first_aggr})')

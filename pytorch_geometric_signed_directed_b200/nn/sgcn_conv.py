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
        # set by the SGCN model wrapper: z = tanh(conv(...)) (SGCN.py:93-96) as the transform's epilogue
        self.fused_tanh = False

    def reset_parameters(self):
        self.lin_b.reset_parameters()
        self.lin_u.reset_parameters()
        self._plans.clear()

    def _plan_for(self, edge_index: Tensor, n_dst: int, n_src: int):
        edge_index = _plan.as_edge_index(edge_index)
        return self._plans.get((edge_index,), (n_dst, n_src),
                               lambda: _plan.build_csr(edge_index, None, n_dst, n_src, "source_to_target"))

    def _plans_for(self, pos_edge_index: Tensor, neg_edge_index: Tensor, n_dst: int, n_src: int):
        # a SparseTensor / torch sparse adjacency (SGCNConv.py:131-134) is read as adj_t[target, source]
        pos_edge_index, neg_edge_index = _plan.as_edge_index(pos_edge_index), _plan.as_edge_index(neg_edge_index)
        mk = lambda ei: ((ei,), (n_dst, n_src), lambda: _plan.build_csr(ei, None, n_dst, n_src, "source_to_target"))
        return self._plans.get_many([mk(pos_edge_index), mk(neg_edge_index)])

    def _folded_weights(self, wb: Tensor, wu: Tensor):
        """[Wb1 | Wu1 | Wb2 | Wu2] and its bias for `_first_layer_folded`; on the inference path the assembled
        tensors are kept until a parameter changes (version counters), so a forward does not re-run the small cats."""
        fi, fo = self.in_dim, self.out_dim
        bb, bu = self.lin_b.bias, self.lin_u.bias
        grad = ag._needs_grad([self.lin_b.weight, self.lin_u.weight, bb, bu])
        key = tuple((t.data_ptr(), t._version) for t in (self.lin_b.weight, self.lin_u.weight, bb, bu) if t is not None)
        hit = getattr(self, "_folded_cache", None)
        if not grad and hit is not None and hit[0] == key:
            return hit[1], hit[2]
        w = torch.cat([wb[:fi], wu[:fi], wb[fi:], wu[fi:]], 1)                 # [in, 4*out]
        bias = None if bb is None else torch.cat([bb.new_zeros(2 * fo), bb, bu])
        if not grad:
            self._folded_cache = (key, w.detach().contiguous(), None if bias is None else bias.detach())
        return w, bias

    def forward(self, x: Union[Tensor, Tuple[Tensor, Tensor]], pos_edge_index: Tensor,
                neg_edge_index: Tensor) -> Tensor:
        x_src, x_dst = (x, x) if isinstance(x, Tensor) else x
        _plan.require_cuda(x_src, "x")
        n_src, n_dst = x_src.size(0), x_dst.size(0)
        fi, fo = self.in_dim, self.out_dim
        wb, wu = self.lin_b.weight.t(), self.lin_u.weight.t()          # [mult*in, out] views
        if self.first_aggr and x_src is x_dst and fo <= fi and (fo * x_src.element_size()) % 16 == 0:
            return self._first_layer_folded(x_src, pos_edge_index, neg_edge_index, wb, wu)
        pos, neg = self._plans_for(pos_edge_index, neg_edge_index, n_dst, n_src)
        m_pos = ag.spmm(pos, [x_src], (0,), mean=True)[0]
        m_neg = ag.spmm(neg, [x_src], (0,), mean=True)[0]
        # Both halves of the output in ONE transform launch: lin_b and lin_u become the two column blocks
        # of block-structured [k, 2*out] weights (zeros where an input does not feed a half), so m_pos,
        # m_neg and x are each read once and `torch.cat([out_b, out_u])` (SGCNConv.py:121) is the
        # kernel's output layout.  The zero blocks cost tensor-core flops only (free: HBM-bound).
        zb = wb.new_zeros((fi, fo))
        if self.first_aggr:
            terms = [(m_pos, torch.cat([wb[:fi], zb], 1), 0),
                     (m_neg, torch.cat([zb, wu[:fi]], 1), 0),
                     (x_dst, torch.cat([wb[fi:], wu[fi:]], 1), 0)]
        else:
            # x = [x_b | x_u]; m_pos = [mean+(x_b) | mean+(x_u)], m_neg = [mean-(x_b) | mean-(x_u)]
            terms = [(m_pos, torch.cat([torch.cat([wb[:fi], zb], 1), torch.cat([zb, wu[:fi]], 1)], 0), 0),
                     (m_neg, torch.cat([torch.cat([zb, wu[fi:2 * fi]], 1), torch.cat([wb[fi:2 * fi], zb], 1)], 0), 0),
                     (x_dst, torch.cat([torch.cat([wb[2 * fi:], zb], 1), torch.cat([zb, wu[2 * fi:]], 1)], 0), 0)]
        bias = None
        if self.lin_b.bias is not None:
            bias = torch.cat([self.lin_b.bias, self.lin_u.bias])
        fuse = self.fused_tanh and not self.norm_emb and not ag._needs_grad(
            [t for t, _, _ in terms] + [self.lin_b.weight, self.lin_u.weight, bias])
        out = ag.dense(terms, 2 * fo, bias=bias, relu_mode=2 if fuse else 0)[0]
        if self.norm_emb:
            out = F.normalize(out, p=2, dim=-1)
        if self.fused_tanh and not fuse:
            out = torch.tanh(out)
        return out

    def _first_layer_folded(self, x: Tensor, pos_edge_index: Tensor, neg_edge_index: Tensor, wb: Tensor,
                            wu: Tensor) -> Tensor:
        """First layer with the Linear folded THROUGH the mean aggregations (both are linear):
            out_b = mean+(x) Wb1 + x Wb2 + bb = mean+(x Wb1) + (x Wb2 + bb),   out_u likewise with mean-.
        One transform writes [x Wb1 | x Wu1 | x Wb2 + bb | x Wu2 + bu]; each sign then aggregates its own
        out-wide block (half the gathered bytes of aggregating x itself when out = in / 2, SGCN.py:68-73) and adds
        the self block in the aggregation epilogue -- the [N, in] means of the reference (SGCNConv.py:100-104) and
        their `cat` never exist.  tanh (applied by SGCN right after the layer) rides in the same epilogue."""
        fi, fo = self.in_dim, self.out_dim
        w, bias = self._folded_weights(wb, wu)
        # the transform is enqueued BEFORE the plans are looked up: validating a cached plan reads a fingerprint of
        # the edge tensors back to the host, and that wait then overlaps this launch instead of idling the GPU
        ps = ag.dense([(x, w, 0)], 4 * fo, bias=bias)[0]
        pos, neg = self._plans_for(pos_edge_index, neg_edge_index, x.size(0), x.size(0))
        grad = ag._needs_grad([ps])
        # hub rows (> 4096 entries) are finished by a second launch with atomics: no tanh epilogue there
        fuse = (self.fused_tanh and not self.norm_emb and not grad
                and pos.hub_rows() is None and neg.hub_rows() is None)
        if grad:
            out = torch.cat([ag.spmm(pos, [ps[:, :fo]], (0,), mean=True, beta=1.0, zs=[ps[:, 2 * fo:3 * fo]])[0],
                             ag.spmm(neg, [ps[:, fo:2 * fo]], (0,), mean=True, beta=1.0, zs=[ps[:, 3 * fo:]])[0]], 1)
        else:
            out = torch.empty((x.size(0), 2 * fo), dtype=x.dtype, device=x.device)
            ops.spmm(pos, [ps[:, :fo]], (0,), mean=True, beta=1.0, zs=[ps[:, 2 * fo:3 * fo]], out=[out[:, :fo]],
                     tanh_out=fuse)
            ops.spmm(neg, [ps[:, fo:2 * fo]], (0,), mean=True, beta=1.0, zs=[ps[:, 3 * fo:]], out=[out[:, fo:]],
                     tanh_out=fuse)
        if self.norm_emb:
            out = F.normalize(out, p=2, dim=-1)
        if self.fused_tanh and not fuse:
            out = torch.tanh(out)
        return out

    def __repr__(self) -> str:
        return (f'{self.__class__.__name__}({self.in_dim}, '
                f'{self.out_dim}, first_aggr={self.first_aggr})')

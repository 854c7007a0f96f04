"""Differentiable wrappers of the two compute primitives.

Every example of the reference trains its model (`examples/magnet_node.py:22-29`), i.e. relies on
autograd through `index_select` / `scatter_add_` / `matmul`.  Here the same derivatives are
expressed with the forward kernels themselves:

  y = alpha * M x (+ beta z + bias)   ->  gx = alpha * M^T gy,  gz = beta * gy,  gbias = colsum(gy)
  y = sum_t X_t W_t (+ b)             ->  gX_t = G W_t^T (pgsd_dense_transform with swapped weight
                                          strides), gW_t = X_t^T G and gb (pgsd_xtg_accumulate)

M^T comes for free for the magnetic plans (real part symmetric, imaginary part antisymmetric:
same plan, `op_scale = (1, -1)`); other plans get a transposed CSR built once and cached on the
plan.  `spmm` / `dense` fall through to the raw kernels when no gradient is required, so the
inference path is unchanged.  A trainable magnetic charge q (MagNetConv.py:58-59) gets its gradient
from `pgsd_magnetic_q_grad` (d val/dq is a rotation of the stored values, so only an SDDMM reduced
to a scalar is needed).  Not differentiated: edge weights.
"""
from __future__ import annotations

from typing import List, Optional, Sequence, Tuple

import torch
from torch import Tensor

from . import ops, plan as _plan
from .plan import CSRPlan


def _needs_grad(tensors) -> bool:
    return torch.is_grad_enabled() and any(t is not None and t.requires_grad for t in tensors)


# ------------------------------------------------------------------------------ transposed plans

def transposed(plan: CSRPlan) -> Tuple[CSRPlan, Tuple[float, float]]:
    """(plan computing M^T, per-operator sign).  Cached on the plan object."""
    cached = getattr(plan, "_transposed", None)
    if cached is not None:
        return cached
    if plan.meta.get("hermitian"):
        res = (plan, (1.0, -1.0))
    else:
        n_rows = plan.n_dst
        counts = (plan.row_ptr[1:] - plan.row_ptr[:-1]).long()
        rows = torch.repeat_interleave(torch.arange(n_rows, device=plan.device), counts)
        ei = torch.stack([rows, plan.col.long()])          # old destination -> new source
        vals, base = [], None
        for v in plan.val:
            t = _plan.build_csr(ei, v, plan.n_src, plan.n_dst, "source_to_target")
            base = base or t
            vals.append(t.val[0])
        if base is None:
            base = _plan.build_csr(ei, None, plan.n_src, plan.n_dst, "source_to_target")
        res = (CSRPlan(plan.n_src, plan.n_dst, plan.nnz, plan.num_input_edges, base.row_ptr, base.col,
                       vals, list(plan.diag), list(plan.diag_const), {}), (1.0, 1.0))
    plan._transposed = res
    return res


# ------------------------------------------------------------------------------------- spmm

class _SpmmFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, plan, op_ids, mean, alpha, beta, n_x, has_z, has_bias, q, *tensors):
        xs = list(tensors[:n_x])
        zs = list(tensors[n_x:2 * n_x]) if has_z else None
        bias = tensors[-1] if has_bias else None
        outs = ops.spmm(plan, xs, op_ids, mean=mean, alpha=alpha, beta=beta, zs=zs, bias=bias)
        ctx.plan, ctx.op_ids, ctx.mean, ctx.alpha, ctx.beta = plan, op_ids, mean, alpha, beta
        ctx.n_x, ctx.has_z, ctx.has_bias = n_x, has_z, has_bias
        ctx.q_shape = None
        if q is not None and q.requires_grad:
            ctx.q_shape = (tuple(q.shape), q.dtype)
            ctx.save_for_backward(*xs)
        return tuple(outs)

    @staticmethod
    def backward(ctx, *gys):
        plan, op_ids = ctx.plan, ctx.op_ids
        gys = [g.contiguous() for g in gys]
        tplan, sign = transposed(plan)
        gin = gys
        if ctx.mean:
            inv = 1.0 / (plan.row_ptr[1:] - plan.row_ptr[:-1]).clamp(min=1).to(gys[0].dtype)
            gin = [g * inv.view(-1, 1) for g in gys]
        gq = None
        if ctx.q_shape is not None and ctx.needs_input_grad[8]:
            gq = ops.magnetic_q_grad(plan, gin, ctx.saved_tensors, ctx.alpha)
            gq = gq.to(ctx.q_shape[1]).reshape(ctx.q_shape[0])
        need_x = any(ctx.needs_input_grad[9 + k] for k in range(ctx.n_x))
        gxs = [None] * ctx.n_x
        if need_x:
            if len(op_ids) == 2:
                gxs = ops.spmm(tplan, gin, op_ids, alpha=ctx.alpha,
                               op_scale=(sign[op_ids[0]], sign[op_ids[1]]))
            else:
                gxs = ops.spmm(tplan, gin, op_ids, alpha=ctx.alpha * sign[op_ids[0]])
        grads = list(gxs)
        if ctx.has_z:
            grads += [ctx.beta * g for g in gys]
        if ctx.has_bias:
            gb = gys[0].float().sum(0)
            for g in gys[1:]:
                gb = gb + g.float().sum(0)
            grads.append(gb)
        return (None,) * 8 + (gq,) + tuple(grads)


def spmm(plan: CSRPlan, xs: Sequence[Tensor], op_ids: Sequence[int] = (0,), *, mean: bool = False,
         alpha: float = 1.0, beta: float = 0.0, zs: Optional[Sequence[Tensor]] = None,
         bias: Optional[Tensor] = None, out=None, q: Optional[Tensor] = None) -> List[Tensor]:
    """`q`: the trainable magnetic charge the plan's values were computed from (plan built with
    `keep_theta`); it then receives d loss / d q."""
    op_ids = tuple(op_ids)
    track = list(xs) + (list(zs) if zs is not None else []) + ([bias] if bias is not None else [])
    if not _needs_grad(track + [q]):
        return ops.spmm(plan, xs, op_ids, mean=mean, alpha=alpha, beta=beta, zs=zs, bias=bias, out=out)
    tensors = list(xs) + (list(zs) if zs is not None else []) + ([bias] if bias is not None else [])
    outs = list(_SpmmFn.apply(plan, op_ids, mean, alpha, beta, len(xs), zs is not None,
                              bias is not None, q, *tensors))
    if out is not None:
        raise RuntimeError("spmm: preallocated outputs are not supported on the autograd path")
    return outs


# ------------------------------------------------------------------------------------ dense

class _DenseFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, n_out, combine, relu_mode, groups, has_bias, n_terms, *tensors):
        xs, ws = tensors[:n_terms], tensors[n_terms:2 * n_terms]
        bias = tensors[-1] if has_bias else None
        terms = [(x, w, g) for x, w, g in zip(xs, ws, groups)]
        if combine and relu_mode == 1:
            # complex ReLU (complex_relu.py:21-22): the mask is `real >= 0` of the PRE-activation; it cannot be
            # recovered from the outputs (real == 0 passes the imaginary part, real < 0 zeroes both), so the
            # training path keeps it: one transform without the fused mask + one elementwise pass
            pre = ops.dense(terms, n_out, bias=bias, combine=True, relu_mode=0)
            mask = pre[0] >= 0
            outs = [pre[0] * mask, pre[1] * mask]
            extra = (mask,)
        else:
            outs = ops.dense(terms, n_out, bias=bias, combine=combine, relu_mode=relu_mode)
            extra = tuple(outs[:1]) if relu_mode else ()
        ctx.save_for_backward(*xs, *ws, *extra)
        ctx.cfg = (n_out, combine, relu_mode, groups, has_bias, n_terms)
        return tuple(outs)

    @staticmethod
    def backward(ctx, *gys):
        n_out, combine, relu_mode, groups, has_bias, n_terms = ctx.cfg
        saved = ctx.saved_tensors
        xs, ws = saved[:n_terms], saved[n_terms:2 * n_terms]
        g0 = gys[0].contiguous()
        if combine:
            g1 = gys[1].contiguous()
            if relu_mode == 1:
                m = saved[-1].to(g0.dtype)             # complex ReLU mask of the forward pass
                g0, g1 = g0 * m, g1 * m
            elif relu_mode:
                raise RuntimeError("dense: only the complex ReLU epilogue is differentiable with combine=True")
            gg = [g0 + g1, g1 - g0]                    # dA, dB of out_real = A-B+b, out_imag = A+B+b
        else:
            if relu_mode == 2:                         # tanh epilogue: d tanh = 1 - out^2
                g0 = g0 * (1.0 - saved[-1].to(g0.dtype) ** 2)
            elif relu_mode:
                raise RuntimeError(f"dense: epilogue {relu_mode} is not differentiable with combine=False")
            gg = [g0]
        dev = g0.device
        gxs, gws = [None] * n_terms, [None] * n_terms
        gb = torch.zeros(n_out, dtype=torch.float32, device=dev) if has_bias else None
        bias_done = not has_bias
        for t in range(n_terms):
            g = gg[groups[t]]
            if ctx.needs_input_grad[6 + t]:
                gxs[t] = ops.dense([(g, ws[t].t(), 0)], ws[t].size(0))[0]
            want_b = (not bias_done) and groups[t] == 0
            if ctx.needs_input_grad[6 + n_terms + t] or want_b:
                gw = torch.zeros((ws[t].size(0), n_out), dtype=torch.float32, device=dev)
                ops.xtg_accumulate(xs[t], g, gw, gb if want_b else None)
                bias_done = bias_done or want_b
                if ctx.needs_input_grad[6 + n_terms + t]:
                    gws[t] = gw.to(ws[t].dtype)
        if has_bias and not bias_done:
            gb = gg[0].float().sum(0)
        return (None,) * 6 + tuple(gxs) + tuple(gws) + ((gb,) if has_bias else ())


def dense(terms: Sequence[Tuple[Tensor, Tensor, int]], n_out: int, *, bias: Optional[Tensor] = None,
          combine: bool = False, relu_mode: int = 0, out=None) -> List[Tensor]:
    track = [t[0] for t in terms] + [t[1] for t in terms] + ([bias] if bias is not None else [])
    if not _needs_grad(track):
        return ops.dense(terms, n_out, bias=bias, combine=combine, relu_mode=relu_mode, out=out)
    if out is not None:
        raise RuntimeError("dense: preallocated outputs are not supported on the autograd path")
    xs = [t[0] for t in terms]
    ws = [t[1] for t in terms]
    groups = tuple(int(t[2]) for t in terms)
    tensors = xs + ws + ([bias] if bias is not None else [])
    return list(_DenseFn.apply(n_out, combine, relu_mode, groups, bias is not None, len(terms), *tensors))


# ----------------------------------------------------------------------------- attention layers
# Training through SNEAConv (nn/signed/SNEAConv.py:135-146) and the GATConv of SDGNN / SiGAT (nn/signed/SDGNN.py:
# 35-64): the forward kernels are reused; the backward is one segment-softmax-backward kernel
# (`pgsd_edge_softmax_backward`), and for the GAT-style weighted sum an SDDMM (`pgsd_sddmm_rows`, dL/dalpha) plus the
# transposed aggregation with the same alpha (`pgsd_spmm_csr` on a transposed pattern cached on the plan).

def transposed_pattern(plan: CSRPlan):
    """(row_ptr_T, col_T, perm): CSR of the transposed pattern and the permutation that carries a per-entry array
    of `plan` into its order.  Built once per plan (stable: entries of a transposed row keep destination order)."""
    cached = getattr(plan, "_tpattern", None)
    if cached is None:
        dev = plan.device
        counts = (plan.row_ptr[1:] - plan.row_ptr[:-1]).long()
        rows = torch.repeat_interleave(torch.arange(plan.n_dst, device=dev), counts)
        col = plan.col[:plan.nnz].long()
        perm = torch.sort(col * plan.n_dst + rows, stable=True).indices
        rp = torch.zeros(plan.n_src + 1, dtype=torch.int32, device=dev)
        if plan.nnz:
            rp[1:] = torch.cumsum(torch.bincount(col, minlength=plan.n_src), 0).int()
        cached = (rp, rows[perm].int().contiguous(), perm)
        plan._tpattern = cached
    return cached


class _EdgeSoftmaxSumFn(torch.autograd.Function):
    """y[i] = sum_t xd_t[i] * (sum of alpha over row i's entries of type t)  -- SNEAConv's aggregation."""

    @staticmethod
    def forward(ctx, plans, act, slope, n_types, *tensors):
        s_src, s_dst, xd = tensors[:n_types], tensors[n_types:2 * n_types], tensors[2 * n_types:]
        y, _ = ops.edge_softmax(plans, list(s_src), list(s_dst), act=act, slope=slope, xd=list(xd))
        ctx.save_for_backward(*tensors)
        ctx.cfg = (plans, act, slope, n_types)
        return y

    @staticmethod
    def backward(ctx, gy):
        plans, act, slope, n_types = ctx.cfg
        t = ctx.saved_tensors
        s_src, s_dst, xd = t[:n_types], t[n_types:2 * n_types], t[2 * n_types:]
        gy = gy.contiguous()
        coef = [(gy * x).sum(1) for x in xd]                       # dL/dalpha of every entry of type t in row i
        g_src, g_dst, sums = ops.edge_softmax_backward(plans, s_src, s_dst, act=act, slope=slope, row_coef=coef,
                                                       want_type_sum=True)
        g_xd = [gy * sm.view(-1, 1) for sm in sums]
        fit = lambda g, like: g[:like.numel()].view(like.shape).to(like.dtype)
        return (None,) * 4 + tuple(fit(g, s) for g, s in zip(g_src, s_src)) \
            + tuple(fit(g, s) for g, s in zip(g_dst, s_dst)) + tuple(g_xd)


def edge_softmax_sum(plans: Sequence[CSRPlan], s_src: Sequence[Tensor], s_dst: Sequence[Tensor],
                     xd: Sequence[Tensor], act: str = "tanh", slope: float = 0.2) -> Tensor:
    tensors = list(s_src) + list(s_dst) + list(xd)
    if not _needs_grad(tensors):
        return ops.edge_softmax(plans, list(s_src), list(s_dst), act=act, slope=slope, xd=list(xd))[0]
    return _EdgeSoftmaxSumFn.apply(list(plans), act, slope, len(plans), *tensors)


class _GatAttendFn(torch.autograd.Function):
    """y = sum_e alpha_e h[col_e] (+ bias),  alpha = softmax_row(leaky_relu(s_src[col] + s_dst[row]))."""

    @staticmethod
    def forward(ctx, plan, slope, has_bias, s_src, s_dst, h, *rest):
        bias = rest[0] if has_bias else None
        _, alphas = ops.edge_softmax([plan], [s_src], [s_dst], act="leaky_relu", slope=slope, want_alpha=True)
        alpha = alphas[0]
        weighted = CSRPlan(plan.n_dst, plan.n_src, plan.nnz, plan.num_input_edges, plan.row_ptr, plan.col, [alpha],
                           [None], [0.0])
        weighted._hubs = plan.hub_rows()
        y = ops.spmm(weighted, [h], (0,), bias=bias)[0]
        ctx.save_for_backward(s_src, s_dst, h, alpha)
        ctx.cfg = (plan, slope, has_bias)
        return y

    @staticmethod
    def backward(ctx, gy):
        plan, slope, has_bias = ctx.cfg
        s_src, s_dst, h, alpha = ctx.saved_tensors
        gy = gy.contiguous()
        dalpha = ops.sddmm_rows(plan, gy, h)
        g_src, g_dst, _ = ops.edge_softmax_backward([plan], [s_src], [s_dst], act="leaky_relu", slope=slope,
                                                    dalpha=[dalpha])
        rp_t, col_t, perm = transposed_pattern(plan)
        tplan = CSRPlan(plan.n_src, plan.n_dst, plan.nnz, plan.num_input_edges, rp_t, col_t, [alpha[perm].contiguous()],
                        [None], [0.0])
        g_h = ops.spmm(tplan, [gy], (0,))[0][:h.size(0)].to(h.dtype)
        fit = lambda g, like: g[:like.numel()].view(like.shape).to(like.dtype)
        grads = (None, None, None, fit(g_src[0], s_src), fit(g_dst[0], s_dst), g_h)
        return grads + ((gy.float().sum(0),) if has_bias else ())


def gat_attend(plan: CSRPlan, s_src: Tensor, s_dst: Tensor, h: Tensor, slope: float = 0.2,
               bias: Optional[Tensor] = None) -> Tensor:
    """Differentiable GAT-style attention aggregation (used on the training path; inference keeps the fused /
    preallocated routes of nn.GATConv.aggregate)."""
    tensors = [s_src, s_dst, h] + ([bias] if bias is not None else [])
    return _GatAttendFn.apply(plan, float(slope), bias is not None, *tensors)

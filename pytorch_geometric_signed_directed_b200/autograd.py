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
        outs = ops.dense(terms, n_out, bias=bias, combine=combine, relu_mode=relu_mode)
        ctx.save_for_backward(*xs, *ws, *(outs[:1] if relu_mode else ()))
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
            if relu_mode:
                m = (saved[-1] > 0).to(g0.dtype)      # complex ReLU mask (treated as a constant)
                g0, g1 = g0 * m, g1 * m
            gg = [g0 + g1, g1 - g0]                    # dA, dB of out_real = A-B+b, out_imag = A+B+b
        else:
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

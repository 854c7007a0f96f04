"""TEST INFRASTRUCTURE ONLY -- torch-only stand-in for the PyTorch-Geometric symbols
the reference's hot-path files import.

Why this exists: the reference (PyGSD 1.1.1) delegates all arithmetic on the hot path to
`torch_geometric` (unpinned, setup.py:12), which is not installed in this image and cannot
be (no network, no wheel).  The reference's own layer files
(`torch_geometric_signed_directed/nn/directed/MagNetConv.py`, ...) DO run unmodified once
the handful of PyG symbols they import exist.  `install()` registers those symbols under
`sys.modules['torch_geometric*']` so that `oracle/load_reference.py` can import the
reference's files by path and `tests/golden/make_golden.py` can produce fixtures from them.

Semantics restated from the PyG 2.x documentation/behaviour:
  * MessagePassing.propagate: flow 'source_to_target' gathers `*_j` from edge_index[0] and
    aggregates on edge_index[1]; 'target_to_source' swaps the roles.  `*_i` is gathered with
    the aggregation index.  Tuple inputs are (source-side tensor, target-side tensor).
  * scatter(sum) = zeros.scatter_add_; scatter(mean) = sum / clamp(count, 1).
  * coalesce = stable sort on row*N+col, duplicates reduced, row-major sorted output.
  * add_self_loops appends arange(N) pairs AFTER the existing edges.
  * add_remaining_self_loops keeps existing self-loop weights, fills the rest.
  * softmax(src, index) = segment softmax with max-shift and +1e-16 in the denominator.

PARITY NOTE: this shim cannot be diffed against a real PyG here ("parity pinned to the
reference's source files + this restated PyG; real PyG unpinned").  Nothing in the product
package imports it.
"""
from __future__ import annotations

import inspect
import sys
import types
from typing import Optional, Tuple, Union

import torch
from torch import Tensor

# --------------------------------------------------------------------------- utils


def maybe_num_nodes(edge_index, num_nodes=None):
    if num_nodes is not None:
        return int(num_nodes)
    if isinstance(edge_index, Tensor):
        return int(edge_index.max()) + 1 if edge_index.numel() > 0 else 0
    raise TypeError("shim: only dense edge_index tensors are supported")


def _broadcast(index: Tensor, ref: Tensor, dim: int) -> Tensor:
    shape = [1] * ref.dim()
    shape[dim] = -1
    return index.view(shape).expand_as(ref)


def scatter(src: Tensor, index: Tensor, dim: int = 0, dim_size: Optional[int] = None,
            reduce: str = 'sum') -> Tensor:
    if dim < 0:
        dim = src.dim() + dim
    if dim_size is None:
        dim_size = int(index.max()) + 1 if index.numel() > 0 else 0
    shape = list(src.shape)
    shape[dim] = dim_size
    if reduce in ('sum', 'add'):
        return src.new_zeros(shape).scatter_add_(dim, _broadcast(index, src, dim), src)
    if reduce == 'mean':
        cnt = src.new_zeros(dim_size).scatter_add_(0, index, src.new_ones(index.numel()))
        cnt = cnt.clamp_(min=1)
        out = src.new_zeros(shape).scatter_add_(dim, _broadcast(index, src, dim), src)
        cshape = [1] * out.dim()
        cshape[dim] = -1
        return out / cnt.view(cshape)
    if reduce in ('max', 'amax'):
        out = src.new_full(shape, float('-inf'))
        out = out.scatter_reduce_(dim, _broadcast(index, src, dim), src, 'amax', include_self=True)
        return out.masked_fill_(out == float('-inf'), 0)
    raise ValueError(f"shim scatter: unsupported reduce {reduce!r}")


def remove_self_loops(edge_index: Tensor, edge_attr: Optional[Tensor] = None):
    keep = edge_index[0] != edge_index[1]
    ei = edge_index[:, keep]
    return (ei, None) if edge_attr is None else (ei, edge_attr[keep])


def add_self_loops(edge_index: Tensor, edge_attr: Optional[Tensor] = None,
                   fill_value=None, num_nodes: Optional[int] = None):
    n = maybe_num_nodes(edge_index, num_nodes)
    loops = torch.arange(n, dtype=edge_index.dtype, device=edge_index.device)
    loops = loops.unsqueeze(0).repeat(2, 1)
    if edge_attr is not None:
        fv = 1.0 if fill_value is None else fill_value
        if isinstance(fv, Tensor):
            extra = fv.to(edge_attr.dtype).expand(n, *edge_attr.shape[1:]).contiguous()
        else:
            extra = edge_attr.new_full((n,) + tuple(edge_attr.shape[1:]), fv)
        edge_attr = torch.cat([edge_attr, extra], dim=0)
    return torch.cat([edge_index, loops], dim=1), edge_attr


def add_remaining_self_loops(edge_index: Tensor, edge_attr: Optional[Tensor] = None,
                             fill_value=None, num_nodes: Optional[int] = None):
    n = maybe_num_nodes(edge_index, num_nodes)
    off = edge_index[0] != edge_index[1]
    loops = torch.arange(n, dtype=edge_index.dtype, device=edge_index.device)
    loops = loops.unsqueeze(0).repeat(2, 1)
    if edge_attr is not None:
        fv = 1.0 if fill_value is None else fill_value
        loop_attr = edge_attr.new_full((n,) + tuple(edge_attr.shape[1:]), fv)
        on = ~off
        # existing self-loop weights win over fill_value (last write wins, like index_put)
        loop_attr[edge_index[0][on]] = edge_attr[on]
        edge_attr = torch.cat([edge_attr[off], loop_attr], dim=0)
    return torch.cat([edge_index[:, off], loops], dim=1), edge_attr


def coalesce(edge_index: Tensor, edge_attr=None, num_nodes: Optional[int] = None,
             reduce: str = 'sum', is_sorted: bool = False, sort_by_row: bool = True):
    n = maybe_num_nodes(edge_index, num_nodes)
    e = edge_index.size(1)
    key = edge_index[0 if sort_by_row else 1] * n + edge_index[1 if sort_by_row else 0]
    if not is_sorted:
        key, perm = torch.sort(key, stable=True)
        edge_index = edge_index[:, perm]
        if isinstance(edge_attr, Tensor):
            edge_attr = edge_attr[perm]
    first = torch.ones(e, dtype=torch.bool, device=edge_index.device)
    if e > 1:
        first[1:] = key[1:] > key[:-1]
    if bool(first.all()):
        return (edge_index, edge_attr) if edge_attr is not None else edge_index
    out_index = edge_index[:, first]
    if edge_attr is None:
        return out_index
    seg = torch.cumsum(first.to(torch.long), 0) - 1
    red = 'sum' if reduce == 'add' else reduce
    return out_index, scatter(edge_attr, seg, 0, out_index.size(1), red)


def to_undirected(edge_index: Tensor, edge_attr: Optional[Tensor] = None, num_nodes: Optional[int] = None,
                  reduce: str = 'add'):
    """PyG to_undirected: both directions of every edge, coalesced (sorted, duplicates merged)."""
    n = maybe_num_nodes(edge_index, num_nodes)
    both = torch.cat([edge_index, edge_index.flip(0)], dim=1)
    if edge_attr is None:
        return coalesce(both, None, n)
    return coalesce(both, torch.cat([edge_attr, edge_attr], dim=0), n, reduce)


def to_scipy_sparse_matrix(edge_index: Tensor, edge_attr: Optional[Tensor] = None,
                           num_nodes: Optional[int] = None):
    import scipy.sparse as sp
    n = maybe_num_nodes(edge_index, num_nodes)
    r, c = edge_index.cpu().numpy()
    if edge_attr is None:
        import numpy as np
        v = np.ones(r.shape[0])
    else:
        v = edge_attr.detach().cpu().numpy()
    return sp.coo_matrix((v, (r, c)), (n, n))


def softmax(src: Tensor, index: Tensor, ptr=None, num_nodes: Optional[int] = None,
            dim: int = 0) -> Tensor:
    n = maybe_num_nodes(index, num_nodes)
    mx = scatter(src.detach(), index, dim, n, 'max')
    ex = (src - mx.index_select(dim, index)).exp()
    den = scatter(ex, index, dim, n, 'sum') + 1e-16
    return ex / den.index_select(dim, index)


class SparseTensor:  # only referenced in isinstance checks / annotations on the COO path
    pass


def spmm(src, other: Tensor, reduce: str = 'sum') -> Tensor:
    raise NotImplementedError("shim: SparseTensor path not provided (COO path only)")


def gcn_norm(edge_index, edge_weight=None, num_nodes=None, improved=False,
             add_self_loops=True, flow="source_to_target", dtype=None):
    fill = 2.0 if improved else 1.0
    n = maybe_num_nodes(edge_index, num_nodes)
    if edge_weight is None:
        edge_weight = torch.ones((edge_index.size(1),), dtype=dtype, device=edge_index.device)
    if add_self_loops:
        edge_index, edge_weight = add_remaining_self_loops(edge_index, edge_weight, fill, n)
    row, col = edge_index[0], edge_index[1]
    idx = col if flow == 'source_to_target' else row
    deg = scatter(edge_weight, idx, 0, n, 'sum')
    dis = deg.pow_(-0.5)
    dis.masked_fill_(dis == float('inf'), 0)
    return edge_index, dis[row] * edge_weight * dis[col]


# --------------------------------------------------------------------------- nn

def glorot(t):
    if t is not None:
        import math
        a = math.sqrt(6.0 / (t.size(-2) + t.size(-1)))
        t.data.uniform_(-a, a)


def zeros(t):
    if t is not None:
        t.data.fill_(0)


class Linear(torch.nn.Linear):
    """y = x W^T + b with weight [out, in]; init differs from real PyG -- tests copy weights."""

    def __init__(self, in_channels: int, out_channels: int, bias: bool = True, **kw):
        super().__init__(in_channels, out_channels, bias=bias)
        self.in_channels, self.out_channels = in_channels, out_channels


class MessagePassing(torch.nn.Module):
    def __init__(self, aggr: Optional[str] = 'add', flow: str = 'source_to_target',
                 node_dim: int = -2, **kwargs):
        super().__init__()
        assert flow in ('source_to_target', 'target_to_source')
        self.aggr, self.flow, self.node_dim = aggr, flow, node_dim
        self._msg_params = [p for p in inspect.signature(self.message).parameters]
        self._upd_params = [p for p in inspect.signature(self.update).parameters][1:]

    # default hooks
    def message(self, x_j):
        return x_j

    def update(self, aggr_out):
        return aggr_out

    def propagate(self, edge_index, size=None, **kwargs):
        if not isinstance(edge_index, Tensor):
            raise NotImplementedError("shim: only COO edge_index tensors are supported")
        i, j = (1, 0) if self.flow == 'source_to_target' else (0, 1)
        sizes = [None, None] if size is None else list(size)
        margs = {}
        for name in self._msg_params:
            if name.endswith('_i') or name.endswith('_j'):
                side = j if name.endswith('_j') else i
                data = kwargs.get(name[:-2])
                if isinstance(data, (tuple, list)):
                    if isinstance(data[1 - side], Tensor) and sizes[1 - side] is None:
                        sizes[1 - side] = data[1 - side].size(self.node_dim)
                    data = data[side]
                if isinstance(data, Tensor):
                    if sizes[side] is None:
                        sizes[side] = data.size(self.node_dim)
                    data = data.index_select(self.node_dim, edge_index[side])
                margs[name] = data
            elif name == 'index':
                margs[name] = edge_index[i]
            elif name == 'ptr':
                margs[name] = None
            elif name == 'size_i':
                margs[name] = sizes[i] if sizes[i] is not None else sizes[j]
            elif name == 'size_j':
                margs[name] = sizes[j] if sizes[j] is not None else sizes[i]
            else:
                margs[name] = kwargs.get(name)
        if sizes[i] is None:
            sizes[i] = sizes[j]
        msg = self.message(**margs)
        red = 'sum' if self.aggr == 'add' else self.aggr
        out = scatter(msg, edge_index[i], self.node_dim, sizes[i], red)
        uargs = {k: kwargs.get(k) for k in self._upd_params}
        return self.update(out, **uargs)


class GATConv(MessagePassing):
    """PyG GATConv restated for heads >= 1, concat=True, no edge features, dropout 0:
    h = lin(x) [N, H, C]; e_ij = leaky_relu(<h_j, att_src> + <h_i, att_dst>, slope);
    alpha = softmax over the incoming edges of i (self-loops removed, then one added per node);
    out_i = sum_j alpha_ij h_j, heads concatenated, + bias.  Parameter names follow PyG >= 2.3
    (`lin`, `att_src`, `att_dst`, `bias`)."""

    def __init__(self, in_channels, out_channels, heads=1, concat=True, negative_slope=0.2, dropout=0.0,
                 add_self_loops=True, bias=True, **kwargs):
        kwargs.setdefault('aggr', 'add')
        super().__init__(node_dim=0, **kwargs)
        assert concat and dropout == 0.0
        self.in_channels, self.out_channels, self.heads = in_channels, out_channels, heads
        self.negative_slope, self.add_self_loops = negative_slope, add_self_loops
        self.lin = Linear(in_channels, heads * out_channels, bias=False)
        self.att_src = torch.nn.Parameter(torch.empty(1, heads, out_channels))
        self.att_dst = torch.nn.Parameter(torch.empty(1, heads, out_channels))
        self.bias = torch.nn.Parameter(torch.empty(heads * out_channels)) if bias else None
        self.reset_parameters()

    def reset_parameters(self):
        glorot(self.lin.weight)
        glorot(self.att_src)
        glorot(self.att_dst)
        zeros(self.bias)

    def forward(self, x, edge_index):
        H, C = self.heads, self.out_channels
        h = self.lin(x).view(-1, H, C)
        a_src = (h * self.att_src).sum(-1)
        a_dst = (h * self.att_dst).sum(-1)
        if self.add_self_loops:
            edge_index, _ = remove_self_loops(edge_index)
            edge_index, _ = add_self_loops(edge_index, num_nodes=x.size(0))
        j, i = edge_index[0], edge_index[1]
        e = torch.nn.functional.leaky_relu(a_src[j] + a_dst[i], self.negative_slope)
        alpha = softmax(e, i, num_nodes=x.size(0))
        out = scatter(h[j] * alpha.unsqueeze(-1), i, 0, x.size(0), 'sum').view(-1, H * C)
        return out if self.bias is None else out + self.bias


# --------------------------------------------------------------------------- install

def install(force: bool = False) -> bool:
    """Register the shim as `torch_geometric` unless a real PyG is importable.
    Returns True when the shim is in use."""
    if not force:
        try:
            import torch_geometric  # noqa: F401
            return getattr(sys.modules['torch_geometric'], '__pgsd_shim__', False)
        except Exception:
            pass

    def mod(name):
        m = types.ModuleType(name)
        m.__pgsd_shim__ = True
        sys.modules[name] = m
        return m

    tg = mod('torch_geometric')
    tg.__path__ = []
    utils = mod('torch_geometric.utils')
    utils.__path__ = []
    nn_ = mod('torch_geometric.nn')
    nn_.__path__ = []
    conv = mod('torch_geometric.nn.conv')
    conv.__path__ = []
    gcn_conv = mod('torch_geometric.nn.conv.gcn_conv')
    inits = mod('torch_geometric.nn.inits')
    dense = mod('torch_geometric.nn.dense')
    dense.__path__ = []
    linear = mod('torch_geometric.nn.dense.linear')
    typing_ = mod('torch_geometric.typing')
    num_nodes = mod('torch_geometric.utils.num_nodes')
    tg.utils, tg.nn, tg.typing = utils, nn_, typing_
    nn_.conv, nn_.inits, nn_.dense = conv, inits, dense
    dense.linear = linear
    conv.gcn_conv = gcn_conv
    utils.num_nodes = num_nodes

    for f in (scatter, coalesce, remove_self_loops, add_self_loops, add_remaining_self_loops,
              to_scipy_sparse_matrix, softmax, spmm, to_undirected):
        setattr(utils, f.__name__, f)
    num_nodes.maybe_num_nodes = maybe_num_nodes
    conv.MessagePassing = MessagePassing
    nn_.MessagePassing = MessagePassing
    conv.GATConv = GATConv
    nn_.GATConv = GATConv
    gcn_conv.gcn_norm = gcn_norm
    inits.glorot, inits.zeros = glorot, zeros
    linear.Linear = Linear
    dense.Linear = Linear
    typing_.OptTensor = Optional[Tensor]
    typing_.PairTensor = Tuple[Tensor, Tensor]
    typing_.OptPairTensor = Tuple[Tensor, Optional[Tensor]]
    typing_.Adj = Union[Tensor, SparseTensor]
    typing_.SparseTensor = SparseTensor
    typing_.Size = Optional[Tuple[int, int]]
    return True

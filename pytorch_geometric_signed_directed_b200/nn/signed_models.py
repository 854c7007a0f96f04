"""SDGNN and SiGAT on the B200 layers.

Reference: nn/signed/SDGNN.py:66-260 and nn/signed/SiGAT.py:11-205 -- same constructors, module / parameter names
(`SDRLayer_{i}.agg_{k}.*`, `SDRLayer_{i}.mlp_layer.{0,2}.*`; `agg_{k}.*`, `mlp_layer.{0,2}.*`), forward() without
arguments.  The motif mining of `build_adj_lists` (Python set loops over every edge upstream) runs on the GPU
(`utils/signed.py` -> `pgsd_signed_triangle_counts`); the attention layers are `nn.GATConv` (edge-softmax kernel +
aggregation kernel); with gradients enabled the layers take their autograd path.  `init_emb` must be given (the TSVD
initialiser is CPU preprocessing outside the hot path) and the loss modules stay with the reference.
"""
from __future__ import annotations

from typing import List, Optional

import torch
from torch import Tensor

from .. import autograd as ag, ops, plan as _plan
from ..utils import signed as _signed
from .sdr_layer import GATConv, SDRLayer, gat_transforms


def _split_signed(edge_index_s: Tensor):
    return (edge_index_s[edge_index_s[:, 2] > 0][:, :2].t(), edge_index_s[edge_index_s[:, 2] < 0][:, :2].t())


def _need_init(init_emb, name):
    if init_emb is None:
        raise NotImplementedError(
            f"{name}: pass init_emb; the TSVD initialisation (create_spectral_features) is CPU preprocessing "
            "outside the B200 hot path -- compute it once with the reference's utility")


class SDGNN(torch.nn.Module):
    def __init__(self, node_num: int, edge_index_s: Tensor, in_dim: int = 20, out_dim: int = 20, layer_num: int = 2,
                 init_emb: Optional[Tensor] = None, init_emb_grad: bool = True, lamb_d: float = 5.0,
                 lamb_t: float = 1.0, **kwargs):
        super().__init__(**kwargs)
        self.node_num, self.in_dim, self.out_dim, self.layer_num = node_num, in_dim, out_dim, layer_num
        self.device, self.lamb_d, self.lamb_t = edge_index_s.device, lamb_d, lamb_t
        self.pos_edge_index, self.neg_edge_index = _split_signed(edge_index_s)
        _need_init(init_emb, "SDGNN")
        self.x = torch.nn.Parameter(init_emb, requires_grad=init_emb_grad)
        self.edge_lists, self.tri_weight_coo = _signed.sdgnn_motifs(edge_index_s, node_num)
        self.layers: List[SDRLayer] = []
        for i in range(layer_num):
            layer = SDRLayer(in_dim if i == 0 else out_dim, out_dim, edge_lists=self.edge_lists)
            self.add_module(f'SDRLayer_{i}', layer)
            self.layers.append(layer)
        self.reset_parameters()

    @property
    def tri_weight(self):
        """The reference's scipy csc matrix of triangle weights (SDGNN.py:248-249), materialised on demand."""
        import scipy.sparse as sp
        r, c, v = (t.cpu().numpy() for t in self.tri_weight_coo)
        return sp.csc_matrix((v, (r, c)), shape=(self.node_num, self.node_num))

    def reset_parameters(self):
        for layer in self.layers:
            layer.reset_parameters()

    def loss(self):
        raise NotImplementedError("SDGNN.loss: the sign / direction / triangle losses live in the reference's "
                                  "utils.signed; apply them to SDGNN.forward()'s embeddings")

    def forward(self) -> Tensor:
        x = self.x
        for layer in self.layers:
            x = layer(x)
        return x


class SiGAT(torch.nn.Module):
    def __init__(self, node_num: int, edge_index_s: Tensor, in_dim: int = 20, out_dim: int = 20,
                 init_emb: Optional[Tensor] = None, init_emb_grad: bool = True, **kwargs):
        super().__init__(**kwargs)
        self.in_dim, self.out_dim, self.node_num = in_dim, out_dim, node_num
        self.device = edge_index_s.device
        self.pos_edge_index, self.neg_edge_index = _split_signed(edge_index_s)
        _need_init(init_emb, "SiGAT")
        self.x = torch.nn.Parameter(init_emb, requires_grad=init_emb_grad)
        self.edge_lists = _signed.sigat_motifs(edge_index_s, node_num)
        self.aggs: List[GATConv] = []
        for i in range(len(self.edge_lists)):
            self.aggs.append(GATConv(in_channels=in_dim, out_channels=out_dim))
            self.add_module('agg_{}'.format(i), self.aggs[-1])
        self.mlp_layer = torch.nn.Sequential(
            torch.nn.Linear(out_dim * (len(self.edge_lists) + 1), out_dim),
            torch.nn.Tanh(),
            torch.nn.Linear(out_dim, out_dim))
        self.reset_parameters()

    def reset_parameters(self):
        for agg in self.aggs:
            agg.reset_parameters()

        def init_weights(m):
            if isinstance(m, torch.nn.Linear):
                torch.nn.init.kaiming_normal_(m.weight)
                m.bias.data.fill_(0.01)
        self.mlp_layer.apply(init_weights)

    def loss(self):
        raise NotImplementedError("SiGAT.loss: Link_Sign_Product_Loss lives in the reference's utils.signed; "
                                  "apply it to SiGAT.forward()'s embeddings")

    def forward(self) -> Tensor:
        x = self.x
        _plan.require_cuda(x, "x")
        if torch.is_grad_enabled() and any(p.requires_grad for p in self.parameters()):
            # training: SiGAT.py:190-199 on the differentiable primitives
            feats = torch.cat([x] + [agg.forward_train(x, e) for e, agg in zip(self.edge_lists, self.aggs)], dim=1)
            l0, l2 = self.mlp_layer[0], self.mlp_layer[2]
            hid = torch.tanh(ag.dense([(feats, l0.weight.t(), 0)], l0.out_features, bias=l0.bias)[0])
            return ag.dense([(hid, l2.weight.t(), 0)], l2.out_features, bias=l2.bias)[0]
        with torch.no_grad():
            n, k, c = x.size(0), len(self.aggs), self.out_dim
            # cat([x0] + neigh_feats) (SiGAT.py:196-197) is the layout of one wide buffer: every GATConv writes
            # its column block in place, and the first Linear reads it as ONE term
            wide = torch.empty((n, x.size(1) + k * c), dtype=x.dtype, device=x.device)
            wide[:, :x.size(1)] = x
            hs, ss = gat_transforms(x, self.aggs)                # 38 transforms + 76 score vectors: two launches
            for i, (edges, agg) in enumerate(zip(self.edge_lists, self.aggs)):
                lo = x.size(1) + i * c
                sl = wide[:, lo:lo + c]
                if (sl.data_ptr() % 16) or (c * x.element_size()) % 16:
                    sl.copy_(agg.aggregate(hs[i], ss[i][0], ss[i][1], edges))   # widths the vector kernels cannot write in place
                else:
                    agg.aggregate(hs[i], ss[i][0], ss[i][1], edges, out=sl)
            l0, l2 = self.mlp_layer[0], self.mlp_layer[2]
            hid = ops.dense([(wide, l0.weight.t(), 0)], l0.out_features, bias=l0.bias, relu_mode=2)[0]
            return ops.dense([(hid, l2.weight.t(), 0)], l2.out_features, bias=l2.bias)[0]

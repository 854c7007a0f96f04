"""Model wrappers of BASELINE configs 3 and 4 on the B200 layers.

* `DiGCN_Inception_Block_node_classification` -- reference
  nn/directed/DiGCN_Inception_Block_node_classification.py:9-74: three inception blocks, `x0 + x1 + x2` after
  each, dropout, log_softmax.  Same constructor, parameter names (`ib{1,2,3}.{ln,conv1,conv2}.*`) and forward
  signature.  Whenever no dropout separates the three parts (eval mode, or dropout = 0) the sum is the
  epilogue of the two aggregation launches (`DiGCN_InceptionBlock.forward_sum`): x1 and x2 never exist.
* `SGCN` -- reference nn/signed/SGCN.py:11-97: conv1 (first_aggr) + (layer_num - 1) deep layers, tanh after each;
  forward() takes no arguments (graph and features are module state).  tanh is the epilogue of each layer's
  transform.  `init_emb` must be given: the TSVD initialisation (utils/signed/create_spectral_features.py,
  scikit-learn on the CPU) and the two losses are outside the hot path (SURVEY §8 "out of scope").
"""
from __future__ import annotations

from typing import Optional, Tuple

import torch
import torch.nn.functional as F
from torch import Tensor

from .digcn_conv import DiGCN_InceptionBlock
from .sgcn_conv import SGCNConv


class DiGCN_Inception_Block_node_classification(torch.nn.Module):
    def __init__(self, num_features: int, hidden: int, label_dim: int, dropout: float = 0.5):
        super().__init__()
        self.ib1 = DiGCN_InceptionBlock(num_features, hidden)
        self.ib2 = DiGCN_InceptionBlock(hidden, hidden)
        self.ib3 = DiGCN_InceptionBlock(hidden, label_dim)
        self._dropout = dropout
        self.reset_parameters()

    def reset_parameters(self):
        self.ib1.reset_parameters()
        self.ib2.reset_parameters()
        self.ib3.reset_parameters()

    def _block(self, ib, x, ei, ew, ei2, ew2):
        if not self.training or self._dropout == 0:
            return ib.forward_sum(x, ei, ew, ei2, ew2)
        x0, x1, x2 = ib(x, ei, ew, ei2, ew2)
        x0 = F.dropout(x0, p=self._dropout, training=True)
        x1 = F.dropout(x1, p=self._dropout, training=True)
        x2 = F.dropout(x2, p=self._dropout, training=True)
        return x0 + x1 + x2

    def forward(self, features: Tensor, edge_index_tuple: Tuple[Tensor, Tensor],
                edge_weight_tuple: Tuple[Tensor, Tensor]) -> Tensor:
        edge_index, edge_index2 = edge_index_tuple
        edge_weight, edge_weight2 = edge_weight_tuple
        x = self._block(self.ib1, features, edge_index, edge_weight, edge_index2, edge_weight2)
        x = F.dropout(x, p=self._dropout, training=self.training)
        x = self._block(self.ib2, x, edge_index, edge_weight, edge_index2, edge_weight2)
        x = F.dropout(x, p=self._dropout, training=self.training)
        x = self._block(self.ib3, x, edge_index, edge_weight, edge_index2, edge_weight2)
        return F.log_softmax(x, dim=1)


class _LinkSignEntropyParams(torch.nn.Module):
    """Parameter container with the names of the reference's `Link_Sign_Entropy_Loss` (utils/signed/link_sign_loss.py:
    `lin = Linear(2 * emb_dim, 3)`), so a reference SGCN state_dict (`lsp_loss.lin.{weight,bias}`) loads with
    strict=True.  The loss itself (PyG negative sampling on the CPU side of the training loop) is outside the hot
    path and stays with the reference."""

    def __init__(self, emb_dim: int):
        super().__init__()
        self.lin = torch.nn.Linear(2 * emb_dim, 3)


class SGCN(torch.nn.Module):
    def __init__(self, node_num: int, edge_index_s: Tensor, in_dim: int = 64, out_dim: int = 64,
                 layer_num: int = 2, init_emb: Optional[Tensor] = None, init_emb_grad: bool = False,
                 lamb: float = 5, norm_emb: bool = False, **kwargs):
        super().__init__(**kwargs)
        self.node_num, self.in_dim, self.out_dim, self.lamb = node_num, in_dim, out_dim, lamb
        self.device = edge_index_s.device
        # bit-exact integer plumbing of SGCN.py:53-54
        self.pos_edge_index = edge_index_s[edge_index_s[:, 2] > 0][:, :2].t()
        self.neg_edge_index = edge_index_s[edge_index_s[:, 2] < 0][:, :2].t()
        if init_emb is None:
            raise NotImplementedError(
                "SGCN: pass init_emb; the TSVD initialisation (create_spectral_features) is CPU preprocessing "
                "outside the B200 hot path -- compute it once with the reference's utility")
        self.x = torch.nn.Parameter(init_emb, requires_grad=init_emb_grad)
        self.conv1 = SGCNConv(in_dim, out_dim // 2, first_aggr=True)
        self.convs = torch.nn.ModuleList()
        for _ in range(layer_num - 1):
            self.convs.append(SGCNConv(out_dim // 2, out_dim // 2, first_aggr=False, norm_emb=norm_emb))
        for conv in [self.conv1, *self.convs]:
            conv.fused_tanh = True
        self.lsp_loss = _LinkSignEntropyParams(out_dim)       # state_dict compatibility (SGCN.py:75)
        self.reset_parameters()

    def reset_parameters(self):
        self.conv1.reset_parameters()
        for conv in self.convs:
            conv.reset_parameters()

    def loss(self):
        raise NotImplementedError(
            "SGCN.loss: Link_Sign_Entropy_Loss / Sign_Structure_Loss live in the reference's utils.signed; "
            "apply them to SGCN.forward()'s embeddings (they are not part of the aggregation hot path)")

    def forward(self) -> Tensor:
        z = self.conv1(self.x, self.pos_edge_index, self.neg_edge_index)       # tanh fused (fused_tanh)
        for conv in self.convs:
            z = conv(z, self.pos_edge_index, self.neg_edge_index)
        return z

"""MagNet_node_classification on the B200 layers (the model-wrapper epilogues that make
BASELINE config 2, "MagNet_node_classification 2-layer, 1M nodes / 20M edges / 64 hidden", end to end).

Reference: nn/directed/MagNet_node_classification.py:39-92 -- same constructor, parameter names
(`Chebs.{i}.weight/bias`, `Conv.weight [label_dim, 2*hidden, 1]`, `Conv.bias`) and forward signature.
What changes underneath:
  * the complex ReLU after every layer (:79-81) is the epilogue of that layer's dense transform;
  * `cat(real, imag) -> Conv1d(kernel_size=1)` (:83-90) is one `pgsd_dense_transform` with two column-block
    terms (no concatenation, no permutes);
  * all layers see the same graph, so the operator plan is built once per forward and shared
    (the reference rebuilds / caches it per layer).
"""
from __future__ import annotations

from typing import Optional

import torch
import torch.nn.functional as F
from torch import Tensor

from .. import autograd as ag
from .complex_relu import complex_relu_layer
from .magnet_conv import MagNetConv


class MagNet_node_classification(torch.nn.Module):
    def __init__(self, num_features: int, hidden: int = 2, q: float = 0.25, K: int = 1, label_dim: int = 2,
                 activation: bool = False, trainable_q: bool = False, layer: int = 2, dropout: float = False,
                 normalization: str = 'sym', cached: bool = False):
        super().__init__()
        chebs = torch.nn.ModuleList()
        chebs.append(MagNetConv(in_channels=num_features, out_channels=hidden, K=K, q=q, trainable_q=trainable_q,
                                normalization=normalization, cached=cached))
        self.normalization = normalization
        self.activation = activation
        if self.activation:
            self.complex_relu = complex_relu_layer()
        for _ in range(1, layer):
            chebs.append(MagNetConv(in_channels=hidden, out_channels=hidden, K=K, q=q, trainable_q=trainable_q,
                                    normalization=normalization, cached=cached))
        self.Chebs = chebs
        self.Conv = torch.nn.Conv1d(2 * hidden, label_dim, kernel_size=1)
        self.dropout = dropout

    def reset_parameters(self):
        for cheb in self.Chebs:
            cheb.reset_parameters()
        self.Conv.reset_parameters()

    def forward(self, real: Tensor, imag: Tensor, edge_index: Tensor,
                edge_weight: Optional[Tensor] = None) -> Tensor:
        shared = None
        for cheb in self.Chebs:
            cheb.fused_complex_relu = bool(self.activation)
            # one operator for the whole stack: a layer that would (re)build its plan takes the one
            # the previous layer just built from the very same inputs
            if shared is not None and not cheb.trainable_q and (not cheb.cached or cheb._plan is None) \
                    and (cheb._q_value(), cheb.normalization) == shared[1]:
                cheb._plan, cheb._cached_result = shared[0], None
                cheb.cached_num_edges, cheb.cached_q = edge_index.size(1), cheb.q
                real, imag = cheb._cheb_forward(real, imag)
            else:
                real, imag = cheb(real, imag, edge_index, edge_weight)
            if not cheb.trainable_q:
                shared = (cheb._plan, (cheb._q_value(), cheb.normalization))
        if self.dropout > 0:
            real = F.dropout(real, self.dropout, training=self.training)
            imag = F.dropout(imag, self.dropout, training=self.training)
        h = real.size(1)
        w = self.Conv.weight[:, :, 0].t()                      # [2*hidden, label_dim] view
        logits = ag.dense([(real, w[:h], 0), (imag, w[h:], 0)], self.Conv.out_channels, bias=self.Conv.bias)[0]
        return F.log_softmax(logits, dim=1)

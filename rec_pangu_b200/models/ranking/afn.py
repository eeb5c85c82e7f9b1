"""AFN (adaptive factorization network) — reference: rec_pangu/models/ranking/afn.py:13-104.

Outside the north-star kernel list: a thin composition.  The embedding rows come from the one gather
launch of this build, the two MLPs and the 2 -> 1 ensemble layer run on the hot-path GEMM kernels; the logarithmic
transformation layer in between (|e| clamped at 1e-5 -> log -> BatchNorm over fields -> learned exponents -> exp ->
BatchNorm over the logarithmic neurons) is a chain of element-wise / [F x L] contraction torch ops on the CUDA tensors."""
from typing import Dict

import torch
from torch import nn

from ... import ops
from ..base_model import BaseModel
from ..layers import EmbeddingLayer, MLP
from ..utils import get_feature_num


class AFN(BaseModel):
    def __init__(self, embedding_dim=32, dnn_hidden_units=[64, 64, 64], afn_hidden_units=[64, 64, 64], ensemble_dnn=True,
                 loss_fun='torch.nn.BCELoss()', logarithmic_neurons=5, enc_dict=None):
        super().__init__(enc_dict, embedding_dim)
        self.dnn_hidden_units = dnn_hidden_units
        self.afn_hidden_units = afn_hidden_units
        self.loss_fun = eval(loss_fun)
        self.enc_dict = enc_dict
        self.num_sparse, self.num_dense = get_feature_num(self.enc_dict)
        self.coefficient_W = nn.Linear(self.num_sparse, logarithmic_neurons, bias=False)
        self.dense_layer = MLP(input_dim=embedding_dim * logarithmic_neurons, output_dim=1, hidden_units=afn_hidden_units,
                               use_bias=True)
        self.log_batch_norm = nn.BatchNorm1d(self.num_sparse)
        self.exp_batch_norm = nn.BatchNorm1d(logarithmic_neurons)
        self.ensemble_dnn = ensemble_dnn
        if ensemble_dnn:
            self.embedding_layer2 = EmbeddingLayer(enc_dict=self.enc_dict, embedding_dim=self.embedding_dim)
            self.dnn = MLP(input_dim=embedding_dim * self.num_sparse, output_dim=1, hidden_units=dnn_hidden_units, use_bias=True)
            self.fc = nn.Linear(2, 1)
        self.reset_parameters()

    def logarithmic_net(self, e: torch.Tensor) -> torch.Tensor:
        """afn.py:93-104: every logarithmic neuron is a learned power-product of the fields, exp(sum_f w_lf * log|e_f|)."""
        log_e = self.log_batch_norm(torch.log(e.abs().clamp(min=1e-5)))                 # [B, F, D], BatchNorm1d over F channels
        powers = torch.einsum('lf,bfd->bld', self.coefficient_W.weight, log_e)          # [B, L, D]
        return self.exp_batch_norm(torch.exp(powers)).flatten(start_dim=1)               # [B, L*D]

    def forward(self, data, is_training=True):
        afn_out = self.dense_layer(self.logarithmic_net(self.embedding_layer(data)))
        if self.ensemble_dnn:
            x2, _, _ = self.embedding_layer2.feature_row(data, with_dense=False)
            dnn_out = self.dnn(x2, K=self.embedding_dim * self.num_sparse)
            logit = ops.linear(torch.cat([afn_out, dnn_out], dim=-1), self.fc.weight, self.fc.bias)
        else:
            logit = afn_out
        return self._finish(logit, data, is_training)

"""AOANet (architecture and operation adaptive network) — reference: rec_pangu/models/ranking/aoanet.py:13-116.

Outside the north-star kernel list: a thin composition.  Feature row and [B, F, D] view from the one
gather launch, the MLP (no output layer) and the final Linear on the hot-path GEMM kernels; the generalized interaction
layers are small contractions (D x D outer products of field pairs weighted by alpha, projected by h) written with einsum on
the CUDA tensors, algebraically re-associated so that the [B, S*F, D, D] outer-product tensor of the reference is never
materialised:  out[b,o,h] = sum_d W[o,h,d] * hvec[o,d] * sum_n alpha[n,o] * B0rep[b,n,h] * Birep[b,n,d]."""
from typing import Dict, List

import torch
from torch import nn

from ... import ops
from ..base_model import BaseModel
from ..layers import MLP
from ..utils import get_feature_num


class GeneralizedInteraction(nn.Module):
    """aoanet.py:95-116.  Parameters and their shapes are the reference's (W [O, D, D], alpha [S*F, O], h [O, D, 1])."""

    def __init__(self, input_subspaces, output_subspaces, num_fields, embedding_dim):
        super().__init__()
        self.input_subspaces, self.num_fields, self.embedding_dim = input_subspaces, num_fields, embedding_dim
        self.W = nn.Parameter(torch.eye(embedding_dim, embedding_dim).unsqueeze(0).repeat(output_subspaces, 1, 1))
        self.alpha = nn.Parameter(torch.ones(input_subspaces * num_fields, output_subspaces))
        self.h = nn.Parameter(torch.ones(output_subspaces, embedding_dim, 1))

    def forward(self, B_0, B_i):
        S, F, D = self.input_subspaces, self.num_fields, self.embedding_dim
        # pair index n = s*F + f of the reference: B_0.repeat(1, S, 1)[n] = B_0[f];  B_i.repeat(1, 1, F).view(B, -1, D)[n] = B_i[n // F]
        left = B_0.repeat(1, S, 1)                                             # [B, S*F, D]
        right = B_i.repeat_interleave(F, dim=1)                                # [B, S*F, D]
        fused = torch.einsum('bnh,bnd,no->bohd', left, right, self.alpha)      # sum over pairs first: [B, O, D, D]
        return torch.einsum('bohd,ohd,od->boh', fused, self.W, self.h.squeeze(-1))


class GeneralizedInteractionNet(nn.Module):
    def __init__(self, num_layers, num_subspaces, num_fields, embedding_dim):
        super().__init__()
        self.layers = nn.ModuleList([GeneralizedInteraction(num_fields if i == 0 else num_subspaces, num_subspaces, num_fields,
                                                            embedding_dim) for i in range(num_layers)])

    def forward(self, B_0):
        B_i = B_0
        for layer in self.layers:
            B_i = layer(B_0, B_i)
        return B_i


class AOANet(BaseModel):
    def __init__(self, embedding_dim: int = 32, dnn_hidden_units: List[int] = [64, 64, 64], num_interaction_layers: int = 3,
                 num_subspaces: int = 4, loss_fun: str = 'torch.nn.BCELoss()', enc_dict: Dict[str, dict] = None):
        super().__init__(enc_dict, embedding_dim)
        self.dnn_hidden_units = dnn_hidden_units
        self.loss_fun = eval(loss_fun)
        self.enc_dict = enc_dict
        self.num_sparse, self.num_dense = get_feature_num(self.enc_dict)
        self.dnn_input_dim = self.embedding_dim * self.num_sparse + self.num_dense
        self.dnn = MLP(input_dim=self.dnn_input_dim, output_dim=None, hidden_units=self.dnn_hidden_units)
        self.gin = GeneralizedInteractionNet(num_interaction_layers, num_subspaces, self.num_sparse, self.embedding_dim)
        self.fc = nn.Linear(dnn_hidden_units[-1] + num_subspaces * self.embedding_dim, 1)
        self.reset_parameters()

    def forward(self, data, is_training=True):
        x, _, _ = self.embedding_layer.feature_row(data, with_dense=True)      # [emb.flatten | dense | 0-pad]
        F, D = self.num_sparse, self.embedding_dim
        emb = x[:, :F * D].view(x.shape[0], F, D)
        dnn_out = self.dnn(x, K=self.dnn_input_dim)
        interact_out = self.gin(emb).flatten(start_dim=1)
        logit = ops.linear(torch.cat([dnn_out, interact_out], dim=-1), self.fc.weight, self.fc.bias)
        return self._finish(logit, data, is_training)

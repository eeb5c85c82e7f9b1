"""LR_Layer (wide part) — reference: rec_pangu/models/layers/shallow.py:14-27.  The D=1 tables ride in the same
gather launch as the main tables when the model passes ``lr_in``; standalone it runs its own gather."""
import torch
from torch import nn

from ... import ops
from ..utils import get_dnn_input_dim
from .embedding import EmbeddingLayer


class LR_Layer(nn.Module):
    def __init__(self, enc_dict):
        super().__init__()
        self.enc_dict = enc_dict
        self.emb_layer = EmbeddingLayer(enc_dict=self.enc_dict, embedding_dim=1)
        self.dnn_input_dim = get_dnn_input_dim(self.enc_dict, 1)
        self.fc = nn.Linear(self.dnn_input_dim, 1)

    def tables(self):
        return self.emb_layer.tables()

    def forward(self, data, lr_in: torch.Tensor = None):
        """lr_in: [B, >=F+Nd] = [lr_table_f[idx_f] | dense] already produced by the fused gather, or None."""
        if lr_in is None:
            lr_in, _, _ = self.emb_layer.feature_row(data, with_dense=True)
        return ops.linear(lr_in, self.fc.weight, self.fc.bias, K=self.dnn_input_dim)

"""AutoInt — reference: rec_pangu/models/ranking/autoint.py:14-88."""
import torch  # noqa: F401  (loss_fun strings such as "torch.nn.BCELoss()" are eval-ed here, as in the reference)
from typing import Dict, List

from torch import nn

from ... import ops
from ..base_model import BaseModel
from ..layers import MLP, LR_Layer, MultiHeadSelfAttention
from ..utils import get_feature_num


class AutoInt(BaseModel):
    def __init__(self, embedding_dim: int = 32, dnn_hidden_units: List[int] = [64, 64, 64], attention_layers: int = 1,
                 num_heads: int = 1, attention_dim: int = 8, loss_fun: str = 'torch.nn.BCELoss()',
                 enc_dict: Dict[str, dict] = None):
        super().__init__(enc_dict, embedding_dim)
        self.dnn_hidden_units = dnn_hidden_units
        self.loss_fun = eval(loss_fun)
        self.enc_dict = enc_dict
        self.num_sparse, self.num_dense = get_feature_num(self.enc_dict)
        self.lr_layer = LR_Layer(enc_dict=enc_dict)
        self.dnn_input_dim = self.embedding_dim * self.num_sparse + self.num_dense
        self.dnn = MLP(input_dim=self.dnn_input_dim, output_dim=1, hidden_units=self.dnn_hidden_units)
        self.self_attention = nn.Sequential(
            *[MultiHeadSelfAttention(self.embedding_dim if i == 0 else num_heads * attention_dim,
                                     attention_dim=attention_dim, num_heads=num_heads, align_to="output")
              for i in range(attention_layers)])
        self.fc = nn.Linear(self.num_sparse * attention_dim * num_heads, 1)
        self.reset_parameters()

    def forward(self, data, is_training=True):
        x, _, lr_in = self.embedding_layer.feature_row(data, with_dense=True, lr_tables=self.lr_layer.tables())
        F, D = self.num_sparse, self.embedding_dim
        emb = x[:, :F * D].view(x.shape[0], F, D)
        att = self.self_attention(emb)                                   # [B, F, H*d]
        logit = ops.linear(att.flatten(start_dim=1), self.fc.weight, self.fc.bias)
        logit = logit + self.dnn(x, K=self.dnn_input_dim) + self.lr_layer(data, lr_in)
        return self._finish(logit, data, is_training)

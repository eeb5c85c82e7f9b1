"""DCN — reference: rec_pangu/models/ranking/dcn.py:14-68 (cross network over [emb | dense], then fc; the
reference builds no deep tower despite the hidden_units kwarg — SURVEY.md App. A-7)."""
import torch  # noqa: F401  (loss_fun strings such as "torch.nn.BCELoss()" are eval-ed here, as in the reference)
from typing import Dict, List

from torch import nn

from ... import ops
from ..base_model import BaseModel
from ..layers import CrossNet
from ..utils import get_feature_num


class DCN(BaseModel):
    def __init__(self, embedding_dim: int = 32, hidden_units: List[int] = [64, 64, 64], crossing_layers: int = 3,
                 loss_fun: str = 'torch.nn.BCELoss()', enc_dict: Dict[str, dict] = None):
        super().__init__(enc_dict, embedding_dim)
        self.dnn_hidden_units = hidden_units
        self.loss_fun = eval(loss_fun)
        self.enc_dict = enc_dict
        self.num_sparse, self.num_dense = get_feature_num(self.enc_dict)
        self.input_dim = self.num_sparse * self.embedding_dim + self.num_dense
        self.crossnet = CrossNet(self.input_dim, crossing_layers)
        self.fc = nn.Linear(self.input_dim, 1)
        self.reset_parameters()

    def forward(self, data, is_training=True):
        x, _, _ = self.embedding_layer.feature_row(data, with_dense=True)
        cross_out = self.crossnet(x, K=self.input_dim)
        logit = ops.linear(cross_out, self.fc.weight, self.fc.bias, K=self.input_dim)
        return self._finish(logit, data, is_training)

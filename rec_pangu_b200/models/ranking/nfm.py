"""NFM — reference: rec_pangu/models/ranking/nfm.py (MLP over the bi-interaction pooled [B,D] vector + LR)."""
import torch  # noqa: F401  (loss_fun strings such as "torch.nn.BCELoss()" are eval-ed here, as in the reference)
from typing import Dict, List

from ..base_model import BaseModel
from ..layers import LR_Layer, MLP, InnerProductLayer
from ..utils import get_dnn_input_dim


class NFM(BaseModel):
    def __init__(self, embedding_dim: int = 32, hidden_units: List[int] = [64, 64, 64],
                 loss_fun: str = 'torch.nn.BCELoss()', enc_dict: Dict[str, dict] = None):
        super().__init__(enc_dict, embedding_dim)
        self.hidden_units = hidden_units
        self.loss_fun = eval(loss_fun)
        self.enc_dict = enc_dict
        self.lr = LR_Layer(enc_dict=self.enc_dict)
        self.inner_product_layer = InnerProductLayer(output="Bi_interaction_pooling")
        self.dnn_input_dim = get_dnn_input_dim(self.enc_dict, self.embedding_dim)
        self.dnn = MLP(input_dim=self.embedding_dim, output_dim=1, hidden_units=self.hidden_units,
                       hidden_activations='relu', dropout_rates=0)
        self.reset_parameters()

    def forward(self, data, is_training: bool = True):
        x, _, lr_in = self.embedding_layer.feature_row(data, with_dense=True, lr_tables=self.lr.tables())
        F, D = len(self.embedding_layer.emb_feature), self.embedding_dim
        emb = x[:, :F * D].view(x.shape[0], F, D)
        bi = self.inner_product_layer(emb)                              # [B, D]
        logit = self.lr(data, lr_in) + self.dnn(bi)
        return self._finish(logit, data, is_training)

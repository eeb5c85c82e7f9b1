"""xDeepFM — reference: rec_pangu/models/ranking/xdeepfm.py:13-79 (LR + CIN + MLP with the default Dropout(0.1))."""
import torch  # noqa: F401  (loss_fun strings such as "torch.nn.BCELoss()" are eval-ed here, as in the reference)
from typing import Dict, List

from ..base_model import BaseModel
from ..layers import MLP, LR_Layer, CompressedInteractionNet
from ..utils import get_feature_num


class xDeepFM(BaseModel):
    def __init__(self, embedding_dim: int = 32, dnn_hidden_units: List[int] = [64, 64, 64],
                 cin_layer_units: List[int] = [16, 16, 16], loss_fun: str = 'torch.nn.BCELoss()',
                 enc_dict: Dict[str, dict] = None) -> None:
        super().__init__(enc_dict, embedding_dim)
        self.embedding_dim = embedding_dim
        self.dnn_hidden_units = dnn_hidden_units
        self.loss_fun = eval(loss_fun)
        self.enc_dict = enc_dict
        self.num_sparse, self.num_dense = get_feature_num(self.enc_dict)
        self.dnn_input_dim = self.num_sparse * self.embedding_dim + self.num_dense
        self.dnn = MLP(input_dim=self.dnn_input_dim, output_dim=1, hidden_units=self.dnn_hidden_units)
        self.lr_layer = LR_Layer(enc_dict=self.enc_dict)
        self.cin = CompressedInteractionNet(self.num_sparse, cin_layer_units, output_dim=1)
        self.reset_parameters()

    def forward(self, data, is_training: bool = True):
        x, _, lr_in = self.embedding_layer.feature_row(data, with_dense=True, lr_tables=self.lr_layer.tables())
        F, D = self.num_sparse, self.embedding_dim
        emb = x[:, :F * D].view(x.shape[0], F, D)
        logit = self.lr_layer(data, lr_in) + self.cin(emb) + self.dnn(x, K=self.dnn_input_dim)
        return self._finish(logit, data, is_training)

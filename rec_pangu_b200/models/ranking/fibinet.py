"""FiBiNet — reference: rec_pangu/models/ranking/fibinet.py:13-77 (one bilinear layer shared by the raw and
the SENET-reweighted embeddings)."""
from typing import Dict, List

import torch

from ..base_model import BaseModel
from ..layers import LR_Layer, MLP, BilinearInteractionLayer, SENET_Layer
from ..utils import get_feature_num


class FiBiNet(BaseModel):
    def __init__(self, embedding_dim: int = 32, hidden_units: List[int] = [64, 64, 64],
                 loss_fun: str = 'torch.nn.BCELoss()', enc_dict: Dict[str, dict] = None):
        super().__init__(enc_dict, embedding_dim)
        self.hidden_units = hidden_units
        self.loss_fun = eval(loss_fun)
        self.enc_dict = enc_dict
        self.num_sparse, self.num_dense = get_feature_num(self.enc_dict)
        self.lr = LR_Layer(enc_dict=self.enc_dict)
        self.senet_layer = SENET_Layer(self.num_sparse, 3)
        self.bilinear_interaction = BilinearInteractionLayer(self.num_sparse, embedding_dim, 'field_interaction')
        input_dim = self.num_sparse * (self.num_sparse - 1) * self.embedding_dim + self.num_dense
        self.dnn = MLP(input_dim=input_dim, output_dim=1, hidden_units=self.hidden_units,
                       hidden_activations='relu', dropout_rates=0)
        self.reset_parameters()

    def forward(self, data, is_training: bool = True):
        from ... import ops
        x, _, lr_in = self.embedding_layer.feature_row(data, with_dense=True, lr_tables=self.lr.tables())
        F, D = self.num_sparse, self.embedding_dim
        # fused SENET + both bilinear passes -> [B, F(F-1)*D + Nd (+pad)] MLP input, dense columns appended in-kernel
        comb = ops.fibinet_interaction(x, F, D, self.num_dense, self.senet_layer.excitation[0].weight,
                                       self.senet_layer.excitation[2].weight,
                                       self.bilinear_interaction.stacked_weight())
        logit = self.lr(data, lr_in) + self.dnn(comb, K=F * (F - 1) * D + self.num_dense)
        return self._finish(logit, data, is_training)

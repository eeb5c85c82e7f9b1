"""ESSM — reference: rec_pangu/models/multi_task/essm.py:12-75.  Two MLP towers (CTR, CVR) over the flattened embeddings
(no dense features: essm.py:26,48), click = sigmoid(ctr), conversion = sigmoid(cvr), and the entire-space loss
BCE(click * conversion, task2_label) + 0.5 * BCE(click, task1_label) (essm.py:52-56,69-75) from one head kernel."""
from torch import nn

from ... import ops
from ..base_model import BaseModel
from ..layers import MLP
from ..utils import get_feature_num


class ESSM(BaseModel):
    def __init__(self, embedding_dim=40, hidden_dim=[128, 64], dropouts=[0.2, 0.2], enc_dict=None, device=None):
        super().__init__(enc_dict, embedding_dim)
        self.enc_dict = enc_dict
        self.hidden_dim = hidden_dim
        self.dropouts = dropouts
        self.num_sparse_fea, self.num_dense_fea = get_feature_num(self.enc_dict)
        hidden_size = self.num_sparse_fea * self.embedding_dim
        self.hidden_size = hidden_size
        self.ctr_layer = MLP(input_dim=hidden_size, output_dim=1, hidden_units=self.hidden_dim,
                             hidden_activations='relu', dropout_rates=self.dropouts)
        self.cvr_layer = MLP(input_dim=hidden_size, output_dim=1, hidden_units=self.hidden_dim,
                             hidden_activations='relu', dropout_rates=self.dropouts)
        self.sigmoid = nn.Sigmoid()
        self.apply(self._init_weights)

    def forward(self, data, is_training=True):
        x, _, _ = self.embedding_layer.feature_row(data, with_dense=False)       # [B, ld]; first F*D columns = hidden
        z_ctr = self.ctr_layer(x, K=self.hidden_size)
        z_cvr = self.cvr_layer(x, K=self.hidden_size)
        self._last_logit = [z_ctr.detach(), z_cvr.detach()]
        if is_training:
            click, conversion, loss = ops.essm_head(z_ctr, z_cvr, data['task1_label'], data['task2_label'], w_ctr=0.5)
            return {'task1_pred': click, 'task2_pred': conversion, 'loss': loss}
        click, conversion = ops.essm_head(z_ctr, z_cvr)
        return {'task1_pred': click, 'task2_pred': conversion}

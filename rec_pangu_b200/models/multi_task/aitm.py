"""AITM — reference: rec_pangu/models/multi_task/aitm.py:13-120 (adaptive information transfer multi-task).

Composition over the hot-path kernels: one gather launch for the [B, F*D] embedding row, both towers and every nn.Linear
on ops.mlp_forward / ops.linear (tcgen05 3xTF32), dropout through rpb_dropout_*; the attention over the TWO tokens
(conversion tower output, transferred click information) — MultiHeadSelfAttention(tower_dims[-1]) with its defaults: one
head, attention_dim = input dim, no scaling, residual without projection, ReLU (layers/attention.py:35-101) — is a 2 x 2
softmax per sample, written as element-wise torch ops on the CUDA tensors."""
from typing import Dict, List

import torch
from torch import nn

from ... import ops
from ..base_model import BaseModel
from ..layers import MLP, MultiHeadSelfAttention
from ..utils import get_feature_num


class AITM(BaseModel):
    def __init__(self, embedding_dim: int = 32, tower_dims: List[int] = [400, 400, 400], drop_prob: List[float] = [0.1, 0.1, 0.1],
                 enc_dict: Dict[str, dict] = None):
        super().__init__(enc_dict, embedding_dim)
        self.enc_dict = enc_dict
        self.tower_dims = tower_dims
        self.drop_prob = drop_prob
        self.num_sparse_fea, self.num_dense_fea = get_feature_num(self.enc_dict)
        self.tower_input_size = self.num_sparse_fea * self.embedding_dim
        self.click_tower = MLP(input_dim=self.tower_input_size, hidden_units=self.tower_dims, hidden_activations='relu',
                               dropout_rates=self.drop_prob)
        self.conversion_tower = MLP(input_dim=self.tower_input_size, hidden_units=self.tower_dims, hidden_activations='relu',
                                    dropout_rates=self.drop_prob)
        self.attention_layer = MultiHeadSelfAttention(self.tower_dims[-1])
        self.info_layer = nn.Sequential(nn.Linear(tower_dims[-1], tower_dims[-1]), nn.ReLU(), nn.Dropout(drop_prob[-1]))
        self.click_layer = nn.Sequential(nn.Linear(tower_dims[-1], 1), nn.Sigmoid())
        self.conversion_layer = nn.Sequential(nn.Linear(tower_dims[-1], 1), nn.Sigmoid())
        self.apply(self._init_weights)

    def _two_token_attention(self, t0: torch.Tensor, t1: torch.Tensor):
        """MultiHeadSelfAttention on the token pair (t0, t1), each [B, d]: out_i = relu(sum_j softmax_j(q_i . k_j) v_j + t_i)."""
        att = self.attention_layer
        if att.num_heads != 1 or att.W_res is not None:
            raise NotImplementedError('AITM attention: only the MultiHeadSelfAttention defaults of the reference are wired')
        q = [ops.linear(t, att.W_q.weight, None) for t in (t0, t1)]
        k = [ops.linear(t, att.W_k.weight, None) for t in (t0, t1)]
        v = [ops.linear(t, att.W_v.weight, None) for t in (t0, t1)]
        outs = []
        for i in range(2):
            s = torch.stack([(q[i] * k[0]).sum(dim=1), (q[i] * k[1]).sum(dim=1)], dim=1)      # [B, 2]
            a = torch.softmax(s, dim=1)
            o = a[:, 0:1] * v[0] + a[:, 1:2] * v[1] + (t0, t1)[i]
            outs.append(torch.relu(o))
        return outs

    def forward(self, data, is_training=True):
        x, _, _ = self.embedding_layer.feature_row(data, with_dense=False)           # [B, ldx], first F*D columns
        K = self.tower_input_size
        tower_click = self.click_tower(x, K=K)
        tower_conversion = self.conversion_tower(x, K=K)
        info = torch.relu(ops.linear(tower_click, self.info_layer[0].weight, self.info_layer[0].bias))
        info = ops.dropout(info, self.info_layer[2].p, self.training)
        a0, a1 = self._two_token_attention(tower_conversion, info)
        ait = a0 + a1
        click_logit = ops.linear(tower_click, self.click_layer[0].weight, self.click_layer[0].bias)
        conv_logit = ops.linear(ait, self.conversion_layer[0].weight, self.conversion_layer[0].bias)
        click, _ = ops.sigmoid_bce(click_logit, None)
        conversion, _ = ops.sigmoid_bce(conv_logit, None)
        click, conversion = click.reshape(-1), conversion.reshape(-1)              # the reference squeezes dim 1: [B]
        out = {'task1_pred': click, 'task2_pred': conversion}
        if is_training:
            out['loss'] = self.loss(data['task1_label'], click, data['task2_label'], conversion)
        return out

    def loss(self, click_label, click_pred, conversion_label, conversion_pred, constraint_weight=0.6):
        """aitm.py:99-120: BCE(click) + BCE(conversion) + 0.6 * sum(max(conversion - click, 0))."""
        click_loss = nn.functional.binary_cross_entropy(click_pred, click_label)
        conversion_loss = nn.functional.binary_cross_entropy(conversion_pred, conversion_label)
        label_constraint = torch.maximum(conversion_pred - click_pred, torch.zeros_like(click_label))
        return click_loss + conversion_loss + constraint_weight * torch.sum(label_constraint)

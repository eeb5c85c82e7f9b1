"""OMOE — reference: rec_pangu/models/multi_task/omoe.py:13-107.  One gate shared by all tasks, and that gate is a
PARAMETER-only softmax over the experts (omoe.py:41,79-80): gate_out = sum_l experts_out[:, :, l] * softmax(gate)[l].
Being batch-independent it folds into the weight: gate_out = hidden @ (experts . g) + experts_bias . g — the [hid, Hh, E]
-> [hid, Hh] contraction is parameter-side plumbing (autograd through torch), the per-sample work is one GEMM kernel."""
import torch

from ... import ops
from ..base_model import BaseModel
from ..utils import get_feature_num
from ._towers import build_towers, run_towers


class OMOE(BaseModel):
    def __init__(self, num_task=2, n_expert=3, embedding_dim=40, omoe_hidden_dim=128, expert_activation=None,
                 hidden_dim=[128, 64], dropouts=[0.2, 0.2], enc_dict=None, device=None):
        super().__init__(enc_dict, embedding_dim)
        self.enc_dict = enc_dict
        self.num_task = num_task
        self.n_expert = n_expert
        self.omoe_hidden_dim = omoe_hidden_dim
        if expert_activation is not None:
            raise NotImplementedError('expert_activation: only the reference default (None) is on the hot path')
        self.expert_activation = expert_activation
        self.hidden_dim = hidden_dim
        self.dropouts = dropouts
        self.num_sparse_fea, self.num_dense_fea = get_feature_num(self.enc_dict)
        hidden_size = self.num_sparse_fea * self.embedding_dim + self.num_dense_fea
        self.hidden_size = hidden_size
        self.experts = torch.nn.Parameter(torch.rand(hidden_size, omoe_hidden_dim, n_expert), requires_grad=True)
        self.experts.data.normal_(0, 1)
        self.experts_bias = torch.nn.Parameter(torch.rand(omoe_hidden_dim, n_expert), requires_grad=True)
        self.gate = torch.nn.Parameter(torch.rand(n_expert, 1), requires_grad=True)
        build_towers(self, num_task, omoe_hidden_dim, hidden_dim, dropouts)
        self.apply(self._init_weights)

    def forward(self, data, is_training=True):
        x, _, _ = self.embedding_layer.feature_row(data, with_dense=True)
        g = torch.softmax(self.gate, dim=0)                                   # [E, 1]
        w_eff = torch.matmul(self.experts, g).squeeze(-1).contiguous()        # [hid, Hh]
        b_eff = torch.matmul(self.experts_bias, g).squeeze(-1).contiguous()   # [Hh]
        gate_out = ops.matmul_kn(x, w_eff, b_eff, K=self.hidden_size)
        return run_towers(self, [gate_out] * self.num_task, data, is_training, eps=0.0)

"""MLMMOE — reference: rec_pangu/models/multi_task/mlmmoe.py:13-143.  MMOE with a second, parameter-only gating level:
level_out[:, :, e] = sum_l experts_out[:, :, l] * softmax(level_gates[e])[l] (mlmmoe.py:96-102).  That level is a fixed
E x E mixing of the experts, so it folds into effective expert weights (parameter-side einsum, autograd through torch);
the per-sample path is then exactly MMOE's: one GEMM over [experts_eff | gates] + the gate-softmax/combine kernel.
level_gates / gates / gates_bias are unregistered python lists as in the reference (mlmmoe.py:40-48,66-72)."""
import torch

from ..base_model import BaseModel
from ..utils import get_feature_num
from ._towers import build_towers, run_towers, moe_mix


class MLMMOE(BaseModel):
    def __init__(self, num_task=2, n_expert=3, embedding_dim=40, mmoe_hidden_dim=128, expert_activation=None,
                 hidden_dim=[128, 64], dropouts=[0.2, 0.2], enc_dict=None, device=None):
        super().__init__(enc_dict, embedding_dim)
        self.enc_dict = enc_dict
        self.num_task = num_task
        self.n_expert = n_expert
        self.mmoe_hidden_dim = mmoe_hidden_dim
        if expert_activation is not None:
            raise NotImplementedError('expert_activation: only the reference default (None) is on the hot path')
        self.expert_activation = expert_activation
        self.hidden_dim = hidden_dim
        self.dropouts = dropouts
        self.num_sparse_fea, self.num_dense_fea = get_feature_num(self.enc_dict)
        hidden_size = self.num_sparse_fea * self.embedding_dim + self.num_dense_fea
        self.hidden_size = hidden_size
        self.experts = torch.nn.Parameter(torch.rand(hidden_size, mmoe_hidden_dim, n_expert), requires_grad=True)
        self.experts.data.normal_(0, 1)
        self.experts_bias = torch.nn.Parameter(torch.rand(mmoe_hidden_dim, n_expert), requires_grad=True)
        self.level_gates = [torch.nn.Parameter(torch.rand(n_expert, 1), requires_grad=True) for _ in range(n_expert)]
        self.gates = [torch.nn.Parameter(torch.rand(hidden_size, n_expert), requires_grad=True) for _ in range(num_task)]
        for gate in self.gates:
            gate.data.normal_(0, 1)
        self.gates_bias = [torch.nn.Parameter(torch.rand(n_expert), requires_grad=True) for _ in range(num_task)]
        build_towers(self, num_task, mmoe_hidden_dim, hidden_dim, dropouts)
        self.set_device(device)
        self.apply(self._init_weights)

    def set_device(self, device):
        for i in range(self.num_task):
            self.gates[i] = self.gates[i].to(device)
            self.gates_bias[i] = self.gates_bias[i].to(device)
        for i in range(self.n_expert):
            self.level_gates[i] = self.level_gates[i].to(device)
        print(f'Successfully set device:{device}')

    def _apply(self, fn, recurse=True):
        # `.to(device)` / `.cuda()` also move the unregistered gate lists (the reference needs a manual set_device call)
        super()._apply(fn, recurse)
        with torch.no_grad():
            for name in ('gates', 'gates_bias', 'level_gates'):
                setattr(self, name, [torch.nn.Parameter(fn(g), requires_grad=g.requires_grad) for g in getattr(self, name)])
        return self

    def forward(self, data, is_training=True):
        x, _, _ = self.embedding_layer.feature_row(data, with_dense=True)
        g2 = torch.cat([torch.softmax(g, dim=0) for g in self.level_gates], dim=1)        # [E, E']: column e' = level gate e'
        experts_eff = torch.matmul(self.experts, g2)                                      # [hid, Hh, E']
        bias_eff = torch.matmul(self.experts_bias, g2)                                    # [Hh, E']
        outs = moe_mix(x, self.hidden_size, experts_eff, bias_eff, self.gates, self.gates_bias)
        return run_towers(self, [outs[i] for i in range(self.num_task)], data, is_training, eps=0.0)

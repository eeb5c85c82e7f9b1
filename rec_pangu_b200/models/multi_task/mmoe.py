"""MMOE — reference: rec_pangu/models/multi_task/mmoe.py:14-130.

Same constructor, parameters and state_dict keys (`experts`, `experts_bias`, `task_{t}_dnn.ctr_hidden_{j}` ...).  As in
the reference the gates are plain Python lists of Parameters — not registered, not optimised, not checkpointed, moved
by `set_device` (mmoe.py:39-47,64-68).  Forward: one gather launch -> one GEMM over the
concatenated [experts | gates] weight -> gate-softmax/combine kernel -> per-task towers (Linear -> BatchNorm1d ->
Dropout, no activation, App. A-10) -> sigmoid + BCE(pred + 1e-6) (mmoe.py:127-128)."""
import numpy as np
import torch
from torch import nn

from ... import ops
from ..base_model import BaseModel
from ..utils import get_feature_num


class MMOE(BaseModel):
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
        self.experts_bias = torch.nn.Parameter(torch.rand(mmoe_hidden_dim, n_expert), requires_grad=True)
        self.gates = [torch.nn.Parameter(torch.rand(hidden_size, n_expert), requires_grad=True) for _ in range(num_task)]
        for gate in self.gates:
            gate.data.normal_(0, 1)
        self.gates_bias = [torch.nn.Parameter(torch.rand(n_expert), requires_grad=True) for _ in range(num_task)]
        for i in range(self.num_task):
            setattr(self, 'task_{}_dnn'.format(i + 1), nn.ModuleList())
            hid_dim = [mmoe_hidden_dim] + hidden_dim
            tower = getattr(self, 'task_{}_dnn'.format(i + 1))
            for j in range(len(hid_dim) - 1):
                tower.add_module('ctr_hidden_{}'.format(j), nn.Linear(hid_dim[j], hid_dim[j + 1]))
                tower.add_module('ctr_batchnorm_{}'.format(j), nn.BatchNorm1d(hid_dim[j + 1]))
                tower.add_module('ctr_dropout_{}'.format(j), nn.Dropout(dropouts[j]))
            tower.add_module('task_last_layer', nn.Linear(hid_dim[-1], 1))
            tower.add_module('task_sigmoid', nn.Sigmoid())
        self.set_device(device)
        self.apply(self._init_weights)

    def set_device(self, device):
        for i in range(self.num_task):
            self.gates[i] = self.gates[i].to(device)
            self.gates_bias[i] = self.gates_bias[i].to(device)
        print(f'Successfully set device:{device}')

    def _apply(self, fn, recurse=True):
        # `.to(device)` / `.cuda()` also move the unregistered gates (the reference needs a manual set_device call)
        super()._apply(fn, recurse)
        with torch.no_grad():
            self.gates = [torch.nn.Parameter(fn(g), requires_grad=g.requires_grad) for g in self.gates]
            self.gates_bias = [torch.nn.Parameter(fn(g), requires_grad=g.requires_grad) for g in self.gates_bias]
        return self

    def forward(self, data, is_training=True):
        x, _, _ = self.embedding_layer.feature_row(data, with_dense=True)            # hidden = [emb | dense]
        Hh, E, T = self.mmoe_hidden_dim, self.n_expert, self.num_task
        w_cat = torch.cat([self.experts.view(self.hidden_size, Hh * E)] + list(self.gates), dim=1)     # [hid, Hh*E + T*E]
        b_cat = torch.cat([self.experts_bias.view(Hh * E)] + list(self.gates_bias), dim=0)
        eo = ops.matmul_kn(x, w_cat, b_cat, K=self.hidden_size)
        outs = ops.mmoe_combine(eo, Hh, E, T)                                          # [T, B, Hh]
        output_dict = dict()
        logits = []
        for i in range(T):
            h = outs[i]
            tower = getattr(self, 'task_{}_dnn'.format(i + 1))
            for mod in tower:
                if isinstance(mod, nn.Linear):
                    h = ops.linear(h, mod.weight, mod.bias)
                elif isinstance(mod, nn.BatchNorm1d):
                    h = ops.batch_norm(h, mod, self.training)
                elif isinstance(mod, nn.Dropout):
                    h = ops.dropout(h, mod.p, self.training)
                elif isinstance(mod, nn.Sigmoid):
                    pass                                                               # fused with the loss below
            logits.append(h)
        loss = 0
        for i in range(T):
            if is_training:
                # mmoe.py:120-128: weight 1/T, BCE(pred + 1e-6, label)
                pred, li = ops.sigmoid_bce(logits[i], data[f'task{i + 1}_label'], eps=1e-6, scale=1.0 / T)
                loss = loss + li
            else:
                pred, _ = ops.sigmoid_bce(logits[i], None)
            output_dict[f'task{i + 1}_pred'] = pred
        self._last_logit = [l.detach() for l in logits]
        if is_training:
            output_dict['loss'] = loss
        return output_dict

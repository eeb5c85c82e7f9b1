"""Per-task towers and the weighted BCE shared by the multi-task models (reference: the identical tower/loss blocks of
multi_task/mmoe.py:49-60,106-130, sharebottom.py:40-52,69-92, omoe.py:44-56,82-107, mlmmoe.py:52-64,118-143).

Tower = [Linear -> BatchNorm1d -> Dropout] x n -> Linear(->1) -> Sigmoid, NO activation between layers (SURVEY.md
App. A-10); loss = sum_t (1/T) * BCE(pred_t (+ eps), task{t}_label) with eps = 1e-6 only in MMOE (mmoe.py:127-128)."""
import torch
from torch import nn

from ... import ops


def build_towers(model: nn.Module, num_task: int, in_dim: int, hidden_dim, dropouts):
    """Same module names / order as the reference => same state_dict keys."""
    for i in range(num_task):
        setattr(model, 'task_{}_dnn'.format(i + 1), nn.ModuleList())
        hid_dim = [in_dim] + list(hidden_dim)
        tower = getattr(model, 'task_{}_dnn'.format(i + 1))
        for j in range(len(hid_dim) - 1):
            tower.add_module('ctr_hidden_{}'.format(j), nn.Linear(hid_dim[j], hid_dim[j + 1]))
            tower.add_module('ctr_batchnorm_{}'.format(j), nn.BatchNorm1d(hid_dim[j + 1]))
            tower.add_module('ctr_dropout_{}'.format(j), nn.Dropout(dropouts[j]))
        tower.add_module('task_last_layer', nn.Linear(hid_dim[-1], 1))
        tower.add_module('task_sigmoid', nn.Sigmoid())


def run_towers(model: nn.Module, tower_inputs, data, is_training: bool, eps: float = 0.0, K=None):
    """tower_inputs[t]: [B, >=in_dim] input of task t's tower (K = valid columns when the buffer is a padded feature row).
    Returns the reference's output dict ({'task{t}_pred', 'loss'})."""
    T = len(tower_inputs)
    logits = []
    for i in range(T):
        h, k = tower_inputs[i], K
        for mod in getattr(model, 'task_{}_dnn'.format(i + 1)):
            if isinstance(mod, nn.Linear):
                h = ops.linear(h, mod.weight, mod.bias, K=k)
                k = None
            elif isinstance(mod, nn.BatchNorm1d):
                h = ops.batch_norm(h, mod, model.training)
            elif isinstance(mod, nn.Dropout):
                h = ops.dropout(h, mod.p, model.training)
            # nn.Sigmoid: fused with the loss below
        logits.append(h)
    out = dict()
    loss = 0
    for i in range(T):
        if is_training:
            pred, li = ops.sigmoid_bce(logits[i], data[f'task{i + 1}_label'], eps=eps, scale=1.0 / T)
            loss = loss + li
        else:
            pred, _ = ops.sigmoid_bce(logits[i], None)
        out[f'task{i + 1}_pred'] = pred
    model._last_logit = [l.detach() for l in logits]
    if is_training:
        out['loss'] = loss
    return out


def moe_mix(x, hidden_size: int, experts, experts_bias, gates, gates_bias):
    """experts_out = einsum('ij,jkl->ikl', hidden, experts) + bias and, per task, softmax(hidden @ gate + b) mixing
    (mmoe.py:86-104) as ONE GEMM over the column-concatenated [experts | gates] weight + the gate-softmax/combine kernel.
    Returns [T, B, Hh]."""
    Hh, E = experts.shape[1], experts.shape[2]
    T = len(gates)
    w_cat = torch.cat([experts.reshape(hidden_size, Hh * E)] + list(gates), dim=1)           # [hid, Hh*E + T*E]
    b_cat = torch.cat([experts_bias.reshape(Hh * E)] + list(gates_bias), dim=0)
    eo = ops.matmul_kn(x, w_cat, b_cat, K=hidden_size)
    return ops.mmoe_combine(eo, Hh, E, T)

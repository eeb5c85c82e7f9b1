"""LR — reference: rec_pangu/models/ranking/lr.py:12-55.  The reference class is an nn.Module that calls a
`reset_parameters()` it does not have (lr.py:28 -> AttributeError at construction); this one
constructs — parameters keep torch's default initialisation, which is what the reference's LR_Layer holds at that line —
and its forward is the reference's: sigmoid(LR_Layer(data)), BCE on squeeze(-1)."""
from typing import Dict

import torch
from torch import nn

from ..layers import LR_Layer
from ..layers.embedding import EmbeddingLayer


class LR(nn.Module):
    def __init__(self, loss_fun: str = 'torch.nn.BCELoss()', enc_dict: Dict[str, dict] = None):
        super().__init__()
        self.loss_fun = eval(loss_fun)
        self.enc_dict = enc_dict
        self.lr_layer = LR_Layer(enc_dict=self.enc_dict)

    def set_grad_mode(self, mode: str):
        assert mode in ('dense', 'persistent')
        self.lr_layer.emb_layer.grad_mode = mode
        return self

    def zero_grad(self, set_to_none: bool = True):
        for m in self.modules():
            if isinstance(m, EmbeddingLayer):
                m.clean_grads()
        super().zero_grad(set_to_none=set_to_none)

    def forward(self, data: Dict[str, torch.Tensor], is_training: bool = True) -> Dict[str, torch.Tensor]:
        from ... import ops
        logit = self.lr_layer(data)
        self._last_logit = logit.detach()
        fused = isinstance(self.loss_fun, torch.nn.BCELoss) and self.loss_fun.reduction == 'mean' and self.loss_fun.weight is None
        if is_training and fused:
            pred, loss = ops.sigmoid_bce(logit, data['label'])
            return {'pred': pred, 'loss': loss}
        pred, _ = ops.sigmoid_bce(logit, None)
        if is_training:
            return {'pred': pred, 'loss': self.loss_fun(pred.squeeze(-1), data['label'])}
        return {'pred': pred}

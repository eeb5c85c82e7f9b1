"""BaseModel — reference: rec_pangu/models/base_model.py:14-90 (same init schemes and set_pretrained_weights)."""
import numpy as np
import torch
from torch import nn
from torch.nn.init import xavier_normal_, constant_

from .layers import EmbeddingLayer


class BaseModel(nn.Module):
    def __init__(self, enc_dict: dict, embedding_dim: int) -> None:
        super().__init__()
        self.enc_dict = enc_dict
        self.embedding_dim = embedding_dim
        self.embedding_layer = EmbeddingLayer(enc_dict=self.enc_dict, embedding_dim=self.embedding_dim)

    def _init_weights(self, module: nn.Module) -> None:
        """base_model.py:28-40 (xavier; used by the multi-task models)."""
        if isinstance(module, nn.Embedding):
            if module.weight.device.type != 'meta':        # dist.deferred_tables(): shards are initialised locally
                xavier_normal_(module.weight.data)
        elif isinstance(module, nn.Linear):
            xavier_normal_(module.weight.data)
            if module.bias is not None:
                constant_(module.bias.data, 0)

    def reset_parameters(self):
        """base_model.py:42-59: kaiming_normal_ on every >=2-D parameter (embedding tables included)."""
        for weight in self.parameters():
            if len(weight.shape) == 1 or weight.device.type == 'meta':      # meta: dist.deferred_tables()
                continue
            torch.nn.init.kaiming_normal_(weight)

    def set_pretrained_weights(self, col_name: str, pretrained_dict: dict, trainable: bool = True) -> None:
        """base_model.py:61-90."""
        assert col_name in self.enc_dict.keys(), "Pretrained Embedding Col: {} must be in the {}".format(
            col_name, self.enc_dict.keys())
        pretrained_emb_dim = len(list(pretrained_dict.values())[0])
        assert self.embedding_dim == pretrained_emb_dim, \
            "Pretrained Embedding Dim:{} must be equal to Model Embedding Dim:{}".format(pretrained_emb_dim, self.embedding_dim)
        # NB: like the reference the matrix has vocab_size rows (not +1): an OOV id then indexes past the table and
        # is reported as IndexError by the gather's bounds check.
        pretrained_emb = np.random.rand(self.enc_dict[col_name]['vocab_size'], pretrained_emb_dim)
        for k, v in self.enc_dict[col_name].items():
            if k == 'vocab_size':
                continue
            pretrained_emb[v, :] = pretrained_dict.get(k, np.random.rand(pretrained_emb_dim))
        embeddings = torch.from_numpy(pretrained_emb).float()
        old = self.embedding_layer.embedding_layer[col_name].weight
        self.embedding_layer.set_weights(col_name=col_name, embedding_matrix=embeddings.to(old.device),
                                         trainable=trainable)

    def set_grad_mode(self, mode: str):
        """'dense' (reference-identical fresh dense table grads) or 'persistent' (see EmbeddingLayer.grad_mode)."""
        assert mode in ('dense', 'persistent')
        for m in self.modules():
            if isinstance(m, EmbeddingLayer):
                m.grad_mode = mode
        return self

    def zero_grad(self, set_to_none: bool = True):
        for m in self.modules():
            if isinstance(m, EmbeddingLayer):
                m.clean_grads()
        super().zero_grad(set_to_none=set_to_none)

    # ---- shared head: sigmoid + loss (ranking models: `y_pred.sigmoid()` + `self.loss_fun(...)`)
    def _finish(self, logit, data, is_training):
        from .. import ops
        loss_fun = getattr(self, 'loss_fun', None)
        self._last_logit = logit.detach()      # parity hook: the reference API only returns post-sigmoid `pred`
        fused = isinstance(loss_fun, torch.nn.BCELoss) and loss_fun.reduction == 'mean' and loss_fun.weight is None
        if is_training and fused:
            pred, loss = ops.sigmoid_bce(logit, data['label'])
            return {'pred': pred, 'loss': loss}
        pred, _ = ops.sigmoid_bce(logit, None)
        if is_training:
            return {'pred': pred, 'loss': loss_fun(pred.squeeze(-1), data['label'])}
        return {'pred': pred}

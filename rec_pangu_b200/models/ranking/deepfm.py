"""DeepFM — reference: rec_pangu/models/ranking/deepfm.py:13-67."""
from typing import Dict, List

import torch

from ..base_model import BaseModel
from ..layers import FM_Layer, MLP
from ..utils import get_dnn_input_dim


class DeepFM(BaseModel):
    def __init__(self, embedding_dim: int = 32, hidden_units: List[int] = [64, 64, 64],
                 loss_fun: str = 'torch.nn.BCELoss()', enc_dict: Dict[str, dict] = None):
        super().__init__(enc_dict, embedding_dim)
        self.hidden_units = hidden_units
        self.loss_fun = eval(loss_fun)
        self.enc_dict = enc_dict
        self.fm = FM_Layer()
        self.dnn_input_dim = get_dnn_input_dim(self.enc_dict, self.embedding_dim)
        self.dnn = MLP(input_dim=self.dnn_input_dim, output_dim=1, hidden_units=self.hidden_units,
                       hidden_activations='relu', dropout_rates=0)
        self.reset_parameters()

    def forward(self, data, is_training=True):
        emb = self.embedding_layer
        from ... import ops
        if (emb._shards is None or ops.SHARDED_FUSED) and self.dnn._out_idx is not None:
            # whole body as one fused autograd node (ops.deepfm_core): gather+FM in one launch, MLP on tcgen05, the
            # logit sum folded into the last row-dot, and in backward the layer-1 dx GEMM scatters the embedding
            # gradients from its epilogue (no dx round trip through HBM)
            Ws, bs, relu, drops = self.dnn.layer_params()
            fused_loss = (is_training and 'label' in data and isinstance(self.loss_fun, torch.nn.BCELoss)
                          and self.loss_fun.reduction == 'mean' and self.loss_fun.weight is None)
            res = ops.deepfm_core(emb.tables(), [data[c] for c in emb.emb_feature], [data[c] for c in emb.dense_feature],
                                  Ws, bs, n_hidden=len(relu), relu=relu, dropout=drops, training=self.training,
                                  grad_store=emb._grad_store if emb.grad_mode == 'persistent' else None,
                                  label=data['label'] if fused_loss else None, shards=emb._shards)
            if isinstance(res, tuple):             # sigmoid + BCE came out of the kernel that finished the MLP
                logit, pred, loss = res
                self._last_logit = logit.detach()
                return {'pred': pred, 'loss': loss}
            if res is not None:
                return self._finish(res, data, is_training)
        # row-sharded tables (shapes the fused core does not take, or RPB_SHARDED_FUSED=0): one launch gathers 26 rows/sample over NVLink -> feature row x + FM second-order term
        x, fm_out, _ = emb.feature_row(data, with_dense=True, want_fm=True)
        dnn_output = self.dnn(x, K=self.dnn_input_dim)                  # [B,1]
        return self._finish(fm_out.unsqueeze(1) + dnn_output, data, is_training)

"""MaskNet — reference: rec_pangu/models/ranking/masknet.py:13-86, MaskBlock rec_pangu/models/layers/interaction.py:254-283.
This is the class examples/ranking/run_ranking_example.py:36 constructs.

A thin composition over the hot-path kernels: the feature row [emb | dense] comes from the one gather launch, every
nn.Linear of the MaskBlocks and of the MLP runs on the tcgen05 3xTF32 GEMM (ops.linear / ops.mlp_forward), LayerNorm is
rpb_layernorm_fwd/bwd; the instance-guided mask product and the mean over parallel blocks are element-wise torch ops on the
CUDA tensors in between."""
from typing import Dict, List

import torch
from torch import nn

from ... import ops
from ..base_model import BaseModel
from ..layers import MLP
from ..utils import get_dnn_input_dim


class MaskBlock(nn.Module):
    """interaction.py:254-283: out = LN_out(W_h (LN_in(net) * mask(mask_input))), mask = W_2 relu(W_1 mask_input)."""

    def __init__(self, input_dim: int, mask_input_dim: int, output_size: int, reduction_factor: float) -> None:
        super().__init__()
        self._input_layer_norm = nn.LayerNorm(input_dim)
        aggregation_size = int(mask_input_dim * reduction_factor)
        self._mask_layer = nn.Sequential(nn.Linear(mask_input_dim, aggregation_size), nn.ReLU(),
                                         nn.Linear(aggregation_size, input_dim))
        self._hidden_layer = nn.Linear(input_dim, output_size)
        self._layer_norm = nn.LayerNorm(output_size)
        self.input_dim, self.mask_input_dim = input_dim, mask_input_dim

    def forward(self, net: torch.Tensor, mask_input: torch.Tensor) -> torch.Tensor:
        net = ops.layer_norm(net, self._input_layer_norm, K=self.input_dim)
        a = torch.relu(ops.linear(mask_input, self._mask_layer[0].weight, self._mask_layer[0].bias, K=self.mask_input_dim))
        mask = ops.linear(a, self._mask_layer[2].weight, self._mask_layer[2].bias)
        hidden = ops.linear(net * mask, self._hidden_layer.weight, self._hidden_layer.bias)
        return ops.layer_norm(hidden, self._layer_norm)


class MaskNet(BaseModel):
    def __init__(self, embedding_dim: int = 32, block_num: int = 3, use_parallel: bool = True, reduction_factor: float = 0.3,
                 hidden_units: List[int] = [64, 64, 64], loss_fun: str = 'torch.nn.BCELoss()',
                 enc_dict: Dict[str, dict] = None):
        super().__init__(enc_dict, embedding_dim)
        self.loss_fun = eval(loss_fun)
        self.enc_dict = enc_dict
        self.block_num = block_num
        self.hidden_units = hidden_units
        self.reduction_factor = reduction_factor
        self.use_parallel = use_parallel
        self.block_output_dim = self.mask_input_dim = self.input_dim = get_dnn_input_dim(self.enc_dict, self.embedding_dim)
        self.mask_block_list = nn.ModuleList()
        for _ in range(self.block_num):
            self.mask_block_list.append(MaskBlock(self.input_dim, self.mask_input_dim, self.block_output_dim,
                                                  self.reduction_factor))
        self.mlp = MLP(self.block_output_dim, hidden_units=self.hidden_units, output_dim=1)
        self.reset_parameters()

    def forward(self, data, is_training=True):
        x, _, _ = self.embedding_layer.feature_row(data, with_dense=True)        # [B, ldx] = [emb.flatten | dense | 0-pad]
        if self.use_parallel:
            mask_output = None
            for layer in self.mask_block_list:
                o = layer(x, x)
                mask_output = o if mask_output is None else mask_output + o
            mask_output = mask_output * (1.0 / self.block_num)                   # torch.mean over the stacked block outputs
        else:
            mask_output = x
            for layer in self.mask_block_list:
                mask_output = layer(mask_output, x)
        return self._finish(self.mlp(mask_output, K=self.block_output_dim), data, is_training)

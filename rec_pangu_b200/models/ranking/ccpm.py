"""CCPM (convolutional click prediction model) — reference: rec_pangu/models/ranking/ccpm.py:13-117, KMaxPooling
rec_pangu/models/layers/sequence.py:63-88.

Outside the north-star kernel list: a thin composition.  The [B, F, D] embedding tensor is a view of
the one gather launch's feature row; the convolution stack over the field axis (zero-pad, (k x 1) Conv2d, k-max pooling
that keeps field order, Tanh) runs as torch CUDA ops with the reference's module layout (same state_dict keys:
`conv_layer.conv_layer.{1,5,9}.*`), the final Linear on the hot-path row-dot kernel."""
from typing import Dict, List

import torch
from torch import nn

from ... import ops
from ..base_model import BaseModel
from ..layers.activation import get_activation
from ..utils import get_feature_num


class KMaxPooling(nn.Module):
    """The k largest entries along `dim`, in their original order."""

    def __init__(self, k: int, dim: int):
        super().__init__()
        self.k, self.dim = k, dim

    def forward(self, X: torch.Tensor) -> torch.Tensor:
        keep = X.topk(self.k, dim=self.dim).indices.sort(dim=self.dim).values
        return X.gather(self.dim, keep)


class CCPM_ConvLayer(nn.Module):
    """Input [B, 1, F, D]; layer i: pad (k_i - 1) rows on both sides of the field axis, Conv2d((k_i, 1)), k-max pooling, Tanh."""

    def __init__(self, num_fields, channels=[3], kernel_heights=[3], activation="Tanh"):
        super().__init__()
        if not isinstance(kernel_heights, list):
            kernel_heights = [kernel_heights] * len(channels)
        elif len(kernel_heights) != len(channels):
            raise ValueError("channels={} and kernel_heights={} should have the same length.".format(channels, kernel_heights))
        self.channels = [1] + channels
        n_layers = len(kernel_heights)
        mods = []
        for i, (c_in, c_out, kh) in enumerate(zip(self.channels[:-1], self.channels[1:], kernel_heights), start=1):
            # pooling size shrinks with depth (ccpm.py:104-107), never below 3
            k = max(3, int((1 - pow(float(i) / n_layers, n_layers - i)) * num_fields)) if i < n_layers else 3
            mods += [nn.ZeroPad2d((0, 0, kh - 1, kh - 1)), nn.Conv2d(c_in, c_out, kernel_size=(kh, 1)), KMaxPooling(k, dim=2),
                     get_activation(activation)]
        self.conv_layer = nn.Sequential(*mods)

    def forward(self, X):
        return self.conv_layer(X)


class CCPM(BaseModel):
    def __init__(self, embedding_dim: int = 32, hidden_units: List[int] = [64, 64, 64], channels: List[int] = [4, 4, 2],
                 kernel_heights: List[int] = [6, 5, 3], loss_fun: str = 'torch.nn.BCELoss()', enc_dict: Dict[str, dict] = None):
        super().__init__(enc_dict, embedding_dim)
        self.dnn_hidden_units = hidden_units
        self.loss_fun = eval(loss_fun)
        self.enc_dict = enc_dict
        self.num_sparse, self.num_dense = get_feature_num(self.enc_dict)
        self.conv_layer = CCPM_ConvLayer(self.num_sparse, channels=channels, kernel_heights=kernel_heights)
        conv_out_dim = 3 * embedding_dim * channels[-1]          # 3 = k-max pooling size of the last layer
        self.fc = nn.Linear(conv_out_dim, 1)
        self.reset_parameters()

    def forward(self, data, is_training=True):
        emb = self.embedding_layer(data)                          # [B, F, D] view of the gathered feature row
        conv_out = self.conv_layer(emb.unsqueeze(1))              # [B, C, 3, D]
        logit = ops.linear(conv_out.flatten(start_dim=1).contiguous(), self.fc.weight, self.fc.bias)
        return self._finish(logit, data, is_training)

"""MLP tower — constructor/state_dict identical to the reference (rec_pangu/models/layers/deep.py:11-84);
forward runs the fused dense-layer kernels (tcgen05 3xTF32 or fp32 SIMT) through ops.mlp_forward."""
from typing import List, Optional, Union

import torch
import torch.nn as nn

from ... import ops
from .activation import get_activation


class MLP(nn.Module):
    def __init__(self, input_dim: int, output_dim: Union[int, None] = None, hidden_units: List[int] = [],
                 hidden_activations: Union[str, List[str]] = "ReLU", output_activation: Union[str, None] = None,
                 dropout_rates: Union[float, List[float]] = 0.1, batch_norm: bool = False, use_bias: bool = True):
        super().__init__()
        if output_dim is not None:
            assert isinstance(output_dim, int) and output_dim > 0, "output_dim must be an integer"
        assert isinstance(input_dim, int) and input_dim > 0, "input_dim must be an integer"
        assert isinstance(hidden_units, list) and all(isinstance(i, int) for i in hidden_units) and len(
            hidden_units) >= 1, "hidden_units must be a list of integers and with at least one element"
        if isinstance(hidden_activations, str):
            hidden_activations = [hidden_activations] * len(hidden_units)
        elif isinstance(hidden_activations, list):
            assert len(hidden_activations) == len(hidden_units), "hidden_activations must have one element per hidden unit"
        else:
            raise TypeError("hidden_activations must be a string or a list of strings")
        if not isinstance(dropout_rates, list):
            dropout_rates = [dropout_rates] * len(hidden_units)
        else:
            assert len(dropout_rates) == len(hidden_units), "dropout_rates must have one element per hidden unit"
        self.input_dim = input_dim
        dims = [input_dim] + hidden_units
        layers = []
        self._plan = []            # per hidden layer: (index of Linear in net, relu?, dropout p)
        for i in range(len(dims) - 1):
            lin_idx = len(layers)
            layers.append(nn.Linear(dims[i], dims[i + 1], bias=use_bias))
            if batch_norm:
                layers.append(nn.BatchNorm1d(dims[i + 1]))
            is_relu = False
            if hidden_activations[i]:
                act = get_activation(hidden_activations[i])
                is_relu = isinstance(act, nn.ReLU)
                if not is_relu:
                    raise NotImplementedError(f'MLP activation {hidden_activations[i]!r}: only ReLU is on the hot path')
                layers.append(act)
            if dropout_rates[i] > 0:
                layers.append(nn.Dropout(p=dropout_rates[i]))
            self._plan.append((lin_idx, is_relu, float(dropout_rates[i])))
        if batch_norm:
            raise NotImplementedError('MLP(batch_norm=True) is not used by the ranking hot path')
        self._out_idx = None
        if output_dim is not None:
            self._out_idx = len(layers)
            layers.append(nn.Linear(dims[-1], output_dim, bias=use_bias))
        if output_activation is not None:
            raise NotImplementedError('MLP output_activation is not used by the ranking hot path')
        self.net = nn.Sequential(*layers)      # same module indices as the reference => same state_dict keys

    def layer_params(self):
        """(weights, biases, relu flags, dropout rates) in layer order, for fused model-level kernels."""
        Ws = [self.net[i].weight for i, _, _ in self._plan]
        bs = [self.net[i].bias for i, _, _ in self._plan]
        if self._out_idx is not None:
            Ws.append(self.net[self._out_idx].weight)
            bs.append(self.net[self._out_idx].bias)
        return Ws, bs, [r for _, r, _ in self._plan], [p for _, _, p in self._plan]

    def forward(self, x: torch.Tensor, K: Optional[int] = None) -> torch.Tensor:
        """x: [B, >=input_dim]; only the first ``K`` (= input_dim) columns are read, so the padded feature row from
        EmbeddingLayer.feature_row can be passed without a cat/copy."""
        K = self.input_dim if K is None else K
        Ws = [self.net[i].weight for i, _, _ in self._plan]
        bs = [self.net[i].bias for i, _, _ in self._plan]
        if self._out_idx is not None:
            Ws.append(self.net[self._out_idx].weight)
            bs.append(self.net[self._out_idx].bias)
        return ops.mlp_forward(x, K, Ws, bs, n_hidden=len(self._plan), has_out=self._out_idx is not None,
                               relu=[r for _, r, _ in self._plan], dropout=[p for _, _, p in self._plan],
                               training=self.training)

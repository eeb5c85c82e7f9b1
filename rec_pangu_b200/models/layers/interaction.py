"""Feature-interaction layers — class names, constructors and state_dict keys follow the reference
(rec_pangu/models/layers/interaction.py); each forward is one fused sm_100a kernel (+ its backward)."""
from itertools import combinations

import torch
from torch import nn

from ... import ops


class InnerProductLayer(nn.Module):
    """interaction.py:12-52.  'product_sum_pooling' ([B,1]) and 'Bi_interaction_pooling' ([B,D]) are the outputs
    the ranking models use (FM/DeepFM/NFM); the pairwise outputs are not on the hot path."""

    def __init__(self, num_fields=None, output="product_sum_pooling"):
        super().__init__()
        self._output_type = output
        if output not in ["product_sum_pooling", "Bi_interaction_pooling", "inner_product", "elementwise_product"]:
            raise ValueError("InnerProductLayer output={} is not supported.".format(output))
        if output in ["inner_product", "elementwise_product"]:
            raise NotImplementedError(f'InnerProductLayer output={output} is outside the B200 hot path')

    def forward(self, feature_emb):
        return ops.fm_interaction(feature_emb, 'sum' if self._output_type == "product_sum_pooling" else 'bi')


class FM_Layer(nn.Module):
    """interaction.py:225-235."""

    def __init__(self, final_activation=None, use_bias=True):
        super().__init__()
        self.inner_product_layer = InnerProductLayer(output="product_sum_pooling")
        self.final_activation = final_activation

    def forward(self, feature_emb_list):
        output = self.inner_product_layer(feature_emb_list)
        if self.final_activation is not None:
            output = self.final_activation(output)
        return output


class CrossInteractionLayer(nn.Module):
    """interaction.py:119-127 — parameter container; the math runs in CrossNet's fused kernel."""

    def __init__(self, input_dim):
        super().__init__()
        self.weight = nn.Linear(input_dim, 1, bias=False)
        self.bias = nn.Parameter(torch.zeros(input_dim))


class CrossNet(nn.Module):
    """interaction.py:130-141: x_{l+1} = x_l + (w_l . x_l) x_0 + b_l, all layers in one kernel."""

    def __init__(self, input_dim, num_layers):
        super().__init__()
        self.num_layers = num_layers
        self.input_dim = input_dim
        self.cross_net = nn.ModuleList(CrossInteractionLayer(input_dim) for _ in range(self.num_layers))

    def forward(self, X_0, K=None):
        K = self.input_dim if K is None else K
        return ops.crossnet(X_0, K, [l.weight.weight for l in self.cross_net], [l.bias for l in self.cross_net])


class CompressedInteractionNet(nn.Module):
    """interaction.py:144-171 (xDeepFM CIN).  Conv1d modules are parameter containers (weight [U, Cin, 1])."""

    def __init__(self, num_fields, cin_layer_units, output_dim=1):
        super().__init__()
        self.cin_layer_units = cin_layer_units
        self.num_fields = num_fields
        self.fc = nn.Linear(sum(cin_layer_units), output_dim)
        self.cin_layer = nn.ModuleDict()
        for i, unit in enumerate(self.cin_layer_units):
            in_channels = num_fields * self.cin_layer_units[i - 1] if i > 0 else num_fields ** 2
            self.cin_layer["layer_" + str(i + 1)] = nn.Conv1d(in_channels, unit, kernel_size=1)

    def forward(self, feature_emb):
        Ws = [self.cin_layer["layer_" + str(i + 1)].weight for i in range(len(self.cin_layer_units))]
        bs = [self.cin_layer["layer_" + str(i + 1)].bias for i in range(len(self.cin_layer_units))]
        pooled = ops.cin(feature_emb, Ws, bs)                       # [B, sum(U)]
        return ops.linear(pooled, self.fc.weight, self.fc.bias)


class SENET_Layer(nn.Module):
    """interaction.py:238-251."""

    def __init__(self, num_fields, reduction_ratio=3):
        super().__init__()
        reduced_size = max(1, int(num_fields / reduction_ratio))
        self.excitation = nn.Sequential(nn.Linear(num_fields, reduced_size, bias=False), nn.ReLU(),
                                        nn.Linear(reduced_size, num_fields, bias=False), nn.ReLU())

    def forward(self, feature_emb):
        return ops.senet(feature_emb, self.excitation[0].weight, self.excitation[2].weight)


class BilinearInteractionLayer(nn.Module):
    """interaction.py:55-81, bilinear_type='field_interaction' (the FiBiNet configuration)."""

    def __init__(self, num_fields, embedding_dim, bilinear_type="field_interaction"):
        super().__init__()
        self.bilinear_type = bilinear_type
        if bilinear_type != "field_interaction":
            raise NotImplementedError(f'bilinear_type={bilinear_type} is not used by the hot path (FiBiNet)')
        self.num_fields = num_fields
        self.bilinear_layer = nn.ModuleList([nn.Linear(embedding_dim, embedding_dim, bias=False)
                                             for _ in combinations(range(num_fields), 2)])

    def stacked_weight(self):
        return torch.stack([l.weight for l in self.bilinear_layer], dim=0)     # [P, D, D]

    def forward(self, feature_emb):
        return ops.bilinear(feature_emb, self.stacked_weight())

"""AutoInt interacting layer — reference: rec_pangu/models/layers/attention.py:35-101
(MultiHeadSelfAttention(align_to='output'), raw-view head regroup, no scale, ReLU)."""
from torch import nn

from ... import ops


class MultiHeadAttention(nn.Module):
    def __init__(self, input_dim, attention_dim=None, num_heads=1, dropout_rate=0., use_residual=True,
                 use_scale=False, layer_norm=False, align_to="input"):
        super().__init__()
        if attention_dim is None:
            attention_dim = input_dim // num_heads
        self.attention_dim = attention_dim
        self.output_dim = num_heads * attention_dim
        self.num_heads = num_heads
        self.use_residual = use_residual
        self.align_to = align_to
        if dropout_rate > 0 or layer_norm or use_scale or not use_residual:
            raise NotImplementedError('only the AutoInt configuration (no dropout/layer-norm/scale, residual) is on the hot path')
        self.W_q = nn.Linear(input_dim, self.output_dim, bias=False)
        self.W_k = nn.Linear(input_dim, self.output_dim, bias=False)
        self.W_v = nn.Linear(input_dim, self.output_dim, bias=False)
        if input_dim != self.output_dim:
            if align_to != "output":
                raise NotImplementedError('align_to="input" (transformer style) is not used by AutoInt')
            self.W_res = nn.Linear(input_dim, self.output_dim, bias=False)
        else:
            self.W_res = None


class MultiHeadSelfAttention(MultiHeadAttention):
    def forward(self, X):
        return ops.autoint_attention(X, self.W_q.weight, self.W_k.weight, self.W_v.weight,
                                     self.W_res.weight if self.W_res is not None else None,
                                     self.num_heads, self.attention_dim)

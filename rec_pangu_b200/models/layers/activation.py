"""Activation lookup (reference: rec_pangu/models/layers/activation.py:37-59).  Only what the ranking /
multi-task hot path uses is wired to kernels (ReLU); other names are returned as torch modules for API parity."""
from torch import nn


def get_activation(activation):
    if isinstance(activation, str):
        a = activation.lower()
        if a == 'relu':
            return nn.ReLU()
        if a == 'sigmoid':
            return nn.Sigmoid()
        if a == 'tanh':
            return nn.Tanh()
        return getattr(nn, activation)()
    return activation

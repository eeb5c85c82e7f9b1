"""MMOE — reference: rec_pangu/models/multi_task/mmoe.py:14-130.  (Body filled in once the expert-GEMM /
gate-combine / BatchNorm kernels land; until then constructing works and forward raises.)"""
import torch
from torch import nn

from ..base_model import BaseModel
from ..utils import get_feature_num


class MMOE(BaseModel):
    def __init__(self, num_task=2, n_expert=3, embedding_dim=40, mmoe_hidden_dim=128, expert_activation=None,
                 hidden_dim=[128, 64], dropouts=[0.2, 0.2], enc_dict=None, device=None):
        super().__init__(enc_dict, embedding_dim)
        raise NotImplementedError('MMOE kernels not built yet')

"""ShareBottom — reference: rec_pangu/models/multi_task/sharebottom.py:12-92.  Shared embedding "bottom", one tower per
task on the concatenated [embeddings | dense] row (the padded feature row of the gather kernel is fed to the first
Linear directly, no cat)."""
from typing import Dict, List

from ..base_model import BaseModel
from ..utils import get_feature_num
from ._towers import build_towers, run_towers


class ShareBottom(BaseModel):
    def __init__(self, num_task: int = 2, embedding_dim: int = 40, hidden_units: List[int] = [128, 64],
                 dropouts: List[float] = [0.2, 0.2], enc_dict: Dict[str, dict] = None):
        super().__init__(enc_dict, embedding_dim)
        self.enc_dict = enc_dict
        self.num_task = num_task
        self.hidden_dim = hidden_units
        self.dropouts = dropouts
        self.num_sparse_fea, self.num_dense_fea = get_feature_num(self.enc_dict)
        self.hidden_size = self.num_sparse_fea * self.embedding_dim + self.num_dense_fea
        self.apply(self._init_weights)            # sharebottom.py:38: called BEFORE the towers exist (they keep torch defaults)
        build_towers(self, num_task, self.hidden_size, hidden_units, dropouts)

    def forward(self, data, is_training=True):
        x, _, _ = self.embedding_layer.feature_row(data, with_dense=True)
        return run_towers(self, [x] * self.num_task, data, is_training, eps=0.0, K=self.hidden_size)

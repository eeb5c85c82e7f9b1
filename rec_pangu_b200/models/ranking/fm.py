"""FM — reference: rec_pangu/models/ranking/fm.py."""
import torch  # noqa: F401  (loss_fun strings such as "torch.nn.BCELoss()" are eval-ed here, as in the reference)
from typing import Dict

from ..base_model import BaseModel
from ..layers import FM_Layer


class FM(BaseModel):
    def __init__(self, embedding_dim: int = 32, loss_fun: str = 'torch.nn.BCELoss()', enc_dict: Dict[str, dict] = None):
        super().__init__(enc_dict, embedding_dim)
        self.loss_fun = eval(loss_fun)
        self.enc_dict = enc_dict
        self.fm = FM_Layer()
        self.reset_parameters()

    def forward(self, data, is_training: bool = True):
        # FM consumes only the second-order sum: in inference the [B,F,D] rows are never written to HBM (want_x=False)
        _, fm_out, _ = self.embedding_layer.feature_row(data, with_dense=False, want_fm=True, want_x=False)
        return self._finish(fm_out.unsqueeze(1), data, is_training)

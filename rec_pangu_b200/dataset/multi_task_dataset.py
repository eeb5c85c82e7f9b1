"""MultiTaskDataset — reference: rec_pangu/dataset/multi_task_dataset.py (the reference crashes at HEAD on a
non-existent self.data(); this one produces the documented wire format:
{col: ..., 'task1_label': ..., 'task2_label': ...})."""
from typing import Dict

import numpy as np
import pandas as pd
import torch

from .base_dataset import BaseDataset


class MultiTaskDataset(BaseDataset):
    def __init__(self, config: dict, df: pd.DataFrame, enc_dict: Dict[str, dict] = None):
        self.label_cols = list(config['label_col'])
        super().__init__(config, df, enc_dict)

    def enc_data(self):
        super().enc_data()
        for i, c in enumerate(self.label_cols):
            if c in self.df.columns:
                self.data_dict[f'task{i + 1}_label'] = torch.tensor(np.asarray(self.df[c], dtype=np.float32))

    def __getitem__(self, index: int):
        data = {col: self.data_dict[col][index] for col in self.dense_cols + self.sparse_cols}
        for i in range(len(self.label_cols)):
            k = f'task{i + 1}_label'
            if k in self.data_dict:
                data[k] = self.data_dict[k][index]
        return data

"""BaseDataset — same schema/enc_dict/wire format as the reference (rec_pangu/dataset/base_dataset.py:14-133):
sparse columns -> sorted-unique string vocabulary, OOV -> vocab_size (base_dataset.py:57-61,92); dense columns ->
(x - min) / (max - min + 1e-5) (base_dataset.py:79-80); __getitem__ yields {col: scalar tensor, 'label': ...}."""
from typing import Dict

import numpy as np
import pandas as pd
import torch
from torch.utils.data import Dataset


class BaseDataset(Dataset):
    def __init__(self, config: dict, df: pd.DataFrame, enc_dict: Dict[str, dict] = None):
        self.config = config
        self.enc_dict = enc_dict
        self.df = df.rename(columns={self.config['label_col']: 'label'}) if isinstance(self.config['label_col'], str) else df
        self.dense_cols = list(set(self.config['dense_cols']))
        self.sparse_cols = list(set(self.config['sparse_cols']))
        self.feature_name = self.dense_cols + self.sparse_cols
        if self.enc_dict is None:
            self.get_enc_dict()
        self.enc_data()

    def get_enc_dict(self) -> Dict[str, dict]:
        self.enc_dict = {c: dict() for c in self.dense_cols + self.sparse_cols}
        for f in self.sparse_cols:
            vals = sorted(self.df[f].astype('str').unique())
            m = dict(zip(vals, range(len(vals))))
            m['vocab_size'] = len(vals)
            self.enc_dict[f] = m
        for f in self.dense_cols:
            self.enc_dict[f]['min'] = self.df[f].min()
            self.enc_dict[f]['max'] = self.df[f].max()
        return self.enc_dict

    def enc_dense_data(self, col: str):
        return (self.df[col] - self.enc_dict[col]['min']) / (self.enc_dict[col]['max'] - self.enc_dict[col]['min'] + 1e-5)

    def enc_sparse_data(self, col: str):
        m = self.enc_dict[col]
        oov = m['vocab_size']
        return self.df[col].astype('str').map(lambda x: m.get(x, oov))

    def enc_data(self):
        self.data_dict = {}
        for col in self.dense_cols:
            self.data_dict[col] = torch.tensor(np.asarray(self.enc_dense_data(col), dtype=np.float32))
        for col in self.sparse_cols:
            self.data_dict[col] = torch.tensor(np.asarray(self.enc_sparse_data(col), dtype=np.int64))
        if 'label' in self.df.columns:
            self.data_dict['label'] = torch.tensor(np.asarray(self.df['label'], dtype=np.float32))

    def __getitem__(self, index: int) -> Dict[str, torch.Tensor]:
        data = {col: self.data_dict[col][index] for col in self.dense_cols + self.sparse_cols}
        if 'label' in self.data_dict:
            data['label'] = self.data_dict['label'][index]
        return data

    def __len__(self) -> int:
        return len(self.df)

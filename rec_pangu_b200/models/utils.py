"""Feature bookkeeping helpers (reference: rec_pangu/models/utils.py:122-170)."""
from typing import Dict, List, Tuple

import torch


def sparse_feature_names(enc_dict: Dict) -> List[str]:
    """Field order = insertion order of enc_dict keys carrying 'vocab_size' (models/layers/embedding.py:28-30)."""
    return [c for c in enc_dict.keys() if 'vocab_size' in enc_dict[c].keys()]


def dense_feature_names(enc_dict: Dict) -> List[str]:
    """Dense order = enc_dict keys carrying 'min' (models/utils.py:133-135)."""
    return [c for c in enc_dict.keys() if 'min' in enc_dict[c].keys()]


def get_linear_input(enc_dict: Dict, data: Dict) -> torch.Tensor:
    """[B, Nd] stack of the dense columns (models/utils.py:122-137).  Model forwards never call this on the hot
    path — the gather kernel writes the dense columns straight into the feature row — it exists for API parity."""
    return torch.stack([data[c] for c in dense_feature_names(enc_dict)], dim=1)


def get_feature_num(enc_dict: Dict) -> Tuple[int, int]:
    """(num_sparse, num_dense) — models/utils.py:154-170 ('min' wins over 'vocab_size' as in the reference)."""
    num_sparse = num_dense = 0
    for col in enc_dict.keys():
        if 'min' in enc_dict[col].keys():
            num_dense += 1
        elif 'vocab_size' in enc_dict[col].keys():
            num_sparse += 1
    return num_sparse, num_dense


def get_dnn_input_dim(enc_dict: Dict, embedding_dim: int) -> int:
    """models/utils.py:140-151."""
    num_sparse, num_dense = get_feature_num(enc_dict)
    return num_sparse * embedding_dim + num_dense


def seed_everything(seed: int = 1029):
    """models/utils.py:16 (default seed 1029)."""
    import os
    import random
    import numpy as np
    random.seed(seed)
    os.environ['PYTHONHASHSEED'] = str(seed)
    np.random.seed(seed)
    torch.manual_seed(seed)
    if torch.cuda.is_available():
        torch.cuda.manual_seed_all(seed)

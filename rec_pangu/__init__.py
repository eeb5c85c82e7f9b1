"""Alias of rec_pangu_b200 under the reference's package name, so that the reference's example scripts
(`from rec_pangu.models.ranking import DeepFM`, `from rec_pangu.trainer import RankTrainer`, ...) run unchanged."""
import importlib
import sys

import rec_pangu_b200 as _impl

__version__ = _impl.__version__

_ALIASES = ['models', 'models.layers', 'models.ranking', 'models.multi_task', 'models.utils', 'models.base_model',
            'trainer', 'model_pipeline', 'dataset', 'utils']
for _name in _ALIASES:
    try:
        _mod = importlib.import_module('rec_pangu_b200.' + _name)
    except ModuleNotFoundError as e:      # sub-module not built yet
        if 'rec_pangu_b200' not in str(e):
            raise
        continue
    sys.modules['rec_pangu.' + _name] = _mod
    if '.' not in _name:
        setattr(sys.modules[__name__], _name, _mod)

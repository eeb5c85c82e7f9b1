"""Multi-task models (reference: rec_pangu/models/multi_task/__init__.py).  MMOE is the north-star config 5; ShareBottom,
OMOE, MLMMOE and ESSM reuse its kernels (gather, [K,N] GEMM, gate-softmax/combine, BatchNorm towers, sigmoid+BCE)."""
from .mmoe import MMOE
from .sharebottom import ShareBottom
from .omoe import OMOE
from .mlmmoe import MLMMOE
from .essm import ESSM


def _unported(name):
    class _Unported:
        def __init__(self, *args, **kwargs):
            raise NotImplementedError(f'{name} is not part of the B200 hot-path scope yet (SURVEY.md §8f rank 4)')
    _Unported.__name__ = name
    return _Unported


AITM = _unported('AITM')

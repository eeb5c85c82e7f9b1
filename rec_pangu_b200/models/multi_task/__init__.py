"""Multi-task models (reference: rec_pangu/models/multi_task/__init__.py).  MMOE is the north-star config 5; ShareBottom,
OMOE, MLMMOE and ESSM reuse its kernels (gather, [K,N] GEMM, gate-softmax/combine, BatchNorm towers, sigmoid+BCE)."""
from .mmoe import MMOE
from .sharebottom import ShareBottom
from .omoe import OMOE
from .mlmmoe import MLMMOE
from .essm import ESSM


from .aitm import AITM

"""Ranking models with the reference's names and constructor signatures (rec_pangu/models/ranking/__init__.py)."""
from .wdl import WDL
from .deepfm import DeepFM
from .nfm import NFM
from .fibinet import FiBiNet
from .autoint import AutoInt
from .fm import FM
from .xdeepfm import xDeepFM
from .dcn import DCN
from .afm import AFM
from .masknet import MaskNet
from .lr import LR
from .afn import AFN
from .aoanet import AOANet
from .ccpm import CCPM

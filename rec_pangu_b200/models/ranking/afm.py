"""AFM — reference: rec_pangu/models/ranking/afm.py:14-67.  In the reference this class is, line for line, the FiBiNet
body under another name (its own "Fixme" at afm.py:12 says so): LR + shared bilinear layer over
the raw and the SENET-reweighted embeddings + MLP.  Same parameters, same state_dict keys, same kernels."""
from .fibinet import FiBiNet


class AFM(FiBiNet):
    pass

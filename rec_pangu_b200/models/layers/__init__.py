from .embedding import EmbeddingLayer
from .interaction import (InnerProductLayer, FM_Layer, CrossNet, CrossInteractionLayer, CompressedInteractionNet,
                          SENET_Layer, BilinearInteractionLayer)
from .deep import MLP
from .shallow import LR_Layer
from .attention import MultiHeadSelfAttention, MultiHeadAttention
from .activation import get_activation

# ruff: noqa: F401
from .activation import ReLU, Softmax, softmax
from .attention import Attention
from .core import Layer, Parameter, weight_init_fn
from .embedding import InputEmbedding, OutputEmbedding, PositionalEncoding
from .linear import Linear
from .loss import softmax_cross_entropy
from .mlp import MLP
from .model import Model
from .normalization import LayerNorm
from .transformer_block import TransformerBlock
from .transformer import Transformer

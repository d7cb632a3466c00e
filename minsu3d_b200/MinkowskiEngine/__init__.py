"""Drop-in for the subset of the MinkowskiEngine Python surface that minsu3d uses (boundary #1).

Symbols: SparseTensor, MinkowskiConvolution, MinkowskiConvolutionTranspose,
MinkowskiBatchNorm, MinkowskiReLU, cat, utils.sparse_quantize, utils.sparse_collate -- call sites
minsu3d/model/module/common.py:12-93, backbone.py:14-38, tiny_unet.py:13-15,
general_model.py:187-191, data/dataset/general_dataset.py:159, data/data_module.py:94.

Parameter names and shapes follow MinkowskiEngine so published checkpoints load: convolution
weight = parameter `kernel` of shape (K, Cin, Cout) (2-D when K == 1), batch norm = child `bn`.
All arithmetic runs in libb2s (hand-written sm_100a CUDA) through minsu3d_b200.ops.
"""
from . import utils  # noqa: F401
from .modules import (MinkowskiBatchNorm, MinkowskiConvolution, MinkowskiConvolutionTranspose,  # noqa: F401
                      MinkowskiGlobalAvgPooling, MinkowskiLinear, MinkowskiReLU, cat)
from .sparse_tensor import CoordinateManager, CoordinateMapKey, SparseTensor  # noqa: F401

__version__ = "0.5.4+b200"

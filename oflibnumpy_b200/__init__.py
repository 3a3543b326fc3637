"""oflibnumpy_b200 -- B200-native (sm_100a CUDA) implementation of oflibnumpy's flow-field hot path, behind the
reference's own API: ``import oflibnumpy_b200 as of`` then ``of.Flow``, ``of.apply_flow``, ``of.combine_flows`` ...

The compute path is liboflib_b200.so (hand-written CUDA behind a C ABI, include/oflib_b200.h) called through ctypes.
There is no CPU fallback: without the library or without a GPU, calls raise.
"""
from .flow import Flow
from .ops import *  # noqa: F401,F403  (mirrors `from .flow_operations import *` of the reference)
from .ops import from_matrix, from_transforms, load_kitti, load_sintel, load_sintel_mask, resize_flow, apply_flow, \
    is_zero_flow, track_pts
from .batch import FlowBatch
from . import device
from ._lib import OflibCudaError

__version__ = '0.1.0'

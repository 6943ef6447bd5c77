"""bevrender_b200 -- B200-native (sm_100a) implementation of BEVRender's deformable-attention hot path.

Host side: Python/PyTorch (device memory, streams, torch.distributed).  Compute: hand-written
CUDA behind the C ABI of ``include/bevr.h`` (``libbevr.so``, loaded with ctypes).
"""
from . import _lib
from .functional import grid_gather, ms_deform_attn, project_points

__all__ = ["ms_deform_attn", "grid_gather", "project_points", "_lib"]

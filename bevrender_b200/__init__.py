"""bevrender_b200 -- B200-native (sm_100a) implementation of BEVRender's deformable-attention hot path.

Host side: Python/PyTorch (device memory, streams, torch.distributed).  Compute: hand-written
CUDA behind the C ABI of ``include/bevr.h`` (``libbevr.so``, loaded with ctypes).
"""
from . import _lib
from .bev_cmr_proj import BEV2CameraProjector
from .deform_attn import SCADeformableAttention, TSADeformableAttention
from .wrappers import SpatialCrossAttn, TemporalSelfAttn
from .functional import affine_warp, correlate, retrieval_recall, retrieval_ranks, grid_gather, ms_deform_attn, project_history_bev, project_points, rpe_attention, offset_head

__all__ = ["SpatialCrossAttn", "TemporalSelfAttn", "TSADeformableAttention", "SCADeformableAttention", "BEV2CameraProjector", "ms_deform_attn", "grid_gather", "project_points", "rpe_attention", "offset_head", "correlate", "affine_warp", "project_history_bev", "retrieval_recall", "retrieval_ranks", "_lib"]

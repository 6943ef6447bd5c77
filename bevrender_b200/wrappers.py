"""Drop-in replacements for the thin attention wrappers model/SCA.py:8-162 and model/TSA.py:7-55
("next" row f-3 of SURVEY.md section 8: the reference-point upload path).

Same constructor keywords, sub-module names (``spatial_deform_attn`` / ``temporal_deform_attn``: state_dict keys
are unchanged) and forward signatures.  What changes is the host side of ``SpatialCrossAttn.forward``
(SCA.py:78-85): the reference stacks the projected points, uploads them host->device and ``repeat``s them over the
batch on EVERY forward, after a ``vehicle_type_idx.item()`` sync; here the projection runs once on the device (K4),
the ``(1, V, N/2, N*D, 2)`` tensor is cached as a buffer and broadcast with ``expand`` (the attention module only
reads it), and the ``.item()`` lookup is skipped when the projector knows a single vehicle type.
"""
import torch
from torch import nn

from .deform_attn import SCADeformableAttention, TSADeformableAttention


class TemporalSelfAttn(nn.Module):
    def __init__(self, bev_feat_shape, dim_embed, n_heads, n_groups, stride, kernel_size, batch_size,
                 scale_offset_range, n_views=3, attn_drop_rate=0.0, proj_drop_rate=0.0, data_type=torch.float32,
                 logger=None):
        super().__init__()
        self.logger = logger
        assert n_heads % n_groups == 0, "n_heads must be divisible by n_groups"
        self.temporal_deform_attn = TSADeformableAttention(
            bev_feat_shape=bev_feat_shape, dim_embed=dim_embed, n_heads=n_heads, n_groups=n_groups, stride=stride,
            kernel_size=kernel_size, scale_offset_range=scale_offset_range, batch_size=batch_size, n_views=n_views,
            attn_drop_rate=attn_drop_rate, proj_drop_rate=proj_drop_rate, data_type=data_type, logger=logger)

    def forward(self, query, prev_bev, wandb_log_dict, return_wandb_log=True):
        return self.temporal_deform_attn(x=prev_bev, query=query, wandb_log_dict=wandb_log_dict,
                                         return_wandb_log=return_wandb_log)


class SpatialCrossAttn(nn.Module):
    def __init__(self, bev_bound, bev2cmr_projector, bev_feat_shape, bev_depth_dim, z_shift, dim_embed, n_heads,
                 n_groups, stride, kernel_size, batch_size, scale_offset_range, n_views=3, attn_drop_rate=0.0,
                 proj_drop_rate=0.0, data_type=torch.float32, logger=None):
        super().__init__()
        self.bev_bound, self.bev_feat_shape, self.bev_depth_dim, self.z_shift = bev_bound, bev_feat_shape, bev_depth_dim, z_shift
        self.logger, self.batch_size, self.num_views = logger, batch_size, n_views
        self.points_2d_dict = bev2cmr_projector.bev_grid_to_camera(self.sample_3d_points())
        assert n_heads % n_groups == 0, "n_heads must be divisible by n_groups"
        self._ref_cache = {}
        self.spatial_deform_attn = SCADeformableAttention(
            bev_feat_shape=bev_feat_shape, bev_depth_dim=bev_depth_dim, dim_embed=dim_embed, n_heads=n_heads,
            n_groups=n_groups, stride=stride, kernel_size=kernel_size, scale_offset_range=scale_offset_range,
            batch_size=batch_size, n_views=n_views, attn_drop_rate=attn_drop_rate, proj_drop_rate=proj_drop_rate,
            data_type=data_type, logger=logger)

    def _reference_points(self, vehicle_type_idx, device, batch):
        if len(self.points_2d_dict) == 1:                       # one vehicle type: no .item() sync needed
            key = next(iter(self.points_2d_dict))
        else:
            key = vehicle_type_idx.item() if torch.is_tensor(vehicle_type_idx) else int(vehicle_type_idx)
        ck = (key, str(device))
        if ck not in self._ref_cache:
            pts = torch.stack(self.points_2d_dict[key], dim=0).to(device)          # (V, 2, N/2, N, D)
            V, _, hh, ww, dd = pts.shape
            self._ref_cache[ck] = pts.permute(0, 2, 3, 4, 1).reshape(1, V, hh, ww * dd, 2).contiguous()
        return self._ref_cache[ck].expand(batch, -1, -1, -1, -1)                   # "v n h w d -> b v h (w d) n"

    def forward(self, query, img_feat, vehicle_type_idx, wandb_log_dict, return_wandb_log=True):
        """query (B, C, N, N); img_feat (B*V, C, Hi, Wi) or (B, V, C, Hi, Wi) -> ((B, C, N, N), wandb_log_dict)."""
        if img_feat.dim() == 4:
            img_feat = img_feat.reshape(self.batch_size, self.num_views, *img_feat.shape[1:])
        ref = self._reference_points(vehicle_type_idx, query.device, self.batch_size)
        return self.spatial_deform_attn(x=img_feat, query=query, reference_points=ref, wandb_log_dict=wandb_log_dict,
                                        return_wandb_log=return_wandb_log)

    def sample_3d_points(self):
        """(4, N/2, N, D) homogeneous pillar points: x in (0, X), y in (-Y, Y), z in (-Z, Z) + z_shift, sampled at cell
        centres with the float ``arange`` bounds of SCA.py:121-148 (the element count depends on them)."""
        X, Y, Z = self.bev_bound["X"], self.bev_bound["Y"], self.bev_bound["Z"]
        hx, hy, hz = X / self.bev_feat_shape, Y / self.bev_feat_shape, Z / self.bev_depth_dim
        xs = torch.arange(0 + hx, X + hx, hx * 2)
        ys = torch.arange(-Y + hy, Y + hy, hy * 2)
        zs = torch.arange(-Z + hz + self.z_shift, Z + hz + self.z_shift, hz * 2)
        gx, gy, gz = torch.meshgrid(xs, ys, zs, indexing="ij")
        return torch.stack((gx, gy, gz, torch.ones_like(gz)), 0)

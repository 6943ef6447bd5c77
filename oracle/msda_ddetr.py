"""Second, independent restatement of the multi-scale deformable-attention operator: the published reference algorithm of
Deformable DETR (Zhu et al., ICLR 2021; fundamentalvision/Deformable-DETR, ``models/ops/functions/ms_deform_attn_func.py``,
function ``ms_deform_attn_core_pytorch`` -- the pure-PyTorch definition its CUDA operator is tested against).

TEST INFRASTRUCTURE ONLY.  The bevrender reference has no multi-level operator and does not depend on
Deformable DETR; the north star names that operator's signature (value, spatial_shapes, level_start_index,
sampling_locations, attention_weights), so this file restates ITS published algorithm -- not code from /root/reference,
not copied source -- to pin ``oracle/msda.py`` (and through it kernels K1/K2) to a definition that exists outside this
repository: ``tests/test_oracle.py::test_msda_oracle_matches_deformable_detr_definition``.

Published algorithm, in its own layout:
  value (N, S, M, D), value_spatial_shapes [(H_l, W_l)], sampling_locations (N, Lq, M, L, P, 2) in [0, 1] as (x, y),
  attention_weights (N, Lq, M, L, P)
  1. split ``value`` along S into the L levels;
  2. ``sampling_grids = 2 * sampling_locations - 1``;
  3. per level: value_l -> (N*M, D, H_l, W_l); grid_l -> (N*M, Lq, P, 2);
     ``F.grid_sample(value_l, grid_l, mode="bilinear", padding_mode="zeros", align_corners=False)`` -> (N*M, D, Lq, P);
  4. stack the levels on the last axis and flatten -> (N*M, D, Lq, L*P); weights -> (N*M, 1, Lq, L*P);
  5. output = (stacked * weights).sum(-1) -> (N, M*D, Lq) -> transpose -> (N, Lq, M*D).
"""
import torch
import torch.nn.functional as F


def ms_deform_attn_core(value, value_spatial_shapes, sampling_locations, attention_weights):
    N, S, M, D = value.shape
    _, Lq, _, L, P, _ = sampling_locations.shape
    value_list = value.split([H * W for H, W in value_spatial_shapes], dim=1)
    sampling_grids = 2 * sampling_locations - 1
    sampled = []
    for lid, (H, W) in enumerate(value_spatial_shapes):
        value_l = value_list[lid].flatten(2).transpose(1, 2).reshape(N * M, D, H, W)
        grid_l = sampling_grids[:, :, :, lid].transpose(1, 2).flatten(0, 1)
        sampled.append(F.grid_sample(value_l, grid_l, mode="bilinear", padding_mode="zeros", align_corners=False))
    weights = attention_weights.transpose(1, 2).reshape(N * M, 1, Lq, L * P)
    out = (torch.stack(sampled, dim=-2).flatten(-2) * weights).sum(-1).view(N, M * D, Lq)
    return out.transpose(1, 2).contiguous()

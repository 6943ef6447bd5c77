"""Oracle for the history-BEV ego-motion warp ("next" row f-1).

Restates ``EncoderLayer.project_history_bev_feat`` (encoder.py:413-466) with the same third-party call the
reference makes -- torchvision's ``F.affine`` (bilinear, fill=0), twice per sample: rotate by the previous
heading + translate by the ego motion, then rotate back by the current heading.  torchvision is unpinned by
the reference; 0.26.0 is installed (its tensor path = ``_gen_affine_grid`` + ``grid_sample(align_corners=False)``
+ the fill mask).  Pinned by ``tests/golden/history_warp.npz`` (output of the real reference method).
"""
import math

import torch
import torchvision.transforms.functional as TF
from torchvision.transforms import InterpolationMode


def project_history_bev(bev, vehicle_pose):
    out = []
    for i in range(bev.shape[0]):
        prev_rot, curr_rot = (float(v) for v in vehicle_pose[i, :, 2])
        dx, dy = (vehicle_pose[i, 0] - vehicle_pose[i, 1])[:2]
        w = TF.affine(bev[i], angle=math.degrees(prev_rot), translate=(dx, dy), scale=1.0, shear=0,
                      interpolation=InterpolationMode.BILINEAR, fill=0)
        w = TF.affine(w, angle=math.degrees(-curr_rot), translate=(0, 0), scale=1.0, shear=0,
                      interpolation=InterpolationMode.BILINEAR, fill=0)
        out.append(w)
    return torch.stack(out, 0)

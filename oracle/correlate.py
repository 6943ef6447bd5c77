"""Oracle for the rotation-sweep BEV<->map correlation (kernel K5).

PARITY UNPINNED: the reference contains no registration / correlation code at all (SURVEY.md section 0.3 -- the
closest thing is a retrieval recall from ``2 - 2*X@Y.T`` in numpy, train.py:551-572), so this file DEFINES the
operator the CUDA kernel implements: rotate the template about its centre with bilinear ``F.grid_sample``
(zeros outside), then valid-mode cross-correlate every rotation with the map tile using ``F.conv2d``.
The torch formulation is checked against plain numpy loops written from the formula in include/bevr.h
(tests/test_oracle.py::test_correlation_oracle_matches_first_principles_loops).
"""
import math

import torch
import torch.nn.functional as F


def rotate_templates(template, angles_rad):
    """template (C, Ht, Wt) -> (R, C, Ht, Wt); rot_r[c,u,v] = bilinear(template[c], centre + R(a_r)(v-cx, u-cy))."""
    Cc, Ht, Wt = template.shape
    cx, cy = 0.5 * (Wt - 1), 0.5 * (Ht - 1)
    u, v = torch.meshgrid(torch.arange(Ht, dtype=template.dtype), torch.arange(Wt, dtype=template.dtype), indexing="ij")
    dx, dy = v - cx, u - cy
    out = []
    for a in angles_rad:
        co, si = math.cos(a), math.sin(a)
        sx, sy = cx + co * dx - si * dy, cy + si * dx + co * dy
        grid = torch.stack((sx / cx - 1.0, sy / cy - 1.0), -1)[None]          # align_corners=True normalisation
        out.append(F.grid_sample(template[None], grid, mode="bilinear", padding_mode="zeros", align_corners=True)[0])
    return torch.stack(out, 0)


def correlate(map_tile, template, angles_rad, operand_dtype=None):
    """(C,Hm,Wm), (C,Ht,Wt) -> (R, Hm-Ht+1, Wm-Wt+1).  ``operand_dtype=torch.bfloat16`` rounds both operands the
    way the tensor-core kernel does (products still accumulate in fp32/fp64)."""
    rot = rotate_templates(template, angles_rad)
    m = map_tile
    if operand_dtype is not None:
        rot, m = rot.to(operand_dtype).to(template.dtype), m.to(operand_dtype).to(template.dtype)
    return F.conv2d(m[None], rot)[0]

"""numpy restatement of ``torch.nn.functional.grid_sample`` (bilinear, zeros).

The reference calls ``F.grid_sample(..., mode="bilinear", align_corners=True)``
with the default ``padding_mode="zeros"`` at TSA_deform_attn.py:210-217 and
:292-302, SCA_deform_attn.py:290-301 and :379-389.  ``grid_sample`` is
third-party (PyTorch, unpinned by the reference; torch 2.11.0 here); this file
restates its published semantics so the oracle does not depend on the library
for the one primitive the CUDA kernels re-implement:

  pixel = (g + 1) / 2 * (size - 1)            align_corners=True
  pixel = ((g + 1) * size - 1) / 2            align_corners=False
  four taps nw/ne/sw/se around floor(pixel); a tap outside the image adds 0.

``tests/test_oracle.py`` checks these functions against ``F.grid_sample`` and
its autograd (forward and both gradients) in float64.
"""
import numpy as np


def unnormalize(g, size, align_corners):
    if align_corners:
        return (g + 1.0) / 2.0 * (size - 1)
    return ((g + 1.0) * size - 1.0) / 2.0


def _taps(px, py, H, W):
    x0 = np.floor(px)
    y0 = np.floor(py)
    wx1 = px - x0
    wx0 = (x0 + 1.0) - px
    wy1 = py - y0
    wy0 = (y0 + 1.0) - py
    x0 = x0.astype(np.int64)
    y0 = y0.astype(np.int64)
    taps = []
    for dy, wy in ((0, wy0), (1, wy1)):
        for dx, wx in ((0, wx0), (1, wx1)):
            xi, yi = x0 + dx, y0 + dy
            ok = (xi >= 0) & (xi < W) & (yi >= 0) & (yi < H)
            taps.append((np.where(ok, xi, 0), np.where(ok, yi, 0), ok, wx, wy, dx, dy))
    return taps


def grid_sample_bilinear(inp, grid, align_corners=True):
    """inp (N,C,H,W), grid (N,Ho,Wo,2) in (x,y) -> (N,C,Ho,Wo)."""
    N, C, H, W = inp.shape
    px = unnormalize(grid[..., 0], W, align_corners)
    py = unnormalize(grid[..., 1], H, align_corners)
    out = np.zeros((N, C) + grid.shape[1:3], dtype=inp.dtype)
    n_idx = np.arange(N)[:, None, None]
    for xi, yi, ok, wx, wy, _, _ in _taps(px, py, H, W):
        v = inp[n_idx, :, yi, xi]                       # (N,Ho,Wo,C)
        v = np.where(ok[..., None], v, 0.0)
        out += np.moveaxis(v * (wx * wy)[..., None], -1, 1)
    return out


def grid_sample_bilinear_backward(grad_out, inp, grid, align_corners=True):
    """Returns (grad_inp, grad_grid) of ``grid_sample_bilinear``.

    d/dpixel_x = sum_c g * (-nw*wy0 + ne*wy0 - sw*wy1 + se*wy1), scaled by
    (W-1)/2 (align_corners) or W/2; out-of-range taps count as zero.
    """
    N, C, H, W = inp.shape
    px = unnormalize(grid[..., 0], W, align_corners)
    py = unnormalize(grid[..., 1], H, align_corners)
    g = np.moveaxis(grad_out, 1, -1)                    # (N,Ho,Wo,C)
    grad_inp = np.zeros_like(inp)
    gx = np.zeros(grid.shape[:3], dtype=inp.dtype)
    gy = np.zeros(grid.shape[:3], dtype=inp.dtype)
    n_idx = np.broadcast_to(np.arange(N)[:, None, None], px.shape)
    for xi, yi, ok, wx, wy, dx, dy in _taps(px, py, H, W):
        v = np.where(ok[..., None], inp[n_idx, :, yi, xi], 0.0)
        dot = (g * v).sum(-1)
        gx += dot * wy * (1.0 if dx else -1.0)
        gy += dot * wx * (1.0 if dy else -1.0)
        contrib = g * (wx * wy * ok)[..., None]
        for c in range(C):
            np.add.at(grad_inp[:, c], (n_idx, yi, xi), contrib[..., c])
    sx = (W - 1) / 2.0 if align_corners else W / 2.0
    sy = (H - 1) / 2.0 if align_corners else H / 2.0
    return grad_inp, np.stack((gx * sx, gy * sy), -1)

"""Oracle for the camera->BEV reference-point projection (kernel K4).

Restates ``SpatialCrossAttn.sample_3d_points`` (SCA.py:112-162) and
``BEV2CameraProjector.bev_grid_to_camera`` / ``get_in_bound_mask``
(bev_cmr_proj.py:61-124, incl. the ``remove_ref_in_gray`` image test) with the same torch CPU
primitives the reference calls (``inverse``, ``@``, ``div``, ``to(int32)``), and
additionally returns the truncated pixel indices and the validity mask that the
reference folds into the coordinates -- those are the integer outputs the CUDA
kernel must reproduce bit-exactly.

Pinned by ``tests/golden/proj_*.npz`` (outputs of the real reference classes).
"""
import numpy as np
import torch


def sample_3d_points(bev_bound, bev_feat_shape, bev_depth_dim, z_shift):
    """(4, N/2, N, D) homogeneous pillar points; x in (0,X), y in (-Y,Y), z in (-Z,Z)+z_shift.
    Uses float ``arange`` with the same start/stop/step expressions as SCA.py:121-148
    (the element COUNT depends on them)."""
    X, Y, Z = bev_bound["X"], bev_bound["Y"], bev_bound["Z"]
    xs_, ys_, zs_ = X / bev_feat_shape, Y / bev_feat_shape, Z / bev_depth_dim
    xs = torch.arange(0 + xs_, X + xs_, xs_ * 2)
    ys = torch.arange(-Y + ys_, Y + ys_, ys_ * 2)
    zs = torch.arange(-Z + zs_ + z_shift, Z + zs_ + z_shift, zs_ * 2)
    gx, gy, gz = torch.meshgrid(xs, ys, zs, indexing="ij")
    return torch.stack((gx, gy, gz, torch.ones_like(gz)), 0)


def scale_intrinsics(K, img_w, img_h, ori_w, ori_h):
    """K[0,0],K[0,2] *= img_w/ori_w ; K[1,1],K[1,2] *= img_h/ori_h   (bev_cmr_proj.py:41-46).
    Returns a float32 copy (the reference mutates the caller's array in place)."""
    K = np.array(K, dtype=np.float64, copy=True)
    sx, sy = img_w / ori_w, img_h / ori_h
    K[0, 0] *= sx
    K[0, 2] *= sx
    K[1, 1] *= sy
    K[1, 2] *= sy
    return torch.tensor(K).float()


def project_points(points_3d, imu_to_cam, K_scaled, img_w, img_h, gray_image=None):
    """points_3d (4,h,w,z) fp32 -> uv (2,h,w,z) in [-1,1], pix (2,h,w,z) int32, mask (h,w,z) bool.
    ``gray_image`` (H, W, 3) uint8: remove_ref_in_gray=True, i.e. points whose truncated pixel is (128,128,128)
    in that image are masked too (bev_cmr_proj.py:114-122)."""
    shp = points_3d.shape[1:]
    p = points_3d.reshape(4, -1)
    T = torch.as_tensor(imu_to_cam).float()
    cam = T.inverse() @ p
    uvw = torch.as_tensor(K_scaled).float()[:, :3] @ cam[:3]
    uv = uvw.div(uvw[-1])[:2]
    pix = uv.to(torch.int32)
    mask = pix[1].ge(0) & pix[1].lt(img_h - 1) & pix[0].ge(0) & pix[0].lt(img_w - 1)
    if gray_image is not None:
        img = torch.as_tensor(np.asarray(gray_image)).permute(2, 0, 1)          # pil_to_tensor layout (3, H, W)
        q = pix.masked_fill(~mask, 0).long()
        mask = mask & ((img[:, q[1], q[0]] == 128).sum(0) != 3)
    uv = uv.masked_fill(~mask, 0)
    uv[0] = uv[0] / (img_w - 1)
    uv[1] = uv[1] / (img_h - 1)
    uv = uv * 2 - 1
    return uv.reshape(2, *shp), pix.reshape(2, *shp), mask.reshape(*shp)

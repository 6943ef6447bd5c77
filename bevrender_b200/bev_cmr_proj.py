"""Drop-in replacement for model/bev_cmr_proj.py:12-124 (``BEV2CameraProjector``).

Same constructor and ``bev_grid_to_camera(points_3d) -> {vehicle_code: [Tensor(2,h,w,z) per camera]}``;
the projection + in-bound mask + normalisation run as one sm_100a kernel (K4) that reproduces the
reference's torch-CPU floats bit for bit.  ``inverse()`` of the 4x4 extrinsics stays on the host in
torch, as in the reference.  Deliberate differences: the caller's ``K`` arrays are NOT scaled in place
(the reference double-scales when two projectors share one config), the result
lives on the CUDA device (``SpatialCrossAttn.forward`` moves it there every step anyway,
SCA.py:78-80), and ``last_mask`` / ``last_pix`` expose the integer outputs the reference discards.
"""
import numpy as np
import torch

from .functional import project_points


class BEV2CameraProjector:
    def __init__(self, imu_to_rgb, K, vehicle_type_code, img_width, img_height, ori_img_width, ori_img_height,
                 remove_ref_in_gray=False, bound_check_img_paths=None, device="cuda", logger=None, use_wandb=False):
        self.scale_x = img_width / ori_img_width
        self.scale_y = img_height / ori_img_height
        self.img_width, self.img_height = img_width, img_height
        self.vehicle_type_code = vehicle_type_code
        self.remove_ref_in_gray = remove_ref_in_gray
        self.bound_check_img_paths = bound_check_img_paths
        self.device, self.logger, self.use_wandb = device, logger, use_wandb
        self.imu_to_cmr = {k: [torch.tensor(np.asarray(v)).float() for v in vals] for k, vals in imu_to_rgb.items()}
        self.K = {}
        for k, vals in K.items():
            self.K[k] = []
            for v in vals:
                v = np.array(v, copy=True)
                v[0, 0] *= self.scale_x
                v[0, 2] *= self.scale_x
                v[1, 1] *= self.scale_y
                v[1, 2] *= self.scale_y
                self.K[k].append(torch.tensor(v).float())
        self.img_paths = {vehicle_type_code: bound_check_img_paths}
        self.last_mask, self.last_pix = [], []

    def bev_grid_to_camera(self, points_3d):
        """points_3d (4, h, w, z) homogeneous IMU-frame points -> {code: [uv (2, h, w, z) in [-1,1]]}."""
        if not torch.cuda.is_available():
            raise RuntimeError("BEV2CameraProjector runs its projection kernel on CUDA (no CPU fallback)")
        dev = torch.device(self.device if str(self.device).startswith("cuda") else "cuda")
        pts = points_3d.to(dev, torch.float32)
        code = self.vehicle_type_code
        out, self.last_mask, self.last_pix = [], [], []
        for cam, (T, K) in enumerate(zip(self.imu_to_cmr[code], self.K[code])):
            uv, pix, mask = project_points(pts, T.inverse(), K, self.img_width, self.img_height)
            if self.remove_ref_in_gray:
                uv, mask = self._drop_gray(uv, pix, mask, code, cam)
            out.append(uv)
            self.last_mask.append(mask)
            self.last_pix.append(pix)
        return {code: out}

    def _drop_gray(self, uv, pix, mask, code, cam):
        """bev_cmr_proj.py:114-122: also mask points that land on a (128,128,128) pixel of a reference image."""
        from PIL import Image
        img = torch.from_numpy(np.asarray(Image.open(self.img_paths[code][cam]))).to(uv.device)
        p = pix.masked_fill(~mask[None], 0).long()
        gray = (img[p[1], p[0], :3] == 128).all(-1)
        mask = mask & ~gray
        return torch.where(mask[None], uv, torch.full_like(uv, -1.0)), mask

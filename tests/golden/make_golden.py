#!/usr/bin/env python
"""Generate the golden fixtures in this directory from the REAL reference.

Runs only in the authoring container (needs /root/reference, read-only); the
fixtures travel, the reference does not.  The reference's ``model/*`` imports
``timm.models.layers.{trunc_normal_, DropPath}`` (TSA_deform_attn.py:7,
encoder.py:5); timm is not installed, so a stand-in module is registered in
``sys.modules`` before import (``trunc_normal_`` is torch's own initialiser;
``DropPath`` is identity here because every fixture uses drop rates of 0).

    python tests/golden/make_golden.py            # rewrites tests/golden/*.npz
"""
import copy
import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.environ.get("BEVR_REFERENCE", "/root/reference")


def install_timm_shim():
    if "timm" in sys.modules:
        return
    layers = types.ModuleType("timm.models.layers")
    layers.trunc_normal_ = torch.nn.init.trunc_normal_

    class DropPath(torch.nn.Module):
        def __init__(self, drop_prob=0.0):
            super().__init__()
            self.drop_prob = drop_prob

        def forward(self, x):
            if self.drop_prob == 0.0 or not self.training:
                return x
            keep = 1.0 - self.drop_prob
            mask = x.new_empty(x.shape[0], *([1] * (x.dim() - 1))).bernoulli_(keep)
            return x * mask / keep

    layers.DropPath = DropPath
    timm = types.ModuleType("timm")
    models = types.ModuleType("timm.models")
    timm.models, models.layers = models, layers
    sys.modules.update({"timm": timm, "timm.models": models, "timm.models.layers": layers})


def load_reference():
    install_timm_shim()
    if REF not in sys.path:
        sys.path.insert(0, REF)
    from model.TSA_deform_attn import TSADeformableAttention
    from model.SCA_deform_attn import SCADeformableAttention
    from model.bev_cmr_proj import BEV2CameraProjector
    from model.SCA import SpatialCrossAttn
    return TSADeformableAttention, SCADeformableAttention, BEV2CameraProjector, SpatialCrossAttn


def np_state(module):
    return {"param." + k: v.detach().numpy().copy() for k, v in module.state_dict().items()}


def np_grads(module):
    return {"gparam." + k: p.grad.numpy().copy() for k, p in module.named_parameters()
            if p.grad is not None}


def stress_init(module, gen, offset_gain):
    """Reference init leaves rpe_table ~ 0.01 and offsets tiny; push both so the bias path,
    its gradients and the zero-padding branches matter."""
    with torch.no_grad():
        module.rpe_table.copy_(torch.randn(module.rpe_table.shape, generator=gen) * 0.5)
        for name, p in module.named_parameters():
            if name.startswith("conv_offset") and name.endswith(".3.weight"):
                p.mul_(offset_gain)


def make_tsa(name, N, C, h, g, k, s, B, self_attn=False, scale_offset_range=True, seed=0):
    TSA, _, _, _ = load_reference()
    torch.manual_seed(seed)
    gen = torch.Generator().manual_seed(seed + 1)
    m = TSA(bev_feat_shape=N, dim_embed=C, n_heads=h, n_groups=g, stride=s, kernel_size=k,
            scale_offset_range=scale_offset_range, batch_size=B, n_views=1)
    stress_init(m, gen, 8.0 if scale_offset_range else 0.5)
    query = torch.randn(B, C, N, N, generator=gen).requires_grad_(True)
    x = None if self_attn else torch.randn(B, C, N, N, generator=gen).requires_grad_(True)
    grad_out = torch.randn(B, C, N, N, generator=gen)
    out, _ = m(x, query, {}, False)
    out.backward(grad_out)
    d = dict(cfg=np.array([N, C, h, g, k, s, B, int(self_attn), int(scale_offset_range)]),
             query=query.detach().numpy(), grad_out=grad_out.numpy(), out=out.detach().numpy(),
             g_query=query.grad.numpy())
    if x is not None:
        d.update(x=x.detach().numpy(), g_x=x.grad.numpy())
    d.update(np_state(m))
    d.update(np_grads(m))
    np.savez_compressed(os.path.join(HERE, name + ".npz"), **d)
    print(name, "out", tuple(out.shape), "|out|max", float(out.abs().max()))


def camera_setup(yaw_deg=0.0, height=1.5, pitch_deg=0.0):
    """Pinhole camera of BASELINE.md config 1: 1.5 m above the IMU, optical axis along +x.
    Returns (imu_to_cam 4x4, K 3x4 at 640x512).  imu_to_cam maps camera -> IMU coordinates
    (the reference applies its inverse, bev_cmr_proj.py:72)."""
    # camera axes in IMU frame: z_cam = +x_imu (forward), x_cam = -y_imu (right), y_cam = -z_imu (down)
    R = np.array([[0.0, 0.0, 1.0], [-1.0, 0.0, 0.0], [0.0, -1.0, 0.0]])
    cy, sy = np.cos(np.radians(yaw_deg)), np.sin(np.radians(yaw_deg))
    cp, sp = np.cos(np.radians(pitch_deg)), np.sin(np.radians(pitch_deg))
    Rz = np.array([[cy, -sy, 0.0], [sy, cy, 0.0], [0.0, 0.0, 1.0]])
    Ry = np.array([[cp, 0.0, sp], [0.0, 1.0, 0.0], [-sp, 0.0, cp]])
    T = np.eye(4)
    T[:3, :3] = Rz @ Ry @ R
    T[:3, 3] = [0.0, 0.0, height]
    K = np.array([[400.0, 0.0, 320.0, 0.0], [0.0, 400.0, 256.0, 0.0], [0.0, 0.0, 1.0, 0.0]])
    return T, K


def make_proj(name, N, D, bound, z_shift, yaw_deg, pitch_deg, img_w=160, img_h=128):
    _, _, Projector, SpatialCrossAttn = load_reference()
    T, K = camera_setup(yaw_deg, pitch_deg=pitch_deg)
    proj = Projector(imu_to_rgb={0: [T.copy()]}, K={0: [K.copy()]}, vehicle_type_code=0,
                     img_width=img_w, img_height=img_h, ori_img_width=640, ori_img_height=512,
                     remove_ref_in_gray=False, bound_check_img_paths=None, device="cpu")
    stub = types.SimpleNamespace(bev_bound=bound, bev_feat_shape=N, bev_depth_dim=D, z_shift=z_shift)
    pts = SpatialCrossAttn.sample_3d_points(stub)
    uv = proj.bev_grid_to_camera(pts)[0][0]
    # the reference folds the mask into uv; recover mask/pix through its own helper
    p = pts.reshape(4, -1)
    cam = proj.imu_to_cmr[0][0].inverse() @ p
    uvw = proj.K[0][0][:, :3] @ cam[:3]
    raw = uvw.div(uvw[-1])[:2]
    mask = proj.get_in_bound_mask(raw, 0, 0)
    np.savez_compressed(
        os.path.join(HERE, name + ".npz"),
        cfg=np.array([N, D, img_w, img_h, 640, 512]), bound=np.array([bound["X"], bound["Y"], bound["Z"]]),
        z_shift=np.array(z_shift), imu_to_cam=T, K=K, points=pts.numpy(), uv=uv.numpy(),
        mask=mask.reshape(pts.shape[1:]).numpy(), pix=raw.to(torch.int32).reshape(2, *pts.shape[1:]).numpy(),
        K_scaled=proj.K[0][0].numpy(), cam_T=proj.imu_to_cmr[0][0].inverse().numpy())
    print(name, "points", tuple(pts.shape), "valid", float(mask.float().mean()))
    return uv


def make_proj_gray(name, N=28, D=4, img_w=160, img_h=128, seed=29):
    """remove_ref_in_gray=True (bev_cmr_proj.py:114-122): points that land on a (128,128,128) pixel of a
    reference image are masked as well.  The image (random colours, a gray band) is stored in the golden."""
    import tempfile
    from PIL import Image
    _, _, Projector, SpatialCrossAttn = load_reference()
    rng = np.random.default_rng(seed)
    img = rng.integers(0, 256, size=(img_h, img_w, 3), dtype=np.uint8)
    img[img_h // 2:, img_w // 3: 2 * img_w // 3] = 128                  # gray band where the ground projects
    img[5:9, 7:30, 0] = 128                                             # one channel only: must NOT be masked
    T, K = camera_setup(0.0)
    bound = {"X": 20.0, "Y": 10.0, "Z": 2.0}
    with tempfile.TemporaryDirectory() as tmp:
        path = os.path.join(tmp, "ref.png")
        Image.fromarray(img).save(path)
        proj = Projector(imu_to_rgb={0: [T.copy()]}, K={0: [K.copy()]}, vehicle_type_code=0, img_width=img_w,
                         img_height=img_h, ori_img_width=640, ori_img_height=512, remove_ref_in_gray=True,
                         bound_check_img_paths=[path], device="cpu")
        stub = types.SimpleNamespace(bev_bound=bound, bev_feat_shape=N, bev_depth_dim=D, z_shift=-1.0)
        pts = SpatialCrossAttn.sample_3d_points(stub)
        uv = proj.bev_grid_to_camera(pts)[0][0]
        p = pts.reshape(4, -1)
        cam = proj.imu_to_cmr[0][0].inverse() @ p
        uvw = proj.K[0][0][:, :3] @ cam[:3]
        mask = proj.get_in_bound_mask(uvw.div(uvw[-1])[:2], 0, 0)
    np.savez_compressed(os.path.join(HERE, name + ".npz"), cfg=np.array([N, D, img_w, img_h, 640, 512]), imu_to_cam=T,
                        K=K, image=img, points=pts.numpy(), uv=uv.numpy(), mask=mask.reshape(pts.shape[1:]).numpy())
    print(name, "valid", float(mask.float().mean()))


def make_sca(name, N, D, C, h, g, B, Hi, Wi, scale_offset_range=True, seed=0):
    _, SCA, Projector, SpatialCrossAttn = load_reference()
    torch.manual_seed(seed)
    gen = torch.Generator().manual_seed(seed + 1)
    m = SCA(bev_feat_shape=N, bev_depth_dim=D, dim_embed=C, n_heads=h, n_groups=g, stride=1,
            kernel_size=1, scale_offset_range=scale_offset_range, batch_size=B, n_views=1)
    stress_init(m, gen, 4.0 if scale_offset_range else 0.2)
    T, K = camera_setup()
    proj = Projector(imu_to_rgb={0: [T]}, K={0: [K]}, vehicle_type_code=0, img_width=Wi * 4,
                     img_height=Hi * 4, ori_img_width=640, ori_img_height=512, device="cpu")
    stub = types.SimpleNamespace(bev_bound={"X": 20.0, "Y": 10.0, "Z": 2.0}, bev_feat_shape=N,
                                 bev_depth_dim=D, z_shift=-1.0)
    uv = proj.bev_grid_to_camera(SpatialCrossAttn.sample_3d_points(stub))[0]
    from einops import repeat
    ref_pts = repeat(torch.stack(uv, 0), "v n h w d -> b v h (w d) n", b=B).contiguous()
    query = torch.randn(B, C, N, N, generator=gen).requires_grad_(True)
    x = torch.randn(B, 1, C, Hi, Wi, generator=gen).requires_grad_(True)
    grad_out = torch.randn(B, C, N, N, generator=gen)
    out, _ = m(x, query, ref_pts, {}, False)
    out.backward(grad_out)
    d = dict(cfg=np.array([N, D, C, h, g, B, Hi, Wi, int(scale_offset_range)]), query=query.detach().numpy(),
             x=x.detach().numpy(), reference_points=ref_pts.numpy(), grad_out=grad_out.numpy(),
             out=out.detach().numpy(), g_query=query.grad.numpy(), g_x=x.grad.numpy())
    d.update(np_state(m))
    d.update(np_grads(m))
    np.savez_compressed(os.path.join(HERE, name + ".npz"), **d)
    print(name, "out", tuple(out.shape), "masked ref", float((ref_pts == -1).all(-1).float().mean()))


def make_history_warp(name, B=3, C=8, N=14, seed=17):
    """EncoderLayer.project_history_bev_feat (encoder.py:413-466) uses no attribute of self."""
    load_reference()
    from model.encoder import EncoderLayer
    gen = torch.Generator().manual_seed(seed)
    bev = torch.randn(B, C, N, N, generator=gen)
    pose = torch.randn(B, 2, 3, generator=gen)
    pose[..., :2] *= 2.0           # translations of a few BEV cells
    pose[..., 2] *= 0.4            # headings in radians
    out = EncoderLayer.project_history_bev_feat(None, bev, pose)
    np.savez_compressed(os.path.join(HERE, name + ".npz"), bev=bev.numpy(), pose=pose.numpy(), out=out.numpy())
    print(name, tuple(out.shape), float(out.abs().mean()))


def make_encoder_layer(name, N=14, C=64, h=2, D=5, B=2, Hi=16, Wi=20, seed=23):
    """One reference EncoderLayer (encoder.py:243-411) in train mode with prev_bev: LPU -> TSA -> MLP -> LPU -> SCA -> MLP."""
    _, _, Projector, _ = load_reference()
    from model.encoder import EncoderLayer
    torch.manual_seed(seed)
    gen = torch.Generator().manual_seed(seed + 1)
    T, K = camera_setup()
    proj = Projector(imu_to_rgb={0: [T]}, K={0: [K]}, vehicle_type_code=0, img_width=Wi * 4, img_height=Hi * 4,
                     ori_img_width=640, ori_img_height=512, device="cpu")
    layer = EncoderLayer(bev_bound={"X": 20.0, "Y": 10.0, "Z": 2.0}, bev2cmr_projector=proj, n_views=1, bev_feat_shape=N,
                         bev_depth_dim=D, z_shift=-1.0, dim_embed=C, expansion=2, stage_idx=0, n_groups=1, n_heads=h,
                         stride=1, kernel_size=3, batch_size=B, scale_offset_range=True, drop_path_rate=0.0)
    stress_init(layer.temporal_self_attn.temporal_deform_attn, gen, 8.0)
    stress_init(layer.spatial_cross_attn.spatial_deform_attn, gen, 4.0)
    layer.train()
    q = torch.randn(B, C, N, N, generator=gen).requires_grad_(True)
    img = torch.randn(B, C, Hi, Wi, generator=gen).requires_grad_(True)
    prev = torch.randn(B, C, N, N, generator=gen)
    go = torch.randn(B, C, N, N, generator=gen)
    out, _ = layer(q, img, prev, torch.zeros(B, 2, 3), torch.tensor(0), {}, False)
    out.backward(go)
    ref_pts = torch.stack(layer.spatial_cross_attn.points_2d_dict[0], 0)           # (V, 2, N/2, N, D)
    d = dict(cfg=np.array([N, C, h, D, B, Hi, Wi]), query=q.detach().numpy(), img=img.detach().numpy(), prev=prev.numpy(),
             grad_out=go.numpy(), out=out.detach().numpy(), g_query=q.grad.numpy(), g_img=img.grad.numpy(),
             ref_points=ref_pts.numpy())
    d.update(np_state(layer))
    d.update(np_grads(layer))
    np.savez_compressed(os.path.join(HERE, name + ".npz"), **d)
    print(name, tuple(out.shape), float(out.abs().max()))


def main():
    torch.set_num_threads(max(1, os.cpu_count() or 1))
    make_tsa("tsa_s1", N=14, C=64, h=2, g=1, k=3, s=1, B=2)
    make_tsa("tsa_s2", N=16, C=64, h=4, g=1, k=5, s=2, B=2, seed=3)  # reference TSA raises for n_groups>1 (TSA_deform_attn.py:278-283)
    make_tsa("tsa_self", N=12, C=32, h=1, g=1, k=3, s=1, B=1, self_attn=True, seed=5)
    make_tsa("tsa_clamp", N=12, C=64, h=2, g=1, k=3, s=1, B=1, scale_offset_range=False, seed=7)
    make_sca("sca_d5", N=14, D=5, C=64, h=2, g=1, B=2, Hi=16, Wi=20)
    make_sca("sca_g2", N=12, D=4, C=64, h=4, g=2, B=1, Hi=16, Wi=20, seed=11)
    make_sca("sca_clamp", N=12, D=3, C=32, h=1, g=1, B=1, Hi=16, Wi=20, scale_offset_range=False, seed=13)
    make_history_warp("history_warp")
    make_encoder_layer("encoder_layer")
    bound = {"X": 20.0, "Y": 10.0, "Z": 2.0}
    make_proj("proj_n50", 50, 5, bound, -1.0, 0.0, 0.0)
    make_proj("proj_n200", 200, 4, bound, -1.0, 0.0, 0.0)
    make_proj("proj_yaw", 56, 5, bound, -1.0, 75.0, 10.0)          # most points off-image / behind
    make_proj("proj_z0", 28, 4, {"X": 20.0, "Y": 10.0, "Z": 2.0}, 1.0, 90.0, 0.0)  # z_cam == 0 rows
    make_proj_gray("proj_gray")
    make_recall("recall")


def reference_get_recall():
    """``Trainer.get_recall`` (train.py:551-572) lifted out of train.py WITHOUT importing the module (its imports need
    pytorch_metric_learning / yourdfpy): the method's source is cut out with ``ast`` and compiled on its own; it only
    uses numpy."""
    import ast
    src = open(os.path.join(REF, "train.py")).read()
    tree = ast.parse(src)
    fn = next(n for cls in tree.body if isinstance(cls, ast.ClassDef) and cls.name == "Trainer"
              for n in cls.body if isinstance(n, ast.FunctionDef) and n.name == "get_recall")
    mod = ast.Module(body=[fn], type_ignores=[])
    ns = {"np": np}
    exec(compile(mod, "train.py:get_recall", "exec"), ns)
    return lambda cam, maps: ns["get_recall"](None, cam, maps)


def make_recall(name="recall"):
    rng = np.random.default_rng(0)
    maps = rng.standard_normal((96, 64)).astype(np.float32)
    maps /= np.linalg.norm(maps, axis=1, keepdims=True)
    cam = maps + 0.45 * rng.standard_normal(maps.shape).astype(np.float32)
    cam /= np.linalg.norm(cam, axis=1, keepdims=True)
    r = reference_get_recall()(cam, maps)
    np.savez_compressed(os.path.join(HERE, name + ".npz"), camera=cam, maps=maps, recall=np.array(r, dtype=np.float64))
    print(name, "recall@1/5/10", r)


if __name__ == "__main__":
    main()

"""Synthetic workloads of BASELINE.json (shape factories only -- no arithmetic of the path)."""
import torch


def make_microbench_inputs(B, seed=0, dtype=torch.float32, n_side=200, h=8, Dh=32, L=4, P=4,
                           sigma_px=2.0, device="cpu", shapes=None):
    """Synthetic inputs of BASELINE.json config 2:
    BEV pyramid [(N,N),(N/2,N/2),...], queries = cell centres of the NxN grid,
    offsets ~ N(0,(2px)^2) per level, weights = softmax(N(0,1)) over L*P.
    ``shapes``: other value levels (the secondary workload of section 8d: the camera pyramid of a 512x640 image)."""
    g = torch.Generator().manual_seed(seed)
    shapes = [(n_side >> l, n_side >> l) for l in range(L)] if shapes is None else [tuple(s) for s in shapes]
    L = len(shapes)
    starts, s = [], 0
    for H, W in shapes:
        starts.append(s)
        s += H * W
    Nq = n_side * n_side
    value = torch.randn(B, s, h, Dh, generator=g)
    ys, xs = torch.meshgrid(torch.arange(n_side), torch.arange(n_side), indexing="ij")
    ref = torch.stack(((xs + 0.5) / n_side, (ys + 0.5) / n_side), -1).reshape(1, Nq, 1, 1, 1, 2)
    wh = torch.tensor([[w, hh] for hh, w in shapes], dtype=torch.float32).reshape(1, 1, 1, L, 1, 2)
    loc = ref + torch.randn(B, Nq, h, L, P, 2, generator=g) * sigma_px / wh
    aw = torch.softmax(torch.randn(B, Nq, h, L * P, generator=g), -1).reshape(B, Nq, h, L, P)
    grad_out = torch.randn(B, Nq, h * Dh, generator=g)
    to = dict(device=device)
    return dict(value=value.to(dtype).to(**to), spatial_shapes=shapes, level_start_index=starts,
                sampling_locations=loc.float().to(**to), attention_weights=aw.to(dtype).to(**to),
                grad_out=grad_out.to(dtype).to(**to), query_w=n_side)


def default_camera():
    """Pinhole camera of BASELINE.md config 1: 1.5 m above the IMU looking along +x, f=400 @ 640x512.
    Returns (imu_to_cam 4x4, K 3x4) as float64 numpy arrays (the format of bev_cmr_proj.py:41-49,66-73)."""
    import numpy as np
    T = np.eye(4)
    T[:3, :3] = [[0.0, 0.0, 1.0], [-1.0, 0.0, 0.0], [0.0, -1.0, 0.0]]
    T[2, 3] = 1.5
    K = np.array([[400.0, 0.0, 320.0, 0.0], [0.0, 400.0, 256.0, 0.0], [0.0, 0.0, 1.0, 0.0]])
    return T, K


def bevrender_config(bev=56, layers=6, batch=8, dim=64, heads=2, depth_dim=5, img_hw=(128, 160),
                     ori_hw=(512, 640), kernel_size=3, stride=1):
    """The config dict ``model.bevrender.BEVRender(config, logger, mode)`` reads (bevrender.py:19-73; the released
    ``get_config()`` is empty): BASELINE.json configs 3 / 5 = 1 stage x ``layers`` EncoderLayers
    of constant shape, 1 view, ResNet18 backbone without FPN, all drop rates 0."""
    T, K = default_camera()
    return {
        "BATCH_SIZE": batch, "DATA_TYPE": torch.float32, "DAT_BEV_SHAPE": [bev], "DAT_EMBED_DIMS": [dim],
        "VEHICLE_TYPE_CODE": 0, "IMU_TO_RGB": {0: [T]}, "INTRINSIC_K": {0: [K]},
        "IMG_HEIGHT": img_hw[0], "IMG_WIDTH": img_hw[1], "ORI_IMG_HEIGHT": ori_hw[0], "ORI_IMG_WIDTH": ori_hw[1],
        "REMOVE_REF_IN_GRAY": False, "BOUND_CHECK_IMG_PATH": None,
        "BEV_BOUND": {"X": 20.0, "Y": 10.0, "Z": 2.0}, "DAT_SCALE_OFFSET_RANGE": True, "DAT_NUM_STAGES": 1,
        "NUM_VIEWS": 1, "DAT_EXPANSION": 2, "DAT_BEV_DEPTH_DIM": depth_dim, "SAMPLE_Z_SHIFT": -1.0,
        "DAT_VIT_DEPTHS": [layers], "DAT_NUM_HEADS": [heads], "DAT_STRIDES": [stride], "DAT_NUM_GROUPS": [1],
        "DAT_K_SIZES": [kernel_size], "DAT_DROP_RATE": 0.0, "DAT_ATTN_DROP_RATE": 0.0, "DAT_DROP_PATH_RATE": 0.0,
        "DAT_BACKBONE_TYPE": "ResNet18", "DECODER_HID_DIM": 64,
    }


def make_train_batch(batch, window=4, img_hw=(128, 160), seed=0):
    """Synthetic batch of config 5: camera (B, T, 1, 3, H, W), pose (B, T, 3) = (x, y, heading), vehicle type (B, 1),
    map target (B, 3, 224, 224)."""
    g = torch.Generator().manual_seed(seed)
    cam = torch.rand(batch, window, 1, 3, *img_hw, generator=g)
    pose = torch.cat((torch.cumsum(torch.rand(batch, window, 2, generator=g) * 0.5, 1),
                      (torch.rand(batch, window, 1, generator=g) - 0.5) * 0.2), -1)
    veh = torch.zeros(batch, 1, dtype=torch.long)
    target = torch.rand(batch, 3, 224, 224, generator=g)
    return cam, pose, veh, target

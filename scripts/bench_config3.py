#!/usr/bin/env python
"""BASELINE.json config 3: full encoder (1 stage x 6 reference EncoderLayers with the drop-ins, prev_bev given) fwd+bwd, and
the reference-point projection K4 alone at 80 000 points.  Prints one JSON object."""
import json
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bevrender_b200 as bb
from bevrender_b200 import compat, workloads


def timeit(fn, warm, iters):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


def encoder(N, C, heads, D, B, layers=6, img_hw=(32, 40)):
    compat.add_reference_to_path()
    compat.install(wrappers=True)
    from model.encoder import BEVEncoderStage
    T, K = workloads.default_camera()
    proj = bb.BEV2CameraProjector(imu_to_rgb={0: [T]}, K={0: [K.copy()]}, vehicle_type_code=0, img_width=160, img_height=128,
                                  ori_img_width=640, ori_img_height=512)
    stage = BEVEncoderStage(bev_bound={"X": 20.0, "Y": 10.0, "Z": 2.0}, bev2cmr_projector=proj, batch_size=B,
                            scale_offset_range=True, stage_idx=0, n_views=1, expansion=2, dims=[C], bev_feat_shapes=[N],
                            bev_depth_dim=D, z_shift=-1.0, depth=layers, n_heads=heads, strides=1, n_groups=1, kernel_size=3,
                            drop_rate=0.0, attn_drop_rate=0.0, drop_path_rate=0.0).cuda().train()
    g = torch.Generator().manual_seed(0)
    q = torch.randn(B, C, N, N, generator=g).cuda().requires_grad_(True)
    img = torch.randn(B, C, *img_hw, generator=g).cuda().requires_grad_(True)
    prev = torch.randn(B, C, N, N, generator=g).cuda()
    go = torch.randn(B, C, N, N, generator=g).cuda()

    def run():
        out, _ = stage(bev_query=q, img_tensor=img, prev_bev=prev, vehicle_pose=None, vehicle_type_idx=torch.tensor(0),
                       wandb_log_dict={}, return_wandb_log=False)
        out.backward(go)
    return run


def main():
    res = {}
    run = encoder(56, 64, 2, 5, 8, img_hw=(128, 160))
    res["encoder_n56_b8_fwd_bwd_ms"] = timeit(run, 2, 5)
    res["encoder_n56"] = "1 stage x 6 reference EncoderLayers (encoder.py:131-238) under compat.install(wrappers=True), N=56, C=64, 2 heads, D=5, B=8, image features 128x160"
    torch.cuda.empty_cache()
    run = encoder(200, 256, 8, 4, 1, img_hw=(128, 160))
    t0 = time.perf_counter()
    res["encoder_n200_b1_fwd_bwd_ms"] = timeit(run, 1, 2)
    res["encoder_n200"] = "same, N=200, C=256, 8 heads, D=4, B=1 (per sample); peak memory %.2f GB" % (torch.cuda.max_memory_allocated() / 1e9)
    # K4 alone: 80 000 points (N=200, D=4 pillars)
    pts = torch.cat((torch.rand(3, 100, 200, 4) * 20 - 5, torch.ones(1, 100, 200, 4)), 0).cuda()
    T, K = workloads.default_camera()
    camT = torch.tensor(T).float().inverse()
    Ks = torch.tensor(K).float()
    res["k4_80000_points_ms_incl_host"] = timeit(lambda: bb.project_points(pts, camT, Ks, 160, 128), 3, 20)
    print(json.dumps(res))


if __name__ == "__main__":
    main()

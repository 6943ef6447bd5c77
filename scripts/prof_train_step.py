#!/usr/bin/env python
"""Where does the config-5 training step (bench.py `train` leg: unmodified reference BEVRender + the drop-ins) spend its
GPU time?  torch.profiler over 3 steps, kernels grouped into this repo's kernels (by name) and everything else."""
import collections
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench                                             # noqa: E402
from bevrender_b200 import losses, workloads             # noqa: E402

OURS = ("rpe_", "msda_", "offset_head", "affine_warp", "project_kernel", "corr_", "recall_", "cast_f32")


def main():
    dev = torch.device("cuda", 0)
    B = 8
    torch.manual_seed(0)
    model, label, fwd = bench.build_train_model(B, 56, 6, dev)
    model = model.to(dev).train()
    cam, pose, veh, target = (t.to(dev) for t in workloads.make_train_batch(B, window=4, seed=1))
    params = [p for p in model.parameters()]
    opt = torch.optim.AdamW(params, lr=1e-4)

    def step():
        opt.zero_grad(set_to_none=True)
        out = fwd(cam, pose, veh)
        loss = torch.nn.functional.mse_loss(out, target) + losses.contrastive_loss(out, target) + losses.triplet_loss(out, target)
        loss.backward()
        opt.step()

    for _ in range(3):
        step()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(3):
        step()
    e1.record()
    torch.cuda.synchronize()
    wall = e0.elapsed_time(e1) / 3
    with torch.profiler.profile(activities=[torch.profiler.ProfilerActivity.CUDA, torch.profiler.ProfilerActivity.CPU]) as prof:
        for _ in range(3):
            step()
        torch.cuda.synchronize()
    tot = collections.Counter()
    cnt = collections.Counter()
    for ev in prof.events():
        if ev.device_type == torch.autograd.DeviceType.CUDA:
            tot[ev.name] += ev.device_time if hasattr(ev, "device_time") else ev.cuda_time
            cnt[ev.name] += 1
    gpu_ms = sum(tot.values()) / 3e3
    ours = {k: v for k, v in tot.items() if any(o in k for o in OURS)}
    top = sorted(tot.items(), key=lambda kv: -kv[1])[:25]
    res = {"model": label, "ms_per_step_events": wall, "gpu_busy_ms_per_step": gpu_ms, "kernel_launches_per_step": sum(cnt.values()) / 3,
           "ours_ms_per_step": sum(ours.values()) / 3e3,
           "ours_by_kernel_ms": {k[:70]: round(v / 3e3, 3) for k, v in sorted(ours.items(), key=lambda kv: -kv[1])},
           "top25_ms": [[k[:90], round(v / 3e3, 3), cnt[k] // 3] for k, v in top]}
    print(json.dumps(res, indent=1))


if __name__ == "__main__":
    main()

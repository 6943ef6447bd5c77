#!/usr/bin/env python
"""BASELINE.json config 1 / SURVEY section 8(d): the reference-style CPU path of one TSA + one SCA attention
(N=50, C=64, h=2, D=5, B=1, image features 32x40), timed on the host cores next to the CUDA modules.

The reference checkout does not exist on the GPU box, so the CPU side is the oracle port
(oracle/deform_attn.py: dense attention + grid_sample, the reference's algorithm) -- kind "port".
"""
import json
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bevrender_b200 as bb
from oracle import deform_attn as oda


def cpu_time(fn, warm=1, n=3):
    for _ in range(warm):
        fn()
    ts = []
    for _ in range(n):
        t0 = time.perf_counter()
        fn()
        ts.append(time.perf_counter() - t0)
    return min(ts), sorted(ts)[len(ts) // 2]


def gpu_time(fn, warm=3, n=20):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e-3


def main():
    torch.set_num_threads(os.cpu_count() or 1)
    torch.manual_seed(0)
    B, N, C, h, D, Hi, Wi = 1, 50, 64, 2, 5, 32, 40
    tsa = bb.TSADeformableAttention(N, C, h, 1, 1, 3, True, B, n_views=1)
    sca = bb.SCADeformableAttention(N, D, C, h, 1, 1, 1, True, B, n_views=1)
    p_tsa = {k: v.detach().clone().requires_grad_(v.dtype.is_floating_point) for k, v in tsa.state_dict().items()}
    p_sca = {k: v.detach().clone().requires_grad_(v.dtype.is_floating_point) for k, v in sca.state_dict().items()}
    q = torch.randn(B, C, N, N, requires_grad=True)
    prev = torch.randn(B, C, N, N)
    img = torch.randn(B, 1, C, Hi, Wi, requires_grad=True)
    ref = torch.rand(B, 1, N // 2, N * D, 2) * 2 - 1
    go = torch.randn(B, C, N, N)

    def cpu_tsa():
        oda.tsa_forward(p_tsa, prev, q, h, 1, 3, 1)[0].backward(go)

    def cpu_sca():
        oda.sca_forward(p_sca, img, q, ref, h, 1, D)[0].backward(go)

    res = {"config": "N=50 C=64 h=2 D=5 B=1 feat 32x40, fwd+bwd", "cores": os.cpu_count(),
           "torch_threads": torch.get_num_threads(), "kind": "port"}
    res["cpu_tsa_s_best_median"] = cpu_time(cpu_tsa)
    res["cpu_sca_s_best_median"] = cpu_time(cpu_sca)
    tsa, sca = tsa.cuda(), sca.cuda()
    qc, pc, ic, rc, gc = q.detach().cuda().requires_grad_(True), prev.cuda(), img.detach().cuda().requires_grad_(True), ref.cuda(), go.cuda()
    res["gpu_tsa_s"] = gpu_time(lambda: tsa(pc, qc, None, False)[0].backward(gc))
    res["gpu_sca_s"] = gpu_time(lambda: sca(ic, qc, rc, None, False)[0].backward(gc))
    print(json.dumps(res))


if __name__ == "__main__":
    main()

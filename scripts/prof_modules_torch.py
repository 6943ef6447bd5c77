#!/usr/bin/env python
"""Kernel-level breakdown (torch.profiler) of one TSA and one SCA module fwd+bwd at N=56, B=8."""
import os, sys
import torch
from torch.profiler import profile, ProfilerActivity
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bevrender_b200 as bb

B, N, C, h, D = 8, 56, 64, 2, 5
torch.manual_seed(0)
dev = "cuda"
tsa = bb.TSADeformableAttention(N, C, h, 1, 1, 3, True, B, n_views=1).to(dev)
sca = bb.SCADeformableAttention(N, D, C, h, 1, 1, 1, True, B, n_views=1).to(dev)
q = torch.randn(B, C, N, N, device=dev, requires_grad=True)
prev = torch.randn(B, C, N, N, device=dev)
img = torch.randn(B, 1, C, 128, 160, device=dev, requires_grad=True)
ref = torch.rand(B, 1, N // 2, N * D, 2, device=dev) * 2 - 1
ref[:, :, :, : N * D // 3] = -1.0
go = torch.randn(B, C, N, N, device=dev)
for name, fn in (("TSA", lambda: tsa(prev, q, None, False)[0].backward(go)), ("SCA", lambda: sca(img, q, ref, None, False)[0].backward(go))):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        for _ in range(3): fn()
        torch.cuda.synchronize()
    print("=====", name, "(3 iterations)")
    print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=14, max_name_column_width=70))

#!/usr/bin/env python
"""One TSA and one SCA fwd+bwd at N=56, B=8 (for ncu captures of the K3 kernels)."""
import os
import sys

import torch

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
for _ in range(int(os.environ.get("REPS", 1))):
    tsa(prev, q, {}, False)[0].backward(go)
    sca(img, q, ref, {}, False)[0].backward(go)
torch.cuda.synchronize()
print("ok")

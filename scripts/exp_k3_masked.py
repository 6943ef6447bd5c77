#!/usr/bin/env python
"""Experiment: what masked reference points (a third of the SCA keys near (-1,-1), bev_cmr_proj.py:106-124) cost the K3
backward, with and without the rpe_table gradient, for several amounts of position jitter (in table cells)."""
import os, sys, json
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bevrender_b200 as bb

def t(fn, n=5):
    for _ in range(2): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n

B, h, N, Dh, Hk, Wk, mul = 8, 2, 56, 32, 28, 280, 5
Ns = Hk * Wk
torch.manual_seed(0)
q = torch.randn(B * h, Dh, N * N, device="cuda", requires_grad=True)
k = torch.randn(B * h, Dh, Ns, device="cuda", requires_grad=True)
v = torch.randn(B * h, Dh, Ns, device="cuda", requires_grad=True)
go = torch.randn(B * h, Dh, N * N, device="cuda")
ys = torch.linspace(-1, 1, Hk); xs = torch.linspace(-1, 1, Wk)
base = torch.stack(torch.meshgrid(ys, xs, indexing="ij"), -1).reshape(1, Hk, Wk, 2).repeat(B, 1, 1, 1).cuda()
res = {}
for name, jitter_cells in (("none_masked", None), ("masked_jitter0.05", 0.05), ("masked_jitter0.75", 0.75), ("masked_jitter2.5", 2.5)):
    pos = base + 0.01 * torch.randn_like(base)
    if jitter_cells is not None:
        m = Wk // 3
        jy = jitter_cells / (0.25 * (2 * N - 2)); jx = jitter_cells / (0.25 * (2 * N * mul - 2))
        pos[:, :, :m, 0] = -1.0 + jy * (2 * torch.rand(B, Hk, m, device="cuda") - 1)
        pos[:, :, :m, 1] = -1.0 + jx * (2 * torch.rand(B, Hk, m, device="cuda") - 1)
    pos = pos.reshape(B, Ns, 2)
    for rg in (True, False):
        rpe = (0.1 * torch.randn(h, 2 * N - 1, 2 * N * mul - 1, device="cuda")).requires_grad_(rg)
        p = pos.clone().requires_grad_(True)
        out = bb.rpe_attention(q, k, v, rpe, p, N, N, 1)
        res[f"{name}_bwd_rpegrad{int(rg)}"] = t(lambda: torch.autograd.grad(out, [q, k, v, p] + ([rpe] if rg else []), go, retain_graph=True))
print(json.dumps(res, indent=1))

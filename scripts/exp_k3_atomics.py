#!/usr/bin/env python
"""Experiment: cost of the rpe_table gradient atomics in the K3 key kernel (TSA / SCA-like shapes, N=56, B=8, h=2)."""
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

res = {}
for name, (Hk, Wk, mul) in {"tsa": (56, 56, 1), "sca": (28, 280, 5)}.items():
    B, h, N, Dh = 8, 2, 56, 32
    Ns = Hk * Wk
    torch.manual_seed(0)
    q = torch.randn(B * h, Dh, N * N, device="cuda", requires_grad=True)
    k = torch.randn(B * h, Dh, Ns, device="cuda", requires_grad=True)
    v = torch.randn(B * h, Dh, Ns, device="cuda", requires_grad=True)
    ys = torch.linspace(-1, 1, Hk); xs = torch.linspace(-1, 1, Wk)
    pos = torch.stack(torch.meshgrid(ys, xs, indexing="ij"), -1).reshape(1, Ns, 2).repeat(B, 1, 1).cuda()
    pos = pos + 0.01 * torch.randn_like(pos)
    go = torch.randn(B * h, Dh, N * N, device="cuda")
    for rg, pg in ((True, True), (False, True), (False, False)):
        rpe = (0.1 * torch.randn(h, 2 * N - 1, 2 * N * mul - 1, device="cuda")).requires_grad_(rg)
        p = pos.clone().requires_grad_(pg)
        out = bb.rpe_attention(q, k, v, rpe, p, N, N, 1)
        res[f"{name}_bwd_rpegrad{int(rg)}_posgrad{int(pg)}"] = t(lambda: torch.autograd.grad(out, [q, k, v] + ([rpe] if rg else []) + ([p] if pg else []), go, retain_graph=True))
    res[f"{name}_fwd"] = t(lambda: bb.rpe_attention(q, k, v, rpe, p, N, N, 1))
print(json.dumps(res))

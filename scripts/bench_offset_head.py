#!/usr/bin/env python
"""Times the fused offset head against the PyTorch composition it replaces (TSA and SCA shapes, N=56, B=8)."""
import json, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bevrender_b200 as bb
from oracle import deform_attn as oda      # timing comparison only (scripts/ is not the product path)

def t(fn, n=20):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n

res = {}
for name, (Bg, Cm, H, W, O, inter, fac) in {"tsa": (8, 64, 56, 56, 2, False, 0.5), "sca": (8, 320, 56, 56, 5, True, 5.0)}.items():
    torch.manual_seed(0)
    y = torch.randn(Bg, Cm, H, W, device="cuda", requires_grad=True)
    lw = torch.ones(Cm, device="cuda", requires_grad=True); lb = torch.zeros(Cm, device="cuda", requires_grad=True)
    w1 = (torch.randn(O, Cm, 1, 1, device="cuda") * 0.1).requires_grad_(True)
    Hk, Wk = (H // 2, W * O) if inter else (H, W)
    ref = torch.rand(Bg, Hk, Wk, 2, device="cuda") * 2 - 1
    g = torch.randn(Bg, Hk, Wk, 2, device="cuda")
    def fused():
        bb.offset_head(y, lw, lb, w1, ref, interleave=inter, factor=fac).backward(g)
    def torch_path():
        oda.offset_head(y, lw, lb, w1, ref, inter, True, fac).backward(g)
    res[name] = {"fused_fwd_bwd_ms": t(fused), "pytorch_fwd_bwd_ms": t(torch_path),
                 "fused_fwd_ms": t(lambda: bb.offset_head(y.detach(), lw.detach(), lb.detach(), w1.detach(), ref, interleave=inter, factor=fac))}
print(json.dumps(res))

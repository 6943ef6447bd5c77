#!/usr/bin/env python
"""Experiment: how the K5 kernel time scales with the number of rotations (UMMA N) -- separates the fixed per-MMA
cost / shared-memory operand bandwidth / math-bound regimes."""
import json, math, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bevrender_b200 as bb
m, t = torch.randn(3, 512, 512, device="cuda"), torch.randn(3, 224, 224, device="cuda")
res = {}
for R in (16, 32, 48, 64, 72, 80, 96, 112, 128):
    ang = [2 * math.pi * r / R for r in range(R)]
    for _ in range(2): bb.correlate(m, t, ang)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(3): bb.correlate(m, t, ang)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 3
    res[R] = {"ms": round(ms, 3), "TFLOPs_useful": round(2.0 * 289 * 289 * R * 3 * 224 * 224 / ms / 1e9, 1)}
print(json.dumps(res))

#!/usr/bin/env python
"""Experiment: what the box's PCIe link gives the e2e leg -- 1.25 GB pinned host <-> device, one direction alone and
both directions at once (two streams)."""
import json
import torch
n = 1254400000 // 4
h_a = torch.empty(n, dtype=torch.float32).pin_memory()
h_b = torch.empty(n, dtype=torch.float32).pin_memory()
d_a = torch.empty(n, dtype=torch.float32, device="cuda")
d_b = torch.empty(n, dtype=torch.float32, device="cuda")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
def timed(fn, reps=5):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    for s in (s1, s2): torch.cuda.current_stream().wait_stream(s)
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps
def h2d():
    with torch.cuda.stream(s1): d_a.copy_(h_a, non_blocking=True)
def d2h():
    with torch.cuda.stream(s2): h_b.copy_(d_b, non_blocking=True)
def both():
    s1.wait_stream(torch.cuda.current_stream()); s2.wait_stream(torch.cuda.current_stream())
    h2d(); d2h()
def chunked(k=32):
    s1.wait_stream(torch.cuda.current_stream()); s2.wait_stream(torch.cuda.current_stream())
    m = n // k
    for i in range(k):
        with torch.cuda.stream(s1): d_a[i*m:(i+1)*m].copy_(h_a[i*m:(i+1)*m], non_blocking=True)
        with torch.cuda.stream(s2): h_b[i*m:(i+1)*m].copy_(d_b[i*m:(i+1)*m], non_blocking=True)
gb = n * 4 / 1e9
r = {"GB": gb}
for name, fn in (("h2d_alone", h2d), ("d2h_alone", d2h), ("both", both), ("both_32_chunks", chunked)):
    ms = timed(fn); r[name + "_ms"] = ms; r[name + "_GBps_per_dir"] = gb / ms * 1e3
print(json.dumps(r))

#!/usr/bin/env python
"""BASELINE.json config 4: rotation-sweep correlation, template (C,224,224) vs map (C,512,512), 72 rotations.
Prints ms per call (CUDA events, incl. operand preparation), the main kernel alone, and useful TFLOP/s."""
import json
import math
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bevrender_b200 as bb


def main():
    R, Ht, Hm = 72, 224, 512
    ang = [2 * math.pi * r / R for r in range(R)]
    res = []
    for C in (3, 64):
        torch.manual_seed(0)
        m, t = torch.randn(C, Hm, Hm, device="cuda"), torch.randn(C, Ht, Ht, device="cuda")
        for _ in range(3):
            out = bb.correlate(m, t, ang)
        torch.cuda.synchronize()
        n = 5 if C == 3 else 2
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(n):
            out = bb.correlate(m, t, ang)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / n
        Ho = Hm - Ht + 1
        flops = 2.0 * Ho * Ho * R * C * Ht * Ht
        res.append({"C": C, "ms": ms, "useful_TFLOPs": flops / ms / 1e9, "frac_of_sustained_bf16_peak": flops / ms / 1e9 / 1324.9,
                    "flop": flops})
    if os.environ.get("CPU_BASELINE"):
        from oracle import correlate as oc
        m, t = torch.randn(3, Hm, Hm), torch.randn(3, Ht, Ht)
        t0 = time.perf_counter()
        oc.correlate(m, t, ang[:8])
        res.append({"cpu_oracle_s_for_8_of_72_rotations_C3": time.perf_counter() - t0, "cores": os.cpu_count()})
    print(json.dumps(res))


if __name__ == "__main__":
    main()

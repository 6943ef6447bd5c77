#!/usr/bin/env python
"""Times the drop-in TSA / SCA modules (fwd+bwd, CUDA events) at the reference-native shape of
BASELINE.json config 3(i): N=56, C=64, h=2, D=5, B=8, image features 128x160."""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bevrender_b200 as bb


def timeit(fn, warm=3, iters=10):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


def main():
    env = lambda k, d: int(os.environ.get(k, d))
    bb.SCADeformableAttention.sort_keys = bool(env("SORT_KEYS", 1))
    B, N, C, h, D = env("B", 8), env("N", 56), env("C", 64), env("H", 2), env("D", 5)
    iters = env("ITERS", 10)
    torch.manual_seed(0)
    dev = "cuda"
    tsa = bb.TSADeformableAttention(N, C, h, 1, 1, 3, True, B, n_views=1).to(dev)
    sca = bb.SCADeformableAttention(N, D, C, h, 1, 1, 1, True, B, n_views=1).to(dev)
    q = torch.randn(B, C, N, N, device=dev, requires_grad=True)
    prev = torch.randn(B, C, N, N, device=dev)
    img = torch.randn(B, 1, C, 128, 160, device=dev, requires_grad=True)
    ref = (torch.rand(B, 1, N // 2, N * D, 2, device=dev) * 2 - 1)
    ref[:, :, :, : N * D // 3] = -1.0          # ~1/3 masked reference points, as the projection produces
    go = torch.randn(B, C, N, N, device=dev)

    def run_tsa(bwd=True):
        out, _ = tsa(prev, q, {}, False)
        if bwd:
            out.backward(go)

    def run_sca(bwd=True):
        out, _ = sca(img, q, ref, {}, False)
        if bwd:
            out.backward(go)

    w = 1 if iters < 5 else 3
    res = {"N": N, "B": B, "C": C, "heads": h, "D": D,
           "tsa_fwd_ms": timeit(lambda: run_tsa(False), w, iters), "tsa_fwd_bwd_ms": timeit(run_tsa, w, iters),
           "sca_fwd_ms": timeit(lambda: run_sca(False), w, iters), "sca_fwd_bwd_ms": timeit(run_sca, w, iters),
           "peak_mem_GB": torch.cuda.max_memory_allocated() / 1e9}
    print(json.dumps(res))


if __name__ == "__main__":
    main()

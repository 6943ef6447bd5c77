"""Shared helpers for the test-suite (golden loading, error metric)."""
import os

import numpy as np
import torch

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load_golden(name):
    z = np.load(os.path.join(GOLDEN, name + ".npz"))
    d = {k: z[k] for k in z.files}
    params = {k[len("param."):]: torch.from_numpy(v) for k, v in d.items() if k.startswith("param.")}
    gparams = {k[len("gparam."):]: torch.from_numpy(v) for k, v in d.items() if k.startswith("gparam.")}
    rest = {k: v for k, v in d.items() if not k.startswith(("param.", "gparam."))}
    return rest, params, gparams


def rel_err(a, b):
    """max|a-b| / max|b|  -- the tolerance of BASELINE.json's north_star is stated on this."""
    a = torch.as_tensor(a).detach().double()
    b = torch.as_tensor(b).detach().double()
    denom = b.abs().max().clamp_min(1e-30)
    return float((a - b).abs().max() / denom)


FP32_TOL = 1e-4     # north_star: "within 1e-4 relative in fp32"
BF16_TOL = 1e-2     # north_star: "1e-2 in bf16"

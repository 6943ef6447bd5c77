#!/usr/bin/env python
"""One call of the K5 correlation at BASELINE config 4 (C=3) for ncu captures."""
import math
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bevrender_b200 as bb

torch.manual_seed(0)
m, t = torch.randn(3, 512, 512, device="cuda"), torch.randn(3, 224, 224, device="cuda")
out = bb.correlate(m, t, [2 * math.pi * r / 72 for r in range(72)])
torch.cuda.synchronize()
print("ok", float(out.abs().max()))

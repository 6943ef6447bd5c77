"""GPU test of K5 (tcgen05 rotation-sweep correlation) against the in-repo oracle ("parity unpinned": the
reference has no correlation code)."""
import math

import pytest
import torch

from oracle import correlate as ocorr
from util import BF16_TOL, rel_err

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("C,Hm,Wm,Ht,Wt,R", [(3, 40, 200, 16, 32, 5), (2, 64, 180, 20, 48, 18), (3, 96, 290, 32, 64, 72)])
def test_correlate_matches_oracle(C, Hm, Wm, Ht, Wt, R):
    import bevrender_b200 as bb
    torch.manual_seed(0)
    m, t = torch.randn(C, Hm, Wm), torch.randn(C, Ht, Wt)
    ang = [2 * math.pi * r / R for r in range(R)]
    out = bb.correlate(m.cuda(), t.cuda(), ang).cpu()
    ref_bf16 = ocorr.correlate(m.double(), t.double(), ang, torch.bfloat16)     # same operand rounding, exact sums
    ref_full = ocorr.correlate(m.double(), t.double(), ang)
    assert out.shape == ref_full.shape
    # the kernel rotates in fp32 then rounds to bf16: an occasional 1-ulp(bf16) difference of a template value
    assert rel_err(out, ref_bf16) < 2e-3
    assert rel_err(out, ref_full) < BF16_TOL


def test_correlate_identity_rotation_peak():
    """Planting the template in the map puts the arg-max of the 0-degree response at the planted offset."""
    import bevrender_b200 as bb
    torch.manual_seed(1)
    m, t = 0.1 * torch.randn(3, 100, 260), torch.randn(3, 32, 64)
    m[:, 41:73, 150:214] += t
    out = bb.correlate(m.cuda(), t.cuda(), [0.0, math.pi / 2]).cpu()
    assert divmod(int(out[0].argmax()), out.shape[2]) == (41, 150)

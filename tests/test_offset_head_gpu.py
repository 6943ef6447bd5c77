"""GPU parity tests of the fused offset head ("next" row f-2) against the oracle's restatement of the reference's
LayerNormProxy -> GELU -> 1x1 conv -> tanh * range * factor (or clamp) -> + reference chain."""
import pytest
import torch

from oracle import deform_attn as oda
from util import FP32_TOL, rel_err

pytestmark = pytest.mark.gpu


def _case(Bg, Cm, H, W, O, interleave, seed, ref_batched):
    torch.manual_seed(seed)
    y = torch.randn(Bg, Cm, H, W) * 1.5 + 0.3
    ln_w, ln_b = 1.0 + 0.2 * torch.randn(Cm), 0.1 * torch.randn(Cm)
    w1 = torch.randn(O, Cm, 1, 1) * (2.0 / Cm ** 0.5)       # raw offsets of order 1: tanh and clamp both active
    Hk, Wk = (H // 2, W * O) if interleave else (H, W)
    ref = torch.rand(Bg if ref_batched else 1, Hk, Wk, 2) * 2.4 - 1.2
    gpos = torch.randn(Bg, Hk, Wk, 2)
    return y, ln_w, ln_b, w1, ref, gpos


@pytest.mark.parametrize("Bg,Cm,H,W,O,interleave,scaled,factor,ref_batched", [
    (2, 64, 14, 14, 2, False, True, 0.5, False),     # TSA, C=64, shared grid
    (3, 30, 7, 9, 2, False, True, 0.5, False),       # Cm not a multiple of 4, ragged pixel count
    (2, 64, 14, 14, 2, False, False, 1.0, False),    # TSA, clamp instead of tanh scaling
    (2, 320, 14, 14, 5, True, True, 5.0, True),      # SCA N=14, D=5: Cm = Cg*D, per-sample reference points
    (1, 128, 6, 10, 4, True, False, 1.0, True),      # SCA D=4, clamp
    (1, 35, 8, 5, 7, True, True, 5.0, True),         # odd Cm, 7 depth bins
    (8, 64, 56, 56, 2, False, True, 0.5, False),     # reference-native TSA shape
])
def test_offset_head_matches_oracle(Bg, Cm, H, W, O, interleave, scaled, factor, ref_batched):
    import bevrender_b200 as bb
    y, ln_w, ln_b, w1, ref, gpos = _case(Bg, Cm, H, W, O, interleave, 0, ref_batched)
    cpu = [t.clone().requires_grad_(True) for t in (y, ln_w, ln_b, w1)]
    want = oda.offset_head(*cpu, ref, interleave, scaled, factor)
    want.backward(gpos)
    dev = [t.cuda().requires_grad_(True) for t in (y, ln_w, ln_b, w1)]
    got = bb.offset_head(*dev, ref.cuda(), interleave=interleave, scaled=scaled, factor=factor)
    got.backward(gpos.cuda())
    assert got.shape == want.shape
    assert rel_err(got.cpu(), want) < FP32_TOL
    for name, a, b in zip(("grad_y", "grad_ln_weight", "grad_ln_bias", "grad_w1"), dev, cpu):
        assert rel_err(a.grad.cpu(), b.grad) < FP32_TOL, name


def test_offset_head_empty_batch_and_errors():
    import bevrender_b200 as bb
    y = torch.zeros(0, 8, 4, 4, device="cuda")
    out = bb.offset_head(y, torch.ones(8, device="cuda"), torch.zeros(8, device="cuda"),
                         torch.zeros(2, 8, device="cuda"), torch.zeros(1, 4, 4, 2, device="cuda"))
    assert out.shape == (0, 4, 4, 2)
    with pytest.raises(RuntimeError):                       # CPU tensors: no fallback
        bb.offset_head(torch.zeros(1, 8, 4, 4), torch.ones(8), torch.zeros(8), torch.zeros(2, 8), torch.zeros(1, 4, 4, 2))
    with pytest.raises(bb._lib.BevrError):                  # plain layout needs O == 2
        bb.offset_head(torch.zeros(1, 8, 4, 4, device="cuda"), torch.ones(8, device="cuda"), torch.zeros(8, device="cuda"),
                       torch.zeros(3, 8, device="cuda"), torch.zeros(1, 4, 4, 2, device="cuda"))

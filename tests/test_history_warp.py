"""History-BEV warp: oracle pinned on CPU, CUDA kernel checked on GPU against the golden output of
the reference's EncoderLayer.project_history_bev_feat (encoder.py:413-466)."""
import pytest
import torch

from oracle import history_warp as ohw
from util import FP32_TOL, load_golden, rel_err


def test_oracle_matches_reference():
    g, _, _ = load_golden("history_warp")
    out = ohw.project_history_bev(torch.from_numpy(g["bev"]), torch.from_numpy(g["pose"]))
    assert rel_err(out, g["out"]) < 1e-6


@pytest.mark.gpu
def test_cuda_warp_matches_reference():
    import bevrender_b200 as bb
    g, _, _ = load_golden("history_warp")
    out = bb.project_history_bev(torch.from_numpy(g["bev"]).cuda(), torch.from_numpy(g["pose"]).cuda())
    assert rel_err(out.cpu(), g["out"]) < FP32_TOL


@pytest.mark.gpu
def test_cuda_warp_identity_and_bigger_batch():
    import bevrender_b200 as bb
    torch.manual_seed(0)
    bev = torch.randn(5, 16, 28, 28)
    pose = torch.zeros(5, 2, 3)
    assert rel_err(bb.project_history_bev(bev.cuda(), pose.cuda()).cpu(), bev) < 1e-5   # zero motion: identity up to fp32 rounding of the grid
    pose = torch.randn(5, 2, 3)
    ref = ohw.project_history_bev(bev, pose)
    assert rel_err(bb.project_history_bev(bev.cuda(), pose.cuda()).cpu(), ref) < FP32_TOL

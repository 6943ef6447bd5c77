"""GPU parity of K4 (reference-point projection): bit-exact against the golden vectors made by the
reference's BEV2CameraProjector (bev_cmr_proj.py:61-124)."""
import numpy as np
import pytest
import torch

from util import load_golden

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("name", ["proj_n50", "proj_n200", "proj_yaw", "proj_z0"])
def test_projection_bit_exact(name):
    import bevrender_b200 as bb
    g, _, _ = load_golden(name)
    N, D, W, H, oW, oH = [int(v) for v in g["cfg"]]
    pts = torch.from_numpy(g["points"]).cuda()
    uv, pix, mask = bb.project_points(pts, torch.from_numpy(g["cam_T"]), torch.from_numpy(g["K_scaled"]), W, H)
    assert np.array_equal(mask.cpu().numpy(), g["mask"])
    assert np.array_equal(pix.cpu().numpy()[:, g["mask"]], g["pix"][:, g["mask"]])
    assert np.array_equal(uv.cpu().numpy().view(np.uint32), g["uv"].view(np.uint32))

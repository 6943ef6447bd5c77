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


def test_projector_class_with_gray_reference_image(tmp_path):
    """Class-level drop-in incl. remove_ref_in_gray (bev_cmr_proj.py:114-122): bit-exact against the reference class."""
    from PIL import Image
    import bevrender_b200 as bb
    g, _, _ = load_golden("proj_gray")
    N, D, W, H, oW, oH = [int(v) for v in g["cfg"]]
    path = str(tmp_path / "ref.png")
    Image.fromarray(g["image"]).save(path)
    K = g["K"].copy()
    proj = bb.BEV2CameraProjector({0: [g["imu_to_cam"]]}, {0: [K]}, 0, W, H, oW, oH, remove_ref_in_gray=True,
                                  bound_check_img_paths=[path], device="cuda")
    assert np.array_equal(K, g["K"])                       # the caller's intrinsics are not scaled in place
    uv = proj.bev_grid_to_camera(torch.from_numpy(g["points"]))[0][0]
    assert np.array_equal(proj.last_mask[0].cpu().numpy(), g["mask"])
    assert np.array_equal(uv.cpu().numpy().view(np.uint32), g["uv"].view(np.uint32))

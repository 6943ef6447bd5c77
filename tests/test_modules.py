"""Drop-in modules: state_dict compatibility with the reference (CPU) and fwd/bwd parity against the golden
outputs of the real reference modules (GPU)."""
import pytest
import torch

from util import FP32_TOL, load_golden, rel_err

# The offset nets and 1x1 projections stay PyTorch/cuDNN.  cuDNN convolutions default to TF32
# on sm_100 (~1e-3 relative), which would mask the 1e-4 parity of the kernels under test -> force fp32 convs.
torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False

TSA_CASES = ["tsa_s1", "tsa_s2", "tsa_self", "tsa_clamp"]
SCA_CASES = ["sca_d5", "sca_g2", "sca_clamp"]


def build_tsa(name):
    import bevrender_b200 as bb
    g, p, gp = load_golden(name)
    N, C, h, grp, k, s, B, self_attn, sor = [int(v) for v in g["cfg"]]
    m = bb.TSADeformableAttention(bev_feat_shape=N, dim_embed=C, n_heads=h, n_groups=grp, stride=s, kernel_size=k,
                                  scale_offset_range=bool(sor), batch_size=B, n_views=1)
    m.load_state_dict(p, strict=True)           # same keys and shapes as the reference module
    return m, g, gp, bool(self_attn)


def build_sca(name):
    import bevrender_b200 as bb
    g, p, gp = load_golden(name)
    N, D, C, h, grp, B, Hi, Wi, sor = [int(v) for v in g["cfg"]]
    m = bb.SCADeformableAttention(bev_feat_shape=N, bev_depth_dim=D, dim_embed=C, n_heads=h, n_groups=grp, stride=1,
                                  kernel_size=1, scale_offset_range=bool(sor), batch_size=B, n_views=1)
    m.load_state_dict(p, strict=True)
    return m, g, gp


@pytest.mark.parametrize("name", TSA_CASES)
def test_tsa_state_dict_matches_reference(name):
    m, g, gp, _ = build_tsa(name)
    _, p, _ = load_golden(name)
    assert set(m.state_dict()) == set(p)


@pytest.mark.parametrize("name", SCA_CASES)
def test_sca_state_dict_matches_reference(name):
    m, g, gp = build_sca(name)
    _, p, _ = load_golden(name)
    assert set(m.state_dict()) == set(p)


def test_modules_fail_loudly_without_cuda():
    if torch.cuda.is_available():
        pytest.skip("CUDA present")
    m, g, _, _ = build_tsa("tsa_s1")
    with pytest.raises(RuntimeError, match="CUDA"):
        m(torch.from_numpy(g["x"]), torch.from_numpy(g["query"]), {}, False)


def _check_grads(m, gp):
    # proj_k.bias has a mathematically ZERO gradient (a key bias shifts every logit of a row equally and the
    # softmax cancels it); the reference value is rounding noise ~1e-8, so such tensors are compared on the
    # scale of the largest parameter gradient instead of on their own.
    scale = max(float(v.abs().max()) for v in gp.values())
    for k_, ref in gp.items():
        got = dict(m.named_parameters())[k_].grad
        assert got is not None, k_
        if float(ref.abs().max()) < 1e-5 * scale:
            assert float(got.abs().max()) < 1e-5 * scale, k_
            continue
        e = rel_err(got.cpu(), ref)
        assert e < FP32_TOL, f"grad {k_}: {e:.3e}"
    for k_, p in m.named_parameters():            # parameters the reference never uses stay grad-free
        if k_ not in gp:
            assert p.grad is None, k_


@pytest.mark.gpu
@pytest.mark.parametrize("name", TSA_CASES)
def test_tsa_module_matches_reference_outputs(name):
    m, g, gp, self_attn = build_tsa(name)
    m = m.cuda()
    query = torch.from_numpy(g["query"]).cuda().requires_grad_(True)
    x = None if self_attn else torch.from_numpy(g["x"]).cuda().requires_grad_(True)
    log = {"k": 1}
    out, log2 = m(x, query, log, False)
    assert log2 is log
    out.backward(torch.from_numpy(g["grad_out"]).cuda())
    assert rel_err(out.cpu(), g["out"]) < FP32_TOL
    assert rel_err(query.grad.cpu(), g["g_query"]) < FP32_TOL
    if x is not None:
        assert rel_err(x.grad.cpu(), g["g_x"]) < FP32_TOL
    _check_grads(m, gp)


@pytest.mark.gpu
@pytest.mark.parametrize("name", SCA_CASES)
def test_sca_module_matches_reference_outputs(name):
    m, g, gp = build_sca(name)
    m = m.cuda()
    query = torch.from_numpy(g["query"]).cuda().requires_grad_(True)
    x = torch.from_numpy(g["x"]).cuda().requires_grad_(True)
    ref_pts = torch.from_numpy(g["reference_points"]).cuda()
    out, _ = m(x, query, ref_pts, {}, False)
    out.backward(torch.from_numpy(g["grad_out"]).cuda())
    assert rel_err(out.cpu(), g["out"]) < FP32_TOL
    assert rel_err(query.grad.cpu(), g["g_query"]) < FP32_TOL
    assert rel_err(x.grad.cpu(), g["g_x"]) < FP32_TOL
    _check_grads(m, gp)


@pytest.mark.gpu
def test_projector_dropin_matches_reference():
    import numpy as np
    import bevrender_b200 as bb
    g, _, _ = load_golden("proj_n50")
    N, D, W, H, oW, oH = [int(v) for v in g["cfg"]]
    K = g["K"].copy()
    proj = bb.BEV2CameraProjector(imu_to_rgb={0: [g["imu_to_cam"]]}, K={0: [K]}, vehicle_type_code=0, img_width=W,
                                  img_height=H, ori_img_width=oW, ori_img_height=oH)
    uv = proj.bev_grid_to_camera(torch.from_numpy(g["points"]))[0][0]
    assert np.array_equal(K, g["K"])                               # caller's intrinsics untouched
    assert np.array_equal(uv.cpu().numpy().view(np.uint32), g["uv"].view(np.uint32))
    assert np.array_equal(proj.last_mask[0].cpu().numpy(), g["mask"])


def test_wrapper_points_and_keys_match_reference():
    """wrappers.SpatialCrossAttn.sample_3d_points reproduces the reference's pillar points bit for bit and the wrapper
    state_dict keys equal the reference's (spatial_cross_attn.* / temporal_self_attn.* of the EncoderLayer golden)."""
    import types
    import numpy as np
    import bevrender_b200 as bb
    g, _, _ = load_golden("proj_n50")
    stub = types.SimpleNamespace(bev_bound=dict(zip("XYZ", [float(v) for v in g["bound"]])), bev_feat_shape=50,
                                 bev_depth_dim=5, z_shift=float(g["z_shift"]))
    assert np.array_equal(bb.SpatialCrossAttn.sample_3d_points(stub).numpy(), g["points"])
    _, p, _ = load_golden("encoder_layer")
    tsa = bb.TemporalSelfAttn(14, 64, 2, 1, 1, 3, 2, True, n_views=1)
    want = {k[len("temporal_self_attn."):] for k in p if k.startswith("temporal_self_attn.")}
    assert set(tsa.state_dict()) == want


@pytest.mark.gpu
def test_spatial_cross_attn_wrapper_matches_module():
    import bevrender_b200 as bb
    g, p, _ = load_golden("encoder_layer")
    N, C, h, D, B, Hi, Wi = [int(v) for v in g["cfg"]]
    from bevrender_b200.harness import default_camera
    T, K = default_camera()
    proj = bb.BEV2CameraProjector({0: [T]}, {0: [K]}, 0, Wi * 4, Hi * 4, 640, 512)
    w = bb.SpatialCrossAttn({"X": 20.0, "Y": 10.0, "Z": 2.0}, proj, N, D, -1.0, C, h, 1, 1, 3, B, True, n_views=1)
    w.load_state_dict({k[len("spatial_cross_attn."):]: v for k, v in p.items() if k.startswith("spatial_cross_attn.")})
    w = w.cuda()
    q = torch.from_numpy(g["query"]).cuda()
    img = torch.from_numpy(g["img"]).cuda()
    out, _ = w(q, img, torch.tensor(0), {}, False)
    ref_pts = torch.from_numpy(g["ref_points"]).cuda()
    V, _, hh, ww, dd = ref_pts.shape
    r = ref_pts.permute(0, 2, 3, 4, 1).reshape(1, V, hh, ww * dd, 2).expand(B, -1, -1, -1, -1)
    assert torch.equal(w._reference_points(torch.tensor(0), q.device, B), r)       # K4 == reference projection, bit-exact
    exp, _ = w.spatial_deform_attn(img.reshape(B, 1, C, Hi, Wi), q, r, {}, False)
    assert torch.equal(out, exp)

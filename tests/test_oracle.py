"""CPU tests: pin the oracle against the golden vectors made from the real reference."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

from oracle import bilinear, deform_attn, msda, projection
from bevrender_b200.workloads import make_microbench_inputs
from util import FP32_TOL, load_golden, rel_err


def _run_tsa(name):
    g, p, gp = load_golden(name)
    N, C, h, grp, k, s, B, self_attn, sor = [int(v) for v in g["cfg"]]
    p = {k_: v.clone().requires_grad_(True) for k_, v in p.items()}
    query = torch.from_numpy(g["query"]).requires_grad_(True)
    x = None if self_attn else torch.from_numpy(g["x"]).requires_grad_(True)
    out, _ = deform_attn.tsa_forward(p, x, query, h, grp, k, s, bool(sor))
    out.backward(torch.from_numpy(g["grad_out"]))
    return g, p, gp, out, query, x


@pytest.mark.parametrize("name", ["tsa_s1", "tsa_s2", "tsa_self", "tsa_clamp"])
def test_tsa_oracle_matches_reference(name):
    g, p, gp, out, query, x = _run_tsa(name)
    assert rel_err(out, g["out"]) < 1e-5
    assert rel_err(query.grad, g["g_query"]) < 1e-4
    if x is not None:
        assert rel_err(x.grad, g["g_x"]) < 1e-4
    for k_, ref in gp.items():
        assert rel_err(p[k_].grad, ref) < 1e-4, k_
    # parameters the reference never uses get no gradient
    assert "proj_q.weight" not in gp and "proj_views.weight" not in gp


@pytest.mark.parametrize("name", ["sca_d5", "sca_g2", "sca_clamp"])
def test_sca_oracle_matches_reference(name):
    g, p, gp = load_golden(name)
    N, D, C, h, grp, B, Hi, Wi, sor = [int(v) for v in g["cfg"]]
    p = {k_: v.clone().requires_grad_(True) for k_, v in p.items()}
    query = torch.from_numpy(g["query"]).requires_grad_(True)
    x = torch.from_numpy(g["x"]).requires_grad_(True)
    out, _ = deform_attn.sca_forward(p, x, query, torch.from_numpy(g["reference_points"]), h, grp, D, bool(sor))
    out.backward(torch.from_numpy(g["grad_out"]))
    assert rel_err(out, g["out"]) < 1e-5
    assert rel_err(query.grad, g["g_query"]) < 1e-4
    assert rel_err(x.grad, g["g_x"]) < 1e-4
    for k_, ref in gp.items():
        assert rel_err(p[k_].grad, ref) < 1e-4, k_


@pytest.mark.parametrize("name", ["proj_n50", "proj_n200", "proj_yaw", "proj_z0"])
def test_projection_oracle_bit_exact(name):
    g, _, _ = load_golden(name)
    N, D, W, H, oW, oH = [int(v) for v in g["cfg"]]
    bound = dict(zip("XYZ", [float(v) for v in g["bound"]]))
    pts = projection.sample_3d_points(bound, N, D, float(g["z_shift"]))
    assert np.array_equal(pts.numpy(), g["points"])
    Ks = projection.scale_intrinsics(g["K"], W, H, oW, oH)
    assert np.array_equal(Ks.numpy(), g["K_scaled"])
    uv, pix, mask = projection.project_points(pts, g["imu_to_cam"], Ks, W, H)
    assert np.array_equal(mask.numpy(), g["mask"])
    assert np.array_equal(pix.numpy()[:, g["mask"]], g["pix"][:, g["mask"]])
    assert np.array_equal(uv.numpy(), g["uv"])            # bit-exact incl. masked -> (-1,-1)


def test_rows_equal_dense():
    g, p, _ = load_golden("tsa_s1")
    torch.manual_seed(0)
    B, h, Dh, N = 2, 2, 32, 14
    q, k, v = (torch.randn(B * h, Dh, N * N) for _ in range(3))
    pos = deform_attn.normalized_grid(N, N).reshape(1, -1, 2).repeat(B, 1, 1) + 0.05 * torch.randn(B, N * N, 2)
    dense = deform_attn.rpe_attention_dense(q, k, v, p["rpe_table"], pos, N, N, 1, Dh ** -0.5)
    rows = deform_attn.rpe_attention_rows(q, k, v, p["rpe_table"], pos, N, N, 1, Dh ** -0.5, 30, 75)
    assert torch.equal(dense[:, :, 30:75], rows)


@pytest.mark.parametrize("align", [True, False])
def test_numpy_bilinear_matches_grid_sample(align):
    torch.manual_seed(1)
    inp = torch.randn(2, 3, 7, 9, dtype=torch.float64, requires_grad=True)
    grid = (torch.rand(2, 5, 6, 2, dtype=torch.float64) * 2.6 - 1.3).requires_grad_(True)
    out = F.grid_sample(inp, grid, mode="bilinear", padding_mode="zeros", align_corners=align)
    go = torch.randn_like(out)
    out.backward(go)
    o = bilinear.grid_sample_bilinear(inp.detach().numpy(), grid.detach().numpy(), align)
    gi, gg = bilinear.grid_sample_bilinear_backward(go.numpy(), inp.detach().numpy(), grid.detach().numpy(), align)
    assert np.allclose(o, out.detach().numpy(), atol=1e-12)
    assert np.allclose(gi, inp.grad.numpy(), atol=1e-12)
    assert np.allclose(gg, grid.grad.numpy(), atol=1e-12)


def test_msda_oracle_matches_numpy_restatement():
    """L>1/P>1 MSDA has no reference counterpart ("parity unpinned"): pin the grid_sample-based
    oracle against the independent numpy restatement instead."""
    d = make_microbench_inputs(1, seed=2, n_side=8, h=2, Dh=4, L=2, P=2, sigma_px=2.0)
    v, loc, w = d["value"].double(), d["sampling_locations"].double(), d["attention_weights"].double()
    out = msda.msda_forward(v, d["spatial_shapes"], d["level_start_index"], loc, w)
    B, S, h, Dh = v.shape
    acc = np.zeros((B * h, Dh, loc.shape[1]))
    for l, (H, W) in enumerate(d["spatial_shapes"]):
        s0 = d["level_start_index"][l]
        vl = v[:, s0:s0 + H * W].permute(0, 2, 3, 1).reshape(B * h, Dh, H, W).numpy()
        gl = (2 * loc[:, :, :, l] - 1).permute(0, 2, 1, 3, 4).reshape(B * h, -1, loc.shape[4], 2).numpy()
        wl = w[:, :, :, l].permute(0, 2, 1, 3).reshape(B * h, 1, -1, loc.shape[4]).numpy()
        acc += (bilinear.grid_sample_bilinear(vl, gl, False) * wl).sum(-1)
    ref = torch.from_numpy(acc).reshape(B, h * Dh, -1).transpose(1, 2)
    assert rel_err(out, ref) < 1e-12


def test_projection_oracle_gray_mask_bit_exact():
    """remove_ref_in_gray=True: golden made by the reference class with a reference image on disk."""
    g, _, _ = load_golden("proj_gray")
    N, D, W, H, oW, oH = [int(v) for v in g["cfg"]]
    Ks = projection.scale_intrinsics(g["K"], W, H, oW, oH)
    pts = torch.from_numpy(g["points"])
    uv, _, mask = projection.project_points(pts, g["imu_to_cam"], Ks, W, H, gray_image=g["image"])
    plain = projection.project_points(pts, g["imu_to_cam"], Ks, W, H)[2]
    assert np.array_equal(mask.numpy(), g["mask"])
    assert np.array_equal(uv.numpy().view(np.uint32), g["uv"].view(np.uint32))
    assert int(plain.sum()) > int(mask.sum()) > 0          # the gray band really removes points


def test_dropout_keep_mask_is_a_fair_coin_per_head_and_seed():
    """oracle.dropout_keep (the restatement of K3's dropout hash): keep rate 1 - p, different per head and per seed,
    deterministic in the seed."""
    from oracle import deform_attn as oda
    a = oda.dropout_keep(7, 4, 50, 60, 0.25)
    assert a.shape == (4, 50, 60) and abs(a.float().mean().item() - 0.75) < 0.02
    assert torch.equal(a, oda.dropout_keep(7, 4, 50, 60, 0.25))
    assert not torch.equal(a[0], a[1]) and not torch.equal(a, oda.dropout_keep(8, 4, 50, 60, 0.25))
    assert oda.dropout_keep(7, 1, 8, 8, 0.0).all()


@pytest.mark.parametrize("B,h,Dh,L,P,n", [(2, 4, 8, 3, 2, 12), (1, 8, 32, 4, 4, 10), (1, 2, 16, 1, 1, 9)])
def test_msda_oracle_matches_deformable_detr_definition(B, h, Dh, L, P, n):
    """oracle/msda.py (the checker of K1/K2) against the published pure-PyTorch definition of the operator it names
    (Deformable DETR's ms_deform_attn_core_pytorch, restated in oracle/msda_ddetr.py): output and all three gradients,
    locations partly outside [0, 1] (zero padding)."""
    from oracle import msda as om
    from oracle import msda_ddetr as dd
    g = torch.Generator().manual_seed(5)
    shapes = [(max(n >> l, 2), max((n + 3) >> l, 2)) for l in range(L)]
    starts, s = [], 0
    for H, W in shapes:
        starts.append(s)
        s += H * W
    Nq = 37
    value = torch.randn(B, s, h, Dh, generator=g, dtype=torch.float64)
    loc = torch.rand(B, Nq, h, L, P, 2, generator=g, dtype=torch.float64) * 1.3 - 0.15
    w = torch.softmax(torch.randn(B, Nq, h, L * P, generator=g, dtype=torch.float64), -1).reshape(B, Nq, h, L, P)
    go = torch.randn(B, Nq, h * Dh, generator=g, dtype=torch.float64)
    a = [t.clone().requires_grad_(True) for t in (value, loc, w)]
    b = [t.clone().requires_grad_(True) for t in (value, loc, w)]
    out_a = om.msda_forward(a[0], shapes, starts, a[1], a[2])
    out_b = dd.ms_deform_attn_core(b[0], shapes, b[1], b[2])
    out_a.backward(go)
    out_b.backward(go)
    assert torch.allclose(out_a, out_b, rtol=1e-12, atol=1e-12)
    for x, y in zip(a, b):
        assert torch.allclose(x.grad, y.grad, rtol=1e-10, atol=1e-12)


def test_correlation_oracle_matches_first_principles_loops():
    """oracle/correlate.py DEFINES K5's operator (the reference has no registration code): check the torch formulation
    (grid_sample rotation + conv2d) against plain numpy loops written from the formula in include/bevr.h,
      out[r, y, x] = sum_{c,u,v} map[c, y+u, x+v] * rot_r(tmpl)[c, u, v],
      rot_r samples tmpl bilinearly (zeros outside) at centre + R(angle_r) (v - cx, u - cy)."""
    import math
    import numpy as np
    from oracle import correlate as oc
    rng = np.random.default_rng(3)
    C_, Hm, Wm, Ht, Wt = 2, 9, 11, 4, 6
    m = rng.standard_normal((C_, Hm, Wm))
    t = rng.standard_normal((C_, Ht, Wt))
    angles = [0.0, 0.4, -1.1, math.pi / 2]
    cx, cy = 0.5 * (Wt - 1), 0.5 * (Ht - 1)

    def tap(c, yy, xx):
        return t[c, yy, xx] if 0 <= yy < Ht and 0 <= xx < Wt else 0.0

    want = np.zeros((len(angles), Hm - Ht + 1, Wm - Wt + 1))
    for r, a in enumerate(angles):
        rot = np.zeros_like(t)
        for u in range(Ht):
            for v in range(Wt):
                sx = cx + math.cos(a) * (v - cx) - math.sin(a) * (u - cy)
                sy = cy + math.sin(a) * (v - cx) + math.cos(a) * (u - cy)
                x0, y0 = math.floor(sx), math.floor(sy)
                fx, fy = sx - x0, sy - y0
                for c in range(C_):
                    rot[c, u, v] = ((1 - fy) * ((1 - fx) * tap(c, y0, x0) + fx * tap(c, y0, x0 + 1)) +
                                    fy * ((1 - fx) * tap(c, y0 + 1, x0) + fx * tap(c, y0 + 1, x0 + 1)))
        for y in range(want.shape[1]):
            for x in range(want.shape[2]):
                want[r, y, x] = np.sum(m[:, y:y + Ht, x:x + Wt] * rot)
    got = oc.correlate(torch.from_numpy(m), torch.from_numpy(t), angles).numpy()
    assert np.allclose(got, want, rtol=1e-10, atol=1e-10)

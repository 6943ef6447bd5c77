"""GPU parity tests of K1/K2 (MSDA gather fwd/bwd) through the C ABI against the oracle."""
import pytest
import torch
import torch.nn.functional as F

from oracle import msda as omsda
from bevrender_b200.workloads import make_microbench_inputs
from util import BF16_TOL, FP32_TOL, load_golden, rel_err

pytestmark = pytest.mark.gpu

BWD_DIRECT = False      # set per test by the fixture below: every case runs on both backward kernels


@pytest.fixture(params=[False, True], ids=["sorted", "direct"], autouse=True)
def backward_kernel(request):
    """K2 has two kernels: the tile-sorted backward (default for Dh <= 128) and the one-RED-per-tap backward
    (BEVR_BWD_DIRECT; also the sorted kernel's per-CTA fallback and the only one for Dh > 128)."""
    global BWD_DIRECT
    BWD_DIRECT = request.param
    yield request.param
    BWD_DIRECT = False


def _run_cuda(d, dtype, align, pm1, need_value=True):
    import bevrender_b200 as bb
    dev = "cuda"
    value = d["value"].to(dev).to(dtype).requires_grad_(need_value)
    loc = d["sampling_locations"].to(dev).requires_grad_(True)
    w = d["attention_weights"]
    w = None if w is None else w.to(dev).to(dtype).requires_grad_(True)
    out = bb.ms_deform_attn(value, d["spatial_shapes"], d["level_start_index"], loc, w,
                            query_w=d.get("query_w", 0), align_corners=align, loc_pm1=pm1, bwd_direct=BWD_DIRECT)
    out.backward(d["grad_out"].to(dev).to(dtype))
    return out, value.grad, loc.grad, (None if w is None else w.grad)


def _check(d, dtype, align=False, pm1=False, tol=None):
    tol = tol or (FP32_TOL if dtype == torch.float32 else BF16_TOL)
    out, gv, gl, gw = _run_cuda(d, dtype, align, pm1)
    # oracle in fp32 on the (possibly bf16-rounded) inputs the kernel saw
    cast = lambda t: None if t is None else t.to(dtype).float()
    ref = omsda.msda_forward_backward(cast(d["value"]), d["spatial_shapes"], d["level_start_index"],
                                      d["sampling_locations"], cast(d["attention_weights"]), cast(d["grad_out"]),
                                      align, pm1)
    names = ("out", "grad_value", "grad_loc", "grad_weight")
    for n, a, b in zip(names, (out, gv, gl, gw), ref):
        if b is None:
            assert a is None
            continue
        e = rel_err(a.float().cpu(), b)
        assert e < tol, f"{n}: rel err {e:.3e} >= {tol}"


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
@pytest.mark.parametrize("align,pm1", [(False, False), (True, False), (True, True), (False, True)])
def test_msda_small_all_coordinate_modes(dtype, align, pm1):
    d = make_microbench_inputs(2, seed=1, n_side=24, h=4, Dh=32, L=3, P=4, sigma_px=2.0)
    if pm1:
        d["sampling_locations"] = d["sampling_locations"] * 2 - 1
    _check(d, dtype, align, pm1)


@pytest.mark.parametrize("h,Dh,L,P", [(1, 4, 1, 1), (2, 8, 2, 3), (3, 16, 1, 5), (8, 32, 4, 4), (2, 64, 2, 2),
                                       (1, 128, 1, 1), (1, 256, 1, 2), (5, 48, 2, 1)])
def test_msda_shapes_fp32(h, Dh, L, P):
    d = make_microbench_inputs(1, seed=3, n_side=16, h=h, Dh=Dh, L=L, P=P, sigma_px=3.0)
    _check(d, torch.float32)


@pytest.mark.parametrize("h,Dh,L,P", [(2, 8, 2, 3), (8, 32, 4, 4), (2, 64, 2, 2), (1, 128, 1, 1)])
def test_msda_shapes_bf16(h, Dh, L, P):
    d = make_microbench_inputs(1, seed=4, n_side=16, h=h, Dh=Dh, L=L, P=P, sigma_px=3.0)
    _check(d, torch.bfloat16)


def test_msda_flat_queries_ragged_and_out_of_range():
    """Flat query list (query_w = 0), Nq not a multiple of the tile, locations far outside the map,
    NaN location: out-of-range taps contribute zero, as grid_sample(padding_mode='zeros')."""
    d = make_microbench_inputs(1, seed=5, n_side=10, h=2, Dh=8, L=2, P=2, sigma_px=1.0)
    d["query_w"] = 0
    Nq = 77
    for k in ("sampling_locations", "attention_weights", "grad_out"):
        d[k] = d[k][:, :Nq].contiguous()
    d["sampling_locations"][0, :5] = 7.5          # far outside
    d["sampling_locations"][0, 5:9] = -3.0
    d["sampling_locations"][0, 9, 0, 0, 0, 0] = 1.0     # exactly on the border
    _check(d, torch.float32)
    import bevrender_b200 as bb
    loc = d["sampling_locations"].clone()
    loc[0, 11] = float("nan")
    out = bb.ms_deform_attn(d["value"].cuda(), d["spatial_shapes"], d["level_start_index"], loc.cuda(),
                            d["attention_weights"].cuda())
    assert torch.isfinite(out).all() and out[0, 11].abs().max() == 0


def test_reference_gather_l1p1_matches_grid_sample():
    """The reference's own use: grid_sample(align_corners=True) on an NCHW map with (y,x) positions."""
    import bevrender_b200 as bb
    torch.manual_seed(0)
    x = torch.randn(3, 64, 14, 18, requires_grad=True)
    pos = (torch.rand(3, 7, 30, 2) * 2.4 - 1.2).requires_grad_(True)
    ref = F.grid_sample(x, pos.flip(-1), mode="bilinear", align_corners=True).reshape(3, 64, -1)
    g = torch.randn_like(ref)
    ref.backward(g)
    xc = x.detach().cuda().requires_grad_(True)
    pc = pos.detach().cuda().requires_grad_(True)
    out = bb.grid_gather(xc, pc)
    out.backward(g.cuda())
    assert rel_err(out.cpu(), ref) < FP32_TOL
    assert rel_err(xc.grad.cpu(), x.grad) < FP32_TOL
    assert rel_err(pc.grad.cpu(), pos.grad) < FP32_TOL


def test_reference_points_plus_offsets_fused():
    import bevrender_b200 as bb
    d = make_microbench_inputs(1, seed=6, n_side=12, h=2, Dh=16, L=2, P=2)
    B, Nq, h, L, P, _ = d["sampling_locations"].shape
    ref_pts = torch.rand(B, Nq, L, 2)
    off = torch.randn(B, Nq, h, L, P, 2) * 2.0
    loc = omsda.locations_from_offsets(ref_pts, off, d["spatial_shapes"], True)
    exp = omsda.msda_forward_backward(d["value"], d["spatial_shapes"], d["level_start_index"], loc,
                                      d["attention_weights"], d["grad_out"])
    offc = off.cuda().requires_grad_(True)
    refc = ref_pts.cuda().requires_grad_(True)
    out = bb.ms_deform_attn(d["value"].cuda(), d["spatial_shapes"], d["level_start_index"], offc,
                            d["attention_weights"].cuda(), reference_points=refc, query_w=12, offset_in_pixels=True,
                            bwd_direct=BWD_DIRECT)
    out.backward(d["grad_out"].cuda())
    assert rel_err(out.cpu(), exp[0]) < FP32_TOL
    wh = torch.tensor([[w, hh] for hh, w in d["spatial_shapes"]]).reshape(1, 1, 1, L, 1, 2)
    assert rel_err(offc.grad.cpu(), exp[2] / wh) < FP32_TOL
    assert rel_err(refc.grad.cpu(), exp[2].sum(dim=(2, 4))) < FP32_TOL


def test_full_size_properties():
    """BASELINE.json config 2 shape (200x200 queries, 4 levels, 8 heads, 4 points, B=1):
    size-independent properties instead of an oracle run -- linearity in value, weights summing to
    one reproduce a constant map, and sum(grad_value) == sum over queries of weighted grad_out."""
    import bevrender_b200 as bb
    d = make_microbench_inputs(1, seed=0)
    dev = "cuda"
    v = d["value"].to(dev)
    loc, w = d["sampling_locations"].to(dev), d["attention_weights"].to(dev)
    args = (d["spatial_shapes"], d["level_start_index"])
    f = lambda val: bb.ms_deform_attn(val, *args, loc, w, query_w=200, bwd_direct=BWD_DIRECT)
    o1, o2 = f(v), f(2.5 * v)
    assert rel_err(o2.cpu(), 2.5 * o1.cpu()) < 1e-6
    # interior queries (all taps in range) of a constant map return the constant
    ones = torch.ones_like(v)
    oc = f(ones).reshape(200, 200, -1)[90:110, 90:110]   # >5 sigma from every level border
    assert (oc - 1).abs().max() < 1e-5
    vg = v.clone().requires_grad_(True)
    go = d["grad_out"].to(dev)
    f(vg).backward(go)
    lhs = vg.grad.double().sum()
    rhs = (f(ones).double() * go.double()).sum()       # <grad_value, 1> == <grad_out, f(1)>
    assert abs(float(lhs - rhs)) / abs(float(rhs)) < 1e-5
    # subset parity against the oracle on a 16-row band of queries
    band = slice(96 * 200, 112 * 200)
    ref = omsda.msda_forward(d["value"], *args, d["sampling_locations"][:, band], d["attention_weights"][:, band])
    assert rel_err(o1[:, band].cpu(), ref) < FP32_TOL


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16], ids=["fp32", "bf16"])
def test_full_size_backward_vs_oracle(dtype):
    """BASELINE.json config 2 at full size (B = 1): out, grad_value, grad_loc and grad_weight of the whole sample
    against the oracle (evaluated with CUDA torch on the box -- the same grid_sample construction, fp32), plus three
    query bands checked element-wise.  bf16: storage dtype bf16, oracle in fp32 on the rounded inputs."""
    from util import assert_close
    d = make_microbench_inputs(1, seed=2)
    tol = FP32_TOL if dtype == torch.float32 else BF16_TOL
    out, gv, gl, gw = _run_cuda(d, dtype, False, False)
    cast = lambda t: t.to(dtype).float().cuda()
    ref = omsda.msda_forward_backward(cast(d["value"]), d["spatial_shapes"], d["level_start_index"],
                                      d["sampling_locations"].cuda(), cast(d["attention_weights"]), cast(d["grad_out"]))
    # d out / d location is piecewise constant: it jumps by O(1) where a pixel coordinate crosses an integer, and the
    # kernel's (fp32, u*W - 0.5) and torch's (fp32, ((2u-1+1)*W - 1)/2) coordinates differ by ~1 ulp.  Of the 5.1 M
    # points of a sample a few hundred sit within 1e-4 px of such a jump: exclude exactly those from grad_loc.
    loc64 = d["sampling_locations"].double()
    wh = torch.tensor([[w, hh] for hh, w in d["spatial_shapes"]], dtype=torch.float64).reshape(1, 1, 1, -1, 1, 2)
    pix = loc64 * wh - 0.5
    near = ((pix - pix.round()).abs() < 1e-4).any(-1).cuda()                   # (B, Nq, h, L, P)
    assert float(near.float().mean()) < 1e-3
    gl_ref = ref[2].masked_fill(near[..., None], 0.0)
    gl = gl.masked_fill(near[..., None], 0.0)
    for n, a, b in zip(("out", "grad_value", "grad_loc", "grad_weight"), (out, gv, gl, gw), (ref[0], ref[1], gl_ref, ref[3])):
        e = rel_err(a.float(), b)
        assert e < tol, f"{n}: rel err {e:.3e} >= {tol}"
    if dtype == torch.float32:
        for r0 in (0, 97, 184):                       # first rows (level borders), middle, last rows
            band = slice(r0 * 200, (r0 + 16) * 200)
            assert_close(out[:, band], ref[0][:, band], tol, f"out rows {r0}")
            assert_close(gl[:, band], gl_ref[:, band], tol, f"grad_loc rows {r0}")
            assert_close(gw[:, band], ref[3][:, band], tol, f"grad_weight rows {r0}")
        assert_close(gv, ref[1], tol, "grad_value")


def test_host_pipeline_matches_device_path():
    """The pinned-host entry point (chunked H2D / compute / D2H pipeline) returns what the device path returns."""
    import bevrender_b200 as bb
    from bevrender_b200.host_pipeline import HostMsdaPipeline
    d = make_microbench_inputs(3, seed=7, n_side=20, h=4, Dh=32, L=2, P=4)
    out, gv, gl, gw = _run_cuda(d, torch.float32, False, False)
    S = sum(h * w for h, w in d["spatial_shapes"])
    pipe = HostMsdaPipeline(d["spatial_shapes"], d["level_start_index"], 4, 32, 400, 4, 20)
    h_in = {k: d[k].contiguous().pin_memory() for k in ("value", "sampling_locations", "attention_weights", "grad_out")}
    h_out = {"out": torch.empty(3, 400, 128).pin_memory(), "grad_value": torch.empty(3, S, 4, 32).pin_memory(),
             "grad_loc": torch.empty(3, 400, 4, 2, 4, 2).pin_memory(), "grad_weight": torch.empty(3, 400, 4, 2, 4).pin_memory()}
    pipe.run(h_in, h_out)
    torch.cuda.synchronize()
    assert torch.equal(h_out["out"], out.detach().cpu())
    assert rel_err(h_out["grad_value"], gv.cpu()) < 1e-6          # atomics: order-dependent rounding only
    # the pipeline always runs the default (sorted) backward: bit-equal to the sorted device path only
    assert rel_err(h_out["grad_loc"], gl.cpu()) < 1e-5
    assert rel_err(h_out["grad_weight"], gw.cpu()) < 1e-5


def test_host_pipeline_back_to_back_calls_without_join():
    """run(join=False) x3 + join(): consecutive calls overlap (the next call uploads while the previous one's last results go
    down; staging slots are handed over with events) and every call's results are complete after the join -- different
    inputs and output buffers per call, compared with joined runs of the same inputs."""
    from bevrender_b200.host_pipeline import HostMsdaPipeline
    ds = [make_microbench_inputs(3, seed=s, n_side=20, h=4, Dh=32, L=2, P=4) for s in (1, 2, 3)]
    S = sum(h * w for h, w in ds[0]["spatial_shapes"])
    pipe = HostMsdaPipeline(ds[0]["spatial_shapes"], ds[0]["level_start_index"], 4, 32, 400, 4, 20)
    names = ("value", "sampling_locations", "attention_weights", "grad_out")
    mk_out = lambda: {"out": torch.zeros(3, 400, 128).pin_memory(), "grad_value": torch.zeros(3, S, 4, 32).pin_memory(),
                      "grad_loc": torch.zeros(3, 400, 4, 2, 4, 2).pin_memory(),
                      "grad_weight": torch.zeros(3, 400, 4, 2, 4).pin_memory()}
    h_ins = [{k: d[k].contiguous().pin_memory() for k in names} for d in ds]
    want = []
    for h_in in h_ins:
        o = mk_out()
        pipe.run(h_in, o)
        torch.cuda.synchronize()
        want.append(o)
    got = [mk_out() for _ in ds]
    for h_in, o in zip(h_ins, got):
        pipe.run(h_in, o, join=False)
    pipe.join()
    torch.cuda.synchronize()
    for w, g in zip(want, got):
        assert torch.equal(g["out"], w["out"])
        assert torch.equal(g["grad_loc"], w["grad_loc"]) and torch.equal(g["grad_weight"], w["grad_weight"])
        assert rel_err(g["grad_value"], w["grad_value"]) < 1e-6          # atomics: order-dependent rounding only


def test_empty_query_list():
    """Nq = 0: forward returns an empty tensor, backward a zero grad_value (no kernel launch)."""
    import bevrender_b200 as bb
    v = torch.randn(2, 16, 2, 8, device="cuda", requires_grad=True)
    loc = torch.zeros(2, 0, 2, 1, 3, 2, device="cuda")
    w = torch.zeros(2, 0, 2, 1, 3, device="cuda")
    out = bb.ms_deform_attn(v, [(4, 4)], [0], loc, w)
    assert out.shape == (2, 0, 16)
    out.sum().backward()
    assert v.grad is not None and float(v.grad.abs().max()) == 0.0


def test_sorted_backward_falls_back_when_windows_do_not_fit():
    """Offsets spread over the whole map: the bounding boxes of a query tile exceed the histogram, every CTA takes the
    direct path inside the sorted kernel; results still match the oracle."""
    d = make_microbench_inputs(1, seed=8, n_side=64, h=2, Dh=32, L=2, P=4, sigma_px=40.0)
    _check(d, torch.float32)


def test_backward_without_grad_value_and_hot_pixel():
    """TSA trains with a detached prev_bev (grad_value = NULL); SCA's masked reference points all sample pixel (0, 0)
    (a third of the queries here): the sorted backward adds them up in registers and issues ONE reduction."""
    d = make_microbench_inputs(1, seed=9, n_side=32, h=1, Dh=64, L=1, P=1, sigma_px=1.0)
    loc = d["sampling_locations"]
    loc.reshape(-1, 2)[torch.randperm(1024)[:340]] = 0.0           # align_corners pixel (0, 0) exactly
    d["attention_weights"] = None
    _check(d, torch.float32, align=True, pm1=False)
    out, gv, gl, gw = _run_cuda(d, torch.float32, True, False, need_value=False)
    assert gv is None and gw is None
    ref = omsda.msda_forward_backward(d["value"], d["spatial_shapes"], d["level_start_index"], loc, None, d["grad_out"], True)
    assert rel_err(gl.cpu(), ref[2]) < FP32_TOL


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16], ids=["fp32", "bf16"])
def test_backward_accumulate_flag(dtype):
    """BEVR_ACCUMULATE through the raw C ABI: a second backward call adds into grad_value instead of overwriting it
    (fp32: the reductions land on the caller's contents; bf16: the final cast adds them), for both kernels."""
    import ctypes as C
    from bevrender_b200 import _lib
    d = make_microbench_inputs(1, seed=11, n_side=16, h=2, Dh=32, L=2, P=4)
    dev = "cuda"
    v, w, go = (d[k].to(dev).to(dtype).contiguous() for k in ("value", "attention_weights", "grad_out"))
    loc = d["sampling_locations"].to(dev).contiguous()
    B, S, h, Dh = v.shape
    Nq, L, P = loc.shape[1], loc.shape[3], loc.shape[4]
    hw = (C.c_int32 * (2 * L))(*[x for s in d["spatial_shapes"] for x in s])
    st = (C.c_int32 * L)(*d["level_start_index"])
    dt = 0 if dtype == torch.float32 else 1
    Lb = _lib.lib()
    ws = torch.empty(max(Lb.bevr_msda_bwd_workspace_bytes(B, S, h, Dh, dt), 16), dtype=torch.uint8, device=dev)
    P_ = lambda t: C.c_void_p(t.data_ptr())
    gv = torch.empty_like(v)
    flags0 = 16 if BWD_DIRECT else 0
    for flags in (flags0, flags0 | 8):
        _lib.check(Lb.bevr_msda_bwd(P_(v), hw, st, P_(loc), None, P_(w), P_(go), P_(gv), None, None, P_(ws), B, S, h, Dh,
                                    Nq, L, P, 16, dt, flags, None), "bevr_msda_bwd")
        if flags == flags0:
            once = gv.clone()
    torch.cuda.synchronize()
    assert float(once.abs().max()) > 0
    assert rel_err(gv.float(), 2 * once.float()) < (1e-6 if dt == 0 else 1e-2)


def test_bf16_with_fp32_main_gradient_flag():
    """BEVR_GRAD_VALUE_F32: bf16 storage, grad_value delivered in fp32 straight from the reductions (no workspace, no
    cast); equals the bf16-output path before its final rounding."""
    import ctypes as C
    from bevrender_b200 import _lib
    d = make_microbench_inputs(1, seed=12, n_side=16, h=2, Dh=32, L=2, P=4)
    dev = "cuda"
    v, w, go = (d[k].to(dev).to(torch.bfloat16).contiguous() for k in ("value", "attention_weights", "grad_out"))
    loc = d["sampling_locations"].to(dev).contiguous()
    B, S, h, Dh = v.shape
    Nq, L, P = loc.shape[1], loc.shape[3], loc.shape[4]
    hw = (C.c_int32 * (2 * L))(*[x for s in d["spatial_shapes"] for x in s])
    st = (C.c_int32 * L)(*d["level_start_index"])
    P_ = lambda t: C.c_void_p(t.data_ptr())
    gv32 = torch.full((B, S, h, Dh), float("nan"), dtype=torch.float32, device=dev)
    flags = (16 if BWD_DIRECT else 0) | 32
    _lib.check(_lib.lib().bevr_msda_bwd(P_(v), hw, st, P_(loc), None, P_(w), P_(go), P_(gv32), None, None, None, B, S, h, Dh,
                                        Nq, L, P, 16, 1, flags, None), "bevr_msda_bwd")
    ref = omsda.msda_forward_backward(v.float().cpu(), d["spatial_shapes"], d["level_start_index"], d["sampling_locations"],
                                      w.float().cpu(), go.float().cpu())
    assert rel_err(gv32.cpu(), ref[1]) < FP32_TOL          # fp32 accumulation of bf16-rounded inputs: fp32-grade

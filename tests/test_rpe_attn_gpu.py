"""GPU parity tests of K3 (fused attention with sampled relative-position bias) against the dense oracle."""
import pytest
import torch

from oracle import deform_attn as oda
from util import FP32_TOL, rel_err

pytestmark = pytest.mark.gpu


FAMILY = 0


@pytest.fixture(params=[1, 2], ids=["ffma2", "mma"], autouse=True)
def kernel_family(request):
    """Every case runs on both K3 kernel families (the default picks by problem size: small grids would only ever
    exercise the FFMA2 kernels, the reference-native and larger ones only the mma.sync 3xTF32 kernels); the family is
    a per-call flag of the C ABI (BEVR_K3_FFMA / BEVR_K3_MMA)."""
    global FAMILY
    FAMILY = request.param
    yield request.param
    FAMILY = 0


def _case(B, h, g, Hq, Wq, Hk, Wk, Rw_mul, seed, spread, Dh=32):
    torch.manual_seed(seed)
    Nq, Ns = Hq * Wq, Hk * Wk
    q = torch.randn(B * h, Dh, Nq)
    k = torch.randn(B * h, Dh, Ns)
    v = torch.randn(B * h, Dh, Ns)
    rpe = torch.randn(h, 2 * Hq - 1, 2 * Wq * Rw_mul - 1) * 0.5
    pos = oda.normalized_grid(Hk, Wk).reshape(1, Ns, 2).repeat(B * g, 1, 1) + spread * torch.randn(B * g, Ns, 2)
    go = torch.randn(B * h, Dh, Nq)
    return q, k, v, rpe, pos, go


@pytest.mark.parametrize("B,h,g,Hq,Wq,Hk,Wk,mul,spread", [
    (2, 2, 1, 14, 14, 14, 14, 1, 0.05),     # TSA stride 1
    (1, 4, 1, 16, 16, 8, 8, 1, 0.1),        # TSA strided keys
    (1, 2, 1, 14, 14, 7, 70, 5, 0.3),       # SCA: keys (H/2, W*D), table 2*W*D-1 wide, many out of range
    (2, 4, 2, 12, 12, 6, 48, 4, 0.2),       # two groups
    (1, 1, 1, 13, 11, 9, 5, 1, 1.5),        # odd sizes, keys far outside -> zero-padded bias
    (1, 2, 2, 50, 50, 10, 13, 1, 0.1),      # Nq not a multiple of the 128-row tile
])
def test_rpe_attention_matches_dense_oracle(B, h, g, Hq, Wq, Hk, Wk, mul, spread):
    import bevrender_b200 as bb
    q, k, v, rpe, pos, go = _case(B, h, g, Hq, Wq, Hk, Wk, mul, 0, spread)
    cpu = [t.clone().requires_grad_(True) for t in (q, k, v, rpe, pos)]
    ref = oda.rpe_attention_dense(*cpu, Hq, Wq, g, 32 ** -0.5)
    ref.backward(go)
    dev = [t.cuda().requires_grad_(True) for t in (q, k, v, rpe, pos)]
    out = bb.rpe_attention(*dev, Hq, Wq, g, family=FAMILY)
    out.backward(go.cuda())
    assert rel_err(out.cpu(), ref) < FP32_TOL
    for name, a, b in zip(("q", "k", "v", "rpe", "pos"), dev, cpu):
        e = rel_err(a.grad.cpu(), b.grad)
        assert e < FP32_TOL, f"grad_{name}: {e:.3e}"


def test_rpe_attention_dh16():
    import bevrender_b200 as bb
    q, k, v, rpe, pos, go = _case(2, 4, 1, 12, 12, 12, 12, 1, 3, 0.1, Dh=16)
    cpu = [t.clone().requires_grad_(True) for t in (q, k, v, rpe, pos)]
    ref = oda.rpe_attention_dense(*cpu, 12, 12, 1, 16 ** -0.5)
    ref.backward(go)
    dev = [t.cuda().requires_grad_(True) for t in (q, k, v, rpe, pos)]
    out = bb.rpe_attention(*dev, 12, 12, 1, family=FAMILY)
    out.backward(go.cuda())
    assert rel_err(out.cpu(), ref) < FP32_TOL
    for name, a, b in zip(("q", "k", "v", "rpe", "pos"), dev, cpu):
        assert rel_err(a.grad.cpu(), b.grad) < FP32_TOL, name


def test_rpe_attention_large_rows_vs_chunked_oracle():
    """N = 100 (Nq = Ns = 10^4): the dense matrix would be 0.4 GB/head; compare sampled row blocks."""
    import bevrender_b200 as bb
    B, h, N = 1, 2, 100
    q, k, v, rpe, pos, _ = _case(B, h, 1, N, N, N, N, 1, 1, 0.01)
    out = bb.rpe_attention(q.cuda(), k.cuda(), v.cuda(), rpe.cuda(), pos.cuda(), N, N, 1, family=FAMILY).cpu()
    for r0 in (0, 4321, 9990):
        ref = oda.rpe_attention_rows(q, k, v, rpe, pos, N, N, 1, 32 ** -0.5, r0, r0 + 10)
        assert rel_err(out[:, :, r0:r0 + 10], ref) < FP32_TOL
    # rows of softmax sum to one: attention of constant values returns the constant
    ones = torch.ones_like(v)
    oc = bb.rpe_attention(q.cuda(), k.cuda(), ones.cuda(), rpe.cuda(), pos.cuda(), N, N, 1, family=FAMILY)
    assert (oc - 1).abs().max() < 1e-5


def test_rpe_attention_masked_reference_points():
    """SCA with masked reference points (bev_cmr_proj.py:106-124): a third of the keys sit at (-1, -1) plus a
    sub-cell offset, so whole warps of the key kernel hit the same four table cells (the pre-reduced atomic path)."""
    import bevrender_b200 as bb
    B, h, Hq, Wq, Hk, Wk = 1, 2, 14, 14, 7, 70
    q, k, v, rpe, pos, go = _case(B, h, 1, Hq, Wq, Hk, Wk, 5, 5, 0.05)
    Ns = Hk * Wk
    pos[:, : Ns // 3] = -1.0 + 1e-3 * torch.randn(B, Ns // 3, 2)
    cpu = [t.clone().requires_grad_(True) for t in (q, k, v, rpe, pos)]
    ref = oda.rpe_attention_dense(*cpu, Hq, Wq, 1, 32 ** -0.5)
    ref.backward(go)
    dev = [t.cuda().requires_grad_(True) for t in (q, k, v, rpe, pos)]
    out = bb.rpe_attention(*dev, Hq, Wq, 1, family=FAMILY)
    out.backward(go.cuda())
    assert rel_err(out.cpu(), ref) < FP32_TOL
    for name, a, b in zip(("q", "k", "v", "rpe", "pos"), dev, cpu):
        assert rel_err(a.grad.cpu(), b.grad) < FP32_TOL, name


def test_rpe_attention_n100_backward_vs_dense_oracle_on_device():
    """N = 100 (Nq = Ns = 10^4, split == 1 paths): all five gradients against the dense oracle, which at this size
    (0.8 GB attention matrix per tensor) is evaluated with CUDA torch on the box (fp32 library math, TF32 off)."""
    import bevrender_b200 as bb
    from util import assert_close
    old = torch.backends.cuda.matmul.allow_tf32
    torch.backends.cuda.matmul.allow_tf32 = False
    try:
        B, h, N = 1, 2, 100
        q, k, v, rpe, pos, go = _case(B, h, 1, N, N, N, N, 1, 11, 0.01)
        ref_in = [t.cuda().requires_grad_(True) for t in (q, k, v, rpe, pos)]
        ref = oda.rpe_attention_dense(*ref_in, N, N, 1, 32 ** -0.5)
        ref.backward(go.cuda())
        dev = [t.cuda().requires_grad_(True) for t in (q, k, v, rpe, pos)]
        out = bb.rpe_attention(*dev, N, N, 1, family=FAMILY)
        out.backward(go.cuda())
        assert_close(out, ref.detach(), FP32_TOL, "out")
        for name, a, b in zip(("q", "k", "v", "rpe", "pos"), dev, ref_in):
            assert_close(a.grad, b.grad, FP32_TOL, "grad_" + name)
    finally:
        torch.backends.cuda.matmul.allow_tf32 = old


@pytest.mark.parametrize("B,h,g,Hq,Wq,Hk,Wk,mul,spread,p", [
    (2, 2, 1, 14, 14, 14, 14, 1, 0.05, 0.1),    # TSA
    (1, 2, 1, 14, 14, 7, 70, 5, 0.3, 0.25),     # SCA shape
    (1, 2, 2, 50, 50, 10, 13, 1, 0.1, 0.5),     # ragged row tile, work split over key ranges
])
def test_attention_dropout_matches_dense_oracle_under_the_same_mask(B, h, g, Hq, Wq, Hk, Wk, mul, spread, p):
    """attn_drop_rate > 0 (TSA_deform_attn.py:313-323): output and all five gradients against the dense reference math
    with the keep mask of the kernels' hash (the backward kernels regenerate it); the keep rate is 1 - p."""
    import bevrender_b200 as bb
    q, k, v, rpe, pos, go = _case(B, h, g, Hq, Wq, Hk, Wk, mul, 3, spread)
    seed = 0x1234_5678_9abc_def1
    keep = oda.dropout_keep(seed, B * h, Hq * Wq, Hk * Wk, p)
    assert abs(keep.float().mean().item() - (1 - p)) < 0.01
    cpu = [t.clone().requires_grad_(True) for t in (q, k, v, rpe, pos)]
    ref = oda.rpe_attention_dense(*cpu, Hq, Wq, g, 32 ** -0.5, drop_p=p, drop_keep=keep)
    ref.backward(go)
    dev = [t.cuda().requires_grad_(True) for t in (q, k, v, rpe, pos)]
    out = bb.rpe_attention(*dev, Hq, Wq, g, family=FAMILY, dropout_p=p, seed=seed)
    out.backward(go.cuda())
    assert rel_err(out.cpu(), ref) < FP32_TOL
    for name, a, b in zip(("q", "k", "v", "rpe", "pos"), dev, cpu):
        e = rel_err(a.grad.cpu(), b.grad)
        assert e < FP32_TOL, f"grad_{name}: {e:.3e}"
    # a different seed gives a different mask; p = 0 is the plain path bit for bit
    out2 = bb.rpe_attention(*[t.detach() for t in dev], Hq, Wq, g, family=FAMILY, dropout_p=p, seed=seed + 1)
    assert not torch.equal(out2, out.detach())
    a0 = bb.rpe_attention(*[t.detach() for t in dev], Hq, Wq, g, family=FAMILY, dropout_p=0.0)
    a1 = bb.rpe_attention(*[t.detach() for t in dev], Hq, Wq, g, family=FAMILY)
    assert torch.equal(a0, a1)


def test_module_attention_dropout_train_vs_eval():
    """The drop-in modules accept attn_drop_rate > 0 like the reference's (TSA.py:30-44): active in training mode (seeded
    from torch's generator, reproducible under torch.manual_seed), inactive in eval mode."""
    import bevrender_b200 as bb
    torch.manual_seed(0)
    N, C = 14, 64
    tsa = bb.TSADeformableAttention(N, C, 2, 1, 1, 3, True, 2, n_views=1, attn_drop_rate=0.2).cuda()
    ref = bb.TSADeformableAttention(N, C, 2, 1, 1, 3, True, 2, n_views=1).cuda()
    ref.load_state_dict(tsa.state_dict())
    q, prev = torch.randn(2, C, N, N, device="cuda", requires_grad=True), torch.randn(2, C, N, N, device="cuda")
    tsa.eval(); ref.eval()
    assert torch.equal(tsa(prev, q, {}, True)[0], ref(prev, q, {}, True)[0])
    tsa.train()
    torch.manual_seed(5); a = tsa(prev, q, {}, True)[0]
    torch.manual_seed(5); b = tsa(prev, q, {}, True)[0]
    torch.manual_seed(6); c = tsa(prev, q, {}, True)[0]
    assert torch.equal(a, b) and not torch.equal(a, c)
    a.sum().backward()
    assert torch.isfinite(q.grad).all()


@pytest.mark.parametrize("sort", [False, True], ids=["any-order", "sorted"])
def test_run_merge_of_coincident_keys_is_exact(sort):
    """BEVR_K3_RUN_MERGE (backward, mma family): keys that address the same bias-table cell -- SCA's masked reference
    points at (-1, -1) plus small offsets -- are summed in the warp before the table-gradient atomics.  The merge must be
    exact for ANY key order (it only merges neighbouring key rows with equal cells); ordered keys are the case it serves."""
    import bevrender_b200 as bb
    if FAMILY != 2:
        pytest.skip("the run merge lives in the mma.sync dK/dV kernel")
    B, h, g, Hq, Wq, Hk, Wk, mul = 1, 2, 1, 28, 28, 14, 140, 5
    q, k, v, rpe, pos, go = _case(B, h, g, Hq, Wq, Hk, Wk, mul, 7, 0.05)
    Ns = Hk * Wk
    masked = torch.rand(B * g, Ns) < 0.4                                   # scattered through the key list
    pos = torch.where(masked[..., None], -1.0 + 0.004 * torch.randn(B * g, Ns, 2), pos)
    if sort:
        Rh, Rw = rpe.shape[1:]
        key = torch.floor((1.0 - pos[..., 0]) * (0.25 * (Rh - 1))) * 8192.0 - pos[..., 1] * (0.25 * (Rw - 1))
        perm = torch.argsort(key, dim=1)
        pos = torch.gather(pos, 1, perm[..., None].expand(-1, -1, 2))
    cpu = [t.clone().requires_grad_(True) for t in (q, k, v, rpe, pos)]
    ref = oda.rpe_attention_dense(*cpu, Hq, Wq, g, 32 ** -0.5)
    ref.backward(go)
    dev = [t.cuda().requires_grad_(True) for t in (q, k, v, rpe, pos)]
    out = bb.rpe_attention(*dev, Hq, Wq, g, family=FAMILY, merge_runs=True)
    out.backward(go.cuda())
    assert rel_err(out.cpu(), ref) < FP32_TOL
    for name, a, b in zip(("q", "k", "v", "rpe", "pos"), dev, cpu):
        e = rel_err(a.grad.cpu(), b.grad)
        assert e < FP32_TOL, f"grad_{name}: {e:.3e}"

"""CPU check of the numerical claim behind the K3 mma.sync kernels (csrc/rpe_attn.cu): a 32-term dot product
evaluated as hi*hi + hi*lo + lo*hi with hi = tf32(x) (round to nearest, 10 explicit mantissa bits), lo = x - hi
truncated to tf32 by the tensor core, and fp32 accumulation, is fp32-grade (error ~1e-6 of |q||k|), whereas a single
TF32 product is not (~5e-4) -- the reason the split is needed for the 1e-4 parity budget of north_star."""
import numpy as np


def tf32_round(x):
    """cvt.rna.tf32.f32: round to nearest (ties away) on the 13 dropped mantissa bits."""
    u = np.asarray(x, dtype=np.float32).view(np.uint32).astype(np.uint64)
    u = (u + 0x1000) & 0xFFFFE000
    return u.astype(np.uint32).view(np.float32)


def tf32_trunc(x):
    """what the tensor core does with the low mantissa bits of an fp32 register fed as a tf32 operand"""
    u = np.asarray(x, dtype=np.float32).view(np.uint32) & np.uint32(0xFFFFE000)
    return u.view(np.float32)


def test_three_product_split_is_fp32_grade():
    rng = np.random.default_rng(0)
    q = rng.standard_normal((4096, 32)).astype(np.float32)
    k = rng.standard_normal((4096, 32)).astype(np.float32)
    exact = (q.astype(np.float64) * k.astype(np.float64)).sum(1)
    qh, kh = tf32_round(q), tf32_round(k)
    ql, kl = tf32_trunc(q - qh), tf32_trunc(k - kh)
    single = (qh.astype(np.float64) * kh).sum(1)
    split = np.zeros(len(q), dtype=np.float32)
    for c in range(32):                                    # fp32 accumulation, three products per term
        for a, b in ((qh, kh), (qh, kl), (ql, kh)):
            split = (split + (a[:, c].astype(np.float64) * b[:, c]).astype(np.float32)).astype(np.float32)
    scale = np.linalg.norm(q, axis=1) * np.linalg.norm(k, axis=1)
    err_single = np.abs(single - exact) / scale
    err_split = np.abs(split - exact) / scale
    assert err_single.max() > 5e-5                         # one TF32 product eats the whole 1e-4 budget on the logits
    assert err_split.max() < 2e-6                          # the split leaves two orders of magnitude of margin
    plain = np.zeros(len(q), dtype=np.float32)
    for c in range(32):
        plain = (plain + q[:, c] * k[:, c]).astype(np.float32)
    assert err_split.max() < 4 * max(np.abs(plain - exact).max() / scale.min(), 5e-7)   # same class as an fp32 FMA chain

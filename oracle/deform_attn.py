"""Oracle for the DAT-style deformable attention of the TSA / SCA layers.

Functional CPU restatement (parameters passed as a ``state_dict``-keyed dict)
of ``TSADeformableAttention.forward`` (TSA_deform_attn.py:128-337) and
``SCADeformableAttention.forward`` (SCA_deform_attn.py:180-421, n_views == 1;
the reference itself fails for more views).  It materialises the
dense (Nq, Ns) attention matrix exactly like the reference, so it is only
usable at small BEV sizes; ``rpe_attention_rows`` evaluates a block of query
rows for larger shapes (row chunking is exact: every row is independent).

Pinned by ``tests/golden/tsa_*.npz`` / ``sca_*.npz`` (outputs and gradients
of the real reference modules, made by ``tests/golden/make_golden.py``).
"""
import torch
import torch.nn.functional as F


def normalized_grid(H, W, dtype=torch.float32, device="cpu"):
    """(H, W, 2) in (y, x): arange -> /(size-1) -> *2 -> -1   (TSA_deform_attn.py:98-109).

    Always evaluated with torch's CPU kernels (IEEE division) and then moved to ``device``: the oracle is the CPU
    restatement, and torch's CUDA kernel for a scalar divisor multiplies by the reciprocal instead, which is one ulp
    off for some indices -- enough to move a bilinear sample across a table cell (tests/util.py, assert_close)."""
    ys = torch.arange(0, H, dtype=dtype).div(H - 1.0).mul(2.0).sub(1.0)
    xs = torch.arange(0, W, dtype=dtype).div(W - 1.0).mul(2.0).sub(1.0)
    gy, gx = torch.meshgrid(ys, xs, indexing="ij")
    return torch.stack((gy, gx), -1).to(device)


def offset_head(y, ln_w, ln_b, w1, ref, interleave, scaled, factor):
    """Tail of the offset network after its depthwise convolution, up to the sampling positions:
    LayerNorm over channels -> GELU -> 1x1 (no bias) -> [rearrange] -> tanh * range * factor (or clamp) -> + ref.
    y (B*g, Cm, H, W); w1 (O, Cm, 1, 1); ref broadcastable to (B*g, Hk, Wk, 2) in (y, x).
    TSA_deform_attn.py:54-68,158-196; SCA_deform_attn.py:56-77,219-277; model_utils.py:51-59."""
    z = y.permute(0, 2, 3, 1)
    z = F.layer_norm(z, (z.shape[-1],), ln_w, ln_b)
    off = F.conv2d(F.gelu(z.permute(0, 3, 1, 2)), w1.reshape(w1.shape[0], -1, 1, 1))
    if interleave:   # "(b g) d (h n) w -> (b g) n h (w d)", n=2: even query rows -> y, odd rows -> x
        Bg, D, Hq, Wq = off.shape
        off = off.reshape(Bg, D, Hq // 2, 2, Wq).permute(0, 3, 2, 4, 1).reshape(Bg, 2, Hq // 2, Wq * D)
    Hk, Wk = off.shape[-2:]
    if scaled:
        rng = torch.tensor([1.0 / (Hk - 1.0), 1.0 / (Wk - 1.0)], device=y.device).reshape(1, 2, 1, 1)
        off = off.tanh().mul(rng).mul(factor)
    pos = off.permute(0, 2, 3, 1) + ref
    return pos if scaled else pos.clamp(-1.0, 1.0)


def offset_positions(p, prefix, x, ref, kernel_size, stride, interleave, scaled, factor):
    """depthwise conv (TSA_deform_attn.py:55-63, SCA_deform_attn.py:57-66) -> offset_head."""
    pad = kernel_size // 2 if kernel_size != stride else 0
    y = F.conv2d(x, p[prefix + ".0.weight"], p[prefix + ".0.bias"], stride=stride, padding=pad, groups=x.shape[1])
    return offset_head(y, p[prefix + ".1.norm.weight"], p[prefix + ".1.norm.bias"], p[prefix + ".3.weight"], ref,
                       interleave, scaled, factor)


def rpe_bias_dense(rpe_table, pos, Hq, Wq, n_groups):
    """Sampled relative-position bias, (B*h, Nq, Ns).
    pos (B*g, Ns, 2) in (y, x).  TSA_deform_attn.py:264-305, SCA_deform_attn.py:360-394."""
    Bg, Ns, _ = pos.shape
    h = rpe_table.shape[0]
    n = h // n_groups
    qg = normalized_grid(Hq, Wq, pos.dtype, pos.device).reshape(1, Hq * Wq, 1, 2)
    disp = (qg - pos[:, None]).mul(0.5)                            # (B*g, Nq, Ns, 2) (y,x)
    table = rpe_table.reshape(1, n_groups, n, *rpe_table.shape[1:]).expand(Bg // n_groups, -1, -1, -1, -1)
    table = table.reshape(Bg, n, *rpe_table.shape[1:])
    bias = F.grid_sample(table, disp.flip(-1), mode="bilinear", align_corners=True)
    return bias.reshape(Bg * n, Hq * Wq, Ns)


def _mix32(x):
    import numpy as np
    x = x.astype(np.uint32)
    x ^= x >> np.uint32(16); x *= np.uint32(0x85ebca6b); x ^= x >> np.uint32(13); x *= np.uint32(0xc2b2ae35)
    x ^= x >> np.uint32(16)
    return x


def dropout_keep(seed, BH, Nq, Ns, p):
    """Keep mask (BH, Nq, Ns) bool of K3's fused attention dropout: restatement of the counter-based hash of
    csrc/rpe_attn.cu (``drop_key`` / ``drop_mul``).  The reference applies nn.Dropout(attn_drop_rate) to the softmax
    matrix (TSA_deform_attn.py:313-323, SCA_deform_attn.py:402-410), i.e. torch's own Philox stream over a matrix K3
    never forms: that stream cannot be reproduced, so for p > 0 parity means "same result as the dense reference math
    under THIS mask" plus the keep rate -- parity unpinned vs reference for the random stream itself."""
    import numpy as np
    with np.errstate(over="ignore"):
        lo, hi = np.uint32(seed & 0xffffffff), np.uint32(seed >> 32)
        bh = np.arange(BH, dtype=np.uint32)
        key = _mix32(lo + bh * np.uint32(0x9e3779b9)) ^ hi                       # (BH,)
        idx = (np.arange(Nq, dtype=np.uint32)[:, None] * np.uint32(Ns) + np.arange(Ns, dtype=np.uint32)[None, :])
        x = _mix32(idx[None] ^ key[:, None, None])
    thresh = np.uint32(np.float32(p) * np.float32(16777216.0))
    return torch.from_numpy((x >> np.uint32(8)) >= thresh)


def rpe_attention_dense(q, k, v, rpe_table, pos, Hq, Wq, n_groups, scale, drop_p=0.0, drop_keep=None):
    """softmax_j(q.k*scale + bias) . v ; q (B*h,Dh,Nq), k/v (B*h,Dh,Ns) -> (B*h,Dh,Nq).
    TSA_deform_attn.py:245-330, SCA_deform_attn.py:331-411.  ``drop_keep`` (B*h, Nq, Ns) bool: the attention dropout
    of TSA_deform_attn.py:313-323 with an explicit keep mask (kept probabilities scaled by 1 / (1 - drop_p))."""
    attn = torch.einsum("bcm,bcn->bmn", q, k).mul(scale)
    attn = attn + rpe_bias_dense(rpe_table, pos, Hq, Wq, n_groups)
    attn = F.softmax(attn, dim=2)
    if drop_keep is not None:
        attn = attn * drop_keep.to(attn) * (1.0 / (1.0 - float(torch.tensor(drop_p, dtype=torch.float32))))
    return torch.einsum("bmn,bcn->bcm", attn, v)


def rpe_attention_rows(q, k, v, rpe_table, pos, Hq, Wq, n_groups, scale, row0, row1):
    """Same as rpe_attention_dense but only for query indices [row0,row1)."""
    Bg, Ns, _ = pos.shape
    h = rpe_table.shape[0]
    n = h // n_groups
    qg = normalized_grid(Hq, Wq, pos.dtype, pos.device).reshape(1, Hq * Wq, 1, 2)[:, row0:row1]
    disp = (qg - pos[:, None]).mul(0.5)
    table = rpe_table.reshape(1, n_groups, n, *rpe_table.shape[1:]).expand(Bg // n_groups, -1, -1, -1, -1)
    table = table.reshape(Bg, n, *rpe_table.shape[1:])
    bias = F.grid_sample(table, disp.flip(-1), mode="bilinear", align_corners=True)
    bias = bias.reshape(Bg * n, row1 - row0, Ns)
    attn = torch.einsum("bcm,bcn->bmn", q[:, :, row0:row1], k).mul(scale) + bias
    return torch.einsum("bmn,bcn->bcm", F.softmax(attn, dim=2), v)


def _attend(p, x_sampled, query, pos, n_heads, n_groups, Hq, Wq):
    B, C = query.shape[:2]
    Dh = C // n_heads
    Ns = x_sampled.shape[-1]
    q = query.reshape(B * n_heads, Dh, Hq * Wq)                     # proj_q is never applied
    k = F.conv2d(x_sampled, p["proj_k.weight"], p["proj_k.bias"]).reshape(B * n_heads, Dh, Ns)
    v = F.conv2d(x_sampled, p["proj_v.weight"], p["proj_v.bias"]).reshape(B * n_heads, Dh, Ns)
    out = rpe_attention_dense(q, k, v, p["rpe_table"], pos, Hq, Wq, n_groups, Dh ** -0.5)
    return out.reshape(B, C, Hq, Wq)


def tsa_forward(p, x, query, n_heads, n_groups, kernel_size, stride, scale_offset_range=True):
    """TSADeformableAttention.forward; returns (out, pos) with pos (B*g,Hk,Wk,2) in (y,x)."""
    if x is None:
        x = query.clone()
    B, C, H, W = x.shape
    Cg = C // n_groups
    pad = kernel_size // 2 if kernel_size != stride else 0
    Hk, Wk = (H + 2 * pad - kernel_size) // stride + 1, (W + 2 * pad - kernel_size) // stride + 1
    pos = offset_positions(p, "conv_offset", query.reshape(B * n_groups, Cg, H, W), normalized_grid(Hk, Wk)[None],
                           kernel_size, stride, False, scale_offset_range, 0.5)
    xs = F.grid_sample(x.reshape(B * n_groups, Cg, H, W), pos.flip(-1), mode="bilinear",
                       align_corners=True).reshape(B, C, 1, Hk * Wk)
    out = _attend(p, xs, query, pos.reshape(B * n_groups, Hk * Wk, 2), n_heads, n_groups, H, W)
    return F.conv2d(out, p["proj_out.weight"], p["proj_out.bias"]), pos


def sca_forward(p, x, query, reference_points, n_heads, n_groups, bev_depth_dim,
                scale_offset_range=True):
    """SCADeformableAttention.forward for one view.
    x (B,1,C,Hi,Wi); reference_points (B,1,Hq/2,Wq*D,2) in (x,y)."""
    B, C, Hq, Wq = query.shape
    D = bev_depth_dim
    Cg = C // n_groups
    Hi, Wi = x.shape[-2:]
    ref = reference_points[:, 0].flip(-1).repeat_interleave(n_groups, 0)          # (y,x)
    pos = offset_positions(p, "conv_offset_m0", query.reshape(B * n_groups, Cg, Hq, Wq), ref, 1, 1, True,
                           scale_offset_range, 5.0)
    Hk, Wk = pos.shape[1:3]
    xs = F.grid_sample(x[:, 0].reshape(B * n_groups, Cg, Hi, Wi), pos.flip(-1), mode="bilinear",
                       align_corners=True).reshape(B, C, 1, Hk * Wk)
    out = _attend(p, xs, query, pos.reshape(B * n_groups, Hk * Wk, 2), n_heads, n_groups, Hq, Wq)
    return F.conv2d(out, p["proj_out.weight"], p["proj_out.bias"]), pos

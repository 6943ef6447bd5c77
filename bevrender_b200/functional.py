"""Host-side operators of the hot path: thin torch.autograd wrappers over the C ABI.

Every operator requires CUDA tensors and raises otherwise -- there is no CPU
fallback (the CPU restatement lives in ``oracle/`` and is test-only).
"""
import ctypes as C
import functools

import torch

from . import _lib

F32, BF16 = 0, 1
ALIGN_CORNERS, LOC_PM1, OFFSET_PIXELS, ACCUMULATE, BWD_DIRECT = 1, 2, 4, 8, 16
_DTYPES = {torch.float32: F32, torch.bfloat16: BF16}


def _ptr(t):
    return None if t is None else C.c_void_p(t.data_ptr())


def _stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def _need_cuda(*tensors):
    for t in tensors:
        if t is not None and not t.is_cuda:
            raise RuntimeError("bevrender_b200 operators run on CUDA tensors only (no CPU fallback)")


@functools.lru_cache(maxsize=256)
def _levels(shapes, starts):
    hw = (C.c_int32 * (2 * len(shapes)))(*[int(v) for s in shapes for v in s])
    st = (C.c_int32 * len(starts))(*[int(v) for v in starts])
    return hw, st


def _as_tuple(x, pair=False):
    if torch.is_tensor(x):
        x = x.tolist()            # a CUDA tensor syncs here: pass CPU tensors / lists instead
    return tuple(tuple(int(v) for v in r) for r in x) if pair else tuple(int(v) for v in x)


class _MSDA(torch.autograd.Function):
    @staticmethod
    def forward(ctx, value, loc, weight, ref, shapes, starts, query_w, flags):
        _need_cuda(value, loc, weight, ref)
        B, S, h, Dh = value.shape
        _, Nq, _, L, P, _ = loc.shape
        dt = _DTYPES[value.dtype]
        value, loc = value.contiguous(), loc.contiguous().float()
        weight = None if weight is None else weight.contiguous().to(value.dtype)
        ref = None if ref is None else ref.contiguous().float()
        out = torch.empty(B, Nq, h * Dh, dtype=value.dtype, device=value.device)
        hw, st = _levels(shapes, starts)
        _lib.check(_lib.lib().bevr_msda_fwd(_ptr(value), hw, st, _ptr(loc), _ptr(ref), _ptr(weight), _ptr(out),
                                            B, S, h, Dh, Nq, L, P, query_w, dt, flags, _stream()), "bevr_msda_fwd")
        ctx.save_for_backward(value, loc, weight, ref)
        ctx.meta = (shapes, starts, query_w, flags, dt)
        return out

    @staticmethod
    def backward(ctx, grad_out):
        value, loc, weight, ref = ctx.saved_tensors
        shapes, starts, query_w, flags, dt = ctx.meta
        B, S, h, Dh = value.shape
        _, Nq, _, L, P, _ = loc.shape
        need_v, need_l, need_w, need_r = ctx.needs_input_grad[:4]
        grad_out = grad_out.contiguous().to(value.dtype)
        gv = torch.empty_like(value) if need_v else None
        gl = torch.empty_like(loc) if (need_l or need_r) else None
        gw = torch.empty_like(weight) if (need_w and weight is not None) else None
        ws = None
        nws = _lib.lib().bevr_msda_bwd_workspace_bytes(B, S, h, Dh, dt)
        if need_v and nws:
            ws = torch.empty(nws, dtype=torch.uint8, device=value.device)
        hw, st = _levels(shapes, starts)
        _lib.check(_lib.lib().bevr_msda_bwd(_ptr(value), hw, st, _ptr(loc), _ptr(ref), _ptr(weight), _ptr(grad_out),
                                            _ptr(gv), _ptr(gl), _ptr(gw), _ptr(ws), B, S, h, Dh, Nq, L, P,
                                            query_w, dt, flags, _stream()), "bevr_msda_bwd")
        gr = None
        if need_r:                      # loc = ref + off * s  ->  d/d ref = sum over heads and points
            gr = gl.sum(dim=(2, 4))
        if gl is not None and ref is not None and (flags & OFFSET_PIXELS):
            wh = torch.tensor([[w, hh] for hh, w in shapes], dtype=gl.dtype, device=gl.device)
            gl = gl / wh[None, None, None, :, None, :]
        return gv, (gl if need_l else None), gw, gr, None, None, None, None


def ms_deform_attn(value, spatial_shapes, level_start_index, sampling_locations, attention_weights=None,
                   reference_points=None, query_w=0, align_corners=False, loc_pm1=False,
                   offset_in_pixels=False, bwd_direct=False):
    """Multi-scale deformable attention gather (kernels K1/K2).

    value (B,S,h,Dh) fp32|bf16 channels-last; sampling_locations (B,Nq,h,L,P,2) fp32 (x,y) --
    or sampling OFFSETS when ``reference_points`` (B,Nq,L,2) is given (loc = ref + off, with
    off / (W_l,H_l) if ``offset_in_pixels``); attention_weights (B,Nq,h,L,P) or None (== 1).
    Returns (B, Nq, h*Dh).  ``query_w``: width of the 2-D query grid (enables 2-D tiling).
    ``bwd_direct``: backward with one RED per tap instead of the tile-sorted kernel (BEVR_BWD_DIRECT).
    """
    flags = (ALIGN_CORNERS if align_corners else 0) | (LOC_PM1 if loc_pm1 else 0) | \
            (OFFSET_PIXELS if offset_in_pixels else 0) | (BWD_DIRECT if bwd_direct else 0)
    shapes = _as_tuple(spatial_shapes, pair=True)
    starts = _as_tuple(level_start_index)
    return _MSDA.apply(value, sampling_locations, attention_weights, reference_points, shapes, starts,
                       int(query_w), flags)


def grid_gather(x, pos_yx):
    """The reference's own gather: ``F.grid_sample(x, pos[..., (1,0)], bilinear, align_corners=True)``
    (TSA_deform_attn.py:210-217, SCA_deform_attn.py:290-301) as MSDA with L = P = 1 and unit weights.

    x (N, C, H, W) NCHW; pos_yx (N, Hk, Wk, 2) in (y, x) -> (N, C, Hk*Wk).
    """
    N, Cc, H, W = x.shape
    Hk, Wk = pos_yx.shape[1:3]
    value = x.permute(0, 2, 3, 1).reshape(N, H * W, 1, Cc)
    loc = pos_yx.flip(-1).reshape(N, Hk * Wk, 1, 1, 1, 2)
    out = ms_deform_attn(value, ((H, W),), (0,), loc, None, query_w=Wk, align_corners=True, loc_pm1=True)
    return out.transpose(1, 2)


def project_points(points, cam_T, K, img_w, img_h):
    """K4: points (4, ...) fp32 CUDA; cam_T = inverse(imu_to_cam) 4x4 and K 3x3 as CPU tensors.
    Returns uv (2, ...) in [-1,1] (masked -> -1), pix (2, ...) int32, mask (...) bool."""
    _need_cuda(points)
    shp = points.shape[1:]
    p = points.reshape(4, -1).contiguous().float()
    n = p.shape[1]
    uv = torch.empty(2, n, dtype=torch.float32, device=p.device)
    pix = torch.empty(2, n, dtype=torch.int32, device=p.device)
    mask = torch.empty(n, dtype=torch.uint8, device=p.device)
    T = (C.c_float * 16)(*torch.as_tensor(cam_T, dtype=torch.float32).reshape(-1).tolist())
    Kc = (C.c_float * 9)(*torch.as_tensor(K, dtype=torch.float32)[:, :3].reshape(-1).tolist())
    _lib.check(_lib.lib().bevr_project_points(_ptr(p), T, Kc, _ptr(uv), _ptr(pix), _ptr(mask), n,
                                              int(img_w), int(img_h), _stream()), "bevr_project_points")
    return uv.reshape(2, *shp), pix.reshape(2, *shp), mask.bool().reshape(*shp)


class _RpeAttention(torch.autograd.Function):
    @staticmethod
    def forward(ctx, q, k, v, rpe, pos, Hq, Wq, hpg, scale, flags, drop_p, seed):
        _need_cuda(q, k, v, rpe, pos)
        q, k, v = q.contiguous().float(), k.contiguous().float(), v.contiguous().float()
        rpe, pos = rpe.contiguous().float(), pos.contiguous().float()
        BH, Dh, Nq = q.shape
        Ns = k.shape[2]
        h, Rh, Rw = rpe.shape
        out = torch.empty_like(q)
        lse = torch.empty(BH, Nq, dtype=torch.float32, device=q.device)
        ws = torch.empty(_lib.lib().bevr_rpe_attn_workspace_bytes(BH, h, Dh, Nq, Ns, Rh, Rw, 0), dtype=torch.uint8, device=q.device)
        _lib.check(_lib.lib().bevr_rpe_attn_fwd_dropout(_ptr(q), _ptr(k), _ptr(v), _ptr(rpe), _ptr(pos), _ptr(out), _ptr(lse),
                                                        _ptr(ws), BH, h, hpg, Dh, Hq, Wq, Ns, Rh, Rw, float(scale), flags,
                                                        float(drop_p), int(seed), _stream()),
                   "bevr_rpe_attn_fwd")
        ctx.save_for_backward(q, k, v, rpe, pos, out, lse)
        ctx.meta = (Hq, Wq, hpg, float(scale), flags, float(drop_p), int(seed))
        return out

    @staticmethod
    def backward(ctx, grad_out):
        q, k, v, rpe, pos, out, lse = ctx.saved_tensors
        Hq, Wq, hpg, scale, flags, drop_p, seed = ctx.meta
        BH, Dh, Nq = q.shape
        Ns = k.shape[2]
        h, Rh, Rw = rpe.shape
        need = ctx.needs_input_grad
        grad_out = grad_out.contiguous().float()
        gq = torch.empty_like(q) if need[0] else None
        gk = torch.empty_like(k) if need[1] else None
        gv = torch.empty_like(v) if need[2] else None
        gr = torch.zeros_like(rpe) if need[3] else None
        gp = torch.zeros_like(pos) if need[4] else None
        delta = torch.empty_like(lse)
        ws = torch.empty(_lib.lib().bevr_rpe_attn_workspace_bytes(BH, h, Dh, Nq, Ns, Rh, Rw, 1), dtype=torch.uint8, device=q.device)
        _lib.check(_lib.lib().bevr_rpe_attn_bwd_dropout(_ptr(q), _ptr(k), _ptr(v), _ptr(rpe), _ptr(pos), _ptr(out), _ptr(lse),
                                                        _ptr(grad_out), _ptr(gq), _ptr(gk), _ptr(gv), _ptr(gr), _ptr(gp),
                                                        _ptr(delta), _ptr(ws), BH, h, hpg, Dh, Hq, Wq, Ns, Rh, Rw, scale,
                                                        flags, drop_p, seed, _stream()),
                   "bevr_rpe_attn_bwd")
        return gq, gk, gv, gr, gp, None, None, None, None, None, None, None


K3_AUTO, K3_FFMA, K3_MMA = 0, 1, 2


K3_RUN_MERGE = 4


def rpe_attention(q, k, v, rpe_table, pos, Hq, Wq, n_groups=1, scale=None, family=K3_AUTO, dropout_p=0.0, seed=None,
                  merge_runs=False):
    """K3: softmax_j(q.k*scale + bilinear(rpe_table[head], (qgrid_i - pos_j)/2)) . v, never materialised.

    q (B*h, Dh, Hq*Wq); k, v (B*h, Dh, Ns); rpe_table (h, Rh, Rw); pos (B*g, Ns, 2) in (y, x);
    returns (B*h, Dh, Hq*Wq).  Replaces TSA_deform_attn.py:245-330 / SCA_deform_attn.py:331-411.
    ``family``: K3_AUTO picks the kernel family by problem size; K3_FFMA / K3_MMA force one (tests, benchmarks).
    ``dropout_p`` > 0: attention dropout fused into the kernels (nn.Dropout on the softmax matrix,
    TSA_deform_attn.py:313-323); the keep mask is a hash of (``seed``, head, query, key) regenerated in backward.
    ``seed`` None draws one from torch's default generator (so ``torch.manual_seed`` makes a run reproducible).
    ``merge_runs``: the keys arrive ordered so that keys addressing the same table cell are consecutive
    (BEVR_K3_RUN_MERGE): the backward sums such runs before the table-gradient atomics.
    """
    h = rpe_table.shape[0]
    if scale is None:
        scale = q.shape[1] ** -0.5
    dropout_p = float(dropout_p)
    if not 0.0 <= dropout_p < 1.0:
        raise ValueError(f"rpe_attention: dropout_p must be in [0, 1), got {dropout_p}")
    if dropout_p > 0.0 and seed is None:
        seed = int(torch.randint(0, 2 ** 62, (1,), dtype=torch.int64).item())
    return _RpeAttention.apply(q, k, v, rpe_table, pos, int(Hq), int(Wq), h // n_groups, scale,
                               int(family) | (K3_RUN_MERGE if merge_runs else 0), dropout_p, int(seed or 0))


class _OffsetHead(torch.autograd.Function):
    @staticmethod
    def forward(ctx, y, ln_w, ln_b, w1, ref, interleave, scaled, factor, eps):
        _need_cuda(y, ln_w, ln_b, w1, ref)
        y, ref = y.contiguous().float(), ref.contiguous().float()
        ln_w, ln_b = ln_w.contiguous().float(), ln_b.contiguous().float()
        Bg, Cm, H, W = y.shape
        O = w1.shape[0]
        w1 = w1.contiguous().float()
        Hk, Wk = (H // 2, W * O) if interleave else (H, W)
        if tuple(ref.shape[1:]) != (Hk, Wk, 2) or ref.shape[0] not in (1, Bg):
            raise ValueError(f"offset_head: ref must be ({Bg} or 1, {Hk}, {Wk}, 2), got {tuple(ref.shape)}")
        stride = 0 if ref.shape[0] == 1 else Hk * Wk * 2
        pos = torch.empty(Bg, Hk, Wk, 2, dtype=torch.float32, device=y.device)
        stats = torch.empty(Bg * H * W, 2, dtype=torch.float32, device=y.device)
        _lib.check(_lib.lib().bevr_offset_head_fwd(_ptr(y), _ptr(ln_w), _ptr(ln_b), _ptr(w1), _ptr(ref), stride,
                                                   _ptr(pos), _ptr(stats), Bg, Cm, H, W, O, int(interleave), int(scaled),
                                                   float(factor), float(eps), _stream()), "bevr_offset_head_fwd")
        ctx.save_for_backward(y, ln_w, ln_b, w1, ref, stats)
        ctx.meta = (stride, int(interleave), int(scaled), float(factor), float(eps))
        return pos

    @staticmethod
    def backward(ctx, grad_pos):
        y, ln_w, ln_b, w1, ref, stats = ctx.saved_tensors
        stride, interleave, scaled, factor, eps = ctx.meta
        Bg, Cm, H, W = y.shape
        O = w1.shape[0]
        need = ctx.needs_input_grad
        grad_pos = grad_pos.contiguous().float()
        gy = torch.empty_like(y) if need[0] else None
        gw = torch.zeros_like(ln_w) if need[1] else None
        gb = torch.zeros_like(ln_b) if need[2] else None
        g1 = torch.zeros_like(w1) if need[3] else None
        _lib.check(_lib.lib().bevr_offset_head_bwd(_ptr(y), _ptr(ln_w), _ptr(ln_b), _ptr(w1), _ptr(ref), stride,
                                                   _ptr(stats), _ptr(grad_pos), _ptr(gy), _ptr(gw), _ptr(gb), _ptr(g1),
                                                   Bg, Cm, H, W, O, interleave, scaled, factor, eps, _stream()),
                   "bevr_offset_head_bwd")
        gr = None
        if need[4]:      # pos = f(raw) + ref (scaled) or clamp(raw + ref): the reference points are never trained
            raise NotImplementedError("offset_head: gradient w.r.t. the reference points is not provided")
        return gy, gw, gb, g1, gr, None, None, None, None


def offset_head(y, ln_weight, ln_bias, w1, ref, interleave=False, scaled=True, factor=0.5, eps=1e-5):
    """Fused tail of the offset network ("next" row f-2): LayerNorm over channels -> GELU -> 1x1 conv (no bias) ->
    ``tanh * (1/(Hk-1), 1/(Wk-1)) * factor`` (or clamp to [-1, 1] when ``scaled`` is False) -> ``+ ref``.

    y (B*g, Cm, H, W): output of the depthwise convolution; w1 (O, Cm[, 1, 1]); ref (B*g or 1, Hk, Wk, 2) in (y, x).
    ``interleave=False`` (TSA): O = 2, pos[b, y, x, :] = (y, x).  ``interleave=True`` (SCA): the reference's
    "(b g) d (h n) w -> (b g) n h (w d)" split, pos (B*g, H/2, W*O, 2).  Returns the sampling positions (y, x).
    Replaces conv_offset[1:] + TSA_deform_attn.py:158-196 / SCA_deform_attn.py:219-277.
    """
    if ref.dim() == 3:
        ref = ref[None]
    w1 = w1.reshape(w1.shape[0], -1)          # (O, Cm, 1, 1) conv weight -> (O, Cm); autograd undoes the view
    return _OffsetHead.apply(y, ln_weight, ln_bias, w1, ref, bool(interleave), bool(scaled), float(factor), float(eps))


def correlate(map_tile, template, angles_rad):
    """K5: rotation-sweep valid-mode correlation on the tcgen05 tensor cores (bf16 operands, fp32 accumulate).

    map_tile (C, Hm, Wm), template (C, Ht, Wt) fp32 CUDA; angles_rad: sequence / CPU tensor of R angles.
    Returns (R, Hm-Ht+1, Wm-Wt+1) fp32.  No reference counterpart exists.
    """
    _need_cuda(map_tile, template)
    m, t = map_tile.contiguous().float(), template.contiguous().float()
    ang = [float(a) for a in (angles_rad.tolist() if torch.is_tensor(angles_rad) else angles_rad)]
    Cc, Hm, Wm = m.shape
    _, Ht, Wt = t.shape
    R = len(ang)
    L = _lib.lib()
    nws = L.bevr_correlate_workspace_bytes(Cc, Hm, Wm, Ht, Wt, R)
    ws = torch.empty(max(nws, 16), dtype=torch.uint8, device=m.device)
    out = torch.empty(R, Hm - Ht + 1, Wm - Wt + 1, dtype=torch.float32, device=m.device)
    _lib.check(L.bevr_correlate(_ptr(m), _ptr(t), (C.c_float * R)(*ang), _ptr(out), _ptr(ws), Cc, Hm, Wm, Ht, Wt, R,
                                _stream()), "bevr_correlate")
    return out


def affine_warp(img, theta):
    """One batched torchvision-style affine resampling pass (bilinear, fill=0): img (B,C,H,W), theta (B,2,3)."""
    _need_cuda(img, theta)
    img, theta = img.contiguous().float(), theta.contiguous().float()
    B, Cc, H, W = img.shape
    out = torch.empty_like(img)
    _lib.check(_lib.lib().bevr_affine_warp(_ptr(img), _ptr(theta), _ptr(out), B, Cc, H, W, _stream()), "bevr_affine_warp")
    return out


@torch.no_grad()
def project_history_bev(bev, vehicle_pose):
    """Drop-in for ``EncoderLayer.project_history_bev_feat(bev, vehicle_pose)`` (encoder.py:413-466): rotate by the
    previous heading and translate by the ego motion, then rotate back by the current heading -- the same TWO
    bilinear resampling passes as the reference, batched, with the matrices built on the device (no ``.item()``
    / ``math.degrees(tensor)`` syncs).  bev (B,C,H,W); vehicle_pose (B,2,3) rows (x, y, heading) of prev / current.
    """
    p = vehicle_pose.to(bev.device).double()
    prev_rot, curr_rot = p[:, 0, 2], p[:, 1, 2]
    d = (vehicle_pose[:, 0, :2] - vehicle_pose[:, 1, :2]).to(bev.device).double()     # fp32 difference, as the reference
    tx, ty = d[:, 0], d[:, 1]
    # the reference converts to degrees and torchvision back to radians: keep that round trip
    r1 = torch.deg2rad(torch.rad2deg(prev_rot))
    r2 = torch.deg2rad(torch.rad2deg(-curr_rot))
    c1, s1, c2, s2 = r1.cos(), r1.sin(), r2.cos(), r2.sin()
    zero = torch.zeros_like(c1)
    th1 = torch.stack((c1, s1, c1 * (-tx) + s1 * (-ty), -s1, c1, (-s1) * (-tx) + c1 * (-ty)), -1)
    th2 = torch.stack((c2, s2, zero, -s2, c2, zero), -1)
    return affine_warp(affine_warp(bev, th1.float()), th2.float())


@torch.no_grad()
def retrieval_ranks(camera_embed, map_embed):
    """rank[k] = #{r : dist[r, k] < dist[k, k]} with dist = 2 - 2 X Y^T: the GEMM is a plain cuBLAS call in fp32 (ranks
    are discontinuous in dist: no TF32), the counting loop of Trainer.get_recall (train.py:559-566) is one kernel."""
    _need_cuda(camera_embed, map_embed)
    prev = torch.backends.cuda.matmul.allow_tf32
    torch.backends.cuda.matmul.allow_tf32 = False
    try:
        dist = (2.0 - 2.0 * (camera_embed.float() @ map_embed.float().T)).contiguous()
    finally:
        torch.backends.cuda.matmul.allow_tf32 = prev
    n = dist.shape[0]
    rank = torch.empty(n, dtype=torch.int32, device=dist.device)
    _lib.check(_lib.lib().bevr_recall_ranks(_ptr(dist), _ptr(rank), n, _stream()), "bevr_recall_ranks")
    return rank


@torch.no_grad()
def retrieval_recall(camera_embed, map_embed, ks=(1, 5, 10)):
    """Recall@k of the retrieval metric ("next" row f-4; Trainer.get_recall, train.py:551-572), on the device:
    recall@i = 100 * mean(rank < i) with the ranks of ``retrieval_ranks``.  The reference evaluates this with numpy and an
    O(11 * N) Python loop on the host."""
    rank = retrieval_ranks(camera_embed, map_embed)
    return tuple(float((rank < k).float().mean() * 100) for k in ks)

"""Oracle for the CALLER of the hot path: one ``EncoderLayer`` of the reference in train mode.

Functional restatement of ``EncoderLayer.forward`` (encoder.py:339-411: LPU -> TSA -> conv-MLP -> LPU -> SCA ->
conv-MLP with one shared LayerNormProxy, identity layer-scale, drop-path 0) together with the reference-point
plumbing of ``SpatialCrossAttn.forward`` (SCA.py:78-94: stack -> repeat "v n h w d -> b v h (w d) n") and
``TransformerMLPWithConv`` (model_utils.py:6-35).  The two attention calls are injectable, so the same glue runs
(a) on CPU with the oracle attention of ``oracle/deform_attn.py`` and (b) on a GPU with the drop-in CUDA modules --
that is the encoder-level drop-in test.  Pinned by ``tests/golden/encoder_layer.npz`` (the real EncoderLayer's
output and gradients).
"""
import torch
import torch.nn.functional as F

from . import deform_attn as oda


def _sub(p, prefix):
    n = len(prefix)
    return {k[n:]: v for k, v in p.items() if k.startswith(prefix)}


def _ln(p, x):
    y = F.layer_norm(x.permute(0, 2, 3, 1), (x.shape[1],), p["layer_norm.norm.weight"], p["layer_norm.norm.bias"])
    return y.permute(0, 3, 1, 2).contiguous()


def _mlp(p, name, x):
    x = F.conv2d(x, p[name + ".linear1.0.weight"], p[name + ".linear1.0.bias"])
    x = x + F.conv2d(x, p[name + ".dwc.weight"], p[name + ".dwc.bias"], padding=1, groups=x.shape[1])
    return F.conv2d(F.gelu(x), p[name + ".linear2.0.weight"], p[name + ".linear2.0.bias"])


def _lpu(p, name, x):
    return F.conv2d(x, p[name + ".weight"], p[name + ".bias"], padding=1, groups=x.shape[1])


def reference_points_for_batch(ref_points, B):
    """(V, 2, N/2, N, D) -> (B, V, N/2, N*D, 2)   ("v n h w d -> b v h (w d) n", SCA.py:83-85)."""
    V, _, hh, ww, dd = ref_points.shape
    r = ref_points.permute(0, 2, 3, 4, 1).reshape(V, hh, ww * dd, 2)
    return r[None].expand(B, -1, -1, -1, -1).contiguous()


def encoder_layer_forward(p, query, img, prev_bev, ref_points, n_heads, bev_depth_dim, kernel_size=3, stride=1,
                          tsa_fn=None, sca_fn=None):
    """p: EncoderLayer state_dict (tensors); img (B*V, C, Hi, Wi) with V = 1; ref_points (V, 2, N/2, N, D)."""
    B = query.shape[0]
    tsa_p = _sub(p, "temporal_self_attn.temporal_deform_attn.")
    sca_p = _sub(p, "spatial_cross_attn.spatial_deform_attn.")
    if tsa_fn is None:
        tsa_fn = lambda x, q: oda.tsa_forward(tsa_p, x, q, n_heads, 1, kernel_size, stride)[0]
    if sca_fn is None:
        sca_fn = lambda x, q, r: oda.sca_forward(sca_p, x, q, r, n_heads, 1, bev_depth_dim)[0]
    x = query
    x = x + _lpu(p, "tsa_local_percept_unit", x)
    x = tsa_fn(prev_bev, _ln(p, x)) + x
    x = _mlp(p, "tsa_mlp", _ln(p, x)) + x
    x = x + _lpu(p, "sca_local_percept_unit", x)
    img5 = img.reshape(B, 1, *img.shape[1:])
    x = sca_fn(img5, _ln(p, x), reference_points_for_batch(ref_points, B)) + x
    return _mlp(p, "sca_mlp", _ln(p, x)) + x

"""Encoder-level drop-in: the reference's EncoderLayer (encoder.py:243-411), restated in oracle/encoder_layer.py
and pinned on CPU against the real layer's golden output, reproduces that output on the GPU when its two attention
calls are served by the drop-in CUDA modules loaded from the reference's own state_dict."""
import pytest
import torch

from oracle import encoder_layer as oel
from util import FP32_TOL, load_golden, rel_err

torch.backends.cudnn.allow_tf32 = False          # the glue convolutions stay PyTorch; keep them fp32 (see test_modules.py)
torch.backends.cuda.matmul.allow_tf32 = False


def test_oracle_layer_matches_reference():
    g, p, gp = load_golden("encoder_layer")
    N, C, h, D, B, Hi, Wi = [int(v) for v in g["cfg"]]
    p = {k: v.clone().requires_grad_(v.dtype.is_floating_point) for k, v in p.items()}
    q = torch.from_numpy(g["query"]).requires_grad_(True)
    img = torch.from_numpy(g["img"]).requires_grad_(True)
    out = oel.encoder_layer_forward(p, q, img, torch.from_numpy(g["prev"]), torch.from_numpy(g["ref_points"]), h, D)
    out.backward(torch.from_numpy(g["grad_out"]))
    assert rel_err(out, g["out"]) < 1e-5
    assert rel_err(q.grad, g["g_query"]) < 1e-4
    assert rel_err(img.grad, g["g_img"]) < 1e-4
    for k in ("layer_norm.norm.weight", "tsa_mlp.dwc.weight", "spatial_cross_attn.spatial_deform_attn.rpe_table",
              "temporal_self_attn.temporal_deform_attn.conv_offset.3.weight"):
        assert rel_err(p[k].grad, gp[k]) < 1e-4, k


@pytest.mark.gpu
def test_encoder_layer_with_cuda_modules_matches_reference():
    import bevrender_b200 as bb
    g, p, gp = load_golden("encoder_layer")
    N, C, h, D, B, Hi, Wi = [int(v) for v in g["cfg"]]
    dev = "cuda"
    tsa = bb.TSADeformableAttention(N, C, h, 1, 1, 3, True, B, n_views=1)
    sca = bb.SCADeformableAttention(N, D, C, h, 1, 1, 1, True, B, n_views=1)
    tsa.load_state_dict(oel._sub(p, "temporal_self_attn.temporal_deform_attn."), strict=True)
    sca.load_state_dict(oel._sub(p, "spatial_cross_attn.spatial_deform_attn."), strict=True)
    tsa, sca = tsa.to(dev), sca.to(dev)
    pc = {k: v.to(dev).requires_grad_(v.dtype.is_floating_point) for k, v in p.items()}
    q = torch.from_numpy(g["query"]).to(dev).requires_grad_(True)
    img = torch.from_numpy(g["img"]).to(dev).requires_grad_(True)
    out = oel.encoder_layer_forward(pc, q, img, torch.from_numpy(g["prev"]).to(dev), torch.from_numpy(g["ref_points"]).to(dev),
                                    h, D, tsa_fn=lambda x, qq: tsa(x, qq, None, False)[0],
                                    sca_fn=lambda x, qq, r: sca(x, qq, r, None, False)[0])
    out.backward(torch.from_numpy(g["grad_out"]).to(dev))
    assert rel_err(out.cpu(), g["out"]) < FP32_TOL
    assert rel_err(q.grad.cpu(), g["g_query"]) < FP32_TOL
    assert rel_err(img.grad.cpu(), g["g_img"]) < FP32_TOL
    tsa_g = dict(tsa.named_parameters())
    sca_g = dict(sca.named_parameters())
    assert rel_err(tsa_g["rpe_table"].grad.cpu(), gp["temporal_self_attn.temporal_deform_attn.rpe_table"]) < FP32_TOL
    assert rel_err(sca_g["rpe_table"].grad.cpu(), gp["spatial_cross_attn.spatial_deform_attn.rpe_table"]) < FP32_TOL
    assert rel_err(sca_g["conv_offset_m0.3.weight"].grad.cpu(),
                   gp["spatial_cross_attn.spatial_deform_attn.conv_offset_m0.3.weight"]) < FP32_TOL
    assert rel_err(pc["layer_norm.norm.weight"].grad.cpu(), gp["layer_norm.norm.weight"]) < FP32_TOL

"""The drop-in really drops in: with compat.install() the reference's own EncoderLayer (encoder.py:243)
builds around the B200 modules and exposes the same state_dict.  Needs /root/reference (authoring
container only); skipped on the GPU box."""
import os
import sys
import types

import pytest
import torch

REF = os.environ.get("BEVR_REFERENCE", "/root/reference")
pytestmark = pytest.mark.skipif(not os.path.isdir(os.path.join(REF, "model")), reason="reference checkout absent")


def _layer(projector_cls):
    from model.encoder import EncoderLayer
    import numpy as np
    T = np.eye(4)
    T[:3, :3] = [[0, 0, 1], [-1, 0, 0], [0, -1, 0]]
    T[2, 3] = 1.5
    K = np.array([[400.0, 0, 320, 0], [0, 400.0, 256, 0], [0, 0, 1, 0]])
    proj = projector_cls(imu_to_rgb={0: [T]}, K={0: [K]}, vehicle_type_code=0, img_width=160, img_height=128,
                         ori_img_width=640, ori_img_height=512, device="cpu")
    return EncoderLayer, proj


def test_install_rebinds_and_state_dict_is_identical():
    import bevrender_b200 as bb
    from bevrender_b200 import compat
    if REF not in sys.path:
        sys.path.insert(0, REF)
    compat.ensure_timm_stub()
    import model.SCA
    import model.TSA
    ref_tsa, ref_sca = model.TSA.TSADeformableAttention, model.SCA.SCADeformableAttention
    from model.bev_cmr_proj import BEV2CameraProjector as RefProj
    EncoderLayer, proj = _layer(RefProj)
    kw = dict(bev_bound={"X": 20.0, "Y": 10.0, "Z": 2.0}, bev2cmr_projector=proj, n_views=1, bev_feat_shape=14,
              bev_depth_dim=5, z_shift=-1.0, dim_embed=64, expansion=2, stage_idx=0, n_groups=1, n_heads=2, stride=1,
              kernel_size=3, batch_size=2, scale_offset_range=True, drop_path_rate=0.0)
    ref_layer = EncoderLayer(**kw)
    try:
        compat.install()
        assert model.TSA.TSADeformableAttention is bb.TSADeformableAttention
        assert model.SCA.SCADeformableAttention is bb.SCADeformableAttention
        new_layer = EncoderLayer(**kw)
    finally:
        model.TSA.TSADeformableAttention, model.SCA.SCADeformableAttention = ref_tsa, ref_sca
    assert isinstance(new_layer.temporal_self_attn.temporal_deform_attn, bb.TSADeformableAttention)
    assert isinstance(new_layer.spatial_cross_attn.spatial_deform_attn, bb.SCADeformableAttention)
    a, b = ref_layer.state_dict(), new_layer.state_dict()
    assert list(a) == list(b)
    assert all(a[k].shape == b[k].shape for k in a)
    new_layer.load_state_dict(a, strict=True)

"""Rebind the reference's names to the B200 modules without editing the checkout.

The reference package (``model/``) must be importable (its directory on ``sys.path``).  ``install()``
replaces the three symbols the callers look up at construction time:
``model.TSA.TSADeformableAttention`` (TSA.py:4,30), ``model.SCA.SCADeformableAttention`` (SCA.py:5,43)
and ``model.bevrender.BEV2CameraProjector`` (bevrender.py:11,31).
"""
import importlib
import os
import sys
import types

import torch

_ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def find_reference():
    """Directory holding the reference's ``model/`` package: ``$BEVR_REFERENCE``, the authoring container's
    ``/root/reference``, or the staged copy that travels to the GPU box (``oracle/_ref/reference``, written by
    ``make -C oracle``).  Returns None when none exists."""
    for cand in (os.environ.get("BEVR_REFERENCE"), "/root/reference", os.path.join(_ROOT, "oracle", "_ref", "reference")):
        if cand and os.path.isdir(os.path.join(cand, "model")):
            return cand
    return None


def add_reference_to_path():
    """Make ``import model.encoder`` (the reference's caller of the hot path) resolvable; raises if there is no checkout."""
    ref = find_reference()
    if ref is None:
        raise ImportError("no reference checkout found (BEVR_REFERENCE, /root/reference, oracle/_ref/reference)")
    if ref not in sys.path:
        sys.path.insert(0, ref)
    ensure_timm_stub()
    return ref


def ensure_timm_stub():
    """``model/*`` imports ``timm.models.layers.{trunc_normal_, DropPath}``; provide them when timm is absent."""
    try:
        import timm.models.layers  # noqa: F401
        return
    except Exception:
        pass
    layers = types.ModuleType("timm.models.layers")
    layers.trunc_normal_ = torch.nn.init.trunc_normal_

    class DropPath(torch.nn.Module):
        def __init__(self, drop_prob=0.0):
            super().__init__()
            self.drop_prob = drop_prob

        def forward(self, x):
            if self.drop_prob == 0.0 or not self.training:
                return x
            keep = 1.0 - self.drop_prob
            mask = x.new_empty(x.shape[0], *([1] * (x.dim() - 1))).bernoulli_(keep)
            return x * mask / keep

    layers.DropPath = DropPath
    timm, models = types.ModuleType("timm"), types.ModuleType("timm.models")
    timm.models, models.layers = models, layers
    sys.modules.update({"timm": timm, "timm.models": models, "timm.models.layers": layers})


def install(projector=True, wrappers=False):
    """``wrappers=True`` additionally rebinds ``model.encoder.{SpatialCrossAttn, TemporalSelfAttn}`` to the versions
    that cache the projected reference points on the device (no per-forward upload / ``.item()`` sync) and routes
    ``EncoderLayer.project_history_bev_feat`` to the batched history-BEV warp kernel (row f-1)."""
    from . import BEV2CameraProjector, SCADeformableAttention, TSADeformableAttention
    ensure_timm_stub()
    tsa = importlib.import_module("model.TSA")
    sca = importlib.import_module("model.SCA")
    tsa.TSADeformableAttention = TSADeformableAttention
    sca.SCADeformableAttention = SCADeformableAttention
    if wrappers:
        from . import SpatialCrossAttn, TemporalSelfAttn, project_history_bev
        enc = importlib.import_module("model.encoder")
        enc.SpatialCrossAttn, enc.TemporalSelfAttn = SpatialCrossAttn, TemporalSelfAttn
        # row f-1: EncoderLayer.project_history_bev_feat (encoder.py:413-466) loops over the samples with two torchvision
        # F.affine calls each and math.degrees(tensor) / tuple(tensor) host syncs (~4 syncs and ~20 tiny kernels per sample,
        # per layer, per history frame).  The batched warp kernel serves the case the model actually runs -- history
        # features carry no gradient (bevrender.py:124-146 computes them under no_grad) and no mask is asked for; every
        # other case goes to the reference's own method.
        if not hasattr(enc.EncoderLayer, "_bevr_reference_project_history"):
            enc.EncoderLayer._bevr_reference_project_history = enc.EncoderLayer.project_history_bev_feat

            def project_history_bev_feat(self, bev, vehicle_pose, return_mask=False):
                if return_mask or not bev.is_cuda or bev.requires_grad or bev.dim() != 4:
                    return self._bevr_reference_project_history(bev, vehicle_pose, return_mask)
                return project_history_bev(bev, vehicle_pose)

            enc.EncoderLayer.project_history_bev_feat = project_history_bev_feat
    if projector:
        try:
            top = importlib.import_module("model.bevrender")
            top.BEV2CameraProjector = BEV2CameraProjector
        except ImportError:          # model.bevrender pulls the decoder/backbone; optional for layer-level use
            pass

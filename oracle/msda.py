"""Oracle for the multi-scale deformable-attention gather (kernels K1/K2).

The reference's own use of this op is the single-level, single-point,
unit-weight case: ``F.grid_sample(x, pos, bilinear, align_corners=True)`` at
TSA_deform_attn.py:210-217 and SCA_deform_attn.py:290-301 (pinned by
``tests/golden/tsa_*.npz`` / ``sca_*.npz`` through ``oracle/deform_attn.py``).

The general (L levels, P points, learned weights) form named by the north
star has NO counterpart in the reference: for L>1 or P>1
this oracle is the same construction the reference uses -- one
``F.grid_sample(bilinear, zeros)`` per level, weighted sum over points and
levels -- and its parity is "unpinned vs reference".  It IS pinned against two
independent definitions: the numpy restatement of grid_sample in
``oracle/bilinear.py`` and the published pure-PyTorch definition of the operator
the north star names, Deformable DETR's ``ms_deform_attn_core_pytorch``
(restated in ``oracle/msda_ddetr.py``; tests/test_oracle.py).

Layouts (channels-last, the B200 kernels' native layout):
  value              (B, S, h, Dh)      S = sum_l H_l*W_l
  sampling_locations (B, Nq, h, L, P, 2) (x, y)
  attention_weights  (B, Nq, h, L, P)   or None (== 1)
  output             (B, Nq, h*Dh)

Coordinate conventions (``flags`` of the C ABI):
  loc_pm1=False: loc in [0,1];  loc_pm1=True: loc in [-1,1] (grid_sample grid)
  align_corners as in ``F.grid_sample``.
"""
import torch
import torch.nn.functional as F


def locations_from_offsets(reference_points, sampling_offsets, spatial_shapes, offset_in_pixels):
    """loc = ref + off [/ (W_l, H_l)]  -- ref (B,Nq,L,2), off (B,Nq,h,L,P,2)."""
    off = sampling_offsets
    if offset_in_pixels:
        wh = torch.tensor([[w, h] for h, w in spatial_shapes], dtype=off.dtype, device=off.device)
        off = off / wh[None, None, None, :, None, :]
    return reference_points[:, :, None, :, None, :] + off


def msda_forward(value, spatial_shapes, level_start_index, sampling_locations,
                 attention_weights=None, align_corners=False, loc_pm1=False):
    B, S, h, Dh = value.shape
    _, Nq, _, L, P, _ = sampling_locations.shape
    out = value.new_zeros(B * h, Dh, Nq)
    for l, (H, W) in enumerate(spatial_shapes):
        s0 = level_start_index[l]
        v = value[:, s0:s0 + H * W].permute(0, 2, 3, 1).reshape(B * h, Dh, H, W)
        g = sampling_locations[:, :, :, l]                       # (B,Nq,h,P,2)
        if not loc_pm1:
            g = 2.0 * g - 1.0
        g = g.permute(0, 2, 1, 3, 4).reshape(B * h, Nq, P, 2)
        sampled = F.grid_sample(v, g, mode="bilinear", padding_mode="zeros",
                                align_corners=align_corners)       # (B*h,Dh,Nq,P)
        if attention_weights is not None:
            w = attention_weights[:, :, :, l].permute(0, 2, 1, 3).reshape(B * h, 1, Nq, P)
            sampled = sampled * w
        out = out + sampled.sum(-1)
    return out.reshape(B, h * Dh, Nq).transpose(1, 2).contiguous()


def msda_forward_backward(value, spatial_shapes, level_start_index, sampling_locations,
                          attention_weights, grad_out, align_corners=False, loc_pm1=False):
    """Returns out, grad_value, grad_loc, grad_weight (autograd of msda_forward)."""
    value = value.detach().clone().requires_grad_(True)
    loc = sampling_locations.detach().clone().requires_grad_(True)
    w = None
    if attention_weights is not None:
        w = attention_weights.detach().clone().requires_grad_(True)
    out = msda_forward(value, spatial_shapes, level_start_index, loc, w, align_corners, loc_pm1)
    out.backward(grad_out)
    return out.detach(), value.grad, loc.grad, (None if w is None else w.grad)

"""Drop-in replacements for the reference's deformable-attention modules.

``TSADeformableAttention`` / ``SCADeformableAttention`` keep the constructor arguments, parameter
names and shapes (so ``state_dict``s interchange) and the forward signatures of
model/TSA_deform_attn.py:14-337 and model/SCA_deform_attn.py:14-421, and therefore drop into
model/TSA.py / model/SCA.py / model/encoder.py unchanged.  The depthwise convolution of the offset
network and the 1x1 K/V/out projections stay PyTorch; the rest of the offset
network up to the sampling positions (fused offset head, row f-2), the feature gather (K1/K2) and the
attention with sampled relative-position bias (K3) run as sm_100a kernels behind the C ABI.

Differences from the reference, all deliberate:
  * nothing of size (Nq, Ns) is materialised (the reference needs 51 GB/sample at 200x200);
  * attention dropout (``attn_drop_rate``, reference default 0.0) is fused into K3: same Bernoulli(1 - p) / rescale
    semantics as ``nn.Dropout`` on the softmax matrix, but the keep mask comes from a counter-based hash seeded from
    torch's default generator, not from torch's Philox stream over the (never materialised) attention matrix;
  * ``TSADeformableAttention`` with ``n_groups > 1`` works here (the reference raises at
    TSA_deform_attn.py:278-283); ``SCADeformableAttention`` with ``n_views > 1`` raises here exactly
    as the reference does (SCA_deform_attn.py:248-255), with a clearer message.
"""
import torch
from torch import nn

from .functional import grid_gather, offset_head, rpe_attention


class LayerNormProxy(nn.Module):
    """LayerNorm over the channel axis of an NCHW tensor (model_utils.py:51-59); the sub-module is
    called ``norm`` so the parameter keys are ``...1.norm.weight`` as in the reference."""

    def __init__(self, dim):
        super().__init__()
        self.norm = nn.LayerNorm(dim)

    def forward(self, x):
        return self.norm(x.permute(0, 2, 3, 1)).permute(0, 3, 1, 2)


class PointwiseConv2d(nn.Conv2d):
    """1x1 convolution evaluated as a batched matmul.  Same parameter (``weight`` of shape (out, in, 1, 1)) as
    ``nn.Conv2d(c_in, c_out, 1, bias=False)``; cuDNN picks ``wgrad2d_grouped_direct_kernel`` for the 2- / D-channel
    outputs of the offset heads, which alone cost 1.7 ms (TSA) / 4.1 ms (SCA) per backward at N=56, B=8."""

    def forward(self, x):
        B, C, H, W = x.shape
        w = self.weight.reshape(self.out_channels, C)
        return torch.matmul(w, x.reshape(B, C, H * W)).reshape(B, self.out_channels, H, W)


def _offset_net(c_in, c_mid, c_out, kernel_size, stride, pad):
    return nn.Sequential(
        nn.Conv2d(c_in, c_mid, kernel_size, stride, pad, groups=c_in),
        LayerNormProxy(c_mid),
        nn.GELU(),
        PointwiseConv2d(c_mid, c_out, 1, 1, 0, bias=False),
    )


class _DeformAttnBase(nn.Module):
    def _common(self, dim_embed, n_heads, n_groups, stride, kernel_size, scale_offset_range, batch_size,
                n_views, attn_drop_rate, proj_drop_rate, data_type, logger):
        self.n_channel_per_head = dim_embed // n_heads
        self.scale = self.n_channel_per_head ** -0.5
        self.n_heads = n_heads
        self.embed_dim = self.n_channel_per_head * n_heads
        self.n_groups = n_groups
        self.n_channel_per_group = self.embed_dim // n_groups
        self.n_heads_per_group = n_heads // n_groups
        self.batch_size = batch_size
        self.scale_offset_range = scale_offset_range
        self.kernel_size, self.stride, self.n_views = kernel_size, stride, n_views
        self.data_type, self.logger = data_type, logger
        self._grids = {}
        self._drop = (attn_drop_rate, proj_drop_rate)

    def _projections(self, out_in_channels):
        """Registered after the offset nets so parameter ORDER (optimizer state) matches the reference too."""
        C, Cg = self.embed_dim, self.n_channel_per_group
        self.proj_q = nn.Conv2d(C, C, 1)          # defined but never applied, as in the reference
        self.proj_k = nn.Conv2d(C, C, 1)
        self.proj_v = nn.Conv2d(C, C, 1)
        self.proj_out = nn.Conv2d(out_in_channels, C, 1)
        self.proj_views = nn.Conv2d(Cg * self.n_views, Cg, 1)     # unused, kept for the state_dict
        self.proj_drop_rate = nn.Dropout(self._drop[1], inplace=True)
        self.attn_drop_rate = nn.Dropout(self._drop[0], inplace=True)

    @torch.no_grad()
    def _get_normalized_grid(self, H, W, B, device):
        """(B*g, H, W, 2) in (y, x), values arange/(n-1)*2-1 with the reference's op order."""
        key = (H, W, str(device))
        if key not in self._grids:
            ys = torch.arange(0, H, dtype=self.data_type, device=device).div_(H - 1.0).mul_(2.0).sub_(1.0)
            xs = torch.arange(0, W, dtype=self.data_type, device=device).div_(W - 1.0).mul_(2.0).sub_(1.0)
            gy, gx = torch.meshgrid(ys, xs, indexing="ij")
            self._grids[key] = torch.stack((gy, gx), -1)
        return self._grids[key][None].expand(B * self.n_groups, -1, -1, -1)

    def _positions(self, net, xq, ref, interleave):
        """Offset network -> sampling positions (B*g, Hk, Wk, 2) in (y, x): the depthwise convolution stays cuDNN,
        everything after it (LayerNorm, GELU, 1x1 conv, tanh * range * factor or clamp, + reference and the two
        rearranges) is the fused offset-head kernel (row f-2).  ``net`` is the reference-shaped nn.Sequential."""
        ln = net[1].norm
        return offset_head(net[0](xq), ln.weight, ln.bias, net[3].weight, ref, interleave=interleave,
                           scaled=bool(self.scale_offset_range),
                           factor=self.offset_range_factor if self.scale_offset_range else 1.0, eps=ln.eps)

    @staticmethod
    def _pointwise(conv, x):
        """conv: nn.Conv2d(C_in, C_out, 1); x (B, C_in, N) -> (B, C_out, N)."""
        w = conv.weight.reshape(conv.out_channels, conv.in_channels)
        return torch.baddbmm(conv.bias.reshape(1, -1, 1), w.expand(x.shape[0], -1, -1), x)

    def _attend(self, x_grouped, pos, query, sort_keys=False):
        """gather (K1/K2) -> K/V projections -> fused RPE attention (K3) -> (B, C, Hq, Wq)."""
        B, C, Hq, Wq = query.shape
        Ns = pos.shape[1] * pos.shape[2]
        sort_keys = sort_keys and self.n_groups == 1
        if sort_keys:
            # Attention is invariant under a permutation of the keys (with several groups the K / V projections mix the
            # groups' channels at equal key index, so a per-group order is only available for n_groups == 1).  SCA's masked reference points all sit at (-1, -1)
            # plus their learned offsets (bev_cmr_proj.py:106-124): in the reference's key order eight consecutive keys
            # then scatter over 3-4 rows of the bias table, and so do the 8 lanes of every table-gradient reduction of the
            # dK/dV kernel (3.3x the L2 RED sectors, 8.3 instead of 5.1 ms per backward).  Keys ordered by the table row
            # they address (the y axis of the table advances exactly one cell per query row, so the row group of a key
            # does not depend on the query) and by x within a row group keep a warp's keys in one row, cells apart.
            with torch.no_grad():
                p = pos.reshape(-1, Ns, 2)
                Rh, Rw = self.rpe_table.shape[1:]
                group = torch.floor((1.0 - p[..., 0]) * (0.25 * (Rh - 1)))
                perm = torch.argsort(group * 8192.0 - p[..., 1] * (0.25 * (Rw - 1)), dim=1)
            pos = torch.gather(pos.reshape(-1, Ns, 2), 1, perm[..., None].expand(-1, -1, 2)).reshape(-1, 1, Ns, 2)
        xs = grid_gather(x_grouped, pos).reshape(B, C, Ns)
        # proj_k / proj_v are 1x1 convolutions over a (B, C, 1, Ns) tensor: evaluate them as batched matmuls
        # (cuDNN selects wgrad2d_grouped_direct_kernel for this H = 1 shape: 1.7 ms (TSA) / 4.1 ms (SCA) per backward)
        k = self._pointwise(self.proj_k, xs).reshape(B * self.n_heads, self.n_channel_per_head, Ns)
        v = self._pointwise(self.proj_v, xs).reshape(B * self.n_heads, self.n_channel_per_head, Ns)
        q = query.reshape(B * self.n_heads, self.n_channel_per_head, Hq * Wq)
        # attention dropout (attn_drop_rate, TSA_deform_attn.py:313-323 / SCA_deform_attn.py:402-410) is fused into K3
        drop = float(self.attn_drop_rate.p) if self.training else 0.0
        out = rpe_attention(q, k, v, self.rpe_table, pos.reshape(-1, Ns, 2), Hq, Wq, self.n_groups, self.scale,
                            dropout_p=drop, merge_runs=sort_keys)
        return out.reshape(B, C, Hq, Wq)


class TSADeformableAttention(_DeformAttnBase):
    def __init__(self, bev_feat_shape, dim_embed, n_heads, n_groups, stride, kernel_size, scale_offset_range,
                 batch_size, n_views=3, attn_drop_rate=0.0, proj_drop_rate=0.0, data_type=torch.float32,
                 logger=None):
        super().__init__()
        self._common(dim_embed, n_heads, n_groups, stride, kernel_size, scale_offset_range, batch_size, n_views,
                     attn_drop_rate, proj_drop_rate, data_type, logger)
        self.bev_h = self.bev_w = bev_feat_shape
        if scale_offset_range:
            self.offset_range_factor = 0.5
        Cg = self.n_channel_per_group
        pad = kernel_size // 2 if kernel_size != stride else 0
        self.conv_offset = _offset_net(Cg, Cg, 2, kernel_size, stride, pad)
        self._projections(self.embed_dim)
        self.rpe_table = nn.Parameter(torch.zeros(n_heads, self.bev_h * 2 - 1, self.bev_w * 2 - 1))
        nn.init.trunc_normal_(self.rpe_table, std=0.01)

    def forward(self, x, query, wandb_log_dict, return_wandb_log=True):
        """x, query (B, C, H, W) -> ((B, C, H, W), wandb_log_dict); x=None means self-attention."""
        if x is None:
            x = query
        B, C, H, W = x.shape
        g, Cg = self.n_groups, self.n_channel_per_group
        pad = self.kernel_size // 2 if self.kernel_size != self.stride else 0
        Hk = (H + 2 * pad - self.kernel_size) // self.stride + 1
        Wk = (W + 2 * pad - self.kernel_size) // self.stride + 1
        pos = self._positions(self.conv_offset, query.reshape(B * g, Cg, H, W),
                              self._get_normalized_grid(Hk, Wk, 1, x.device)[:1], interleave=False)
        out = self._attend(x.reshape(B * g, Cg, H, W), pos, query)
        return self.proj_drop_rate(self.proj_out(out)), wandb_log_dict


class SCADeformableAttention(_DeformAttnBase):
    sort_keys = True      # order the keys by the bias-table row they address (see _attend); class-level switch for measurements

    def __init__(self, bev_feat_shape, bev_depth_dim, dim_embed, n_heads, n_groups, stride, kernel_size,
                 scale_offset_range, batch_size, n_views=3, attn_drop_rate=0.0, proj_drop_rate=0.0,
                 data_type=torch.float32, logger=None):
        super().__init__()
        self._common(dim_embed, n_heads, n_groups, stride, kernel_size, scale_offset_range, batch_size, n_views,
                     attn_drop_rate, proj_drop_rate, data_type, logger)
        self.query_height = self.query_width = bev_feat_shape
        self.bev_depth_dim = D = bev_depth_dim
        if scale_offset_range:
            self.offset_range_factor = 5.0
        Cg = self.n_channel_per_group
        # only conv_offset_m0 is ever usable (the reference's rearrange fixes d = D); m1/m2 exist for the keys
        self.conv_offset_m0 = _offset_net(Cg, Cg * D, D, 1, 1, 0)
        self.conv_offset_m1 = _offset_net(Cg, Cg * D, 2 * D, 1, 1, 0)
        self.conv_offset_m2 = _offset_net(Cg, Cg * D, 2 * D, 1, 1, 0)
        self._projections(self.embed_dim * n_views)
        self.rpe_table = nn.Parameter(
            torch.zeros(n_heads, self.query_height * 2 - 1, self.query_width * D * 2 - 1))
        nn.init.trunc_normal_(self.rpe_table, std=0.01)

    def forward(self, x, query, reference_points, wandb_log_dict, return_wandb_log=True):
        """x (B, V, C, Hi, Wi); query (B, C, Hq, Wq); reference_points (B, V, Hq/2, Wq*D, 2) in (x, y)."""
        if self.n_views != 1:
            raise NotImplementedError(
                "SCADeformableAttention: n_views > 1 is unsupported (the reference itself raises an EinopsError "
                "for views 1 and 2: conv_offset_m1/m2 emit 2*D channels, SCA_deform_attn.py:91-121,248-255)")
        B, C, Hq, Wq = query.shape
        g, Cg, D = self.n_groups, self.n_channel_per_group, self.bev_depth_dim
        Hi, Wi = x.shape[-2:]
        ref = reference_points[:, 0].flip(-1)                         # (x, y) -> (y, x)
        if g > 1:
            ref = ref.repeat_interleave(g, 0)
        # "(b g) d (h n) w -> (b g) n h (w d)", n = 2: even query rows carry y-, odd rows x-offsets (interleave=True)
        pos = self._positions(self.conv_offset_m0, query.reshape(B * g, Cg, Hq, Wq), ref, interleave=True)
        out = self._attend(x[:, 0].reshape(B * g, Cg, Hi, Wi), pos, query, sort_keys=self.sort_keys)
        return self.proj_drop_rate(self.proj_out(out)), wandb_log_dict

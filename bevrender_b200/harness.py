"""Caller-side harness for BASELINE.json configs 3 and 5 (full encoder / end-to-end training step).

NOT part of the product path and not a re-implementation of the reference model: the reference's
``model/encoder.py`` is the real caller of the drop-in modules, but the reference checkout
does not exist on the GPU box, so the benchmarks need a stand-in caller written from scratch.  It keeps the
*sequence* in which the reference invokes the hot path -- per layer ``x += dw3x3(x)``; TSA(LN(x), prev_bev);
conv-MLP; ``x += dw3x3(x)``; SCA(LN(x), image features, projected reference points); conv-MLP
(encoder.py:339-411) and, per training step, T-1 history passes without gradient followed by one pass with
gradient (bevrender.py:124-146) -- around a small strided-conv stem and a conv/upsample render head.
"""
import numpy as np
import torch
import torch.nn.functional as F
from torch import nn

from .bev_cmr_proj import BEV2CameraProjector
from .deform_attn import LayerNormProxy, SCADeformableAttention, TSADeformableAttention


def default_camera():
    """Pinhole camera of BASELINE.md config 1: 1.5 m above the IMU looking along +x, f=400 @ 640x512."""
    T = np.eye(4)
    T[:3, :3] = [[0.0, 0.0, 1.0], [-1.0, 0.0, 0.0], [0.0, -1.0, 0.0]]
    T[2, 3] = 1.5
    K = np.array([[400.0, 0.0, 320.0, 0.0], [0.0, 400.0, 256.0, 0.0], [0.0, 0.0, 1.0, 0.0]])
    return T, K


def pillar_points(bound, n, depth, z_shift):
    """(4, n/2, n, depth) homogeneous sampling points, the layout of SCA.py:112-162."""
    X, Y, Z = bound["X"], bound["Y"], bound["Z"]
    xs = torch.arange(n // 2, dtype=torch.float32) * (2 * X / n) + X / n
    ys = torch.arange(n, dtype=torch.float32) * (2 * Y / n) + (-Y + Y / n)
    zs = torch.arange(depth, dtype=torch.float32) * (2 * Z / depth) + (-Z + Z / depth + z_shift)
    gx, gy, gz = torch.meshgrid(xs, ys, zs, indexing="ij")
    return torch.stack((gx, gy, gz, torch.ones_like(gz)), 0)


class ConvMLP(nn.Module):
    def __init__(self, c, expansion):
        super().__init__()
        self.fc1 = nn.Conv2d(c, c * expansion, 1)
        self.dw = nn.Conv2d(c * expansion, c * expansion, 3, 1, 1, groups=c * expansion)
        self.fc2 = nn.Conv2d(c * expansion, c, 1)

    def forward(self, x):
        x = self.fc1(x)
        return self.fc2(F.gelu(x + self.dw(x)))


class EncoderLayer(nn.Module):
    def __init__(self, n, c, heads, depth, batch, ref_points, expansion=2, kernel_size=3, stride=1):
        super().__init__()
        self.norm = LayerNormProxy(c)
        self.tsa_lpu = nn.Conv2d(c, c, 3, 1, 1, groups=c)
        self.sca_lpu = nn.Conv2d(c, c, 3, 1, 1, groups=c)
        self.tsa = TSADeformableAttention(n, c, heads, 1, stride, kernel_size, True, batch, n_views=1)
        self.sca = SCADeformableAttention(n, depth, c, heads, 1, 1, 1, True, batch, n_views=1)
        self.tsa_mlp, self.sca_mlp = ConvMLP(c, expansion), ConvMLP(c, expansion)
        self.register_buffer("ref_points", ref_points, persistent=False)     # (1, 1, n/2, n*depth, 2)

    def forward(self, x, img_feat, prev_bev):
        B = x.shape[0]
        x = x + self.tsa_lpu(x)
        x = x + self.tsa(prev_bev, self.norm(x).contiguous(), None, False)[0]
        x = x + self.tsa_mlp(self.norm(x))
        x = x + self.sca_lpu(x)
        ref = self.ref_points.expand(B, -1, -1, -1, -1)
        x = x + self.sca(img_feat[:, None], self.norm(x).contiguous(), ref, None, False)[0]
        return x + self.sca_mlp(self.norm(x))


class BevHarness(nn.Module):
    """stem(image) -> L encoder layers over a learned BEV query -> render head (B, 3, 224, 224)."""

    def __init__(self, n=56, c=64, heads=2, depth=5, layers=6, batch=8, img_hw=(128, 160), out_hw=(224, 224),
                 bound=None, z_shift=-1.0, device="cuda"):
        super().__init__()
        bound = bound or {"X": 20.0, "Y": 10.0, "Z": 2.0}
        T, K = default_camera()
        proj = BEV2CameraProjector({0: [T]}, {0: [K]}, 0, img_hw[1], img_hw[0], 640, 512, device=device)
        uv = proj.bev_grid_to_camera(pillar_points(bound, n, depth, z_shift))[0][0]          # (2, n/2, n, depth)
        ref = uv.permute(1, 2, 3, 0).reshape(1, 1, n // 2, n * depth, 2).contiguous()       # "v n h w d -> b v h (w d) n"
        self.n, self.c, self.out_hw = n, c, out_hw
        self.bev_query = nn.Parameter(torch.rand(c, n, n))
        self.stem = nn.Sequential(nn.Conv2d(3, c // 2, 3, 2, 1), nn.BatchNorm2d(c // 2), nn.ReLU(inplace=True),
                                  nn.Conv2d(c // 2, c, 3, 2, 1), nn.BatchNorm2d(c), nn.ReLU(inplace=True),
                                  nn.Conv2d(c, c, 3, 1, 1), nn.BatchNorm2d(c), nn.ReLU(inplace=True))
        self.layers = nn.ModuleList([EncoderLayer(n, c, heads, depth, batch, ref) for _ in range(layers)])
        self.head = nn.Sequential(nn.Conv2d(c, c, 3, 1, 1), nn.BatchNorm2d(c), nn.ReLU(inplace=True),
                                  nn.Upsample(scale_factor=2, mode="bilinear", align_corners=False),
                                  nn.Conv2d(c, c // 2, 3, 1, 1), nn.BatchNorm2d(c // 2), nn.ReLU(inplace=True),
                                  nn.Conv2d(c // 2, 3, 3, 1, 1))

    def encode(self, img, prev_bev):
        feat = self.stem(img)
        x = self.bev_query[None].expand(img.shape[0], -1, -1, -1).contiguous()
        for layer in self.layers:
            x = layer(x, feat, prev_bev)
        return x

    def forward(self, frames):
        """frames (B, T, 3, H, W): T-1 history passes without gradient, then the final pass with gradient."""
        prev = None
        with torch.no_grad():
            for t in range(frames.shape[1] - 1):
                prev = self.encode(frames[:, t], prev)
        bev = self.encode(frames[:, -1], prev)
        out = F.interpolate(self.head(bev), size=self.out_hw, mode="bilinear", align_corners=False)
        return torch.sigmoid(out)

"""Synthetic workloads of BASELINE.json (shape factories only -- no arithmetic of the path)."""
import torch


def make_microbench_inputs(B, seed=0, dtype=torch.float32, n_side=200, h=8, Dh=32, L=4, P=4,
                           sigma_px=2.0, device="cpu"):
    """Synthetic inputs of BASELINE.json config 2:
    BEV pyramid [(N,N),(N/2,N/2),...], queries = cell centres of the NxN grid,
    offsets ~ N(0,(2px)^2) per level, weights = softmax(N(0,1)) over L*P."""
    g = torch.Generator().manual_seed(seed)
    shapes = [(n_side >> l, n_side >> l) for l in range(L)]
    starts, s = [], 0
    for H, W in shapes:
        starts.append(s)
        s += H * W
    Nq = n_side * n_side
    value = torch.randn(B, s, h, Dh, generator=g)
    ys, xs = torch.meshgrid(torch.arange(n_side), torch.arange(n_side), indexing="ij")
    ref = torch.stack(((xs + 0.5) / n_side, (ys + 0.5) / n_side), -1).reshape(1, Nq, 1, 1, 1, 2)
    wh = torch.tensor([[w, hh] for hh, w in shapes], dtype=torch.float32).reshape(1, 1, 1, L, 1, 2)
    loc = ref + torch.randn(B, Nq, h, L, P, 2, generator=g) * sigma_px / wh
    aw = torch.softmax(torch.randn(B, Nq, h, L * P, generator=g), -1).reshape(B, Nq, h, L, P)
    grad_out = torch.randn(B, Nq, h * Dh, generator=g)
    to = dict(device=device)
    return dict(value=value.to(dtype).to(**to), spatial_shapes=shapes, level_start_index=starts,
                sampling_locations=loc.float().to(**to), attention_weights=aw.to(dtype).to(**to),
                grad_out=grad_out.to(dtype).to(**to), query_w=n_side)

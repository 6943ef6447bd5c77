"""GPU smoke test of the caller-side harness + flat-bucket gradient reducer (world size 1): one training step
runs through the drop-in modules, every live parameter receives a finite gradient in the flat buffer, and the
parameters the reference never uses stay out of the optimiser."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def test_harness_training_step():
    from bevrender_b200.dp import FlatGradAllReduce, live_named_parameters
    from bevrender_b200.harness import BevHarness
    torch.manual_seed(0)
    model = BevHarness(n=14, c=64, heads=2, depth=5, layers=2, batch=2, img_hw=(64, 80), out_hw=(56, 56)).cuda()
    live = live_named_parameters(model)
    assert not any("proj_q" in n or "proj_views" in n or "conv_offset_m1" in n for n, _ in live)
    red = FlatGradAllReduce([p for _, p in live], n_buckets=2)
    frames = torch.rand(2, 3, 3, 64, 80, device="cuda")
    target = torch.rand(2, 3, 56, 56, device="cuda")
    red.reset()
    loss = torch.nn.functional.mse_loss(model(frames), target)
    loss.backward()
    red.finish()
    assert torch.isfinite(loss)
    assert torch.isfinite(red.flat).all() and red.flat.abs().sum() > 0
    for n, p in live:
        if n.endswith("proj_k.bias"):          # mathematically zero gradient (softmax shift invariance)
            continue
        assert p.grad is not None and p.grad.abs().sum() > 0, n

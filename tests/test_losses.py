"""Plain-torch retrieval losses of config 5 (PML restatement, parity unpinned): hand-computed known answers."""
import torch

from bevrender_b200 import losses


def test_contrastive_known_answer():
    # two classes; embeddings on the unit circle at 0, 90 degrees (camera) and 0, 180 degrees (map)
    cmr = torch.tensor([[1.0, 0.0], [0.0, 1.0]])
    mp = torch.tensor([[1.0, 0.0], [0.0, -1.0]])
    # positives: (c0,m0) d=0 -> 0 ; (c1,m1) d=2 -> 2 (both directions); mean over non-zero = 2
    # negatives: d(c0,c1)=sqrt2, d(c0,m1)=sqrt2, d(m0,c1)=sqrt2, d(m0,m1)=sqrt2 -> relu(1-sqrt2)=0 -> 0
    assert abs(float(losses.contrastive_loss(cmr, mp)) - 2.0) < 1e-6


def test_triplet_runs_and_backpropagates():
    torch.manual_seed(0)
    a = torch.randn(6, 32, requires_grad=True)
    b = (a.detach() + 0.5 * torch.randn(6, 32)).requires_grad_(True)
    loss = losses.triplet_loss(a, b)
    loss.backward()
    assert torch.isfinite(loss) and a.grad is not None and b.grad is not None
    # no semihard triplet when the positives coincide and negatives are far: loss is exactly the (thresholded) regulariser
    e = torch.eye(4) * 0.1
    assert abs(float(losses.triplet_loss(e, e.clone())) - 0.1) < 1e-6

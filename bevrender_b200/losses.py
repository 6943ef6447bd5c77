"""Retrieval losses of BASELINE.json config 5, restated in plain PyTorch.  PARITY UNPINNED.

The reference wraps ``pytorch_metric_learning`` (loss/contrastive_loss.py:8, loss/triplet_loss_metric.py:11-16); that
package is neither vendored in the reference nor installed here and its version is not pinned, so these functions
restate the PUBLISHED algorithm of PML 2.x for exactly the configurations the reference constructs.  They are caller-side
glue (not on the hot path): they exist so that the backward graph of the end-to-end training step matches the
configuration BASELINE.json quotes.  The shipped ``BEVRender`` returns a 4-D image, which PML itself would reject
(SURVEY.md section 8c); following config 5 the embeddings are the flattened outputs.

Both take the camera-side and map-side embeddings (B, D) and label row i of either side with class i, as the
reference's ``get_loss`` does (``labels = cat(arange(B), arange(B))``).
"""
import torch
import torch.nn.functional as F


def _pairs(B, device):
    labels = torch.arange(B, device=device).repeat(2)
    same = labels[:, None] == labels[None, :]
    eye = torch.eye(2 * B, dtype=torch.bool, device=device)
    return same & ~eye, ~same


def _avg_nonzero(x):
    return x.sum() / (x > 0).sum().clamp_min(1)          # no host sync: mean over the non-zero terms, 0 if none


def contrastive_loss(cmr, mp, pos_margin=0.0, neg_margin=1.0):
    """losses.ContrastiveLoss() defaults: L2 distance between L2-normalised embeddings; every positive pair costs
    relu(d - pos_margin), every negative pair relu(neg_margin - d); each group is averaged over its non-zero terms
    (AvgNonZeroReducer) and the two averages are summed."""
    e = F.normalize(torch.cat((cmr.flatten(1), mp.flatten(1)), 0), dim=1)
    d = torch.cdist(e, e)
    pos, neg = _pairs(cmr.shape[0], e.device)
    return _avg_nonzero(F.relu(d[pos] - pos_margin)) + _avg_nonzero(F.relu(neg_margin - d[neg]))


def triplet_loss(cmr, mp, margin=0.05, miner_margin=0.2, high=0.3):
    """TripletMarginLoss(distance=CosineSimilarity(), reducer=ThresholdReducer(high=0.3),
    embedding_regularizer=LpRegularizer()) on the triplets of TripletMarginMiner(margin=0.2, "semihard").
    Miner (default L2 distance on normalised embeddings): triplets with 0 < d_an - d_ap <= 0.2.  Loss per mined
    triplet: relu(sim_an - sim_ap + margin); the reducer averages the terms below ``high`` (zero if none).  The
    regulariser adds the embeddings' L2 norms through the same reducer."""
    raw = torch.cat((cmr.flatten(1), mp.flatten(1)), 0)
    e = F.normalize(raw, dim=1)
    pos, neg = _pairs(cmr.shape[0], e.device)
    with torch.no_grad():
        d = torch.cdist(e, e)
        gap = d[:, None, :] - d[:, :, None]                       # [a, p, n] = d_an - d_ap
        mined = pos[:, :, None] & neg[:, None, :] & (gap > 0) & (gap <= miner_margin)
    sim = e @ e.T
    viol = F.relu(sim[:, None, :] - sim[:, :, None] + margin)

    def threshold_mean(x, keep):
        return (x * keep).sum() / keep.sum().clamp_min(1)

    norms = raw.norm(dim=1)
    return threshold_mean(viol, mined & (viol < high)) + threshold_mean(norms, norms < high)

"""Retrieval recall: the oracle is pinned by ``tests/golden/recall.npz`` -- the output of the
reference's own ``Trainer.get_recall`` (train.py:551-572), lifted out of train.py with ``ast`` by
tests/golden/make_golden.py because the module itself cannot be imported -- and the device implementation (cuBLAS GEMM
+ the rank-counting kernel) is checked against the oracle."""
import numpy as np
import pytest
import torch

from oracle import recall as orec


def test_oracle_recall_sanity():
    x = np.eye(6, 16, dtype=np.float32)
    assert orec.get_recall(x, x) == (100.0, 100.0, 100.0)           # perfect retrieval
    y = np.roll(x, 1, axis=0)
    assert orec.get_recall(x, y)[0] == 0.0                           # every true match is beaten by another row


def test_oracle_matches_reference_get_recall():
    from util import load_golden
    g, _, _ = load_golden("recall")
    assert np.allclose(orec.get_recall(g["camera"], g["maps"]), g["recall"], atol=0, rtol=0)
    assert 0.0 < g["recall"][0] < g["recall"][2] <= 100.0          # a non-trivial case: R@1 < R@10


@pytest.mark.gpu
def test_recall_matches_oracle():
    import bevrender_b200 as bb
    rng = np.random.default_rng(0)
    maps = rng.standard_normal((200, 512)).astype(np.float32)
    maps /= np.linalg.norm(maps, axis=1, keepdims=True)
    cam = maps + 0.5 * rng.standard_normal(maps.shape).astype(np.float32)
    cam /= np.linalg.norm(cam, axis=1, keepdims=True)
    ref = orec.get_recall(cam, maps)
    got = bb.retrieval_recall(torch.from_numpy(cam).cuda(), torch.from_numpy(maps).cuda())
    assert 0.0 < ref[0] < 100.0
    # the ranks themselves: identical unless two fp32 distances tie to the last bit (cuBLAS vs numpy summation order)
    dist = 2.0 - 2.0 * np.matmul(cam, maps.T)
    want = np.array([np.sum(dist[:, k] < dist[k, k]) for k in range(dist.shape[0])])
    rank = bb.retrieval_ranks(torch.from_numpy(cam).cuda(), torch.from_numpy(maps).cuda()).cpu().numpy()
    assert rank.dtype == np.int32 and (rank != want).sum() <= 1
    assert np.allclose(got, ref, atol=100.0 / cam.shape[0])          # at most that one query


@pytest.mark.gpu
def test_recall_golden_and_odd_sizes():
    import bevrender_b200 as bb
    from util import load_golden
    g, _, _ = load_golden("recall")
    got = bb.retrieval_recall(torch.from_numpy(g["camera"]).cuda(), torch.from_numpy(g["maps"]).cuda())
    assert np.allclose(got, g["recall"], atol=100.0 / 96)
    rng = np.random.default_rng(1)
    for n in (1, 33, 257):                                             # ragged column blocks / row splits
        x = rng.standard_normal((n, 40)).astype(np.float32)
        y = x + 0.3 * rng.standard_normal(x.shape).astype(np.float32)
        dist = 2.0 - 2.0 * np.matmul(x, y.T)
        want = np.array([np.sum(dist[:, k] < dist[k, k]) for k in range(n)])
        rank = bb.retrieval_ranks(torch.from_numpy(x).cuda(), torch.from_numpy(y).cuda()).cpu().numpy()
        assert (rank != want).sum() <= max(1, n // 100)

"""Retrieval recall: device implementation vs the numpy restatement of train.py:551-572."""
import numpy as np
import pytest
import torch

from oracle import recall as orec


def test_oracle_recall_sanity():
    x = np.eye(6, 16, dtype=np.float32)
    assert orec.get_recall(x, x) == (100.0, 100.0, 100.0)           # perfect retrieval
    y = np.roll(x, 1, axis=0)
    assert orec.get_recall(x, y)[0] == 0.0                           # every true match is beaten by another row


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
    assert np.allclose(got, ref, atol=1.0)       # a rank can flip on fp32 ties: 1 query of 200 = 0.5 %

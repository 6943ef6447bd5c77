"""Oracle for the retrieval recall metric ("next" row f-4): numpy restatement of ``Trainer.get_recall``
(train.py:551-572).  ``train.py`` itself cannot be imported (pytorch_metric_learning / yourdfpy missing); the method
is numpy-only, so ``tests/golden/make_golden.py`` cuts it out of train.py with ``ast``, runs it, and commits its output
as ``tests/golden/recall.npz``: this restatement is pinned against that (``tests/test_recall.py``)."""
import numpy as np


def get_recall(camera, maps):
    dist = 2.0 - 2.0 * np.matmul(camera, maps.T)
    n = dist.shape[0]
    rank = np.array([np.sum(dist[:, k] < dist[k, k]) for k in range(n)])
    acc = [float(np.mean(rank < i)) for i in range(11)]
    return acc[1] * 100, acc[5] * 100, acc[10] * 100

"""CPU (gloo, world_size 2) test of the flat-bucket gradient allreduce: an N-rank step on a sharded batch
yields the gradients of a 1-rank step on the full batch."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp
from torch import nn


class TSADeformableAttention(nn.Module):
    """Stand-in with the reference's dead projection: the filter keys on the owning class, not on the bare name."""

    def __init__(self):
        super().__init__()
        self.proj_q = nn.Linear(4, 4)            # never applied, as in the reference
        self.proj_k = nn.Linear(16, 8)

    def forward(self, x):
        return self.proj_k(x)


class Tiny(nn.Module):
    def __init__(self):
        super().__init__()
        self.a = nn.Linear(6, 16)
        self.attn = TSADeformableAttention()
        self.proj_q = nn.Linear(8, 8)            # a caller's own, LIVE, proj_q
        self.unused = nn.Linear(3, 3)            # built but never called (the reference's FeedForwardLayer)
        self.c = nn.Linear(8, 3)

    def forward(self, x):
        return self.c(torch.tanh(self.proj_q(self.attn(torch.relu(self.a(x))))))


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from bevrender_b200.dp import FlatGradAllReduce, find_live_parameters, live_named_parameters
    torch.manual_seed(rank)                       # ranks start from DIFFERENT weights: the reducer broadcasts rank 0's
    model = Tiny()
    torch.manual_seed(0)
    x = torch.randn(8, 6)
    y = torch.randn(8, 3)
    names = [n for n, _ in live_named_parameters(model)]
    assert "proj_q.weight" in names and "attn.proj_q.weight" not in names and "unused.weight" in names
    live_named = find_live_parameters(model, lambda: ((model(x) - y) ** 2).mean().backward())
    assert sorted(n for n, _ in live_named) == sorted(n for n in names if not n.startswith("unused"))
    live = [p for _, p in live_named]
    red = FlatGradAllReduce(live, n_buckets=3)
    assert len(red.buckets) >= 2
    shard = slice(rank * 4, rank * 4 + 4)
    for step in range(2):                         # second step checks reset()/re-arming
        red.reset()
        ((model(x[shard]) - y[shard]) ** 2).mean().backward()
        red.finish()
    if rank == 0:
        torch.save({n: p.grad.clone() for n, p in live_named}, out)
    dist.destroy_process_group()


def test_flat_bucket_allreduce_matches_full_batch(tmp_path):
    out = str(tmp_path / "grads.pt")
    mp.spawn(_worker, args=(2, _free_port(), out), nprocs=2, join=True)
    got = torch.load(out)
    torch.manual_seed(0)                          # rank 0's weights
    model = Tiny()
    torch.manual_seed(0)
    x = torch.randn(8, 6)
    y = torch.randn(8, 3)
    ((model(x) - y) ** 2).mean().backward()
    for n, p in model.named_parameters():
        if n.startswith(("attn.proj_q", "unused")):
            assert n not in got
            continue
        assert torch.allclose(got[n], p.grad, atol=1e-6), n

"""Batch-data-parallel gradient exchange for the training step (K6).

The reference wraps the model in ``DistributedDataParallel(find_unused_parameters=True)`` +
``SyncBatchNorm`` (train.py:128-141).  The gradient it exchanges is 8-16 MB of fp32, i.e. a
latency-bound message on NVLink 5 / NVSwitch -- so instead of DDP's 25 MB bucket machinery, unused-
parameter bitmap allreduce and per-iteration buffer broadcast, the B200 design is:

  * every LIVE parameter's ``.grad`` is a view into ONE flat, pre-allocated fp32 buffer (autograd
    accumulates in place, no flatten/unflatten copies);
  * the buffer is cut into a few contiguous buckets in backward order; a post-accumulate-grad hook counts
    a bucket's parameters down and launches one NCCL ``all_reduce(AVG)`` for the bucket as soon as it is
    complete, so the wire time of the decoder/late-layer buckets overlaps the rest of backward;
  * parameters that never receive a gradient (the reference's dead ``proj_q``, ``proj_views``,
    ``conv_offset_m1/m2``) are excluded up front, so there is no unused-parameter
    detection at all.

One process per GPU; NCCL over NVLink for CUDA tensors, gloo for the CPU tests.
"""
import torch
import torch.distributed as dist

# sub-modules of TSADeformableAttention / SCADeformableAttention that exist only for state_dict compatibility
# (SURVEY.md appendix B): never applied in forward, so their parameters never receive a gradient
DEAD_SUBMODULES = ("proj_q", "proj_views", "conv_offset_m1", "conv_offset_m2")


def live_named_parameters(module):
    """Trainable parameters minus the ones the deformable-attention modules never use.  Only parameters OWNED by a
    TSADeformableAttention / SCADeformableAttention instance (this package's or the reference's) are filtered: a
    caller's own ``proj_q`` elsewhere in the model stays live."""
    dead = set()
    for m in module.modules():
        if type(m).__name__ in ("TSADeformableAttention", "SCADeformableAttention"):
            for name in DEAD_SUBMODULES:
                sub = getattr(m, name, None)
                if isinstance(sub, torch.nn.Module):
                    dead.update(id(p) for p in sub.parameters())
    return [(n, p) for n, p in module.named_parameters() if p.requires_grad and id(p) not in dead]


def find_live_parameters(module, run_step):
    """Static replacement for DDP's ``find_unused_parameters=True`` (train.py:134): run ONE forward + backward
    (``run_step()`` must call ``.backward()``) and return the (name, parameter) pairs that received a gradient.  The
    reference model carries parameters no forward ever touches (FeedForwardLayer is built but not called,
    encoder.py:291-296,383,404; the dead projections above): DDP re-discovers them with a bitmap allreduce every
    iteration, here they are excluded once, before the flat bucket is laid out."""
    for p in module.parameters():
        p.grad = None
    run_step()
    live = [(n, p) for n, p in module.named_parameters() if p.requires_grad and p.grad is not None]
    for p in module.parameters():
        p.grad = None
    return live


class FlatGradAllReduce:
    def __init__(self, params, n_buckets=2, group=None, average=True):
        self.params = list(params)
        self.group = group
        self.average = average
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        if self.world > 1:                     # as DDP does at construction: every rank starts from rank 0's weights
            with torch.no_grad():
                for p in self.params:
                    dist.broadcast(p.data, src=dist.get_global_rank(group, 0) if group is not None else 0, group=group)
        dev, total = self.params[0].device, sum(p.numel() for p in self.params)
        self.flat = torch.zeros(total, dtype=torch.float32, device=dev)
        # lay parameters out in REVERSE registration order: the last layers' grads are ready first and
        # occupy the front of the buffer, so each bucket is one contiguous slice
        self.slices, off = {}, 0
        order = list(reversed(self.params))
        for p in order:
            n = p.numel()
            p.grad = self.flat[off:off + n].view_as(p)
            self.slices[p] = (off, off + n)
            off += n
        per = (total + n_buckets - 1) // n_buckets
        self.buckets, cur, start = [], [], 0
        for p in order:
            cur.append(p)
            end = self.slices[p][1]
            if end - start >= per or p is order[-1]:
                self.buckets.append({"params": cur, "range": (start, end)})
                cur, start = [], end
        self._bucket_of = {p: i for i, b in enumerate(self.buckets) for p in b["params"]}
        self._pending = [0] * len(self.buckets)
        self._works = []
        self._hooks = [p.register_post_accumulate_grad_hook(self._on_grad) for p in self.params]
        self.reset()

    def reset(self):
        """Call before backward: zero the flat gradient and re-arm the buckets."""
        self.flat.zero_()
        self._pending = [len(b["params"]) for b in self.buckets]
        self._works = []

    def _on_grad(self, p):
        if p.grad.data_ptr() != self.flat.data_ptr() + self.slices[p][0] * 4:
            raise RuntimeError("a parameter's .grad was replaced; FlatGradAllReduce needs in-place accumulation "
                               "(use reset() instead of zero_grad(set_to_none=True))")
        i = self._bucket_of[p]
        self._pending[i] -= 1
        if self._pending[i] == 0 and self.world > 1:
            lo, hi = self.buckets[i]["range"]
            op = dist.ReduceOp.AVG if (self.average and self.flat.is_cuda) else dist.ReduceOp.SUM
            self._works.append((dist.all_reduce(self.flat[lo:hi], op=op, group=self.group, async_op=True), op, lo, hi))

    def finish(self):
        """Call after backward: wait for the bucket allreduces (and average on backends without AVG)."""
        if any(n > 0 for n in self._pending) and self.world > 1:
            missing = [i for i, n in enumerate(self._pending) if n > 0]
            raise RuntimeError(f"buckets {missing} never completed: a 'live' parameter got no gradient")
        for work, op, lo, hi in self._works:
            work.wait()
            if self.average and op == dist.ReduceOp.SUM:
                self.flat[lo:hi].div_(self.world)
        self._works = []

    def remove(self):
        for h in self._hooks:
            h.remove()

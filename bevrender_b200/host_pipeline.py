"""Host-buffer entry point of the MSDA gather: pinned HOST inputs -> device -> fwd+bwd -> pinned HOST results.

This is the call a user holding host-resident tensors makes (``bench.py``'s ``e2e`` leg).  The batch flows through
three CUDA streams (H2D, compute through the C ABI, D2H) over ``nslots`` device staging slots, so the PCIe
uploads of the next piece and the downloads of the previous one overlap the kernels of the current one (PCIe is
full duplex; the kernels are ~1/10 of the transfer time).  Two slots suffice: three or four measure the same
30.8 ms/step (the link itself gives 26.1 ms for 1.25 GB each way at once, 22.6 ms one way, scripts/exp_pcie.py).

Pieces: a sample's ``value`` goes up first, then its queries in ``qsplit`` row blocks (coordinates, weights,
``grad_out`` up; ``out``, ``grad_loc``, ``grad_weight`` down); every block's backward accumulates into the sample's
``grad_value`` (``BEVR_ACCUMULATE``), which goes down after the last block.  Finer pieces shorten the two half-duplex
phases of the pipeline (first upload, last download) from one sample to ``1/qsplit`` of one.
"""
import ctypes as C

import torch

from . import _lib

_NAMES_IN = ("value", "sampling_locations", "attention_weights", "grad_out")
_NAMES_OUT = ("out", "grad_value", "grad_loc", "grad_weight")
_Q_IN = ("sampling_locations", "attention_weights", "grad_out")       # per-query inputs
_Q_OUT = ("out", "grad_loc", "grad_weight")                           # per-query results
_ACCUMULATE = 8


class HostMsdaPipeline:
    def __init__(self, spatial_shapes, level_start_index, h, Dh, Nq, P, query_w, dtype=torch.float32,
                 device=None, chunk=1, flags=0, qsplit=4, nslots=2):
        self.dev = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        self.shapes, self.L = list(spatial_shapes), len(spatial_shapes)
        self.S = sum(hh * ww for hh, ww in spatial_shapes)
        self.h, self.Dh, self.Nq, self.P, self.query_w, self.flags = h, Dh, Nq, P, query_w, flags
        self.dtype, self.dt, self.chunk = dtype, (0 if dtype == torch.float32 else 1), chunk
        self.hw = (C.c_int32 * (2 * self.L))(*[v for s in self.shapes for v in s])
        self.st = (C.c_int32 * self.L)(*level_start_index)
        # query row blocks: whole rows of the 2-D query grid so the kernels keep their 2-D tiling; bf16 results need
        # the fp32 workspace to be cast once per sample, so only fp32 storage splits the queries
        rows = Nq // query_w if query_w > 0 else 1
        nblk = max(1, min(qsplit, rows)) if (self.dt == 0 and query_w > 0) else 1
        per = -(-rows // nblk)
        per = -(-per // 8) * 8 if per > 8 else per          # multiples of the 8-row query tile
        self.qblocks = [(r * query_w, min(rows, r + per) * query_w) for r in range(0, rows, per)] if query_w > 0 \
            else [(0, Nq)]
        c, f32 = chunk, torch.float32
        shp_in = {"value": ((c, self.S, h, Dh), dtype), "sampling_locations": ((c, Nq, h, self.L, P, 2), f32),
                  "attention_weights": ((c, Nq, h, self.L, P), dtype), "grad_out": ((c, Nq, h * Dh), dtype)}
        shp_out = {"out": ((c, Nq, h * Dh), dtype), "grad_value": ((c, self.S, h, Dh), dtype),
                   "grad_loc": ((c, Nq, h, self.L, P, 2), f32), "grad_weight": ((c, Nq, h, self.L, P), dtype)}
        mk = lambda d: {k: torch.empty(s, dtype=t, device=self.dev) for k, (s, t) in d.items()}
        self.nslots = nslots
        self.slots = [{"in": mk(shp_in), "out": mk(shp_out)} for _ in range(nslots)]
        nws = _lib.lib().bevr_msda_bwd_workspace_bytes(c, self.S, h, Dh, self.dt)
        self.ws = [torch.empty(max(nws, 16), dtype=torch.uint8, device=self.dev) for _ in range(nslots)]
        self.s_in, self.s_cmp, self.s_out = (torch.cuda.Stream(self.dev) for _ in range(3))
        self.ev_out = [None] * nslots          # per staging slot: its last download (also across calls of run)
        self._forked = False

    def bytes_per_sample(self):
        n = lambda d: sum(t.numel() * t.element_size() for t in d.values()) // self.chunk
        return n(self.slots[0]["in"]), n(self.slots[0]["out"])

    def join(self):
        """Make the caller's current stream wait for everything ``run(..., join=False)`` has enqueued."""
        cur = torch.cuda.current_stream(self.dev)
        for s in (self.s_in, self.s_cmp, self.s_out):
            cur.wait_stream(s)
        self._forked = False

    def run(self, h_in, h_out, join=True):
        """h_in / h_out: dicts of PINNED host tensors with a leading batch dimension (B % chunk == 0).

        ``join=False`` returns without making the current stream wait for the pipeline: a following ``run`` then starts
        uploading while this call's last results are still going down (the staging slots are handed over with events), as
        a training loop that prefetches its next batch does.  Call ``join()`` (or ``run(join=True)``) before reading
        ``h_out`` of the earlier calls from the current stream / after a synchronize."""
        L = _lib.lib()
        B = h_in["value"].shape[0]
        P = lambda t: C.c_void_p(t.data_ptr())
        cur = torch.cuda.current_stream(self.dev)
        if not self._forked:                                       # first call since the last join: order after the caller
            for s in (self.s_in, self.s_cmp, self.s_out):
                s.wait_stream(cur)
            self._forked = True
        ns = self.nslots
        ev_out = self.ev_out
        c = self.chunk
        split = len(self.qblocks) > 1 and c == 1
        blocks = self.qblocks if split else [(0, self.Nq)]
        cs = C.c_void_p(self.s_cmp.cuda_stream)
        for i in range(B // c):
            slot, sl = self.slots[i % ns], slice(i * c, (i + 1) * c)
            a, o = slot["in"], slot["out"]
            with torch.cuda.stream(self.s_in):
                if ev_out[i % ns] is not None:
                    self.s_in.wait_event(ev_out[i % ns])           # slot drained by chunk i - nslots
                a["value"].copy_(h_in["value"][sl], non_blocking=True)
            for bi, (q0, q1) in enumerate(blocks):
                with torch.cuda.stream(self.s_in):
                    for k in _Q_IN:
                        a[k][:, q0:q1].copy_(h_in[k][sl, q0:q1], non_blocking=True)
                    ev_in = self.s_in.record_event()
                with torch.cuda.stream(self.s_cmp):
                    self.s_cmp.wait_event(ev_in)
                    nq = q1 - q0
                    loc, w, go = a["sampling_locations"][:, q0:q1], a["attention_weights"][:, q0:q1], a["grad_out"][:, q0:q1]
                    _lib.check(L.bevr_msda_fwd(P(a["value"]), self.hw, self.st, P(loc), None, P(w), P(o["out"][:, q0:q1]),
                                               c, self.S, self.h, self.Dh, nq, self.L, self.P, self.query_w, self.dt,
                                               self.flags, cs), "bevr_msda_fwd")
                    _lib.check(L.bevr_msda_bwd(P(a["value"]), self.hw, self.st, P(loc), None, P(w), P(go),
                                               P(o["grad_value"]), P(o["grad_loc"][:, q0:q1]),
                                               P(o["grad_weight"][:, q0:q1]), P(self.ws[i % ns]), c, self.S, self.h,
                                               self.Dh, nq, self.L, self.P, self.query_w, self.dt,
                                               self.flags | (_ACCUMULATE if bi > 0 else 0), cs), "bevr_msda_bwd")
                    ev_cmp = self.s_cmp.record_event()
                with torch.cuda.stream(self.s_out):
                    self.s_out.wait_event(ev_cmp)
                    for k in _Q_OUT:
                        h_out[k][sl, q0:q1].copy_(o[k][:, q0:q1], non_blocking=True)
                    if bi == len(blocks) - 1:
                        h_out["grad_value"][sl].copy_(o["grad_value"], non_blocking=True)
                        ev_out[i % ns] = self.s_out.record_event()
        if join:
            self.join()

"""Host-buffer entry point of the MSDA gather: pinned HOST inputs -> device -> fwd+bwd -> pinned HOST results.

This is the call a user holding host-resident tensors makes (``bench.py``'s ``e2e`` leg).  The batch is cut
into per-sample chunks that flow through three CUDA streams (H2D, compute through the C ABI, D2H) over a
double-buffered device staging area, so the PCIe uploads of chunk i+1 and the downloads of chunk i-1 overlap
the kernels of chunk i (PCIe is full duplex; the kernels are ~1/10 of the transfer time).
"""
import ctypes as C

import torch

from . import _lib

_NAMES_IN = ("value", "sampling_locations", "attention_weights", "grad_out")
_NAMES_OUT = ("out", "grad_value", "grad_loc", "grad_weight")


class HostMsdaPipeline:
    def __init__(self, spatial_shapes, level_start_index, h, Dh, Nq, P, query_w, dtype=torch.float32,
                 device=None, chunk=1, flags=0):
        self.dev = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        self.shapes, self.L = list(spatial_shapes), len(spatial_shapes)
        self.S = sum(hh * ww for hh, ww in spatial_shapes)
        self.h, self.Dh, self.Nq, self.P, self.query_w, self.flags = h, Dh, Nq, P, query_w, flags
        self.dtype, self.dt, self.chunk = dtype, (0 if dtype == torch.float32 else 1), chunk
        self.hw = (C.c_int32 * (2 * self.L))(*[v for s in self.shapes for v in s])
        self.st = (C.c_int32 * self.L)(*level_start_index)
        c, f32 = chunk, torch.float32
        shp_in = {"value": ((c, self.S, h, Dh), dtype), "sampling_locations": ((c, Nq, h, self.L, P, 2), f32),
                  "attention_weights": ((c, Nq, h, self.L, P), dtype), "grad_out": ((c, Nq, h * Dh), dtype)}
        shp_out = {"out": ((c, Nq, h * Dh), dtype), "grad_value": ((c, self.S, h, Dh), dtype),
                   "grad_loc": ((c, Nq, h, self.L, P, 2), f32), "grad_weight": ((c, Nq, h, self.L, P), dtype)}
        mk = lambda d: {k: torch.empty(s, dtype=t, device=self.dev) for k, (s, t) in d.items()}
        self.slots = [{"in": mk(shp_in), "out": mk(shp_out)} for _ in range(2)]
        nws = _lib.lib().bevr_msda_bwd_workspace_bytes(c, self.S, h, Dh, self.dt)
        self.ws = [torch.empty(max(nws, 16), dtype=torch.uint8, device=self.dev) for _ in range(2)]
        self.s_in, self.s_cmp, self.s_out = (torch.cuda.Stream(self.dev) for _ in range(3))

    def bytes_per_sample(self):
        n = lambda d: sum(t.numel() * t.element_size() for t in d.values()) // self.chunk
        return n(self.slots[0]["in"]), n(self.slots[0]["out"])

    def run(self, h_in, h_out):
        """h_in / h_out: dicts of PINNED host tensors with a leading batch dimension (B % chunk == 0)."""
        L = _lib.lib()
        B = h_in["value"].shape[0]
        P = lambda t: C.c_void_p(t.data_ptr())
        cur = torch.cuda.current_stream(self.dev)
        for s in (self.s_in, self.s_cmp, self.s_out):
            s.wait_stream(cur)
        ev_out = [None, None]
        c = self.chunk
        for i in range(B // c):
            slot, sl = self.slots[i % 2], slice(i * c, (i + 1) * c)
            with torch.cuda.stream(self.s_in):
                if ev_out[i % 2] is not None:
                    self.s_in.wait_event(ev_out[i % 2])            # slot drained by chunk i-2
                for k in _NAMES_IN:
                    slot["in"][k].copy_(h_in[k][sl], non_blocking=True)
                ev_in = self.s_in.record_event()
            with torch.cuda.stream(self.s_cmp):
                self.s_cmp.wait_event(ev_in)
                a, o = slot["in"], slot["out"]
                cs = C.c_void_p(self.s_cmp.cuda_stream)
                _lib.check(L.bevr_msda_fwd(P(a["value"]), self.hw, self.st, P(a["sampling_locations"]), None,
                                           P(a["attention_weights"]), P(o["out"]), c, self.S, self.h, self.Dh,
                                           self.Nq, self.L, self.P, self.query_w, self.dt, self.flags, cs),
                           "bevr_msda_fwd")
                _lib.check(L.bevr_msda_bwd(P(a["value"]), self.hw, self.st, P(a["sampling_locations"]), None,
                                           P(a["attention_weights"]), P(a["grad_out"]), P(o["grad_value"]),
                                           P(o["grad_loc"]), P(o["grad_weight"]), P(self.ws[i % 2]), c, self.S,
                                           self.h, self.Dh, self.Nq, self.L, self.P, self.query_w, self.dt,
                                           self.flags, cs), "bevr_msda_bwd")
                ev_cmp = self.s_cmp.record_event()
            with torch.cuda.stream(self.s_out):
                self.s_out.wait_event(ev_cmp)
                for k in _NAMES_OUT:
                    h_out[k][sl].copy_(slot["out"][k], non_blocking=True)
                ev_out[i % 2] = self.s_out.record_event()
        for s in (self.s_in, self.s_cmp, self.s_out):
            cur.wait_stream(s)

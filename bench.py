#!/usr/bin/env python
"""Benchmark of the BEVRender deformable-attention hot path (BASELINE.json config 2).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--dtype fp32|bf16] [--impl reference]

A "step" is one forward + backward pass of the multi-scale deformable-attention gather over one
batch of B=8 synthetic samples per GPU (200x200 BEV queries, 4 levels, 8 heads, 4 points, Dh=32),
through the C ABI of libbevr.so.  Prints ONE JSON line (see DESIGN.md "Measurement").
"""
import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

N_SIDE, HEADS, DH, LEVELS, POINTS = 200, 8, 32, 4, 4
METRIC = "BEV queries/s (deform-attn fwd+bwd)"


def algorithmic_bytes(B, esize):
    """every input read once, every output written once at storage dtype
    (locations always fp32).  Returns (fwd, bwd) bytes for B samples."""
    Nq = N_SIDE * N_SIDE
    S = sum((N_SIDE >> l) ** 2 for l in range(LEVELS))
    value = S * HEADS * DH * esize
    loc = Nq * HEADS * LEVELS * POINTS * 2 * 4
    w = Nq * HEADS * LEVELS * POINTS * esize
    out = Nq * HEADS * DH * esize
    fwd = value + loc + w + out
    bwd = (out + value + loc + w) + (value + loc + w)
    return B * fwd, B * bwd


def peaks():
    try:
        p = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        return float(p["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


class ClockSampler:
    """Samples SM clocks / throttle reasons through NVML while the timed region runs (nvidia-smi as fallback)."""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
    BITS = {"hw_slowdown": 0x8, "hw_thermal_slowdown": 0x40, "sw_thermal_slowdown": 0x20, "sw_power_cap": 0x4}

    def __init__(self, index):
        self.index, self.sm, self.mx, self.reasons, self.stop = index, [], [], set(), False
        self.nvml = None
        try:
            import pynvml
            pynvml.nvmlInit()
            vis = os.environ.get("CUDA_VISIBLE_DEVICES")
            phys = int(vis.split(",")[index]) if vis and vis.split(",")[index].isdigit() else index
            self.h = pynvml.nvmlDeviceGetHandleByIndex(phys)
            self.nvml = pynvml
        except Exception:
            self.nvml = None
        self.t = threading.Thread(target=self.run, daemon=True)

    def sample(self):
        if self.nvml is not None:
            n = self.nvml
            self.sm.append(n.nvmlDeviceGetClockInfo(self.h, n.NVML_CLOCK_SM))
            self.mx.append(n.nvmlDeviceGetMaxClockInfo(self.h, n.NVML_CLOCK_SM))
            try:
                r = n.nvmlDeviceGetCurrentClocksEventReasons(self.h)
            except Exception:
                r = n.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
            self.reasons |= {k for k, bit in self.BITS.items() if r & bit}
            return
        o = subprocess.run(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                            "--format=csv,noheader,nounits"], capture_output=True, text=True, timeout=5)
        c = [v.strip() for v in o.stdout.strip().split(",")]
        if len(c) >= 6 and c[0].isdigit():
            self.sm.append(int(c[0]))
            self.mx.append(int(c[1]))
            self.reasons |= {k for k, v in zip(self.BITS, c[2:6]) if v == "Active"}

    def run(self):
        while not self.stop:
            try:
                self.sample()
            except Exception:
                pass
            time.sleep(0.002)

    def __enter__(self):
        self.t.start()
        return self

    def __exit__(self, *a):
        self.stop = True
        self.t.join(timeout=6)

    def summary(self):
        sm = sorted(self.sm)
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(self.mx) if self.mx else None,
                "reasons": sorted(self.reasons), "samples": len(sm), "source": "nvml" if self.nvml else "nvidia-smi"}


_JSON_OUT = None


def claim_stdout():
    """stdout carries exactly ONE JSON line: keep a private handle on the real stdout and point file descriptor 1 at
    stderr, so anything a library prints there (NCCL's "NCCL version ..." banner under torchrun) cannot get in."""
    global _JSON_OUT
    if _JSON_OUT is None:
        sys.stdout.flush()
        _JSON_OUT = os.fdopen(os.dup(1), "w")
        os.dup2(2, 1)


def emit(line):
    out = _JSON_OUT if _JSON_OUT is not None else sys.stdout
    out.write(json.dumps(line) + "\n")
    out.flush()


def bind_to_gpu_numa_node(local_rank):
    """Pin this process to the CPUs NVML reports as closest to its GPU BEFORE any pinned host buffer is
    allocated, so the staging pages of the e2e leg are first-touched on the GPU's own NUMA node (with 8 ranks
    on one box, remote-socket staging memory halves the aggregate PCIe throughput)."""
    try:
        import pynvml
        pynvml.nvmlInit()
        vis = os.environ.get("CUDA_VISIBLE_DEVICES")
        phys = int(vis.split(",")[local_rank]) if vis and vis.split(",")[local_rank].isdigit() else local_rank
        pynvml.nvmlDeviceSetCpuAffinity(pynvml.nvmlDeviceGetHandleByIndex(phys))
        return sorted(os.sched_getaffinity(0))
    except Exception:
        return None


def cpu_reference_step(B_sample, dtype_name, threads):
    """One fwd+bwd of the CPU oracle (grid_sample-based MSDA, the construction the reference uses)
    on B_sample samples of the config-2 workload.  Returns seconds."""
    from bevrender_b200.workloads import make_microbench_inputs
    from oracle import msda as omsda          # cpu_baseline / --impl reference leg only
    torch.set_num_threads(threads)
    d = make_microbench_inputs(B_sample, seed=0)
    t0 = time.perf_counter()
    omsda.msda_forward_backward(d["value"], d["spatial_shapes"], d["level_start_index"], d["sampling_locations"],
                                d["attention_weights"], d["grad_out"])
    return time.perf_counter() - t0


def run_reference(args):
    """--impl reference: the reference-style CPU path (torch grid_sample + autograd, all host threads),
    timed on the box's host cores.  The reference itself has no multi-level op: this is
    the oracle port, kind='port'."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    Nq = N_SIDE * N_SIDE
    for _ in range(min(args.warmup, 1)):
        cpu_reference_step(1, args.dtype, cores)
    steps = max(1, min(args.steps, 3))
    ts = [cpu_reference_step(1, args.dtype, cores) for _ in range(steps)]
    t = sum(ts) / len(ts)
    v = Nq / t
    line = {"impl": "reference", "metric": METRIC, "value": v, "unit": "queries/s", "n_gpus": args.gpus,
            "steps": steps, "warmup": min(args.warmup, 1), "ms_per_step": t * 1e3, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": workload_config(1, "fp32"),
            "cpu_baseline": {"value": v, "unit": "queries/s", "cores": cores, "kind": "port",
                             "sample": "1 of 8 samples per step (40000 queries), fwd+bwd, torch CPU grid_sample path"},
            "e2e": {"value": v, "unit": "queries/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    emit(line)


def workload_config(B, dtype):
    return {"workload": "config2: MSDA fwd+bwd, 200x200 BEV queries, levels 200/100/50/25, 8 heads x 32 ch, 4 points",
            "batch_per_gpu": B, "storage_dtype": dtype, "l2": "working set per step (3.4 GB fp32) exceeds the 126 MB L2; "
            "no explicit flush", "parallelism": "batch-sharded, no data-path collective"}


def run_train(args):
    """--workload train (BASELINE.json config 5): end-to-end training step of the caller-side harness
    (stem -> 6 x [TSA, SCA] encoder layers with the drop-in modules -> render head; MSE loss, AdamW, clip 1.0;
    window of 4 frames = 3 history passes without gradient + 1 pass with gradient), batch-data-parallel with the
    flat-bucket NCCL gradient allreduce.  Metric: training samples/s, whole job."""
    import torch.distributed as dist
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    from bevrender_b200.dp import FlatGradAllReduce, live_named_parameters
    from bevrender_b200.harness import BevHarness
    torch.backends.cudnn.allow_tf32 = True           # the PyTorch convolutions of the harness, as torch defaults
    torch.manual_seed(0)
    B, T, n = args.batch, 4, args.bev
    model = BevHarness(n=n, c=64, heads=2, depth=5, layers=6, batch=B, device=dev).to(dev)
    live = [p for _, p in live_named_parameters(model)]
    for name, p in model.named_parameters():
        if all(p is not q for q in live):
            p.requires_grad_(False)
    opt = torch.optim.AdamW(live, lr=1e-4, weight_decay=1e-2)
    red = FlatGradAllReduce(live, n_buckets=2)
    g = torch.Generator().manual_seed(1 + rank)
    frames_h = torch.rand(B, T, 3, 128, 160, generator=g).pin_memory()
    target_h = torch.rand(B, 3, 224, 224, generator=g).pin_memory()

    def step():
        frames = frames_h.to(dev, non_blocking=True)
        target = target_h.to(dev, non_blocking=True)
        red.reset()
        loss = torch.nn.functional.mse_loss(model(frames), target)
        loss.backward()
        red.finish()
        torch.nn.utils.clip_grad_norm_(live, 1.0)
        opt.step()
        return loss

    for _ in range(max(args.warmup, 3)):
        step()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with ClockSampler(local) as clk:
        e0.record()
        for _ in range(args.steps):
            loss = step()
        e1.record()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    if world > 1:
        t = torch.tensor([ms], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    ms /= args.steps
    if rank == 0:
        nparam = sum(p.numel() for p in live)
        emit({"metric": "train samples/s", "value": world * B / (ms * 1e-3), "unit": "samples/s",
                          "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": ms,
                          "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
                          "data": "synthetic", "loss": float(loss.detach()),
                          "config": {"workload": "config5: harness training step, bev %d, 6 layers, window 4, "
                                     "image 128x160, C=64, 2 heads, D=5" % n, "batch_per_gpu": B,
                                     "parallelism": "dp%d, flat fp32 gradient bucket (%d params, %.1f MB), 2 NCCL "
                                     "allreduce(AVG) launches overlapped with backward" % (world, nparam, nparam * 4 / 1e6)},
                          "clocks": clk.summary()})
    if world > 1:
        dist.destroy_process_group()


def main():
    claim_stdout()
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--dtype", default="fp32", choices=["fp32", "bf16"])
    ap.add_argument("--batch", type=int, default=8)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true", help="tuning sweeps only: skip the host-buffer leg")
    ap.add_argument("--tile", default="", help="tuning: tile_w,tile_h,threads")
    ap.add_argument("--workload", default="msda", choices=["msda", "train"],
                    help="msda = BASELINE config 2 (the contract line); train = config 5 harness training step")
    ap.add_argument("--bev", type=int, default=56, help="--workload train: BEV grid size")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)
    if args.workload == "train":
        return run_train(args)

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the product path has no CPU fallback")
    cpus = bind_to_gpu_numa_node(local)
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=dev)

    from bevrender_b200 import _lib
    from bevrender_b200.workloads import make_microbench_inputs
    L = _lib.lib()
    if args.tile:
        tw, th, thr = [int(v) for v in args.tile.split(",")]
        L.bevr_msda_set_tile(tw, th, thr)
    B = args.batch
    tdt = torch.float32 if args.dtype == "fp32" else torch.bfloat16
    dt = 0 if args.dtype == "fp32" else 1
    esize = 4 if dt == 0 else 2
    warm = max(args.warmup, 3)

    # ---- synthetic inputs (pinned host copies for the e2e leg, device copies for the kernel leg)
    host = make_microbench_inputs(B, seed=rank, dtype=tdt)
    shapes, starts = host["spatial_shapes"], host["level_start_index"]
    S = sum(h * w for h, w in shapes)
    Nq = N_SIDE * N_SIDE
    names = ("value", "sampling_locations", "attention_weights", "grad_out")
    h_in = {k: host[k].contiguous().pin_memory() for k in names}
    d_in = {k: h_in[k].to(dev) for k in names}
    out = torch.empty(B, Nq, HEADS * DH, dtype=tdt, device=dev)
    gv = torch.empty(B, S, HEADS, DH, dtype=tdt, device=dev)
    gl = torch.empty(B, Nq, HEADS, LEVELS, POINTS, 2, dtype=torch.float32, device=dev)
    gw = torch.empty(B, Nq, HEADS, LEVELS, POINTS, dtype=tdt, device=dev)
    nws = L.bevr_msda_bwd_workspace_bytes(B, S, HEADS, DH, dt)
    ws = torch.empty(max(nws, 16), dtype=torch.uint8, device=dev)
    hw = (C.c_int32 * (2 * LEVELS))(*[v for s in shapes for v in s])
    st = (C.c_int32 * LEVELS)(*starts)
    P = lambda t: C.c_void_p(t.data_ptr())

    def fwd(dd, stream):
        _lib.check(L.bevr_msda_fwd(P(dd["value"]), hw, st, P(dd["sampling_locations"]), None, P(dd["attention_weights"]),
                                   P(out), B, S, HEADS, DH, Nq, LEVELS, POINTS, N_SIDE, dt, 0, stream), "msda_fwd")

    def bwd(dd, stream):
        _lib.check(L.bevr_msda_bwd(P(dd["value"]), hw, st, P(dd["sampling_locations"]), None, P(dd["attention_weights"]),
                                   P(dd["grad_out"]), P(gv), P(gl), P(gw), P(ws), B, S, HEADS, DH, Nq, LEVELS, POINTS,
                                   N_SIDE, dt, 0, stream), "msda_bwd")

    cur = torch.cuda.current_stream()
    cs = C.c_void_p(cur.cuda_stream)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(ms):
        if world == 1:
            return ms
        t = torch.tensor([ms], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    # ---- device-resident leg: the headline `value`
    for _ in range(warm):
        fwd(d_in, cs)
        bwd(d_in, cs)
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with ClockSampler(local) as clk:
        e0.record()
        for _ in range(args.steps):
            fwd(d_in, cs)
            bwd(d_in, cs)
        e1.record()
        barrier()
    ms_total = max_over_ranks(e0.elapsed_time(e1))
    ms_step = ms_total / args.steps
    value = world * B * Nq / (ms_step * 1e-3)

    # ---- per-kernel durations for the roofline (CUDA events around each launch, same stream)
    ev = [[torch.cuda.Event(enable_timing=True) for _ in range(3)] for _ in range(args.steps)]
    for i in range(args.steps):
        ev[i][0].record()
        fwd(d_in, cs)
        ev[i][1].record()
        bwd(d_in, cs)
        ev[i][2].record()
    torch.cuda.synchronize()
    t_fwd = sum(e[0].elapsed_time(e[1]) for e in ev) / args.steps
    t_bwd = sum(e[1].elapsed_time(e[2]) for e in ev) / args.steps
    ab_fwd, ab_bwd = algorithmic_bytes(B, esize)
    peak, peak_src = peaks()
    dom = ("msda_bwd_kernel", ab_bwd, t_bwd) if t_bwd >= t_fwd else ("msda_fwd_kernel", ab_fwd, t_fwd)
    achieved = dom[1] / (dom[2] * 1e-3) / 1e9
    # dram__bytes_read.sum + dram__bytes_write.sum per launch of the same kernels at B=8 fp32, from the ncu --set full
    # capture committed as profiles/r1_ncu_msda.md (bwd: 1.709 GB + 0.907 GB; fwd: 0.927 GB + 0.313 GB)
    ncu_traffic = {"msda_bwd_kernel": 2.616e9, "msda_fwd_kernel": 1.240e9} if (dt == 0 and B == 8) else {}
    roofline = {"bound": "hbm", "kernel": dom[0], "achieved": achieved, "peak": peak, "unit": "GB/s",
                "frac": achieved / peak, "traffic": ncu_traffic.get(dom[0]), "peak_source": peak_src,
                "limiter": "bwd: L2 atomic units (655 M RED sectors/launch; RED.v4, TMA bulk reduce and tensor TMA reduce all "
                           "measure 3.03 ms for the scatter alone); fwd: L1 data pipe (pure-gather ceiling 0.85 ms) -- "
                           "profiles/r1_msda_ceilings.md, DESIGN.md section 4",
                "algorithmic_bytes_per_launch": dom[1], "launch_ms": dom[2],
                "fwd": {"ms": t_fwd, "GBps": ab_fwd / (t_fwd * 1e-3) / 1e9},
                "bwd": {"ms": t_bwd, "GBps": ab_bwd / (t_bwd * 1e-3) / 1e9,
                        "note": "includes the grad_value zero-fill (and the bf16 cast) the entry point enqueues"},
                "fwd_bwd_frac": (ab_fwd + ab_bwd) / (ms_step * 1e-3) / 1e9 / peak}

    # ---- end-to-end leg: HOST pinned inputs -> H2D -> fwd+bwd -> D2H of every result, per step, through the
    # host-buffer entry point (per-sample chunks pipelined over H2D / compute / D2H streams)
    from bevrender_b200.host_pipeline import HostMsdaPipeline
    h_out = {k: torch.empty(t.shape, dtype=t.dtype).pin_memory() for k, t in
             (("out", out), ("grad_value", gv), ("grad_loc", gl), ("grad_weight", gw))}
    h2d = sum(t.numel() * t.element_size() for t in h_in.values())
    d2h = sum(t.numel() * t.element_size() for t in h_out.values())
    pipe = None if args.no_e2e else HostMsdaPipeline(shapes, starts, HEADS, DH, Nq, POINTS, N_SIDE, tdt, dev,
                                                    qsplit=int(os.environ.get("BEVR_QSPLIT", "4")),
                                                    nslots=int(os.environ.get("BEVR_NSLOTS", "2")))

    def e2e_step():
        pipe.run(h_in, h_out)

    e2e_steps = 0 if args.no_e2e else max(2, min(args.steps, 5))
    for _ in range(2 if e2e_steps else 0):
        e2e_step()
    barrier()
    e0.record()
    for _ in range(e2e_steps):
        e2e_step()
    e1.record()
    barrier()
    ms_e2e = max_over_ranks(e0.elapsed_time(e1)) / max(e2e_steps, 1)
    e2e = None if args.no_e2e else {"value": world * B * Nq / (ms_e2e * 1e-3), "unit": "queries/s", "ms_per_step": ms_e2e,
           "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h, "steps": e2e_steps,
           "host_cpus_bound": None if cpus is None else len(cpus)}

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cores = os.cpu_count() or 1
        cpu_reference_step(1, args.dtype, cores)
        t = min(cpu_reference_step(1, args.dtype, cores) for _ in range(2))
        cpu = {"value": Nq / t, "unit": "queries/s", "cores": cores, "kind": "port", "seconds": t,
               "sample": "1 of the step's 8 samples (40000 queries), fwd+bwd, torch CPU grid_sample path, best of 2"}

    if rank == 0:
        n_k = 2 if dt == 0 else 3
        line = {"metric": METRIC, "value": value, "unit": "queries/s", "n_gpus": world, "steps": args.steps,
                "warmup": warm, "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak",
                "vs_baseline": None, "dtype": "f32" if dt == 0 else "bf16", "data": "synthetic",
                "config": workload_config(B, args.dtype), "roofline": roofline, "cpu_baseline": cpu, "e2e": e2e,
                "gpu_launches": args.steps * n_k, "clocks": clk.summary()}
        emit(line)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()

#!/usr/bin/env python
"""Benchmark of the BEVRender deformable-attention hot path (BASELINE.json config 2).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--dtype fp32|bf16] [--impl reference]

A "step" is one forward + backward pass of the multi-scale deformable-attention gather over one
batch of B=8 synthetic samples per GPU (200x200 BEV queries, 4 levels, 8 heads, 4 points, Dh=32),
through the C ABI of libbevr.so.  Prints ONE JSON line (see DESIGN.md "Measurement").  The same line
carries, as extra objects, the config-5 training step of the reference BEVRender with the drop-ins
("train": samples/s with the NCCL gradient allreduce and SyncBatchNorm inside the timed region, every
rank), and on single-GPU runs K3 alone ("k3"), the correlation ("k5") and config 1 against the real
reference on the host cores ("config1").
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


def timeit_cuda(fn, warm=3, iters=10):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


def build_train_model(B, bev, layers, dev):
    """BASELINE.json config 5 model: the UNMODIFIED reference ``BEVRender`` (model/bevrender.py:14-221: ResNet18 backbone,
    1 stage x `layers` EncoderLayers, render decoder) constructed under ``compat.install(wrappers=True)``, i.e. with the
    drop-in attention modules / projector / wrappers of this repo behind the reference's own encoder.py.  The reference
    checkout is found at /root/reference or, on the GPU box, at oracle/_ref/reference (staged by ``make -C oracle``).
    Falls back to the stand-in harness (labelled) only when no checkout exists."""
    from bevrender_b200 import compat, workloads
    if compat.find_reference() is None:
        from bevrender_b200.harness import BevHarness
        m = BevHarness(n=bev, c=64, heads=2, depth=5, layers=layers, batch=B, device=dev)
        return m, "stand-in harness (no reference checkout on this box)", lambda cam, pose, veh: m(cam[:, :, 0])
    compat.add_reference_to_path()
    compat.install(wrappers=True)
    from model.bevrender import BEVRender
    m = BEVRender(workloads.bevrender_config(bev=bev, layers=layers, batch=B), None, "train")
    return m, "reference BEVRender (model/bevrender.py, unmodified) + compat.install(wrappers=True)", \
        lambda cam, pose, veh: m(cam, pose, veh, {}, True)[0]


def train_leg(args, dev, rank, world, steps, warmup):
    """One training step of config 5, batch-data-parallel: H2D of the batch -> forward (window of 4 frames: 3 history
    passes in eval mode without gradient + 1 training pass, bevrender.py:124-146) -> MSE render loss + plain-torch
    contrastive + triplet losses on the flattened outputs -> backward with the flat-bucket NCCL gradient allreduce
    overlapped (dp.FlatGradAllReduce, replaces DDP train.py:133-135) -> clip_grad_norm_(1.0) -> AdamW
    (train.py:236-247, 706-711).  BatchNorm layers are converted to SyncBatchNorm when world > 1, as train.py:129-131."""
    import torch.distributed as dist
    from bevrender_b200 import losses, workloads
    from bevrender_b200.dp import FlatGradAllReduce, find_live_parameters
    B, T, bev, layers = args.batch, 4, args.bev, 6
    tf32 = torch.backends.cudnn.allow_tf32
    torch.backends.cudnn.allow_tf32 = True            # the reference's own convolutions: torch defaults
    torch.manual_seed(0)
    model, label, fwd = build_train_model(B, bev, layers, dev)
    n_bn = sum(isinstance(m, torch.nn.modules.batchnorm._BatchNorm) for m in model.modules())
    if world > 1:
        model = torch.nn.SyncBatchNorm.convert_sync_batchnorm(model)
    model = model.to(dev).train()
    cam, pose, veh, target = workloads.make_train_batch(B, window=T, seed=1 + rank)
    h_in = [t.pin_memory() for t in (cam, pose, veh, target)]
    h_loss = torch.zeros(1).pin_memory()

    def loss_of(cam_d, pose_d, veh_d, target_d):
        out = fwd(cam_d, pose_d, veh_d)
        return (torch.nn.functional.mse_loss(out, target_d) + losses.contrastive_loss(out, target_d) +
                losses.triplet_loss(out, target_d))

    d0 = [t.to(dev) for t in h_in]
    live_named = find_live_parameters(model, lambda: loss_of(*d0).backward())
    live = [p for _, p in live_named]
    live_ids = {id(p) for p in live}
    n_dead = 0
    for p in model.parameters():
        if id(p) not in live_ids:
            p.requires_grad_(False)
            n_dead += 1
    opt = torch.optim.AdamW(live, lr=1e-4, weight_decay=1e-2, eps=1e-8)
    red = FlatGradAllReduce(live, n_buckets=2)

    def step():
        cam_d, pose_d, veh_d, target_d = (t.to(dev, non_blocking=True) for t in h_in)
        red.reset()
        loss = loss_of(cam_d, pose_d, veh_d, target_d)
        loss.backward()
        red.finish()
        torch.nn.utils.clip_grad_norm_(live, 1.0)
        opt.step()
        h_loss.copy_(loss.detach().reshape(1), non_blocking=True)

    for _ in range(warmup):
        step()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        step()
    e1.record()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    if world > 1:
        t = torch.tensor([ms], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    ms /= steps
    torch.backends.cudnn.allow_tf32 = tf32
    nparam = sum(p.numel() for p in live)
    red.remove()
    return {"samples_per_s": world * B / (ms * 1e-3), "ms_per_step": ms, "steps": steps, "warmup": warmup,
            "nranks": world, "batch_per_gpu": B, "allreduce_bytes": nparam * 4, "allreduce_launches_per_step": len(red.buckets),
            "sync_batchnorm": world > 1, "batchnorm_layers": n_bn, "live_params": nparam, "excluded_param_tensors": n_dead,
            "loss": float(h_loss), "h2d_bytes_per_step": sum(t.numel() * t.element_size() for t in h_in),
            "model": label, "config": "config5: bev %d, 1 stage x %d layers, window %d, image 128x160, C=64, 2 heads, D=5; "
            "loss = MSE + contrastive + triplet (plain torch, bevrender_b200/losses.py)" % (bev, layers, T)}


def k3_leg():
    """K3 (fused attention with sampled RPE bias) alone, at the reference-native shapes of config 3(i): N=56, C=64, 2 heads,
    B=8 -> BH=16; TSA keys 56x56, SCA keys 28x280 (D=5).  Forward and full backward (all five gradients) in ms,
    (query, key) pairs/s -- SURVEY.md section 8(d): K3 is reported as achieved gathers/s, one bilinear gather per pair."""
    import bevrender_b200 as bb
    res = {}
    for name, Hk, Wk, mul in (("tsa", 56, 56, 1), ("sca", 28, 280, 5)):
        torch.manual_seed(0)
        BH, Dh, N = 16, 32, 56
        Nq, Ns = N * N, Hk * Wk
        mk = lambda *s: torch.randn(*s, device="cuda")
        q, k, v = mk(BH, Dh, Nq).requires_grad_(True), mk(BH, Dh, Ns).requires_grad_(True), mk(BH, Dh, Ns).requires_grad_(True)
        rpe = (mk(2, 2 * N - 1, 2 * N * mul - 1) * 0.5).requires_grad_(True)
        ys = torch.linspace(-1, 1, Hk, device="cuda")
        xs = torch.linspace(-1, 1, Wk, device="cuda")
        pos = torch.stack(torch.meshgrid(ys, xs, indexing="ij"), -1).reshape(1, Ns, 2).repeat(8, 1, 1)
        pos = (pos + 0.02 * mk(8, Ns, 2)).requires_grad_(True)
        go = mk(BH, Dh, Nq)
        f = lambda: bb.rpe_attention(q, k, v, rpe, pos, N, N, 1)
        t_f = timeit_cuda(f)

        def fb():
            for t in (q, k, v, rpe, pos):
                t.grad = None
            f().backward(go)
        t_fb = timeit_cuda(fb)
        pairs = BH * Nq * Ns
        # algorithmic HBM bytes of the forward: q, out (BH, Dh, Nq); k, v (BH, Dh, Ns); pos; table -- the (Nq, Ns) matrix
        # the reference materialises (pairs * 4 B, three times over) never exists
        hbm_fwd = 4 * (2 * BH * Dh * Nq + 2 * BH * Dh * Ns + pos.numel() + rpe.numel())
        res[name] = {"fwd_ms": t_f, "bwd_ms": t_fb - t_f, "pairs": pairs, "fwd_pairs_per_s": pairs / (t_f * 1e-3),
                     "fwd_bwd_pairs_per_s": pairs / (t_fb * 1e-3),
                     "fwd_bilinear_gathers_per_s": pairs / (t_f * 1e-3), "fwd_table_tap_loads_per_s": 4 * pairs / (t_f * 1e-3),
                     "fwd_hbm_bytes": hbm_fwd, "fwd_hbm_GBps": hbm_fwd / (t_f * 1e-3) / 1e9,
                     "reference_attention_matrix_bytes": 4 * pairs,
                     # executed tensor work: 2 contractions x 3 TF32 products x 2*Dh flops per pair, against the measured
                     # legacy mma.sync TF32 rate of this chip (278 TFLOP/s, profiles/r1_exp_mma_sync.txt)
                     "fwd_mma_TFLOPs": 2 * 3 * 2 * Dh * pairs / (t_f * 1e-3) / 1e12, "mma_sync_tf32_peak_TFLOPs": 278.0}
    res["shape"] = "N=56, C=64, 2 heads, B=8 (BH=16), Dh=32; TSA Ns=3136, SCA Ns=7840 (D=5)"
    return res


def config2_variants_leg():
    """The other config-2 numbers SURVEY.md section 8(d) asks for beside the B=8 headline: B=1 (working set 430 MB, still
    larger than L2) and the secondary camera-pyramid workload (value levels of a 512x640 image at /8../64, S=6800:
    L2-resident), fp32, device-resident, fwd+bwd through the C ABI."""
    from bevrender_b200 import _lib
    from bevrender_b200.workloads import make_microbench_inputs
    L = _lib.lib()
    res = {}
    for name, B, shp in (("b1", 1, None), ("camera_pyramid_b8", 8, [(64, 80), (32, 40), (16, 20), (8, 10)])):
        d = make_microbench_inputs(B, seed=0, shapes=shp, device="cuda")
        shapes, starts = d["spatial_shapes"], d["level_start_index"]
        S, Nq = sum(h * w for h, w in shapes), N_SIDE * N_SIDE
        out = torch.empty(B, Nq, HEADS * DH, device="cuda")
        gv = torch.empty(B, S, HEADS, DH, device="cuda")
        gl = torch.empty(B, Nq, HEADS, LEVELS, POINTS, 2, device="cuda")
        gw = torch.empty(B, Nq, HEADS, LEVELS, POINTS, device="cuda")
        hw = (C.c_int32 * (2 * LEVELS))(*[v for s_ in shapes for v in s_])
        st = (C.c_int32 * LEVELS)(*starts)
        P = lambda t: C.c_void_p(t.data_ptr())
        cs = C.c_void_p(torch.cuda.current_stream().cuda_stream)

        def f():
            _lib.check(L.bevr_msda_fwd(P(d["value"]), hw, st, P(d["sampling_locations"]), None, P(d["attention_weights"]),
                                       P(out), B, S, HEADS, DH, Nq, LEVELS, POINTS, N_SIDE, 0, 0, cs), "msda_fwd")

        def fb():
            f()
            _lib.check(L.bevr_msda_bwd(P(d["value"]), hw, st, P(d["sampling_locations"]), None, P(d["attention_weights"]),
                                       P(d["grad_out"]), P(gv), P(gl), P(gw), None, B, S, HEADS, DH, Nq, LEVELS, POINTS,
                                       N_SIDE, 0, 0, cs), "msda_bwd")
        t_f, t_fb = timeit_cuda(f, 3, 20), timeit_cuda(fb, 3, 20)
        nbytes = lambda *ts: sum(t.numel() * t.element_size() for t in ts)
        ab = nbytes(d["value"], d["sampling_locations"], d["attention_weights"], out) + \
            nbytes(d["grad_out"], d["value"], d["sampling_locations"], d["attention_weights"], gv, gl, gw)
        res[name] = {"fwd_ms": t_f, "bwd_ms": t_fb - t_f, "queries_per_s": B * Nq / (t_fb * 1e-3),
                     "algorithmic_GBps": ab / (t_fb * 1e-3) / 1e9, "levels": [list(s_) for s_ in shapes], "batch": B}
    return res


def k5_leg():
    """K5 (rotation-sweep correlation, tcgen05) at config 4: template (3,224,224) vs map (3,512,512), 72 rotations."""
    import math
    import bevrender_b200 as bb
    R, Ht, Hm, C = 72, 224, 512, 3
    ang = [2 * math.pi * r / R for r in range(R)]
    torch.manual_seed(0)
    m, t = torch.randn(C, Hm, Hm, device="cuda"), torch.randn(C, Ht, Ht, device="cuda")
    ms = timeit_cuda(lambda: bb.correlate(m, t, ang), 3, 5)
    flops = 2.0 * (Hm - Ht + 1) ** 2 * R * C * Ht * Ht
    try:
        peak = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["bf16_tflops_sustained"])
        src = "measured (MEASURED_PEAKS.json bf16_tflops_sustained)"
    except Exception:
        peak, src = 1324.9, "fallback 1324.9 TFLOP/s"
    return {"ms": ms, "useful_TFLOPs": flops / ms / 1e9, "peak_TFLOPs": peak, "frac": flops / ms / 1e9 / peak,
            "peak_source": src, "bound": "tensor", "config": "config4: C=3, 224x224 template, 512x512 map, 72 rotations"}


def config1_leg(cores):
    """BASELINE.json config 1, the reference's own CPU-runnable case: ONE reference ``EncoderLayer`` (encoder.py:243-411;
    N=50, C=64, 2 heads, D=5, image features 32x40, B=1, prev_bev given), forward + backward.  The real reference runs on
    the box's host cores (kind='reference'); the same layer class built under compat.install() runs on the GPU."""
    import copy
    import importlib
    from bevrender_b200 import compat, workloads
    if compat.find_reference() is None:
        return {"unavailable": "no reference checkout on this box"}
    compat.add_reference_to_path()
    import bevrender_b200 as bb
    enc = importlib.import_module("model.encoder")
    tsa_mod, sca_mod = importlib.import_module("model.TSA"), importlib.import_module("model.SCA")
    RefProj = importlib.import_module("model.bev_cmr_proj").BEV2CameraProjector
    saved = (tsa_mod.TSADeformableAttention, sca_mod.SCADeformableAttention, enc.SpatialCrossAttn, enc.TemporalSelfAttn)
    N, B = 50, 1
    T, K = workloads.default_camera()
    pk = dict(imu_to_rgb={0: [T]}, vehicle_type_code=0, img_width=160, img_height=128, ori_img_width=640, ori_img_height=512)
    kw = dict(bev_bound={"X": 20.0, "Y": 10.0, "Z": 2.0}, n_views=1, bev_feat_shape=N, bev_depth_dim=5, z_shift=-1.0,
              dim_embed=64, expansion=2, stage_idx=0, n_groups=1, n_heads=2, stride=1, kernel_size=3, batch_size=B,
              scale_offset_range=True, drop_path_rate=0.0)
    torch.manual_seed(0)
    try:
        if type(saved[0]).__module__.startswith("bevrender_b200") or saved[0].__module__.startswith("bevrender_b200"):
            importlib.reload(tsa_mod), importlib.reload(sca_mod), importlib.reload(enc)   # undo an earlier install()
        ref_layer = enc.EncoderLayer(bev2cmr_projector=RefProj(K={0: [K.copy()]}, device="cpu", **pk), **kw)
        compat.install(wrappers=True)
        new_layer = enc.EncoderLayer(bev2cmr_projector=bb.BEV2CameraProjector(K={0: [K.copy()]}, **pk), **kw)
    finally:
        tsa_mod.TSADeformableAttention, sca_mod.SCADeformableAttention, enc.SpatialCrossAttn, enc.TemporalSelfAttn = saved
    new_layer.load_state_dict(ref_layer.state_dict(), strict=True)
    new_layer = new_layer.cuda().train()
    ref_layer.train()
    g = torch.Generator().manual_seed(3)
    q0, img0, prev, go = (torch.randn(*s, generator=g) for s in ((B, 64, N, N), (B, 64, 32, 40), (B, 64, N, N), (B, 64, N, N)))

    def run(layer, dev):
        q, img = q0.to(dev).requires_grad_(True), img0.to(dev).requires_grad_(True)
        out, _ = layer(bev_query=q, img_tensor=img, prev_bev=prev.to(dev), vehicle_pose=None, vehicle_type_idx=torch.tensor(0),
                       wandb_log_dict={}, return_wandb_log=False)
        out.backward(go.to(dev))
        return out
    torch.set_num_threads(cores)
    run(ref_layer, "cpu")
    ts = []
    for _ in range(2):
        t0 = time.perf_counter()
        run(ref_layer, "cpu")
        ts.append(time.perf_counter() - t0)
    gpu_ms = timeit_cuda(lambda: run(new_layer, "cuda"), 3, 10)
    return {"workload": "config1: reference EncoderLayer fwd+bwd, N=50, C=64, 2 heads, D=5, img feat 32x40, B=1",
            "reference_cpu_ms": min(ts) * 1e3, "cores": cores, "kind": "reference", "ours_gpu_ms": gpu_ms,
            "speedup": min(ts) * 1e3 / gpu_ms, "queries": N * N,
            "note": "GPU time includes the host->device copies of the inputs and the layer's PyTorch glue (launch-bound at B=1)"}


def run_train(args):
    """--workload train: the config-5 training step alone, as its own JSON line (the contract line carries the same numbers
    under "train")."""
    import torch.distributed as dist
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    with ClockSampler(local) as clk:
        tr = train_leg(args, dev, rank, world, args.steps, max(args.warmup, 3))
    if rank == 0:
        emit({"metric": "train samples/s", "value": tr["samples_per_s"], "unit": "samples/s", "n_gpus": world,
              "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": tr["ms_per_step"], "higher_is_better": True,
              "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
              "config": {"workload": tr["config"], "model": tr["model"], "batch_per_gpu": args.batch,
                         "parallelism": "dp%d, flat fp32 gradient bucket (%.1f MB, %d NCCL allreduce(AVG) launches overlapped "
                         "with backward), SyncBatchNorm=%s" % (world, tr["allreduce_bytes"] / 1e6,
                                                                tr["allreduce_launches_per_step"], tr["sync_batchnorm"])},
              "train": tr, "clocks": clk.summary()})
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
    ap.add_argument("--bf16-out-grad", dest="bf16_grad_f32", action="store_false",
                    help="--dtype bf16: deliver grad_value in bf16 (fp32 workspace + cast) instead of fp32 main gradients")
    ap.add_argument("--no-train", action="store_true", help="skip the config-5 training-step leg")
    ap.add_argument("--no-extra", action="store_true", help="skip the K3 / K5 / config-1 legs")
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
    B = args.batch
    tdt = torch.float32 if args.dtype == "fp32" else torch.bfloat16
    dt = 0 if args.dtype == "fp32" else 1
    esize = 4 if dt == 0 else 2
    warm = max(args.warmup, 3)

    # ---- synthetic inputs (pinned host copies for the e2e leg, device copies for the kernel leg)
    host = make_microbench_inputs(B, seed=rank, dtype=tdt, sigma_px=float(os.environ.get("BEVR_SIGMA_PX", "2.0")))
    shapes, starts = host["spatial_shapes"], host["level_start_index"]
    S = sum(h * w for h, w in shapes)
    Nq = N_SIDE * N_SIDE
    names = ("value", "sampling_locations", "attention_weights", "grad_out")
    h_in = {k: host[k].contiguous().pin_memory() for k in names}
    d_in = {k: h_in[k].to(dev) for k in names}
    out = torch.empty(B, Nq, HEADS * DH, dtype=tdt, device=dev)
    # bf16 storage keeps fp32 main gradients for value (BEVR_GRAD_VALUE_F32): "grad_value accumulated fp32" of BASELINE.json
    # config 2 without the workspace + cast pass the bf16-output path needs
    gv = torch.empty(B, S, HEADS, DH, dtype=torch.float32 if args.bf16_grad_f32 else tdt, device=dev)
    bwd_flags = 32 if (dt == 1 and args.bf16_grad_f32) else 0
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
                                   N_SIDE, dt, bwd_flags, stream), "msda_bwd")

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
    bwd_name = "msda_bwd_sorted_kernel"
    dom = (bwd_name, ab_bwd, t_bwd) if t_bwd >= t_fwd else ("msda_fwd_kernel", ab_fwd, t_fwd)
    achieved = dom[1] / (dom[2] * 1e-3) / 1e9
    # `traffic`: dram__bytes_read.sum + dram__bytes_write.sum per launch from the committed ncu --set full capture of the
    # SAME kernels at this configuration (profiles/traffic.json names the capture); null when no capture matches
    traffic, traffic_src = None, None
    try:
        tj = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))
        if tj.get("dtype") == args.dtype and tj.get("batch") == B:
            traffic, traffic_src = tj["kernels"].get(dom[0]), tj.get("capture")
    except Exception:
        pass
    roofline = {"bound": "hbm", "kernel": dom[0], "achieved": achieved, "peak": peak, "unit": "GB/s",
                "frac": achieved / peak, "traffic": traffic, "traffic_source": traffic_src, "peak_source": peak_src,
                "limiter": "bwd: issue slots of the tile-sorted segmented reduction (profiles/r2_ncu_msda.md); "
                           "fwd: L1 data pipe (pure-gather ceiling 0.85 ms, profiles/r1_msda_ceilings.md); DESIGN.md section 4",
                "algorithmic_bytes_per_launch": dom[1], "launch_ms": dom[2],
                "fwd": {"ms": t_fwd, "GBps": ab_fwd / (t_fwd * 1e-3) / 1e9},
                "bwd": {"ms": t_bwd, "GBps": ab_bwd / (t_bwd * 1e-3) / 1e9,
                        "note": "includes the grad_value zero-fill (and the bf16 cast) the entry point enqueues"},
                "fwd_bwd_frac": (ab_fwd + ab_bwd) / (ms_step * 1e-3) / 1e9 / peak}

    # ---- end-to-end leg: HOST pinned inputs -> H2D -> fwd+bwd -> D2H of every result, per step, through the
    # host-buffer entry point (per-sample chunks pipelined over H2D / compute / D2H streams)
    from bevrender_b200.host_pipeline import HostMsdaPipeline
    h_out = {k: torch.empty(t.shape, dtype=(tdt if k == "grad_value" else t.dtype)).pin_memory() for k, t in
             (("out", out), ("grad_value", gv), ("grad_loc", gl), ("grad_weight", gw))}
    h2d = sum(t.numel() * t.element_size() for t in h_in.values())
    d2h = sum(t.numel() * t.element_size() for t in h_out.values())
    pipe = None if args.no_e2e else HostMsdaPipeline(shapes, starts, HEADS, DH, Nq, POINTS, N_SIDE, tdt, dev,
                                                    qsplit=int(os.environ.get("BEVR_QSPLIT", "4")),
                                                    nslots=int(os.environ.get("BEVR_NSLOTS", "2")))

    # consecutive steps are enqueued back to back (join=False): step i+1's uploads start while step i's last results are
    # still going down, as a loop that prefetches its next batch does; the pipeline is joined inside the timed region
    def e2e_step(last=True):
        pipe.run(h_in, h_out, join=last)

    e2e_steps = 0 if args.no_e2e else max(2, min(args.steps, 5))
    for _ in range(2 if e2e_steps else 0):
        e2e_step()
    barrier()
    e0.record()
    for i in range(e2e_steps):
        e2e_step(last=(i == e2e_steps - 1) or bool(int(os.environ.get("BEVR_E2E_JOIN_EVERY_STEP", "0"))))
    e1.record()
    barrier()
    ms_e2e = max_over_ranks(e0.elapsed_time(e1)) / max(e2e_steps, 1)
    e2e = None if args.no_e2e else {"value": world * B * Nq / (ms_e2e * 1e-3), "unit": "queries/s", "ms_per_step": ms_e2e,
           "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h, "steps": e2e_steps,
           "host_cpus_bound": None if cpus is None else len(cpus),
           "pipelining": "per-sample / per-row-block pieces over H2D, compute and D2H streams; consecutive steps are enqueued "
                         "back to back and joined inside the timed region (BEVR_E2E_JOIN_EVERY_STEP=1 joins after every step)"}

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cores = os.cpu_count() or 1
        cpu_reference_step(1, args.dtype, cores)
        t = min(cpu_reference_step(1, args.dtype, cores) for _ in range(2))
        cpu = {"value": Nq / t, "unit": "queries/s", "cores": cores, "kind": "port", "seconds": t,
               "sample": "1 of the step's 8 samples (40000 queries), fwd+bwd, torch CPU grid_sample path, best of 2"}

    if e2e is not None:
        e2e["aggregate_host_device_GBps"] = world * (h2d + d2h) / (ms_e2e * 1e-3) / 1e9

    # ---- config 5: the end-to-end training step (all ranks; NCCL gradient allreduce + SyncBatchNorm inside the timed region)
    del pipe, h_out, d_in, out, gv, gl, gw, ws
    torch.cuda.empty_cache()
    train = None
    if not args.no_train:
        try:
            train = train_leg(args, dev, rank, world, steps=max(3, min(args.steps, 8)), warmup=3)
        except Exception as e:                       # never lose the contract line to the secondary legs
            train = {"error": "%s: %s" % (type(e).__name__, str(e)[:300])}
    # ---- K3 / K5 alone and config 1 against the REAL reference on the host cores (single-GPU runs only)
    k3 = k5 = config1 = config2v = None
    if rank == 0 and world == 1 and not args.no_extra:
        for name, fn in (("k3", k3_leg), ("k5", k5_leg), ("config1", lambda: config1_leg(os.cpu_count() or 1)),
                         ("config2_variants", config2_variants_leg)):
            try:
                r = fn()
            except Exception as e:
                r = {"error": "%s: %s" % (type(e).__name__, str(e)[:300])}
            if name == "k3":
                k3 = r
            elif name == "k5":
                k5 = r
            elif name == "config1":
                config1 = r
            else:
                config2v = r

    if rank == 0:
        # kernels of libbevr.so launched inside the timed device-resident region: msda_fwd + msda_bwd (sorted) per step,
        # plus the bf16 cast kernel; the cudaMemsetAsync zero-fill of grad_value is a driver memset, not one of ours
        n_k = 2 if (dt == 0 or bwd_flags) else 3
        line = {"metric": METRIC, "value": value, "unit": "queries/s", "n_gpus": world, "steps": args.steps,
                "warmup": warm, "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak",
                "vs_baseline": None, "dtype": "f32" if dt == 0 else "bf16", "data": "synthetic",
                "config": dict(workload_config(B, args.dtype), grad_value="fp32 main gradients (BEVR_GRAD_VALUE_F32)"
                               if bwd_flags else "storage dtype"), "roofline": roofline, "cpu_baseline": cpu, "e2e": e2e,
                "gpu_launches": args.steps * n_k, "train": train, "k3": k3, "k5": k5, "config1": config1,
                "config2_variants": config2v, "clocks": clk.summary()}
        emit(line)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()

#!/usr/bin/env python
"""Benchmark of the OT knowledge-distillation hot path (BASELINE.json metric: KD-loss fwd+bwd images/s).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl kdot|reference] [--workload ape_b64|dense_b32]

One "step" = one pass of the fused loss (forward + analytic backward) over one synthetic LINEMOD-ape shaped
mini-batch.  Default workload = BASELINE.json configs[1]: ape shape (8 keypoint slots, ~10 student / ~10
teacher cells per image), batch 64 per GPU.  Prints ONE JSON line (rank 0).  See DESIGN.md "Measurement".

* ``value``  device-resident throughput: inputs already in HBM, CUDA events around each step, L2 flushed and
  inputs restored between steps (outside the timed events), max over ranks.
* ``e2e``    same metric through the C ABI's host-buffer entry point (``kdot_sinkhorn_fwd_bwd_host``): pack +
  H2D + kernel + D2H + sync every step, wall clock.
* ``roofline``  algorithmic FLOPs of the launch / event time vs the FP32 FMA peak measured live on this GPU
  (the path is FP32/SFU bound, not HBM bound: the N x M cost matrix never leaves registers); HBM fraction
  reported beside it against MEASURED_PEAKS.json.
* ``cpu_baseline``  the oracle port of the reference formulation (torch CPU ops, fp32, autograd) on a bounded
  sample, host cores of this box.  ``--impl reference`` times only that, as the reference arm.
"""
from __future__ import annotations

import argparse
import ctypes
import json
import os
import statistics
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "kd_loss_fwd_bwd_images_per_sec"
UNIT = "images/s"
CFG = dict(p=2.0, blur=0.001, scaling=0.5, reach=0.5, w=640.0, h=480.0)
WORKLOADS = {
    # BASELINE.json configs[1]: the configuration the metric is quoted on
    "ape_b64": dict(nimg=64, dense=None, desc="LINEMOD-ape shape: B=8 keypoint slots, N~U{8..12} student / "
                    "M~U{8..12} teacher cells per image (5% empty teachers), D=2, batch 64 per GPU, "
                    "sinkhorn p=2 blur=1e-3 scaling=0.5 reach=0.5"),
    # multi-object images (e.g. 5-12 objects x ~10 cells): the 33..256-point range of the CTA-resident tiled kernel
    "multi_b64": dict(nimg=64, dense=None, n_range=(40, 120), m_range=(40, 120), desc="multi-object shape: N, M ~ U{40..120} "
                      "cells per image (5% empty teachers), B=8, D=2, batch 64 per GPU"),
    # BASELINE.json configs[2] variant 3b: every cell of the darknet_tiny / darknet53 grids
    "dense_b32": dict(nimg=32, dense=(1360, 1364), desc="all cells: N=1360 student / M=1364 teacher cells per "
                      "image, B=8, D=2, batch 32 per GPU"),
    # BASELINE.json configs[3]: dense ZebraPose-style 16-D per-cell code distributions on a 64x64 grid, one slot
    "zebra_b8": dict(nimg=8, dense=(4096, 4096), B=1, D=16, blur=0.05, desc="ZebraPose-style: N=M=4096 cells (64x64 "
                     "grid), D=16 code probabilities, B=1, blur=0.05, batch 8 per GPU"),
}


def workload_cfg(workload, args=None):
    """Solver knobs of a workload; ``--scaling`` / ``--blur`` override them (BASELINE.json configs[4]: the
    iteration / epsilon sweep)."""
    w = WORKLOADS[workload]
    cfg = dict(CFG, blur=w.get("blur", CFG["blur"]))
    if args is not None and args.scaling is not None:
        cfg["scaling"] = args.scaling
    if args is not None and args.blur is not None:
        cfg["blur"] = args.blur
    return cfg


def make_batch(workload, rank, nimg=None):
    from kd_6d_pose_adlp_b200.synthetic import ot_batch

    w = WORKLOADS[workload]
    extra = {k: w[k] for k in ("n_range", "m_range") if k in w}
    return ot_batch(nimg or w["nimg"], seed=1234 + rank, dense=w["dense"], sigma=0.05 if w["dense"] is None else 0.1,
                    B=w.get("B", 8), D=w.get("D", 2), **extra)


def algorithmic_work(batch, nits):
    """FLOPs / exps / bytes of one step from BASELINE.md section 3 (B slots, R = nits + 2 rounds per image)."""
    B, D = batch["xs"].shape[1], batch["xs"].shape[2]
    flops = exps = byts = 0.0
    for n, m, it in zip(batch["pos_per_img"], batch["pos_per_img_t"], nits):
        if n == 0 or m == 0:
            continue
        R = int(it) + 2
        flops += B * (R * (n + m) ** 2 * (3 * D + 5) + n * (n + m) * 2 * D)
        exps += B * R * (n + m) ** 2
        byts += 4 * B * ((n + m) * (D + 1) + n * (D + 1) + 1)
    return flops, exps, byts


class ClockSampler(threading.Thread):
    """Polls SM clock / throttle reasons through NVML while the timed region runs."""

    def __init__(self, index, period=0.01):
        super().__init__(daemon=True)
        self.index, self.period = index, period
        self.samples, self.reasons, self.stop_flag, self.max_mhz = [], set(), False, None
        try:
            import pynvml

            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None

    def run(self):
        if self.nv is None:
            return
        nv = self.nv
        names = {
            getattr(nv, "nvmlClocksThrottleReasonHwSlowdown", 0x8): "hw_slowdown",
            getattr(nv, "nvmlClocksThrottleReasonHwThermalSlowdown", 0x40): "hw_thermal_slowdown",
            getattr(nv, "nvmlClocksThrottleReasonSwThermalSlowdown", 0x20): "sw_thermal_slowdown",
            getattr(nv, "nvmlClocksThrottleReasonSwPowerCap", 0x4): "sw_power_cap",
        }
        while not self.stop_flag:
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for bit, name in names.items():
                    if r & bit:
                        self.reasons.add(name)
            except Exception:
                pass
            time.sleep(self.period)

    def result(self):
        self.stop_flag = True
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": [], "samples": 0}
        return {"sm_mhz": float(statistics.median(self.samples)), "sm_max_mhz": self.max_mhz,
                "reasons": sorted(self.reasons), "samples": len(self.samples)}


def cpu_step_fn(batch, cfg):
    """One fwd+bwd pass of the reference formulation on host cores (oracle port: fp32 torch CPU ops + autograd).
    D == 2 goes through the restated ``kd_loss_2d`` driver (normalise, per-image loop, transposes); other D call the
    restated ``SamplesLoss`` once per image on the ``(B, N, D)`` layout."""
    import torch

    from oracle import geomloss_ref, kd_loss_ref  # test infrastructure: allowed in the cpu_baseline / reference legs only

    L = geomloss_ref.SamplesLoss("sinkhorn", p=cfg["p"], blur=cfg["blur"], scaling=cfg["scaling"], reach=cfg["reach"])
    B, D = batch["xs"].shape[1], batch["xs"].shape[2]
    pn, pm = batch["pos_per_img"], batch["pos_per_img_t"]
    wt = torch.from_numpy(batch["wt"])
    if D == 2 and B == 8:
        xt0 = torch.from_numpy(batch["xt"].reshape(-1, 2))

        def one_pass():
            xs = torch.from_numpy(batch["xs"].reshape(-1, 2)).clone().requires_grad_(True)
            ws = torch.from_numpy(batch["ws"]).clone().requires_grad_(True)
            losses = kd_loss_ref.kd_loss_2d_ref(xs.clone(), xt0.clone(), ws, wt, cfg["w"], cfg["h"], "point", L, dim=2,
                                                pos_per_img=pn, pos_per_img_t=pm)
            (sum(losses) / len(losses)).backward()
            return xs.grad
    else:
        xt0 = torch.from_numpy(batch["xt"])

        def one_pass():
            xs = torch.from_numpy(batch["xs"]).clone().requires_grad_(True)
            ws = torch.from_numpy(batch["ws"]).clone().requires_grad_(True)
            losses, s0, t0 = [], 0, 0
            for n, m in zip(pn, pm):
                if n > 0 and m > 0:
                    losses.append(L(ws[s0:s0 + n].transpose(0, 1).contiguous(), xs[s0:s0 + n].transpose(0, 1).contiguous(),
                                    wt[t0:t0 + m].transpose(0, 1).contiguous(),
                                    xt0[t0:t0 + m].transpose(0, 1).contiguous()).sum())
                s0, t0 = s0 + n, t0 + m
            (sum(losses) / len(losses)).backward()
            return xs.grad
    return one_pass


def cpu_port_images_per_sec(batch, cfg, budget_s=12.0, min_passes=2, threads=None):
    """Times the oracle port of the reference formulation on `batch` for about `budget_s` seconds."""
    import torch

    if threads:
        torch.set_num_threads(threads)
    one_pass = cpu_step_fn(batch, cfg)
    nimg = len(batch["pos_per_img"])
    one_pass()  # warm-up
    t0 = time.perf_counter()
    passes = 0
    while passes < min_passes or (time.perf_counter() - t0) < budget_s:
        one_pass()
        passes += 1
        if passes >= 1000:
            break
    dt = time.perf_counter() - t0
    return nimg * passes / dt, passes, dt, torch.get_num_threads()


def run_reference_arm(args, rank, world):
    """--impl reference: the reference's formulation on the box's host cores (oracle port; geomloss itself is
    not installable offline).  Rank 0 only; each step is a bounded sample of the workload."""
    if rank != 0:
        return
    import torch

    workload = args.workload
    cfg = workload_cfg(workload, args)
    sample_img = min(args.images or 64, 64) if WORKLOADS[workload]["dense"] is None else 1
    batch = make_batch(workload, 0, nimg=sample_img)
    cores = len(os.sched_getaffinity(0))
    torch.set_num_threads(cores)
    step = cpu_step_fn(batch, cfg)
    for _ in range(args.warmup):
        step()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step()
    dt = time.perf_counter() - t0
    ms = dt / args.steps * 1e3
    value = sample_img / (ms * 1e-3)
    sample = f"{sample_img} images of workload {workload} per step, fwd+bwd, fp32 torch CPU ops"
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {"workload": workload, "desc": WORKLOADS[workload]["desc"], **cfg},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": torch.get_num_threads(), "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
        "note": "reference formulation (kd_loss_2d loop + geomloss-0.2.4 restatement) on host cores; "
                "geomloss is not installable offline, see DESIGN.md",
    }
    print(json.dumps(line), flush=True)


class DeviceBench:
    """Device-resident runner: preallocated buffers, direct C-ABI calls on the current stream."""

    def __init__(self, batch, dev, cfg=None):
        import torch

        self.cfg = cfg or CFG

        from kd_6d_pose_adlp_b200 import _lib
        from kd_6d_pose_adlp_b200.ops import cu_seqlens

        self.torch, self.L, self._lib = torch, _lib.lib(), _lib
        self.dev = dev
        self.batch = batch
        t = lambda a: torch.from_numpy(a).to(dev)
        self.xs0, self.xt0 = t(batch["xs"]), t(batch["xt"])
        self.xs, self.xt = self.xs0.clone(), self.xt0.clone()
        self.ws, self.wt = t(batch["ws"]), t(batch["wt"])
        self.nimg = len(batch["pos_per_img"])
        self.B, self.D = batch["xs"].shape[1], batch["xs"].shape[2]
        self.cu_n = cu_seqlens(batch["pos_per_img"], dev)
        self.cu_m = cu_seqlens(batch["pos_per_img_t"], dev)
        self.max_n, self.max_m = max(batch["pos_per_img"]), max(batch["pos_per_img_t"])
        self.loss = torch.empty(self.nimg, dtype=torch.float32, device=dev)
        self.valid = torch.empty(self.nimg, dtype=torch.int32, device=dev)
        self.nits = torch.empty(self.nimg, dtype=torch.int32, device=dev)
        self.gx = torch.empty_like(self.xs)
        self.gw = torch.empty_like(self.ws)
        nb = int(self.L.kdot_workspace_bytes(self.nimg, self.max_n, self.max_m, self.B, self.D))
        self.wsp = torch.empty(max(nb, 16), dtype=torch.uint8, device=dev)
        self.wsp_bytes = nb
        self.flush_buf = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)  # > 126 MB L2

    def restore_and_flush(self):
        self.xs.copy_(self.xs0)
        self.xt.copy_(self.xt0)
        self.flush_buf.fill_(1)
        # GPU-side delay (~100 us) so the host has enqueued <start event, kernel, stop event> before the GPU reaches
        # the start event: otherwise the host's launch latency would be counted as kernel time for a ~10 us kernel
        self.torch.cuda._sleep(200000)

    def step(self):
        torch = self.torch
        rc = self.L.kdot_sinkhorn_fwd_bwd(
            self.xs.data_ptr(), self.ws.data_ptr(), self.xt.data_ptr(), self.wt.data_ptr(), self.cu_n.data_ptr(),
            self.cu_m.data_ptr(), self.nimg, self.B, self.D, self.max_n, self.max_m, 0, self.cfg["p"], self.cfg["blur"],
            self.cfg["reach"], self.cfg["scaling"], self.cfg["w"], self.cfg["h"], 1 if self.D == 2 else 0,
            self.loss.data_ptr(), None, self.valid.data_ptr(),
            self.gx.data_ptr(), self.gw.data_ptr(), self.nits.data_ptr(), self.wsp.data_ptr(), self.wsp_bytes,
            torch.cuda.current_stream(self.dev).cuda_stream)
        self._lib.check(rc, "kdot_sinkhorn_fwd_bwd")

    def timed(self, steps, warmup, barrier):
        torch = self.torch
        for _ in range(warmup):
            self.restore_and_flush()
            self.step()
        evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
        barrier()
        torch.cuda.synchronize(self.dev)
        l0 = self._lib.launch_count()
        for e0, e1 in evs:
            self.restore_and_flush()
            e0.record()
            self.step()
            e1.record()
        torch.cuda.synchronize(self.dev)
        barrier()
        launches = self._lib.launch_count() - l0
        per = [e0.elapsed_time(e1) for e0, e1 in evs]
        return sum(per), per, launches


def host_e2e(batch, dev_index, steps, warmup, barrier, cfg=CFG):
    """End to end through kdot_sinkhorn_fwd_bwd_host: host numpy buffers in, host numpy buffers out."""
    from kd_6d_pose_adlp_b200 import _lib

    L = _lib.lib()
    nimg = len(batch["pos_per_img"])
    B, D = batch["xs"].shape[1], batch["xs"].shape[2]
    sn, sm = batch["xs"].shape[0], batch["xt"].shape[0]
    ctx = L.kdot_host_ctx_create(dev_index, nimg, sn, sm, B, D)
    if not ctx:
        raise RuntimeError("kdot_host_ctx_create: " + L.kdot_last_error().decode())
    xs, xt = batch["xs"].copy(), batch["xt"].copy()
    ws, wt = batch["ws"], batch["wt"]
    pn = np.asarray(batch["pos_per_img"], np.int32)
    pm = np.asarray(batch["pos_per_img_t"], np.int32)
    loss = np.empty(nimg, np.float32)
    valid = np.empty(nimg, np.int32)
    nits = np.empty(nimg, np.int32)
    gx = np.empty_like(xs)
    gw = np.empty_like(ws)
    p = lambda a: a.ctypes.data
    normalize = 1 if D == 2 else 0
    # raw host addresses of the caller's NumPy buffers, taken once (as any caller holding fixed buffers would)
    args = (ctx, p(xs), p(ws), p(xt), p(wt), p(pn), p(pm), nimg, cfg["p"], cfg["blur"], cfg["reach"], cfg["scaling"],
            cfg["w"], cfg["h"], normalize, 0, p(loss), p(valid), p(gx), p(gw), p(nits))
    call = L.kdot_sinkhorn_fwd_bwd_host

    def step():
        rc = call(*args)
        if rc != 0:
            _lib.check(rc, "kdot_sinkhorn_fwd_bwd_host")

    for _ in range(warmup):
        step()
    barrier()
    t0 = time.perf_counter()
    for _ in range(steps):
        step()
    dt = time.perf_counter() - t0
    barrier()
    h2d, d2h = ctypes.c_size_t(0), ctypes.c_size_t(0)
    L.kdot_host_ctx_last_traffic(ctx, ctypes.byref(h2d), ctypes.byref(d2h))
    phases = (ctypes.c_double * 4)()
    L.kdot_host_ctx_last_timing(ctx, phases)
    host_e2e.last_phases = dict(zip(("pack_us", "enqueue_us", "sync_us", "unpack_us"), [round(v, 2) for v in phases]))
    n_valid = int((valid == 1).sum())
    mean_loss = float(loss.sum() / max(n_valid, 1))
    L.kdot_host_ctx_destroy(ctx)
    return dt, int(h2d.value), int(d2h.value), mean_loss


def kernel_name(max_n, max_m, launches_per_step):
    if max_n + max_m <= 32:
        return "kdot_small_fast_kernel"
    return "kdot_tiled_kernel" if launches_per_step == 2 else "kdot_stream_kernel"


def ncu_traffic_bytes(kernel, workload):
    """DRAM bytes (read + write) of one launch of `kernel` from the committed `ncu --set full` summary of the same
    workload (profiles/r02_prof_*_ncu_summary.txt), or None when no capture of that kernel ON THAT WORKLOAD is on file."""
    tag = {("kdot_small_fast_kernel", "ape_b64"): "small_fast", ("kdot_stream_kernel", "dense_b32"): "stream",
           ("kdot_stream_kernel", "zebra_b8"): "stream_zebra", ("kdot_tiled_kernel", "multi_b64"): "tiled"}.get((kernel, workload))
    path = os.path.join(ROOT, "profiles", f"r02_prof_{tag}_ncu_summary.txt")
    if tag is None or not os.path.exists(path):
        return None
    unit = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    total = 0.0
    for line in open(path):
        f = line.split()
        if len(f) >= 3 and f[0] in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
            total += float(f[1].replace(",", "")) * unit.get(f[2], 1.0)
    return total or None


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as fh:
            return json.load(fh), "measured (MEASURED_PEAKS.json)"
    return {"hbm_gbs": 6650.0}, "fallback (B200_PROFILING.md)"


def allreduce_leg(dev, bench, barrier, max_over_ranks, world, rank):
    """The exchange a data-parallel run adds around the path (reference strategy
    train_kd.py:50,137 + libs/train_libs.py:124-130): NCCL all-reduce of the student's gradient bucket, in place on
    the persistent flat buffer of kd_6d_pose_adlp_b200.dist.GradBucket, plus the two-scalar loss exchange.  Device
    timed (CUDA events on the launching stream, which waits for NCCL's stream), max over ranks, median of 50."""
    import torch
    import torch.distributed as dist

    from kd_6d_pose_adlp_b200.dist import GradBucket, global_mean_loss

    out = {"backend": dist.get_backend(), "op": "all_reduce AVG, in place on the flat gradient bucket (zero copies)",
           "timing": "CUDA events, 20 warm-up + 50 timed, median, max over ranks"}

    def time_call(fn, iters=50, warm=20):
        for _ in range(warm):
            fn()
        torch.cuda.synchronize(dev)
        barrier()
        ts = []
        for _ in range(iters):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            fn()
            e1.record()
            e1.synchronize()
            ts.append(e0.elapsed_time(e1) * 1e3)
        return max_over_ranks(statistics.median(ts))

    buckets = {}
    for name, n in (("darknet_tiny_h", 2304468), ("darknet_tiny", 8486076)):
        prm = torch.nn.Parameter(torch.zeros(n, device=dev))
        bucket = GradBucket([prm])
        bucket.flat.fill_(float(rank + 1))
        us = time_call(lambda b=bucket: b.allreduce(average=True))
        nbytes = n * 4
        out[name] = {"elements": n, "bytes": nbytes, "us": us, "alg_GBps": nbytes / us / 1e3,
                     "bus_GBps": 2.0 * (world - 1) / world * nbytes / us / 1e3}
        buckets[name] = bucket
    loss_sum = torch.ones((), device=dev)
    out["global_mean_loss_us"] = time_call(lambda: global_mean_loss(loss_sum, 60))

    # the loss kernel of the headline workload with the student-gradient all-reduce of the PREVIOUS step in flight
    # (issued first: NCCL's stream then runs beside the kernel's stream; both are waited for before the stop event)
    bucket = buckets["darknet_tiny_h"]

    def overlapped():
        work = bucket.allreduce(average=True, async_op=True)
        bench.step()
        if work is not None:
            work.wait()

    def serial():
        bucket.allreduce(average=True)
        bench.step()

    out["step_with_allreduce"] = {
        "workload": "ape_b64 + darknet_tiny_h bucket", "overlapped_us": time_call(overlapped), "serial_us": time_call(serial),
        "kernel_only_us": time_call(bench.step),
        "note": "no L2 flush between these launches (steady-state training step); the all-reduce, not the 2x-us loss "
                "kernel, bounds an ape-shaped data-parallel step"}
    return out


def b0_and_select_legs(dev, nimg=64):
    """The callers either side of the kernel, through the reference-facing Python seams:
    KDPoseLoss.__call__ + backward with the REAL SSC target assignment on the device, and PostProcessorKD's selection
    kernel.  Host wall clock around a synchronised call (these paths are launch / host bound), median of 20."""
    import types

    import torch

    from kd_6d_pose_adlp_b200.losses.kd_loss import make_kd_pose_loss
    from kd_6d_pose_adlp_b200.postprocess.postprocess_kd import select_cells
    from kd_6d_pose_adlp_b200.target_coder import TargetCoder, grid_anchors
    from tests import doubles, scenario

    def timed(fn, warm=5, iters=20):
        for _ in range(warm):
            fn()
        ts = []
        for _ in range(iters):
            torch.cuda.synchronize(dev)
            t0 = time.perf_counter()
            fn()
            torch.cuda.synchronize(dev)
            ts.append((time.perf_counter() - t0) * 1e3)
        return statistics.median(ts)

    s_hw = [(32, 32), (16, 16), (8, 8), (4, 4)]
    t_hw = s_hw + [(2, 2)]
    arr = scenario.make_target_arrays(nimg, 0)
    tt = lambda a: torch.tensor(a).to(dev)
    targets = [types.SimpleNamespace(keypoints_3d=tt(arr["keypoints_3d"]), K=tt(arr["K"]), mask=tt(arr["mask"][i]),
                                     class_ids=tt(arr["class_ids"][i]), rotations=tt(arr["rotations"][i]),
                                     translations=tt(arr["translations"][i]), bbox_trans=tt(arr["bbox_trans"][i])) for i in range(nimg)]
    s_cls, s_reg = scenario.make_head_outputs(nimg, s_hw, 200, teacher=False, target_seed=0)
    t_cls, t_reg = scenario.make_head_outputs(nimg, t_hw, 100, teacher=True, target_seed=0)
    d_cls = [torch.from_numpy(a).to(dev).requires_grad_(True) for a in s_cls]
    d_reg = [torch.from_numpy(a).to(dev).requires_grad_(True) for a in s_reg]
    tc = [torch.from_numpy(a).to(dev) for a in t_cls]
    tr = [torch.from_numpy(a).to(dev) for a in t_reg]
    sel = select_cells(tc, tr, scenario.ANCHOR_SIZES, scenario.ANCHOR_STRIDES, 0.1, 10, 1.0)
    torch.cuda.synchronize(dev)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    sel_us = []
    for _ in range(25):
        torch.cuda._sleep(400000)   # GPU-side delay: the host enqueues <e0, kernel, e1> before the GPU reaches e0
        e0.record()
        select_cells(tc, tr, scenario.ANCHOR_SIZES, scenario.ANCHOR_STRIDES, 0.1, 10, 1.0)
        e1.record()
        e1.synchronize()
        sel_us.append(e0.elapsed_time(e1) * 1e3)
    sel_us = statistics.median(sel_us[5:])
    logit_bytes = sum(a.size for a in t_cls) * 4
    peaks, _src = measured_peaks()
    # teacher dict for the loss from the selection itself (level-0 cells of class 0; PnP is host work outside this leg)
    cnt = sel["sel_count"].view(nimg, -1)[:, 0].cpu().tolist()
    kp = sel["sel_kpts"].view(nimg, -1, sel["cap"], 16)[:, 0]
    sc = sel["sel_score"].view(nimg, -1, sel["cap"])[:, 0]
    kp2d = torch.cat([kp[i, :cnt[i]].view(-1, 2, 8).transpose(1, 2) for i in range(nimg)]).contiguous()
    kcls = torch.cat([sc[i, :cnt[i]].view(-1, 1).repeat(1, 8) for i in range(nimg)]).contiguous()
    KDPoseLoss = make_kd_pose_loss(doubles.ReplayBase)
    lv = grid_anchors(s_hw, scenario.ANCHOR_SIZES, scenario.ANCHOR_STRIDES, device=dev)
    anchors = [lv for _ in range(nimg)]
    out = {"images": nimg, "unit": "ms per forward+backward of [cls, reg, kd] through KDPoseLoss.__call__ (seam B0), host wall clock"}
    for mode in ("philox", "parity"):
        fn = KDPoseLoss(2.0, 0.25, scenario.ANCHOR_SIZES, scenario.ANCHOR_STRIDES, "SSC", 10, 1.0, 9, scenario.INTERNAL_K,
                        scenario.MESH_DIAMETERS, TargetCoder("POINT", scenario.ANCHOR_SIZES, scenario.ANCHOR_STRIDES, target_type="3D"),
                        dict(scenario.CFG_KD, DEVICE_TARGETS=mode))

        def run():
            for t in d_cls + d_reg:
                t.grad = None
            c, r, k = fn(d_cls, d_reg, targets, anchors, {"post_kp_2d": kp2d.clone(), "post_kp_cls": kcls, "post_pos_per_img": cnt})
            (0.1 * c + r + 5.0 * k).backward()

        out[f"device_targets_{mode}_ms"] = timed(run)
    out["images_per_s_philox"] = nimg / (out["device_targets_philox_ms"] * 1e-3)
    out["note"] = ("real (not replayed) SSC target assignment on the device; the reference's own KDPoseLoss on the same GPU takes "
                   "~530 ms for this batch (tools/time_kd_pose_loss.py, profiles/r02_kd_pose_loss_timing.json)")
    select = {"kernel": "kdot_select_kernel", "images": nimg, "us": sel_us, "logit_bytes": logit_bytes,
              "GBps": logit_bytes / sel_us / 1e3, "hbm_peak_GBps": peaks.get("hbm_gbs"),
              "frac_of_hbm": logit_bytes / sel_us / 1e3 / peaks.get("hbm_gbs", 6552.0),
              "note": "one streaming pass over nimg x cells x 15 teacher logits is the bound; the launch is latency-bound at this size"}
    return out, select


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=20)
    ap.add_argument("--impl", default="kdot", choices=["kdot", "reference"])
    ap.add_argument("--workload", default="ape_b64", choices=sorted(WORKLOADS))
    ap.add_argument("--images", type=int, default=None, help="images per GPU (default: the workload's batch)")
    ap.add_argument("--scaling", type=float, default=None, help="override the epsilon-scaling ratio (iteration sweep)")
    ap.add_argument("--blur", type=float, default=None, help="override the blur (final temperature = blur**p)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-dense", action="store_true", help="skip the secondary dense-workload roofline leg")
    ap.add_argument("--no-b0", action="store_true", help="skip the KDPoseLoss.__call__ / selection-kernel legs")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))

    if args.impl == "reference":
        run_reference_arm(args, rank, world)
        return

    import torch
    import torch.distributed as dist

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- the kdot path has no CPU fallback (use --impl reference for the CPU arm)")
    dev = torch.device("cuda", local_rank)
    torch.cuda.set_device(dev)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    def barrier():
        if world > 1:
            dist.barrier()

    def max_over_ranks(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    from kd_6d_pose_adlp_b200 import _lib

    L = _lib.lib()
    peaks, peak_src = measured_peaks()
    fp32_peak = float(L.kdot_measure_fp32_peak_tflops(local_rank, 2000))

    batch = make_batch(args.workload, rank, nimg=args.images)
    cfg = workload_cfg(args.workload, args)
    nimg = len(batch["pos_per_img"])
    bench = DeviceBench(batch, dev, cfg)
    sampler = ClockSampler(local_rank)
    sampler.start()
    total_ms, per, launches = bench.timed(args.steps, args.warmup, barrier)
    clocks = sampler.result()
    total_ms = max_over_ranks(total_ms)
    ms_per_step = total_ms / args.steps
    value = nimg * world / (ms_per_step * 1e-3)

    nits = bench.nits.cpu().numpy()
    flops, exps, byts = algorithmic_work(batch, nits)
    med_ms = statistics.median(per)
    roofline = {
        "bound": "fp32", "kernel": kernel_name(bench.max_n, bench.max_m, launches // max(args.steps, 1)),
        "achieved": flops / (ms_per_step * 1e-3) / 1e12, "peak": fp32_peak, "unit": "TFLOP/s",
        "frac": flops / (ms_per_step * 1e-3) / 1e12 / fp32_peak if fp32_peak > 0 else None,
        "peak_source": "FP32 FMA chain measured live on this GPU (kdot_measure_fp32_peak_tflops)",
        "traffic": None,  # filled below from the committed ncu capture of this kernel / workload
        "sfu_exp_per_s": exps / (ms_per_step * 1e-3),
        "hbm": {"achieved": byts / (ms_per_step * 1e-3) / 1e9, "peak": peaks.get("hbm_gbs"), "unit": "GB/s",
                "frac": byts / (ms_per_step * 1e-3) / 1e9 / peaks.get("hbm_gbs", 6552.0), "peak_source": peak_src},
        "algorithmic_flops_per_step": flops, "algorithmic_bytes_per_step": byts, "median_ms_per_step": med_ms,
        "note": ("ape-shaped problems are ~20 x 20 points: the launch is bound by the latency of its dependent Sinkhorn "
                 "rounds, not by a pipe; see the `dense` object for the roofline-relevant configuration")
        if args.workload == "ape_b64" else None,
    }

    roofline["traffic"] = ncu_traffic_bytes(roofline["kernel"], args.workload)
    roofline["traffic_source"] = "dram__bytes_read.sum + dram__bytes_write.sum per launch, ncu --set full (profiles/)"
    if roofline["kernel"] == "kdot_stream_kernel":
        roofline["traffic_note"] = ("includes the write-back / re-fetch of the kernel's L2-resident scratch (staged clouds, potentials, "
                                    "float64 potentials, fp32 head + tail of h, tile maxima: ~70 B per point, rewritten every round) under ncu's cache control; < 0.5 % of HBM bandwidth, the kernel is SFU/FP32 bound")

    # end to end through the host-buffer C-ABI call
    e2e_dt, h2d, d2h, mean_loss = host_e2e(batch, local_rank, args.steps, args.warmup, barrier, cfg)
    e2e_dt = max_over_ranks(e2e_dt)
    e2e = {"value": nimg * world * args.steps / e2e_dt, "unit": UNIT, "h2d_bytes_per_step": h2d,
           "d2h_bytes_per_step": d2h, "ms_per_step": e2e_dt / args.steps * 1e3, "mean_kd_loss": mean_loss,
           "last_call_phases": getattr(host_e2e, "last_phases", None)}

    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
        "data": "synthetic",
        "config": {"workload": args.workload, "desc": WORKLOADS[args.workload]["desc"], "images_per_gpu": nimg,
                   "l2": "256 MiB buffer written between timed steps (L2 flush), inputs restored outside the events",
                   "parallelism": f"images sharded over {world} rank(s), no data-path collective",
                   "softmin_rounds_per_image": int(np.median(nits[nits > 0])) + 2 if (nits > 0).any() else 0, **cfg},
        "roofline": roofline, "e2e": e2e, "gpu_launches": int(launches), "clocks": clocks,
    }

    if not args.no_dense and args.workload != "dense_b32":
        # secondary leg on EVERY rank: the roofline-relevant dense configuration (BASELINE.json configs[2], variant 3b:
        # "batch 256 sharded over 8 x B200" is exactly this leg at N = 8 -- 32 images per GPU).  Own clock sampler,
        # device-timed, max over ranks; same L2 flush / restore protocol as the main leg.
        d_steps = max(min(args.steps, 20), 3)
        db = make_batch("dense_b32", rank)
        dbench = DeviceBench(db, dev)
        d_sampler = ClockSampler(local_rank)
        d_sampler.start()
        d_total, d_per, d_launch = dbench.timed(d_steps, 3, barrier)
        d_clocks = d_sampler.result()
        d_ms = max_over_ranks(d_total) / d_steps
        d_fl, d_ex, d_by = algorithmic_work(db, dbench.nits.cpu().numpy())
        d_kernel = kernel_name(dbench.max_n, dbench.max_m, d_launch // d_steps)
        line["dense"] = {
            "workload": "dense_b32", "desc": WORKLOADS["dense_b32"]["desc"], "n_gpus": world, "steps": d_steps, "warmup": 3,
            "value": len(db["pos_per_img"]) * world / (d_ms * 1e-3), "unit": UNIT, "ms_per_step": d_ms, "scaling": "weak",
            "clocks": d_clocks,
            "roofline": {"bound": "fp32", "kernel": d_kernel, "achieved": d_fl / (d_ms * 1e-3) / 1e12,
                         "peak": fp32_peak, "unit": "TFLOP/s", "frac": d_fl / (d_ms * 1e-3) / 1e12 / fp32_peak,
                         "per": "GPU (rank 0's algorithmic FLOPs over the max-over-ranks step time)",
                         "sfu_exp_per_s": d_ex / (d_ms * 1e-3),
                         "traffic": ncu_traffic_bytes(d_kernel, "dense_b32"),
                         "algorithmic_bytes_per_step": d_by,
                         "hbm_gbs": d_by / (d_ms * 1e-3) / 1e9},
        }
        del dbench

    if world > 1:
        line["allreduce"] = allreduce_leg(dev, bench, barrier, max_over_ranks, world, rank)

    if rank == 0 and world == 1 and not args.no_b0 and args.workload == "ape_b64":
        try:
            line["e2e_b0"], line["select"] = b0_and_select_legs(dev)
        except Exception as exc:  # these legs import tests/scenario.py (synthetic head outputs); never fail the headline for them
            line["e2e_b0"] = {"unavailable": repr(exc)}

    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cb = batch if WORKLOADS[args.workload]["dense"] is None else make_batch(args.workload, 0, nimg=1)
        v, passes, dt, cores = cpu_port_images_per_sec(cb, cfg, threads=len(os.sched_getaffinity(0)),
                                                       min_passes=1 if cb is not batch else 2)
        line["cpu_baseline"] = {"value": v, "unit": UNIT, "cores": cores, "kind": "port",
                                "sample": f"{passes} passes over a {len(cb['pos_per_img'])}-image {args.workload} batch in {dt:.1f} s "
                                          "(fp32 torch CPU ops + autograd: the reference formulation)"}
    if rank == 0:
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()

"""BASELINE.json configs[1] / SURVEY.md section 8(d) "Config 2": the fused kernel against the reference formulation
issued as ATen ops ON THE SAME GPU (restated ``kd_loss_2d`` loop + geomloss restatement with autograd), both timed
with CUDA events; kernel launches per step counted for both.  The numbers are written to
``gpurun_out/aten_vs_fused.json`` (copied to ``profiles/`` by hand); the assertions only pin the order of magnitude.
"""
import json
import os
import statistics

import numpy as np
import pytest
import torch

from kd_6d_pose_adlp_b200 import _lib
from kd_6d_pose_adlp_b200.ops import OTConfig, ot_loss_batched
from kd_6d_pose_adlp_b200.synthetic import ot_batch
from oracle import geomloss_ref, kd_loss_ref

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _aten_step(batch, dev, L):
    xs = torch.from_numpy(batch["xs"].reshape(-1, 2)).to(dev).requires_grad_(True)
    ws = torch.from_numpy(batch["ws"]).to(dev).requires_grad_(True)
    xt = torch.from_numpy(batch["xt"].reshape(-1, 2)).to(dev)
    wt = torch.from_numpy(batch["wt"]).to(dev)

    def step():
        xs.grad = ws.grad = None
        losses = kd_loss_ref.kd_loss_2d_ref(xs.clone(), xt.clone(), ws, wt, 640.0, 480.0, "point", L, dim=2,
                                            pos_per_img=batch["pos_per_img"], pos_per_img_t=batch["pos_per_img_t"])
        loss = sum(losses) / len(losses)
        loss.backward()
        return loss

    return step, xs, ws


def _event_ms(fn, warmup, iters):
    for _ in range(warmup):
        fn()
    out = []
    for _ in range(iters):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        out.append(e0.elapsed_time(e1))
    return statistics.median(out)


def _count_cuda_kernels(fn):
    from torch.profiler import ProfilerActivity, profile

    fn()
    torch.cuda.synchronize()
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        fn()
        torch.cuda.synchronize()
    return sum(1 for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA)


def test_fused_kernel_vs_aten_reference_on_gpu():
    dev = torch.device("cuda:0")
    torch.backends.cuda.matmul.allow_tf32 = False
    nimg = 64
    batch = ot_batch(nimg, seed=1234)
    cfg = OTConfig()
    L = geomloss_ref.SamplesLoss("sinkhorn", p=cfg.p, blur=cfg.blur, scaling=cfg.scaling, reach=cfg.reach)
    aten, xs_a, ws_a = _aten_step(batch, dev, L)
    aten_ms = _event_ms(aten, 2, 5)
    loss_a = float(aten())
    try:
        aten_launches = _count_cuda_kernels(aten)
    except Exception:  # profiler (CUPTI) unavailable on the box: the timing comparison still stands
        aten_launches = None

    t = {k: torch.from_numpy(batch[k]).to(dev) for k in ("xs", "ws", "xt", "wt")}
    xs0 = t["xs"].clone()

    def fused():
        t["xs"].copy_(xs0)
        return ot_loss_batched(t["xs"], t["ws"], t["xt"].clone(), t["wt"], batch["pos_per_img"], batch["pos_per_img_t"], cfg)

    l0 = _lib.launch_count()
    out = fused()
    fused_launches = _lib.launch_count() - l0
    fused_ms = _event_ms(fused, 5, 50)  # includes the public wrapper's allocations and the cu_seqlens upload
    n_valid = int((out["valid"] == 1).sum())
    loss_f = float(out["loss_per_img"].sum()) / n_valid
    gx_f = (out["grad_xs"] / n_valid).reshape(-1, 2)

    # the same reference formulation on the HOST (fp32 torch CPU ops): how far the reference is from itself across
    # its two backends is the yardstick for what "matches the reference" can mean for d/dx at eps = 1e-6
    xs_c = torch.from_numpy(batch["xs"].reshape(-1, 2)).clone().requires_grad_(True)
    ws_c = torch.from_numpy(batch["ws"]).clone().requires_grad_(True)
    losses_c = kd_loss_ref.kd_loss_2d_ref(xs_c.clone(), torch.from_numpy(batch["xt"].reshape(-1, 2)).clone(), ws_c,
                                          torch.from_numpy(batch["wt"]), 640.0, 480.0, "point", L, dim=2,
                                          pos_per_img=batch["pos_per_img"], pos_per_img_t=batch["pos_per_img_t"])
    (sum(losses_c) / len(losses_c)).backward()
    gmax = float(xs_c.grad.abs().max())
    ref_cpu_vs_gpu = float((xs_c.grad - xs_a.grad.cpu()).abs().max()) / gmax
    fused_vs_cpu = float((gx_f.cpu() - xs_c.grad).abs().max()) / gmax

    assert abs(loss_f - loss_a) <= 1e-4 * max(abs(loss_a), 1.0)
    # d/dx: both are fp32 evaluations of an eps = 1e-6 problem; they agree to the fp32 formulation's own accuracy
    rel = float((gx_f - xs_a.grad).abs().max() / xs_a.grad.abs().max())
    assert rel < 2e-2
    # ... and the kernel is no farther from either backend of the reference than they are from each other
    assert max(rel, fused_vs_cpu) <= 1.5 * ref_cpu_vs_gpu + 1e-4
    assert fused_launches == 1
    assert fused_ms * 100 < aten_ms

    rec = {"workload": "ape_b64 (BASELINE.json configs[1])", "images": nimg,
           "aten_reference_gpu": {"ms_per_step": aten_ms, "images_per_s": nimg / aten_ms * 1e3, "cuda_kernels_per_step": aten_launches},
           "fused_public_api": {"ms_per_step": fused_ms, "images_per_s": nimg / fused_ms * 1e3, "cuda_kernels_per_step": fused_launches,
                                "note": "ot_loss_batched incl. output allocation + cu_seqlens H2D; bench.py times the bare C-ABI call"},
           "speedup": aten_ms / fused_ms, "mean_loss": {"aten": loss_a, "fused": loss_f}, "grad_xs_rel_maxnorm_diff": {"fused_vs_reference_gpu": rel, "fused_vs_reference_cpu": fused_vs_cpu,
                                        "reference_cpu_vs_reference_gpu": ref_cpu_vs_gpu},
           "gpu": torch.cuda.get_device_name(0)}
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    with open(os.path.join(ROOT, "gpurun_out", "aten_vs_fused.json"), "w") as fh:
        json.dump(rec, fh, indent=1)
    print(json.dumps(rec))

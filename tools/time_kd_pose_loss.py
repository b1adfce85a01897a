#!/usr/bin/env python
"""Seam-B0 timing on the GPU box: ``KDPoseLoss.__call__`` + backward at the ape shape, batch 64, with the fused
gather/decode prologue (what ``__call__`` runs) against the reference's op sequence for that tensor (flatten all of pred_reg,
index the positives, ``TargetCoder.decode``) swapped into the same ``__call__``.  Target assignment is
replayed (synthetic: 10 positive cells per image), so the number isolates the device-side loss path.

    python tools/time_kd_pose_loss.py [nimg]
"""
import json
import os
import statistics
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from kd_6d_pose_adlp_b200.losses.kd_loss import make_kd_pose_loss  # noqa: E402
from kd_6d_pose_adlp_b200.target_coder import TargetCoder, grid_anchors  # noqa: E402
from tests import doubles, scenario  # noqa: E402

HW = [(32, 32), (16, 16), (8, 8), (4, 4)]


def main():
    nimg = int(sys.argv[1]) if len(sys.argv) > 1 else 64
    dev = torch.device("cuda:0")
    g = torch.Generator().manual_seed(0)
    cells = sum(h * w for h, w in HW)
    s_cls = [(torch.randn(nimg, 15, h, w, generator=g) - 2).to(dev).requires_grad_(True) for h, w in HW]
    s_reg = [(torch.randn(nimg, 240, h, w, generator=g) * 0.3).to(dev).requires_grad_(True) for h, w in HW]
    labels = torch.zeros(nimg, cells, dtype=torch.int64)
    for i in range(nimg):
        labels[i, torch.randperm(cells, generator=g)[:10]] = 1
    bt = torch.tensor([[1.6, 0.0, -300.0], [0.0, 1.6, -200.0]]).repeat(nimg * cells, 1, 1)
    a3 = torch.randn(nimg * cells, 8, 3, generator=g) * 50 + torch.tensor([0.0, 0.0, 900.0])
    split = lambda t: list(torch.split(t.to(dev), [cells] * nimg))
    doubles.ReplayBase.recorded = dict(labels=split(labels.view(-1)), reg_targets=split(torch.zeros(nimg * cells, 16)),
                                       aux_raw_boxes=split(torch.zeros(nimg * cells, 4)), aux_3d=split(a3),
                                       aux_bbox_trans=split(bt))
    KDPoseLoss = make_kd_pose_loss(doubles.ReplayBase)
    loss_fn = KDPoseLoss(2.0, 0.25, scenario.ANCHOR_SIZES, scenario.ANCHOR_STRIDES, "SSC", 10, 1.0, 9,
                         scenario.INTERNAL_K, scenario.MESH_DIAMETERS,
                         TargetCoder("POINT", scenario.ANCHOR_SIZES, scenario.ANCHOR_STRIDES, target_type="3D"),
                         dict(scenario.CFG_KD))
    lv = grid_anchors(HW, scenario.ANCHOR_SIZES, scenario.ANCHOR_STRIDES, device=dev)
    anchors = [lv for _ in range(nimg)]
    kp = (torch.rand(nimg * 10, 8, 2, generator=g) * torch.tensor([640.0, 480.0])).to(dev)
    kc = (0.3 + 0.6 * torch.rand(nimg * 10, 1, generator=g)).repeat(1, 8).to(dev)

    def teacher():
        return {"post_kp_2d": kp.clone(), "post_kp_cls": kc, "post_pos_per_img": [10] * nimg}

    def fused():
        c, r, k = loss_fn(s_cls, s_reg, None, anchors, teacher())
        (0.1 * c + r + 5.0 * k).backward()
        return float(k)

    from kd_6d_pose_adlp_b200.losses import kd_loss as kd_mod

    coder = loss_fn.target_coder
    fused_op = kd_mod.gather_decode

    def reference_ops(pred_reg, pos_inds, cls_label, anchors_pos, bt):
        """The reference's op sequence for the same tensor (loss.py:62-96, kd_loss.py:156,47-50, model.py:144-166)."""
        flat = kd_mod.flatten_level_list(pred_reg)[pos_inds]
        n = flat.shape[0]
        picked = flat.view(n, -1, 16)[torch.arange(n, device=flat.device), cls_label]
        return coder.decode(picked, anchors_pos, bt).view(-1, 2, 8).transpose(1, 2).contiguous().view(-1, 2)

    def unfused():
        kd_mod.gather_decode = reference_ops
        try:
            return fused()
        finally:
            kd_mod.gather_decode = fused_op

    def timed(fn, warm=5, iters=20):
        for _ in range(warm):
            fn()
        out = []
        for _ in range(iters):
            for t in s_cls + s_reg:
                t.grad = None
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            fn()
            torch.cuda.synchronize()
            out.append((time.perf_counter() - t0) * 1e3)
        return statistics.median(out)

    kf, ku = fused(), unfused()
    rec = {"nimg": nimg, "kd_loss": {"fused": kf, "unfused": ku},
           "ms_fwd_bwd": {"fused_gather_decode": timed(fused), "flatten_index_decode": timed(unfused)},
           "gpu": torch.cuda.get_device_name(0)}
    rec["speedup"] = rec["ms_fwd_bwd"]["flatten_index_decode"] / rec["ms_fwd_bwd"]["fused_gather_decode"]
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    with open(os.path.join(ROOT, "gpurun_out", "kd_pose_loss_timing.json"), "w") as fh:
        json.dump(rec, fh, indent=1)
    print(json.dumps(rec))


if __name__ == "__main__":
    main()

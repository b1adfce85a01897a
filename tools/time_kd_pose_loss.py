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

    # ---- REAL (not replayed) target assignment on the device: cfg_kd["DEVICE_TARGETS"] = parity | philox ----
    import types

    arr = scenario.make_target_arrays(nimg, 0)
    tt = lambda a: torch.tensor(a).to(dev)
    targets = [types.SimpleNamespace(keypoints_3d=tt(arr["keypoints_3d"]), K=tt(arr["K"]), mask=tt(arr["mask"][i]),
                                     class_ids=tt(arr["class_ids"][i]), rotations=tt(arr["rotations"][i]),
                                     translations=tt(arr["translations"][i]), bbox_trans=tt(arr["bbox_trans"][i])) for i in range(nimg)]
    h_cls, h_reg = scenario.make_head_outputs(nimg, HW, 200, teacher=False, target_seed=0)
    d_cls = [torch.from_numpy(a).to(dev).requires_grad_(True) for a in h_cls]
    d_reg = [torch.from_numpy(a).to(dev).requires_grad_(True) for a in h_reg]

    def make_real(mode):
        fn = KDPoseLoss(2.0, 0.25, scenario.ANCHOR_SIZES, scenario.ANCHOR_STRIDES, "SSC", 10, 1.0, 9,
                        scenario.INTERNAL_K, scenario.MESH_DIAMETERS,
                        TargetCoder("POINT", scenario.ANCHOR_SIZES, scenario.ANCHOR_STRIDES, target_type="3D"),
                        dict(scenario.CFG_KD, DEVICE_TARGETS=mode))

        def run():
            for t in d_cls + d_reg:
                t.grad = None
            c, r, k = fn(d_cls, d_reg, targets, anchors, teacher())
            (0.1 * c + r + 5.0 * k).backward()
            return float(k.detach())
        return run

    real = {m: timed(make_real(m)) for m in ("parity", "philox")}
    ref_ms = None
    ref_root = os.path.join(ROOT, "baseline", "_ref")
    if os.path.isdir(os.path.join(ref_root, "losses")):
        # the reference's own KDPoseLoss (its prepare_targets loops + the restated geomloss as stock torch ops) on the same GPU
        os.environ["KDOT_REFERENCE_ROOT"] = ref_root
        from oracle import ref_loader

        ref_loader.REFERENCE_ROOT = ref_root
        ref = ref_loader.load()
        r_targets = [ref.PoseAnnot(torch.tensor(arr["keypoints_3d"]), torch.tensor(arr["K"]), torch.tensor(arr["mask"][i]),
                                   torch.tensor(arr["class_ids"][i]), torch.tensor(arr["rotations"][i]),
                                   torch.tensor(arr["translations"][i]), 256, 256, bbox_scale=torch.tensor(1.0),
                                   bbox_trans=torch.tensor(arr["bbox_trans"][i])).to(dev) for i in range(nimg)]
        gen = ref.modules["model"].make_anchor_generator_atss(scenario.ANCHOR_SIZES[:4], scenario.ANCHOR_STRIDES[:4]).to(dev)
        il = types.SimpleNamespace(sizes=[(256, 256)] * nimg, image_sizes=[(256, 256)] * nimg)
        r_anchors = gen(il, [torch.zeros(nimg, 1, h, w, device=dev) for h, w in HW])
        r_fn = ref.KDPoseLoss(2.0, 0.25, scenario.ANCHOR_SIZES, scenario.ANCHOR_STRIDES, "SSC", 10, 1.0, 9, scenario.INTERNAL_K,
                              scenario.MESH_DIAMETERS, ref.TargetCoder("POINT", scenario.ANCHOR_SIZES, scenario.ANCHOR_STRIDES, target_type="3D"),
                              dict(scenario.CFG_KD, vis_dir="/tmp/kdot_vis_time"))
        r_fn.step = 5   # past the step-0 plot

        def run_ref():
            for t in d_cls + d_reg:
                t.grad = None
            c, r, k = r_fn(d_cls, d_reg, r_targets, r_anchors, teacher())
            (0.1 * c + r + 5.0 * k).backward()
            return float(k.detach())

        ref_ms = timed(run_ref, warm=2, iters=5)

    kf, ku = fused(), unfused()
    rec = {"nimg": nimg, "kd_loss": {"fused": kf, "unfused": ku},
           "ms_fwd_bwd": {"replayed_targets_fused": timed(fused), "replayed_targets_flatten_index_decode": timed(unfused),
                          "device_targets_parity": real["parity"], "device_targets_philox": real["philox"],
                          "reference_KDPoseLoss_on_this_gpu": ref_ms},
           "note": "forward + backward of [cls, reg, kd] through seam B0, batch %d; 'device_targets_*' include the REAL SSC label "
                   "assignment (kdot_ssc_*), 'reference' is the reference's own KDPoseLoss (host-loop prepare_targets, restated geomloss "
                   "as stock torch ops) on the same GPU" % nimg,
           "gpu": torch.cuda.get_device_name(0)}
    rec["speedup_vs_reference"] = None if ref_ms is None else ref_ms / real["philox"]
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    with open(os.path.join(ROOT, "gpurun_out", "kd_pose_loss_timing.json"), "w") as fh:
        json.dump(rec, fh, indent=1)
    print(json.dumps(rec))


if __name__ == "__main__":
    main()

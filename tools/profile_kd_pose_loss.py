#!/usr/bin/env python
"""Where the host time of seam B0 goes: torch.profiler over ``KDPoseLoss.__call__`` + backward with the device-side SSC
assignment (``DEVICE_TARGETS = philox``), ape shape, batch 64.  Prints the launch count per step and the ops ordered by
host time.

    python tools/profile_kd_pose_loss.py [nimg] [mode]
"""
import os
import sys
import time
import types

import torch
from torch.profiler import ProfilerActivity, profile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from kd_6d_pose_adlp_b200.losses.kd_loss import make_kd_pose_loss  # noqa: E402
from kd_6d_pose_adlp_b200.target_coder import TargetCoder, grid_anchors  # noqa: E402
from tests import doubles, scenario  # noqa: E402

HW = [(32, 32), (16, 16), (8, 8), (4, 4)]


def main():
    nimg = int(sys.argv[1]) if len(sys.argv) > 1 else 64
    mode = sys.argv[2] if len(sys.argv) > 2 else "philox"
    dev = torch.device("cuda:0")
    g = torch.Generator().manual_seed(0)
    arr = scenario.make_target_arrays(nimg, 0)
    tt = lambda a: torch.tensor(a).to(dev)
    kp3d, K = tt(arr["keypoints_3d"]), tt(arr["K"])
    targets = [types.SimpleNamespace(keypoints_3d=kp3d, K=K, mask=tt(arr["mask"][i]), class_ids=tt(arr["class_ids"][i]),
                                     rotations=tt(arr["rotations"][i]), translations=tt(arr["translations"][i]),
                                     bbox_trans=tt(arr["bbox_trans"][i])) for i in range(nimg)]
    h_cls, h_reg = scenario.make_head_outputs(nimg, HW, 200, teacher=False, target_seed=0)
    d_cls = [torch.from_numpy(a).to(dev).requires_grad_(True) for a in h_cls]
    d_reg = [torch.from_numpy(a).to(dev).requires_grad_(True) for a in h_reg]
    KDPoseLoss = make_kd_pose_loss(doubles.ReplayBase)
    fn = KDPoseLoss(2.0, 0.25, scenario.ANCHOR_SIZES, scenario.ANCHOR_STRIDES, "SSC", 10, 1.0, 9, scenario.INTERNAL_K,
                    scenario.MESH_DIAMETERS, TargetCoder("POINT", scenario.ANCHOR_SIZES, scenario.ANCHOR_STRIDES, target_type="3D"),
                    dict(scenario.CFG_KD, DEVICE_TARGETS=mode))
    lv = grid_anchors(HW, scenario.ANCHOR_SIZES, scenario.ANCHOR_STRIDES, device=dev)
    anchors = [lv for _ in range(nimg)]
    kp = (torch.rand(nimg * 10, 8, 2, generator=g) * torch.tensor([640.0, 480.0])).to(dev)
    kc = (0.3 + 0.6 * torch.rand(nimg * 10, 1, generator=g)).repeat(1, 8).to(dev)

    def step():
        for t in d_cls + d_reg:
            t.grad = None
        c, r, k = fn(d_cls, d_reg, targets, anchors, {"post_kp_2d": kp.clone(), "post_kp_cls": kc, "post_pos_per_img": [10] * nimg})
        (0.1 * c + r + 5.0 * k).backward()

    for _ in range(5):
        step()
    torch.cuda.synchronize()
    ts = []
    for _ in range(20):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        step()
        torch.cuda.synchronize()
        ts.append((time.perf_counter() - t0) * 1e3)
    print("median ms/step", sorted(ts)[len(ts) // 2])
    if os.environ.get("KDOT_CPROFILE"):
        import cProfile, pstats
        pr = cProfile.Profile()
        pr.enable()
        for _ in range(20):
            step()
        torch.cuda.synchronize()
        pr.disable()
        st = pstats.Stats(pr)
        st.sort_stats("cumulative").print_stats(45)
        return
    with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
        for _ in range(5):
            step()
        torch.cuda.synchronize()
    ka = prof.key_averages()
    launches = sum(e.count for e in ka if e.key in ("cudaLaunchKernel", "cudaLaunchKernelExC", "cuLaunchKernel", "cudaMemcpyAsync", "cudaMemsetAsync", "cuLaunchKernelEx"))
    print("launch-type API calls per step:", launches / 5)
    print(ka.table(sort_by="self_cpu_time_total", row_limit=45, max_name_column_width=60))


if __name__ == "__main__":
    main()

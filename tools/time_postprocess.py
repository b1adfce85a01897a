#!/usr/bin/env python
"""Teacher knowledge extraction (``PostProcessorKD.forward``) on the GPU box: selection kernel + one D2H copy +
RANSAC-EPnP per image, serial (the reference's loop) against the host thread pool.

    python tools/time_postprocess.py [nimg]
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
from kd_6d_pose_adlp_b200.postprocess.postprocess_kd import PostProcessorKD  # noqa: E402
from kd_6d_pose_adlp_b200.target_coder import TargetCoder  # noqa: E402
from tests import doubles, scenario  # noqa: E402

T_HW = [(32, 32), (16, 16), (8, 8), (4, 4), (2, 2)]


def main():
    nimg = int(sys.argv[1]) if len(sys.argv) > 1 else 64
    z = np.load(os.path.join(ROOT, "tests", "golden", "postprocess_kd.npz"))
    seed = int(z["seed"])
    t_cls, t_reg = scenario.make_head_outputs(nimg, T_HW, seed + 100, teacher=True, target_seed=seed)
    dev = torch.device("cuda:0")
    bts = z["bbox_trans"]
    targets = [doubles.Target(torch.tensor(z["K"]), torch.tensor(z["keypoints_3d"]), torch.tensor(bts[i % len(bts)]))
               for i in range(nimg)]
    pp = PostProcessorKD(0.1, TargetCoder("POINT", scenario.ANCHOR_SIZES, scenario.ANCHOR_STRIDES), 10, 1.0, {})
    cls = [torch.from_numpy(a).to(dev) for a in t_cls]
    reg = [torch.from_numpy(a).to(dev) for a in t_reg]

    def timed(threads):
        pp.pnp_threads = threads
        out = []
        for it in range(8):
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            res = pp(cls, reg, targets, None)
            torch.cuda.synchronize()
            if it >= 2:
                out.append((time.perf_counter() - t0) * 1e3)
        return statistics.median(out), sum(len(r) for r in res[0])

    cores = len(os.sched_getaffinity(0))
    serial, cells = timed(1)
    pooled, cells2 = timed(min(16, cores))
    assert cells == cells2
    rec = {"nimg": nimg, "selected_cells": cells, "host_cores": cores, "ms_serial_pnp": serial, "ms_thread_pool_pnp": pooled,
           "speedup": serial / pooled}
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    with open(os.path.join(ROOT, "gpurun_out", "postprocess_timing.json"), "w") as fh:
        json.dump(rec, fh, indent=1)
    print(json.dumps(rec))


if __name__ == "__main__":
    main()

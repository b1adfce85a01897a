#!/usr/bin/env python
"""A/B of the streaming kernel's tile skipping: the shipped library against a -DKDOT_NO_TILE_SKIP build (same Morton
order, seeds and arithmetic, every tile evaluated), outputs compared bit for bit.  Run on the GPU box:

    python tools/ab_tile_skip.py /path/to/noskip.so
"""
import json
import os
import subprocess
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CASES = {"dense": dict(nimg=4, dense=(1360, 1364), sigma=0.1, B=8), "tight": dict(nimg=2, dense=(600, 640), sigma=0.005, B=2),
         "wide": dict(nimg=2, dense=(900, 300), sigma=0.3, B=1)}


def child(out_path):
    import torch

    sys.path.insert(0, ROOT)
    from kd_6d_pose_adlp_b200.ops import OTConfig, ot_loss_batched
    from kd_6d_pose_adlp_b200.synthetic import ot_batch

    dev = torch.device("cuda:0")
    res = {}
    for name, c in CASES.items():
        b = ot_batch(c["nimg"], seed=7, dense=c["dense"], sigma=c["sigma"], B=c["B"])
        t = {k: torch.from_numpy(b[k]).to(dev) for k in ("xs", "ws", "xt", "wt")}
        o = ot_loss_batched(t["xs"], t["ws"], t["xt"], t["wt"], b["pos_per_img"], b["pos_per_img_t"], OTConfig())
        torch.cuda.synchronize()
        for k in ("loss_per_img", "grad_xs", "grad_ws"):
            res[f"{name}.{k}"] = o[k].cpu().numpy()
    np.savez(out_path, **res)


def main():
    if sys.argv[1] == "--child":
        return child(sys.argv[2])
    outs = {}
    for tag, lib in (("skip", None), ("noskip", sys.argv[1])):
        env = dict(os.environ)
        if lib:
            env["KDOT_LIB"] = os.path.abspath(lib)
        path = f"/tmp/ab_{tag}.npz"
        subprocess.run([sys.executable, os.path.abspath(__file__), "--child", path], check=True, env=env)
        outs[tag] = np.load(path)
    rec = {}
    for k in outs["skip"].files:
        a, b = outs["skip"][k], outs["noskip"][k]
        rec[k] = {"bit_identical": bool(np.array_equal(a.view(np.uint32), b.view(np.uint32))),
                  "max_abs_diff": float(np.abs(a - b).max()), "max_abs": float(np.abs(b).max())}
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    json.dump(rec, open(os.path.join(ROOT, "gpurun_out", "ab_tile_skip.json"), "w"), indent=1)
    print(json.dumps(rec, indent=1))


if __name__ == "__main__":
    main()

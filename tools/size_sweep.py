#!/usr/bin/env python
"""Kernel time / achieved FP32 throughput over problem sizes (N = M cells per image, B = 8, D = 2) on the GPU box:
shows where each kernel family (small-fast <= 32 points, CTA-resident tiled 33..256, streaming above) sits against the roofline.

    python tools/size_sweep.py [nimg] [n1,n2,...]
"""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from kd_6d_pose_adlp_b200 import _lib  # noqa: E402
from kd_6d_pose_adlp_b200.synthetic import ot_batch  # noqa: E402


def main():
    nimg = int(sys.argv[1]) if len(sys.argv) > 1 else 64
    dev = torch.device("cuda:0")
    torch.cuda.set_device(dev)
    peak = float(_lib.lib().kdot_measure_fp32_peak_tflops(0, 2000))
    rows = []
    sizes = [int(v) for v in sys.argv[2].split(",")] if len(sys.argv) > 2 else (10, 16, 24, 32, 48, 64, 96, 128, 192, 256, 384, 512, 768, 1024, 1360)
    for n in sizes:
        imgs = nimg if n <= 512 else max(8, nimg // 4)
        b = ot_batch(imgs, seed=n, dense=(n, n), sigma=0.1)
        db = bench.DeviceBench(b, dev)
        steps = 20 if n <= 256 else 5
        total, per, launches = db.timed(steps, 3, lambda: None)
        ms = total / steps
        fl, ex, by = bench.algorithmic_work(b, db.nits.cpu().numpy())
        rows.append(dict(n=n, images=imgs, ms_per_step=ms, us_per_image=ms * 1e3 / imgs, tflops=fl / (ms * 1e-3) / 1e12,
                         frac=fl / (ms * 1e-3) / 1e12 / peak, kernel=bench.kernel_name(n, n, launches // steps)))
        print(f"N=M={n:5d} images={imgs:3d} {ms:9.4f} ms/step {ms * 1e3 / imgs:10.2f} us/img {rows[-1]['tflops']:7.2f} TFLOP/s "
              f"frac {rows[-1]['frac']:.3f}  {rows[-1]['kernel']}", flush=True)
        del db
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    with open(os.path.join(ROOT, "gpurun_out", "size_sweep.json"), "w") as fh:
        json.dump(dict(fp32_peak_tflops=peak, rows=rows), fh, indent=1)


if __name__ == "__main__":
    main()

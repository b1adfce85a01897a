#!/usr/bin/env python
"""globaltimer timeline of problem 0 inside the streaming kernel (profiling aid; GPU box).

    python tools/stream_timeline.py [n] [nimg]
"""
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
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 128
    nimg = int(sys.argv[2]) if len(sys.argv) > 2 else 64
    dev = torch.device("cuda:0")
    torch.cuda.set_device(dev)
    L = _lib.lib()
    b = ot_batch(nimg, seed=n, dense=(n, n + 4 if n == 1360 else n), sigma=0.1)
    db = bench.DeviceBench(b, dev)
    for _ in range(3):
        db.restore_and_flush(); db.step()
    buf = torch.zeros(1 << 16, dtype=torch.int64, device=dev)
    L.kdot_debug_set_clock_buffer(buf.data_ptr())
    db.restore_and_flush(); db.step(); torch.cuda.synchronize()
    L.kdot_debug_set_clock_buffer(None)
    t = buf.cpu().numpy()
    t0 = t[0]
    hdr = (t[:5] - t0) / 1e3
    print("float64 sub-tiles: gradient-round screened %d -> evaluated %d; potential rounds evaluated %d" % (t[5], t[7], t[6]))
    print(f"N=M={n} images={nimg}: phase0a end {hdr[1]:.1f} us, phase0b end {hdr[2]:.1f}, rounds end {hdr[3]:.1f}, after grid.sync {hdr[4]:.1f}")
    rounds = int(db.nits.cpu().numpy().max()) + 2
    R = 2
    upp = 2 * ((n + 32 * R - 1) // (32 * R)) + 2 * ((max(b["pos_per_img_t"]) + 32 * R - 1) // (32 * R))
    st = t[8:8 + rounds * upp * 4].reshape(rounds, upp, 4).astype(np.float64)
    st = (st - t0) / 1e3
    print("round: ticket(min..max) waitdone(min..max) computed(min..max) published(max) | compute us (mean)")
    for r in range(rounds):
        s = st[r]
        s = s[s[:, 2] > 0]
        if len(s) == 0:
            continue
        print(f"{r:3d}: {s[:,0].min():8.1f}..{s[:,0].max():8.1f}  {s[:,1].min():8.1f}..{s[:,1].max():8.1f}  {s[:,2].min():8.1f}..{s[:,2].max():8.1f}  "
              f"{s[:,3].max():8.1f} | {np.mean(s[:,2]-s[:,1]):6.2f}  fence+atomic {np.mean(s[:,3]-s[:,2]):5.2f}")
    if os.environ.get("KDOT_TIMELINE_UNITS"):
        for r in [int(v) for v in os.environ["KDOT_TIMELINE_UNITS"].split(",")]:
            d = st[r][:, 2] - st[r][:, 1]
            print(f"round {r} unit compute us:", " ".join(f"{v:.0f}" for v in d))


if __name__ == "__main__":
    main()

"""Distance of the fused kernels from the fp64 oracle, per kernel family and cloud spread (GPU box).

    python tools/accuracy_report.py [--out profiles/rNN_accuracy.json]

Prints max-norm relative errors of loss_per_img, d/dx and d/dalpha against oracle/sinkhorn_analytic.py
(test infrastructure) for the shapes the parity tests use; the north-star bar is 1e-4.
"""
import argparse
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from kd_6d_pose_adlp_b200.ops import OTConfig, ot_loss_batched  # noqa: E402
from kd_6d_pose_adlp_b200.synthetic import cu_seqlens, ot_batch  # noqa: E402
from oracle import sinkhorn_analytic  # noqa: E402  (checker only)


def rel(a, b):
    den = np.abs(b).max()
    return float(np.abs(a - b).max() / den) if den > 0 else float(np.abs(a - b).max())


CASES = [
    ("small s=0.005", dict(nimg=16, seed=3, sigma=0.005)),
    ("small s=0.05", dict(nimg=16, seed=3, sigma=0.05)),
    ("small s=0.15", dict(nimg=16, seed=3, sigma=0.15)),
    ("small s=0.3", dict(nimg=16, seed=4, sigma=0.3)),
    ("tiled s=0.05", dict(nimg=4, seed=1, n_range=(70, 90), m_range=(70, 90), p_empty_teacher=0.0)),
    ("tiled s=0.15", dict(nimg=4, seed=1, n_range=(70, 90), m_range=(70, 90), p_empty_teacher=0.0, sigma=0.15)),
    ("tiled s=0.005", dict(nimg=4, seed=1, n_range=(70, 90), m_range=(70, 90), p_empty_teacher=0.0, sigma=0.005)),
    ("stream 300 s=0.05", dict(nimg=2, seed=2, n_range=(300, 320), m_range=(280, 300), p_empty_teacher=0.0)),
    ("stream 300 s=0.15", dict(nimg=2, seed=2, n_range=(300, 320), m_range=(280, 300), p_empty_teacher=0.0, sigma=0.15)),
    ("stream dense 1360", dict(nimg=1, seed=5, dense=(1360, 1364), sigma=0.1)),
]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default=None)
    ap.add_argument("--scaling", type=float, default=0.5)
    ap.add_argument("--cases", default="", help="only the cases whose name contains this substring")
    args = ap.parse_args()
    dev = torch.device("cuda:0")
    rows = []
    for name, kw in CASES:
        if args.cases not in name:
            continue
        b = ot_batch(**kw)
        xs, xt = torch.from_numpy(b["xs"]).to(dev), torch.from_numpy(b["xt"]).to(dev)
        ws, wt = torch.from_numpy(b["ws"]).to(dev), torch.from_numpy(b["wt"]).to(dev)
        out = ot_loss_batched(xs, ws, xt, wt, b["pos_per_img"], b["pos_per_img_t"], OTConfig(scaling=args.scaling))
        torch.cuda.synchronize()
        ref = sinkhorn_analytic.kdot_fwd_bwd_f64(b["xs"], b["ws"], b["xt"], b["wt"], cu_seqlens(b["pos_per_img"]),
                                                 cu_seqlens(b["pos_per_img_t"]), 8, 2, scaling=args.scaling)
        row = dict(case=name, loss=rel(out["loss_per_img"].cpu().numpy(), ref["loss_per_img"]),
                   grad_xs=rel(out["grad_xs"].cpu().numpy(), ref["grad_xs"]),
                   grad_ws=rel(out["grad_ws"].cpu().numpy(), ref["grad_ws"]),
                   nits_equal=bool(np.array_equal(out["nits"].cpu().numpy(), ref["nits"])))
        rows.append(row)
        print("%-20s loss %.2e  d/dx %.2e  d/dalpha %.2e  nits_equal %s" % (name, row["loss"], row["grad_xs"], row["grad_ws"], row["nits_equal"]), flush=True)
    if args.out:
        with open(args.out, "w") as fh:
            json.dump(rows, fh, indent=1)


if __name__ == "__main__":
    main()

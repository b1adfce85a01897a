#!/usr/bin/env python
"""Why the streaming kernel folds its fp32 row sums into a compensated total every 32 columns (``urow_fold``,
``csrc/kdot_stream.cu``).

A numpy restatement of the float64 solver (``oracle/sinkhorn_analytic.py``, one slot of one image) in which ONLY the row
sums of the soft-min are degraded:

* ``fp32seq``   two running fp32 accumulators over the whole row (what the kernel did before),
* ``subtile``   fp32 accumulators over 32 columns, folded into a float64 (the kernel: a Kahan pair in fp32),

everything else -- pair arguments, exponentials, potentials -- exact.  Prints the error of d/dx and of the final offsets
against the all-float64 run.  Needs no GPU:

    python tools/rowsum_study.py [points]          # default: the dense 1360 x 1364 problem, ~1 min
"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from kd_6d_pose_adlp_b200.synthetic import ot_batch  # noqa: E402
from oracle import sinkhorn_analytic as sa  # noqa: E402  (checker-side tool)


def lse_rows(A, mode):
    m = A.max(axis=1)
    if mode == "exact":
        return m + np.log(np.exp(A - m[:, None]).sum(axis=1))
    E = np.exp(A - m[:, None]).astype(np.float32)   # fp32 exponentials; their own 1e-7 error is random and averages out
    n = E.shape[1]

    def two_partials(blk):
        s0 = np.zeros(blk.shape[0], np.float32)
        s1 = np.zeros(blk.shape[0], np.float32)
        for j in range(0, blk.shape[1] - 1, 2):
            s0 += blk[:, j]
            s1 += blk[:, j + 1]
        if blk.shape[1] % 2:
            s0 += blk[:, -1]
        return (s0 + s1).astype(np.float64)

    if mode == "fp32seq":
        tot = two_partials(E)
    elif mode == "subtile":
        tot = np.zeros(E.shape[0], np.float64)
        for c in range(0, n, 32):
            tot += two_partials(E[:, c:c + 32])
    else:
        raise ValueError(mode)
    return m + np.log(tot)


def solve(x, y, a, b, eps_s, rho, mode):
    cost = lambda u, v: 0.5 * ((u[:, None, :] - v[None, :, :]) ** 2).sum(-1)
    C_xx, C_yy, C_xy = cost(x, x), cost(y, y), cost(x, y)
    C_yx = C_xy.T
    la, lb = np.log(a), np.log(b)
    lam_of = lambda e: 1.0 / (1.0 + e / rho)
    softmin = lambda e, C, h: -e * lse_rows(h[None, :] - C / e, mode)
    e = eps_s[0]
    lam = lam_of(e)
    a_x, b_y, a_y, b_x = (lam * softmin(e, C_xx, la), lam * softmin(e, C_yy, lb), lam * softmin(e, C_yx, la),
                          lam * softmin(e, C_xy, lb))
    for e in eps_s:
        lam = lam_of(e)
        at_x, bt_y = lam * softmin(e, C_xx, la + a_x / e), lam * softmin(e, C_yy, lb + b_y / e)
        at_y, bt_x = lam * softmin(e, C_yx, la + b_x / e), lam * softmin(e, C_xy, lb + a_y / e)
        a_x, b_y, a_y, b_x = 0.5 * (a_x + at_x), 0.5 * (b_y + bt_y), 0.5 * (a_y + at_y), 0.5 * (b_x + bt_x)
    h_xx, h_xy = la + a_x / e, lb + a_y / e
    sm = lambda A: (lambda E: E / E.sum(1, keepdims=True))(np.exp(A - A.max(1, keepdims=True)))
    W_xx, W_xy = sm(h_xx[None, :] - C_xx / e), sm(h_xy[None, :] - C_xy / e)
    a_xf, b_xf = lam * softmin(e, C_xx, h_xx), lam * softmin(e, C_xy, h_xy)
    k = rho + e / 2
    ea, eb = np.exp(-a_xf / rho), np.exp(-b_xf / rho)
    g = (a * k / rho * lam)[:, None] * (eb[:, None] * (x - W_xy @ y) - ea[:, None] * (x - W_xx @ x))
    return g, h_xy


def study(n=1360, m=None, seed=5, sigma=0.1, blur=0.001, reach=0.5, scaling=0.5):
    m = n + 4 if m is None else m
    b = ot_batch(nimg=1, seed=seed, dense=(n, m), sigma=sigma)
    sc = np.array([640.0, 480.0], np.float32)
    xs, xt = (b["xs"] / sc).astype(np.float32), (b["xt"] / sc).astype(np.float32)
    x, y = xs[:, 0].astype(np.float64), xt[:, 0].astype(np.float64)
    a, bb = b["ws"][:, 0].astype(np.float64), b["wt"][:, 0].astype(np.float64)
    diam = sa.diameter_f32(xs[:, :1].transpose(1, 0, 2), xt[:, :1].transpose(1, 0, 2), np.float32)
    eps_s = sa.eps_schedule(diam, 2, blur, scaling)
    g_ref, h_ref = solve(x, y, a, bb, eps_s, reach ** 2, "exact")
    out = {}
    for mode in ("fp32seq", "subtile"):
        g, h = solve(x, y, a, bb, eps_s, reach ** 2, mode)
        out[mode] = (float(np.abs(g - g_ref).max() / np.abs(g_ref).max()), float(np.abs(h - h_ref).max()))
    return out


if __name__ == "__main__":
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 1360
    for mode, (eg, eh) in study(n).items():
        print("%-8s d/dx rel err %.2e   max |dh| of the final offsets %.2e" % (mode, eg, eh))

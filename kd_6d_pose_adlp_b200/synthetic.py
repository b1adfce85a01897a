"""Synthetic inputs at the OT boundary and at the head-output boundary.

There is no network for LINEMOD or checkpoints, so every benchmark / parity input is synthetic and
LINEMOD-ape shaped: 8 keypoint slots per cell, 2-D keypoints in full-image pixels (640x480),
~10 student cells (``POSITIVE_NUM = 10``, reference ``arguments/argument.py:81``) and ~9-11 teacher
cells per image, segmentation-probability masses shared by the 8 slots of a cell
(reference ``losses/kd_loss.py:82-83``).  Pure numpy: usable with or without a GPU.
"""
from __future__ import annotations

import numpy as np

IMG_W, IMG_H = 640.0, 480.0


def ot_batch(nimg, seed=0, n_range=(8, 12), m_range=(8, 12), sigma=0.05, p_empty_teacher=0.05,
             B=8, D=2, dense=None, in_pixels=True, teacher_shift=0.01):
    """OT-boundary batch in the C-ABI's flat cell-major layout.

    Returns dict(xs (sumN,B,D) f32, ws (sumN,B) f32, xt (sumM,B,D) f32, wt (sumM,B) f32,
    pos_per_img list[int], pos_per_img_t list[int]).  ``dense=(N, M)`` fixes the cell counts
    (e.g. ``(1360, 1364)``: every cell of the darknet_tiny student / darknet53 teacher grids).
    For ``D == 2`` coordinates are full-image pixels (so the (w, h) normalisation is exercised);
    for other D they are code probabilities in (0, 1) (ZebraPose-style config).
    """
    rng = np.random.default_rng(seed)
    ns, ms = [], []
    for _ in range(nimg):
        if dense is not None:
            n, m = dense
        else:
            n = int(rng.integers(n_range[0], n_range[1] + 1))
            m = int(rng.integers(m_range[0], m_range[1] + 1))
            if rng.random() < p_empty_teacher:
                m = 0
        ns.append(n)
        ms.append(m)
    sn, sm = sum(ns), sum(ms)
    xs = np.empty((sn, B, D), np.float32)
    xt = np.empty((sm, B, D), np.float32)
    ws = np.empty((sn, B), np.float32)
    wt = np.empty((sm, B), np.float32)
    o_s = o_t = 0
    for n, m in zip(ns, ms):
        if D == 2:
            centre = rng.uniform(0.3, 0.7, size=(1, B, D))
            s = centre + sigma * rng.standard_normal((n, B, D))
            t = centre + teacher_shift * rng.standard_normal((1, B, D)) + sigma * rng.standard_normal((m, B, D))
            if in_pixels:
                s = s * np.array([IMG_W, IMG_H])
                t = t * np.array([IMG_W, IMG_H])
        else:
            s = 1.0 / (1.0 + np.exp(-rng.standard_normal((n, B, D))))
            t = 1.0 / (1.0 + np.exp(-rng.standard_normal((m, B, D))))
        xs[o_s:o_s + n] = s
        xt[o_t:o_t + m] = t
        ws[o_s:o_s + n] = np.clip(rng.uniform(0.05, 0.95, size=(n, 1)), 1e-3, 1 - 1e-3)
        wt[o_t:o_t + m] = rng.uniform(0.1, 0.95, size=(m, 1))
        o_s += n
        o_t += m
    return dict(xs=xs, ws=ws, xt=xt, wt=wt, pos_per_img=ns, pos_per_img_t=ms)


def cu_seqlens(counts):
    """Exclusive prefix sums (int32, len nimg+1) of a per-image count list."""
    cu = np.zeros(len(counts) + 1, np.int32)
    np.cumsum(np.asarray(counts, np.int64), out=cu[1:])
    return cu

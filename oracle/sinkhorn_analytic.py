"""fp64 numpy restatement of the whole OT section with the ANALYTIC backward.

TEST INFRASTRUCTURE ONLY (see ``oracle/__init__.py``); **parity unpinned** for the geomloss part.

This is the second, independent oracle: where ``oracle/geomloss_ref.py`` follows geomloss op by
op and lets autograd differentiate it, this file writes down the closed-form gradients the CUDA
kernel implements ("Resulting analytic backward") and works directly on
the flat cell-major layout of the C-ABI (``include/kdot.h``):

    xs [sumN][B][D]  ws [sumN][B]   student cells   (reference ``losses/kd_loss.py:50,83``)
    xt [sumM][B][D]  wt [sumM][B]   teacher cells   (reference ``losses/kd_loss.py:79,82``)
    cu_n / cu_m      exclusive prefix sums of pos_per_img / pos_per_img_t

It covers what ``losses/loss_libs.py:1-51`` + ``SamplesLoss("sinkhorn", p=2)`` compute per image:
in-place normalisation by (w, h), skip of empty images, the eps-scaling loop, the debiased
(un)balanced cost summed over the B keypoint slots, and d/d(xs), d/d(ws) w.r.t. the
UN-normalised student inputs.  Inputs are taken as float32 and normalised in float32 (exactly the
bits the reference and the kernel see) and then promoted to float64.
"""
from __future__ import annotations

import numpy as np


def _lse(v, axis):
    m = np.max(v, axis=axis, keepdims=True)
    return (m + np.log(np.sum(np.exp(v - m), axis=axis, keepdims=True))).squeeze(axis)


def _softmax(v, axis):
    m = np.max(v, axis=axis, keepdims=True)
    e = np.exp(v - m)
    return e / np.sum(e, axis=axis, keepdims=True)


def _cost(u, v):
    """0.5*|u_i - v_j|^2, (B,N,D),(B,M,D) -> (B,N,M); direct differences (exact in fp64)."""
    d = u[:, :, None, :] - v[:, None, :, :]
    return 0.5 * np.sum(d * d, axis=-1)


def eps_schedule(diam, p, blur, scaling):
    """geomloss ``epsilon_schedule`` in float64; len == 2 + ceil(ln(blur/diam)/ln(scaling))."""
    mid = [float(np.exp(e)) for e in np.arange(p * np.log(diam), p * np.log(blur), p * np.log(scaling))]
    return [diam ** p] + mid + [blur ** p]


def diameter_f32(x, y, dtype=np.float32):
    """geomloss ``max_diameter`` on clouds (B,N,D)/(B,M,D): bbox diagonal in ``dtype`` -> float.

    The reference evaluates it on its fp32 tensors (``(maxs - mins).norm().item()``), hence fp32 here."""
    pts = np.concatenate([x.reshape(-1, x.shape[-1]), y.reshape(-1, y.shape[-1])], 0).astype(dtype)
    ext = (pts.max(0) - pts.min(0)).astype(dtype)
    return float(np.sqrt(np.sum(ext * ext, dtype=dtype), dtype=dtype))


def sinkhorn_image_f64(a, x, b, y, blur, reach, scaling, diam, p=2):
    """One image: a (B,N), x (B,N,D), b (B,M), y (B,M,D) float64.

    Returns (F (B,), dF/dx (B,N,D), dF/da (B,N), nits).
    """
    assert p == 2
    eps_s = eps_schedule(diam, p, blur, scaling)
    rho = None if reach is None else reach ** p
    lam_of = (lambda e: 1.0) if rho is None else (lambda e: 1.0 / (1.0 + e / rho))
    with np.errstate(divide="ignore"):
        la = np.where(a > 0, np.log(np.where(a > 0, a, 1.0)), -100000.0)
        lb = np.where(b > 0, np.log(np.where(b > 0, b, 1.0)), -100000.0)
    C_xx, C_yy, C_xy = _cost(x, x), _cost(y, y), _cost(x, y)
    C_yx = np.swapaxes(C_xy, 1, 2)

    def softmin(e, C, h):  # -e * LSE_j(h_j - C_ij/e)
        return -e * _lse(h[:, None, :] - C / e, axis=2)

    e = eps_s[0]
    lam = lam_of(e)
    a_x = lam * softmin(e, C_xx, la)
    b_y = lam * softmin(e, C_yy, lb)
    a_y = lam * softmin(e, C_yx, la)
    b_x = lam * softmin(e, C_xy, lb)
    for e in eps_s:
        lam = lam_of(e)
        at_x = lam * softmin(e, C_xx, la + a_x / e)
        bt_y = lam * softmin(e, C_yy, lb + b_y / e)
        at_y = lam * softmin(e, C_yx, la + b_x / e)
        bt_x = lam * softmin(e, C_xy, lb + a_y / e)
        a_x, b_y, a_y, b_x = 0.5 * (a_x + at_x), 0.5 * (b_y + bt_y), 0.5 * (a_y + at_y), 0.5 * (b_x + bt_x)
    # last extrapolation, all four from the same old potentials
    h_xx, h_yy, h_yx, h_xy = la + a_x / e, lb + b_y / e, la + b_x / e, lb + a_y / e
    W_xx = _softmax(h_xx[:, None, :] - C_xx / e, axis=2)
    W_xy = _softmax(h_xy[:, None, :] - C_xy / e, axis=2)
    a_x, b_y, a_y, b_x = (lam * softmin(e, C_xx, h_xx), lam * softmin(e, C_yy, h_yy),
                          lam * softmin(e, C_yx, h_yx), lam * softmin(e, C_xy, h_xy))
    bary_xy = x - np.einsum("bnm,bmd->bnd", W_xy, y)  # x_i - sum_j W_ij y_j
    bary_xx = x - np.einsum("bnk,bkd->bnd", W_xx, x)
    if rho is None:
        F = np.sum(a * (b_x - a_x), 1) + np.sum(b * (a_y - b_y), 1)
        g_a = b_x - a_x
        g_x = a[..., None] * (bary_xy - bary_xx)
    else:
        k = rho + e / 2
        ea, eb = np.exp(-a_x / rho), np.exp(-b_x / rho)
        F = np.sum(a * k * (ea - eb), 1) + np.sum(b * k * (np.exp(-b_y / rho) - np.exp(-a_y / rho)), 1)
        g_a = k * (ea - eb)
        g_x = (a * k / rho * lam)[..., None] * (eb[..., None] * bary_xy - ea[..., None] * bary_xx)
    return F, g_x, g_a, len(eps_s)


def kdot_fwd_bwd_f64(xs, ws, xt, wt, cu_n, cu_m, B, D, blur=0.001, reach=0.5, scaling=0.5,
                     w=640.0, h=480.0, normalize=True, p=2, diam_dtype=np.float32):
    """Flat-layout oracle with the C-ABI's semantics (``kdot_sinkhorn_fwd_bwd``).

    ``ws``/``wt`` None -> uniform 1/N, 1/M masses (``losses/loss_libs.py:49``).
    Returns dict: loss_per_img (nimg,) f64 [sum over B slots; 0 if skipped], valid (nimg,) i32,
    nits (nimg,) i32, grad_xs (sumN,B,D) f64 w.r.t. the un-normalised xs, grad_ws (sumN,B) f64,
    xs_norm / xt_norm float32 (the in-place side effect of ``loss_libs.py:8-12``).
    """
    xs = np.asarray(xs, np.float32).reshape(-1, B, D).copy()
    xt = np.asarray(xt, np.float32).reshape(-1, B, D).copy()
    scale = np.ones(D, np.float32)
    if normalize:
        assert D == 2, "reference normalises only the 2-D case (loss_libs.py:7)"
        scale = np.array([w, h], np.float32)
        xs = (xs / scale).astype(np.float32)
        xt = (xt / scale).astype(np.float32)
    nimg = len(cu_n) - 1
    out = dict(loss_per_img=np.zeros(nimg), valid=np.zeros(nimg, np.int32), nits=np.zeros(nimg, np.int32),
               grad_xs=np.zeros(xs.shape), grad_ws=np.zeros(xs.shape[:2]), xs_norm=xs, xt_norm=xt)
    for i in range(nimg):
        n0, n1, m0, m1 = cu_n[i], cu_n[i + 1], cu_m[i], cu_m[i + 1]
        if n1 == n0 or m1 == m0:
            continue
        x32 = np.ascontiguousarray(xs[n0:n1].transpose(1, 0, 2))
        y32 = np.ascontiguousarray(xt[m0:m1].transpose(1, 0, 2))
        N, M = n1 - n0, m1 - m0
        a = np.full((B, N), 1.0 / N) if ws is None else np.asarray(ws, np.float32)[n0:n1].T.astype(np.float64)
        b = np.full((B, M), 1.0 / M) if wt is None else np.asarray(wt, np.float32)[m0:m1].T.astype(np.float64)
        diam = diameter_f32(x32, y32, diam_dtype)
        F, g_x, g_a, nits = sinkhorn_image_f64(a, x32.astype(np.float64), b, y32.astype(np.float64),
                                               blur, reach, scaling, diam, p)
        out["loss_per_img"][i] = F.sum()
        out["valid"][i] = 1
        out["nits"][i] = nits
        out["grad_xs"][n0:n1] = g_x.transpose(1, 0, 2) / scale.astype(np.float64)
        out["grad_ws"][n0:n1] = g_a.T
    return out

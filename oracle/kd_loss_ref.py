"""CPU restatement (torch) of the reference's per-image OT driver and its caller's OT section.

TEST INFRASTRUCTURE ONLY (see ``oracle/__init__.py``).  Follows

* ``/root/reference/losses/loss_libs.py:1-51``  (``kd_loss_2d``: in-place normalise, per-image split,
  ``(N,8,.) -> (8,N,.)`` transposes, ``SamplesLoss(...).sum()`` per non-empty image), and
* ``/root/reference/losses/kd_loss.py:73-103``  (masses, optional detach, mean over non-empty images).

It is validated against the reference's own ``kd_loss_2d`` in ``tests/test_oracle_vs_reference.py``
(run in the authoring container, where ``/root/reference`` is mounted) and is what ``bench.py`` times
as the CPU baseline (``cpu_baseline.kind == "port"``): the reference's formulation, issued as stock
torch ops on the host cores.
"""
from __future__ import annotations

import torch

from .geomloss_ref import SamplesLoss


def kd_loss_2d_ref(pred_xy, target_xy, pred_cls, target_cls, w, h, level, kd_loss, dim,
                   pos_per_img=None, pos_per_img_t=None, normalize=True):
    """Same contract as ``loss_libs.py:1``: mutates ``pred_xy``/``target_xy`` in place, returns a list
    with one 0-dim loss per image that has both student and teacher cells."""
    if dim == 2 and normalize:
        for t in (pred_xy, target_xy):
            t[:, 0] = t[:, 0] / w
            t[:, 1] = t[:, 1] / h
    if level != "point":
        raise NotImplementedError(level)
    xs = pred_xy.view(-1, 8, dim)
    xt = target_xy.view(-1, 8, dim)
    out = []
    s0 = t0 = 0
    for n_s, n_t in zip(pos_per_img, pos_per_img_t):
        s1, t1 = s0 + n_s, t0 + n_t
        if n_s > 0 and n_t > 0:
            x_s = xs[s0:s1].transpose(0, 1).contiguous()
            x_t = xt[t0:t1].transpose(0, 1).contiguous()
            if target_cls is not None:
                a_s = pred_cls[s0:s1].transpose(0, 1).contiguous()
                a_t = target_cls[t0:t1].transpose(0, 1).contiguous()
                out.append(kd_loss(a_s, x_s, a_t, x_t).sum())
            else:
                out.append(kd_loss(x_s, x_t).sum())
        s0, t0 = s1, t1
    return out


def kd_ot_section_ref(pred_xy, s_cls, pred_t, pos_per_img, cfg_kd, w=640, h=480, kd_loss=None):
    """``kd_loss.py:73-103`` given decoded student keypoints ``pred_xy (sumN*8, 2)`` (full-image px),
    student masses ``s_cls (sumN, 8)`` (or None when not weighted) and the teacher dict."""
    if kd_loss is None:
        kd_loss = SamplesLoss(cfg_kd["GTYPE"], p=cfg_kd["GP"], blur=cfg_kd["GBLUR"],
                              scaling=cfg_kd["SCALING"], reach=cfg_kd["REACH"])
    pred_t_xy = pred_t["post_kp_2d"].view(-1, 2)
    if cfg_kd["WEIGHTED_OT"]:
        t_cls = pred_t["post_kp_cls"].pow(2)
        if cfg_kd["DETACH"]:
            s_cls = s_cls.detach()
    else:
        t_cls = s_cls = None
    losses = kd_loss_2d_ref(pred_xy, pred_t_xy, s_cls, t_cls, w, h, cfg_kd["GLEVEL"], kd_loss, dim=2,
                            pos_per_img=pos_per_img, pos_per_img_t=pred_t["post_pos_per_img"])
    if len(losses) > 0:
        return (sum(losses) / len(losses)).sum()
    return torch.tensor(0.0)

"""``kd_loss_2d`` -- drop-in for ``/root/reference/losses/loss_libs.py:1`` (seam B1).

Same signature, same side effects (``pred_xy`` and ``target_xy`` are normalised IN PLACE, images without
student or teacher cells are skipped, one 0-dim loss per remaining image is returned) -- but the Python
loop over images with its per-image ``SamplesLoss`` call (``loss_libs.py:22-50``) is replaced by ONE fused
kernel launch for the whole mini-batch.
"""
from __future__ import annotations

import os

from ..ops import OTLossFunction, normalize_in_place
from ..samples_loss import SamplesLoss


def kd_loss_2d(pred_xy, target_xy, pred_cls, target_cls, w, h, level, kd_loss, dim, pos_per_img=None,
               pos_per_img_t=None, normalize=True):
    """
    pred_xy: (npos*8, dim) student keypoints, target_xy: (npos_t*8, dim) teacher keypoints,
    pred_cls / target_cls: (npos, 8) / (npos_t, 8) masses or None (uniform), kd_loss: :class:`SamplesLoss`.
    """
    if not isinstance(kd_loss, SamplesLoss):
        raise TypeError("kd_loss must be kd_6d_pose_adlp_b200.SamplesLoss (no fallback to foreign solvers)")
    if level != "point":
        raise NotImplementedError(f"kd_loss_2d(level={level!r}): the reference only defines 'point'")
    if pos_per_img is None or pos_per_img_t is None:
        raise ValueError("pos_per_img and pos_per_img_t are required")
    if (pred_cls is None) != (target_cls is None):
        raise ValueError("pred_cls and target_cls must both be given or both be None")
    if dim == 2 and normalize:
        pred_xy = normalize_in_place(pred_xy, w, h)
        normalize_in_place(target_xy, w, h)
    pos_per_img = [int(v) for v in pos_per_img]
    pos_per_img_t = [int(v) for v in pos_per_img_t]
    loss_per_img, valid, nits = OTLossFunction.apply(
        pred_xy.view(-1, 8, dim), pred_cls, target_xy.view(-1, 8, dim), target_cls,
        pos_per_img, pos_per_img_t, kd_loss.config, float(w), float(h), False)
    # Per-image kernel status (include/kdot.h KDOT_IMG_*): a degenerate image (all key-points coincide) or an
    # epsilon schedule longer than KDOT_MAX_ROUNDS yields NaN loss and gradients for that image where geomloss would
    # raise.  The status stays on the device (no host sync here); it is left on the solver object, and
    # KDOT_SYNC_STATUS=1 turns it into an exception right away.
    kd_loss.last_valid, kd_loss.last_nits = valid, nits
    if os.environ.get("KDOT_SYNC_STATUS", "0") == "1":
        from .kd_loss import _status_error

        _status_error(valid.cpu().tolist(), "kd_loss_2d")
    keep = [i for i, (n, m) in enumerate(zip(pos_per_img, pos_per_img_t)) if n > 0 and m > 0]
    return [loss_per_img[i] for i in keep]

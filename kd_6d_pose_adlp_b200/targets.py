"""Device-side SSC target assignment -- what ``PoseLossDzi.prepare_targets`` (reference ``losses/loss.py:164-268``,
``POSITIVE_TYPE == 'SSC'``) produces for the loss, from three launches over the whole mini-batch
(``kdot_ssc_count`` / ``kdot_ssc_pick`` / ``kdot_ssc_assign``, ``csrc/kdot_targets.cu``) instead of a Python walk over
images x levels x objects with a ``nonzero`` + ``randperm`` + host sync each.

Two draw modes:

* ``"parity"``  the random draw of ``loss.py:227`` is REPLAYED: ``count`` comes back to the host in one copy, the CPU
  generator is advanced with exactly the reference's calls (``torch.randperm(len(valid_pos))`` for every image, level and
  object, in that order), the first ``min(nk, count)`` entries go to the device.  With the same ``torch.manual_seed`` the
  labels are bit-identical to a reference run (``tests/test_targets_gpu.py`` against ``tests/golden/kd_pose_loss.npz``).
* ``"philox"``  the draw happens on the device from a counter-based generator (no host round trip at all): uniform
  without replacement like ``randperm(n)[:k]``, but a different stream -- same distribution, different cells
  (the test checks budgets, mask membership, distinctness and uniformity).

``targets`` are the reference's ``PoseAnnot`` objects (``libs/poses.py``) or anything with the attributes read below.
"""
from __future__ import annotations

import ctypes as C
from typing import List, Sequence

import numpy as np
import torch

from . import _lib

MAX_GT = 8


def _stack_targets(targets, dev):
    """Per-image object lists padded to ``maxgt`` (<= 8) and stacked -- a dozen launches for the whole mini-batch (one
    ``cat`` / ``stack`` per field and one scatter into the padded layout), independent of the number of images."""
    nimg = len(targets)
    ngt = [int(t.class_ids.shape[0]) for t in targets]
    maxgt = max(1, max(ngt))
    if maxgt > MAX_GT:
        raise ValueError(f"kdot_ssc_count handles up to {MAX_GT} objects per image, got {maxgt}")
    f32 = dict(dtype=torch.float32, device=dev)
    with_obj = [t for t, g in zip(targets, ngt) if g > 0]
    rot = torch.zeros(nimg, maxgt, 3, 3, **f32)
    trans = torch.zeros(nimg, maxgt, 3, **f32)
    kp3d = torch.zeros(nimg, maxgt, 8, 3, **f32)
    cls1 = torch.zeros(nimg, maxgt, dtype=torch.int64, device=dev)
    if with_obj:
        ids = torch.cat([t.class_ids.reshape(-1) for t in with_obj]).to(dev).long()
        rflat = torch.cat([t.rotations.reshape(-1, 3, 3) for t in with_obj]).to(**f32)
        tflat = torch.cat([t.translations.reshape(-1, 3) for t in with_obj]).to(**f32)
        shared = all(t.keypoints_3d is with_obj[0].keypoints_3d for t in with_obj)
        if shared:
            kflat = with_obj[0].keypoints_3d.to(**f32)[ids]
        else:
            kflat = torch.cat([t.keypoints_3d.to(**f32)[t.class_ids.to(dev).long()] for t in with_obj])
        where = np.asarray([(i, g) for i, n in enumerate(ngt) for g in range(n)], np.int64)
        idx = torch.from_numpy(where).to(dev, non_blocking=True)
        ii, gg = idx[:, 0], idx[:, 1]
        rot[ii, gg], trans[ii, gg], kp3d[ii, gg], cls1[ii, gg] = rflat, tflat, kflat, ids + 1
    mask = torch.stack([t.mask for t in targets]).to(**f32)
    same_k = all(t.K is targets[0].K for t in targets)
    K = (targets[0].K.to(**f32).view(1, 3, 3).expand(nimg, 3, 3).contiguous() if same_k
         else torch.stack([t.K.reshape(3, 3) for t in targets]).to(**f32))
    has_bt = [getattr(t, "bbox_trans", None) is not None for t in targets]
    if any(has_bt) and not all(has_bt):
        raise ValueError("either every target carries bbox_trans or none does")
    bt = torch.stack([t.bbox_trans.reshape(2, 3) for t in targets]).to(**f32) if all(has_bt) else None
    return dict(nimg=nimg, ngt=ngt, maxgt=maxgt, rot=rot, trans=trans, kp3d=kp3d, cls1=cls1, mask=mask.contiguous(), K=K, bt=bt,
                num_gt=torch.tensor(ngt, dtype=torch.int32).to(dev, non_blocking=True))


def ssc_assign(targets, anchors_one_image: torch.Tensor, level_hw: Sequence[int], anchor_sizes: Sequence[float],
               positive_num: int, positive_lambda: float, mode: str = "parity", seed: int = 0):
    """Labels of every cell of the mini-batch.

    ``anchors_one_image (cells, 4)`` xyxy, level-major (identical for every image: ``cat_boxlist(anchors[0]).bbox``);
    ``level_hw[l] = H_l * W_l``.  Returns a dict of device tensors: ``labels (nimg * cells,) int64``,
    ``owner (nimg * cells,) int32`` (object index of a positive cell), ``npos (nimg,) int32`` plus ``count`` / ``nk``
    ``(nimg, nlvl, maxgt)``, ``span (nimg, maxgt)`` and the stacked per-image object tensors (``st``)."""
    if mode not in ("parity", "philox"):
        raise ValueError("mode must be 'parity' or 'philox'")
    L = _lib.lib()
    dev = anchors_one_image.device
    if dev.type != "cuda":
        raise ValueError("ssc_assign runs on a CUDA device (libkdot has no CPU fallback)")
    st = _stack_targets(targets, dev)
    nimg, maxgt, nlvl = st["nimg"], st["maxgt"], len(level_hw)
    cells = int(sum(level_hw))
    anchors = anchors_one_image.to(torch.float32).contiguous()
    if anchors.shape != (cells, 4):
        raise ValueError("anchors_one_image must be (sum(level_hw), 4)")
    cap = int(positive_num) + 1
    i32 = dict(dtype=torch.int32, device=dev)
    gtid = torch.empty(nimg, cells, dtype=torch.uint8, device=dev)
    cnt_nk = torch.empty(2, nimg, nlvl, maxgt, **i32)
    span = torch.empty(nimg, maxgt, dtype=torch.float32, device=dev)
    hw = (C.c_int32 * nlvl)(*[int(v) for v in level_hw])
    sz = (C.c_float * nlvl)(*[float(v) for v in anchor_sizes[:nlvl]])
    stream = torch.cuda.current_stream(dev).cuda_stream
    mh, mw = int(st["mask"].shape[1]), int(st["mask"].shape[2])
    with torch.cuda.device(dev):
        rc = L.kdot_ssc_count(st["mask"].data_ptr(), mh, mw, anchors.data_ptr(), hw, sz, nlvl, nimg, maxgt,
                              st["num_gt"].data_ptr(), st["rot"].data_ptr(), st["trans"].data_ptr(), st["kp3d"].data_ptr(),
                              st["K"].data_ptr(), None if st["bt"] is None else st["bt"].data_ptr(), int(positive_num),
                              float(positive_lambda), gtid.data_ptr(), cnt_nk[0].data_ptr(), cnt_nk[1].data_ptr(),
                              span.data_ptr(), stream)
        _lib.check(rc, "kdot_ssc_count")
        if mode == "philox":
            picks = torch.empty(nimg, nlvl, maxgt, cap, **i32)
            rc = L.kdot_ssc_pick(cnt_nk[0].data_ptr(), cnt_nk[1].data_ptr(), nimg, nlvl, maxgt, cap, int(seed) & (2 ** 64 - 1),
                                 picks.data_ptr(), stream)
            _lib.check(rc, "kdot_ssc_pick")
        else:
            host = cnt_nk.cpu().numpy()                      # ONE device->host copy (the reference syncs per level and object)
            cnt_h, nk_h = host[0], host[1]
            picks_h = np.full((nimg, nlvl, maxgt, cap), -1, np.int32)
            for i in range(nimg):                            # the reference's loop order (loss.py:216-230): level, then object
                for l in range(nlvl):
                    for g in range(st["ngt"][i]):
                        n = int(cnt_h[i, l, g])
                        k = min(int(nk_h[i, l, g]), n, cap)
                        perm = torch.randperm(n)             # the CPU generator advances exactly as in the reference
                        picks_h[i, l, g, :k] = perm[:k].numpy()
            picks = torch.from_numpy(picks_h).to(dev, non_blocking=True)
        labels = torch.empty(nimg * cells, dtype=torch.int64, device=dev)
        owner = torch.empty(nimg * cells, **i32)
        npos = torch.empty(nimg, **i32)
        rc = L.kdot_ssc_assign(gtid.data_ptr(), picks.data_ptr(), st["cls1"].data_ptr(), hw, nlvl, nimg, maxgt, cap,
                               labels.data_ptr(), owner.data_ptr(), npos.data_ptr(), stream)
        _lib.check(rc, "kdot_ssc_assign")
    return dict(labels=labels, owner=owner, npos=npos, count=cnt_nk[0], nk=cnt_nk[1], span=span, gtid=gtid, st=st,
                cells=cells)


def positives_aux(res, pos_inds: torch.Tensor):
    """What the loss needs at the positive cells (``loss.py:255-266`` restricted to ``pos_inds``): class label (0-based),
    the 3-D key-points in the camera frame ``R X + T`` ``(npos, 8, 3)`` and the crop affine ``(npos, 2, 3)``."""
    st, cells = res["st"], res["cells"]
    img = torch.div(pos_inds, cells, rounding_mode="floor")
    g = res["owner"][pos_inds].long()
    cls_label = st["cls1"][img, g] - 1
    R, T, X = st["rot"][img, g], st["trans"][img, g], st["kp3d"][img, g]
    aux_3d = torch.baddbmm(T.unsqueeze(1), X, R.transpose(1, 2))            # (R X^T + T)^T
    bt = None if st["bt"] is None else st["bt"][img]
    return cls_label, aux_3d, bt

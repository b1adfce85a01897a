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

Object counts: the reference's loop indexes ``bbox_trans.unsqueeze(0)`` -- a ``(1, 2, 3)`` tensor -- with the object index of
every cell (``loss.py:184,253``), so it runs for ONE object per (cropped) image, which is what its data pipeline produces and
what the golden fixtures hold.  The kernels and the packing below take up to 8 objects per image (one crop affine per image);
that part has no reference run to compare with and is covered on the host side only (``tests/test_targets_host.py``).
"""
from __future__ import annotations

import ctypes as C
from typing import List, Sequence

import numpy as np
import torch

from . import _lib

MAX_GT = 8


def _stack_targets(targets, dev):
    """Per-image object lists padded to ``maxgt`` (<= 8) and stacked -- a handful of launches for the whole mini-batch
    (one ``cat`` / ``stack`` per field, one gather into the padded layout; none at all for the padding when every image
    holds the same number of objects), independent of the number of images."""
    nimg = len(targets)
    ngt = [int(t.class_ids.shape[0]) for t in targets]
    maxgt = max(1, max(ngt))
    if maxgt > MAX_GT:
        raise ValueError(f"kdot_ssc_count handles up to {MAX_GT} objects per image, got {maxgt}")
    f32 = dict(dtype=torch.float32, device=dev)
    with_obj = [t for t, g in zip(targets, ngt) if g > 0]
    if with_obj:
        ids = torch.cat([t.class_ids for t in with_obj]).reshape(-1).to(dev).long()
        rflat = torch.cat([t.rotations for t in with_obj]).reshape(-1, 3, 3).to(**f32)
        tflat = torch.cat([t.translations for t in with_obj]).reshape(-1, 3).to(**f32)
        kp0 = with_obj[0].keypoints_3d
        if all(t.keypoints_3d is kp0 for t in with_obj):
            kflat = kp0.to(**f32)[ids]
        else:
            kflat = torch.cat([t.keypoints_3d.to(**f32)[t.class_ids.to(dev).long()] for t in with_obj])
        if min(ngt) == maxgt:      # no padding needed: the flat arrays ARE the padded layout
            rot, trans = rflat.view(nimg, maxgt, 3, 3), tflat.view(nimg, maxgt, 3)
            kp3d, cls1 = kflat.view(nimg, maxgt, 8, 3), (ids + 1).view(nimg, maxgt)
        else:                      # gather through an index with a sentinel row of zeros appended to every flat array
            total = int(ids.shape[0])
            where = np.full((nimg, maxgt), total, np.int64)
            o = 0
            for i, n in enumerate(ngt):
                where[i, :n] = np.arange(o, o + n)
                o += n
            idx = torch.from_numpy(where.reshape(-1)).to(dev, non_blocking=True)
            pad = lambda a: torch.cat([a, a.new_zeros((1,) + tuple(a.shape[1:]))])[idx]
            rot, trans = pad(rflat).view(nimg, maxgt, 3, 3), pad(tflat).view(nimg, maxgt, 3)
            kp3d, cls1 = pad(kflat).view(nimg, maxgt, 8, 3), pad(ids + 1).view(nimg, maxgt)
    else:
        rot = torch.zeros(nimg, maxgt, 3, 3, **f32)
        trans = torch.zeros(nimg, maxgt, 3, **f32)
        kp3d = torch.zeros(nimg, maxgt, 8, 3, **f32)
        cls1 = torch.zeros(nimg, maxgt, dtype=torch.int64, device=dev)
    mask = torch.stack([t.mask for t in targets]).to(**f32)
    same_k = all(t.K is targets[0].K for t in targets)
    K = (targets[0].K.to(**f32).view(1, 3, 3).expand(nimg, 3, 3).contiguous() if same_k
         else torch.stack([t.K for t in targets]).reshape(nimg, 3, 3).to(**f32))
    has_bt = [getattr(t, "bbox_trans", None) is not None for t in targets]
    if any(has_bt) and not all(has_bt):
        raise ValueError("either every target carries bbox_trans or none does")
    bt = torch.stack([t.bbox_trans for t in targets]).reshape(nimg, 2, 3).to(**f32) if all(has_bt) else None
    return dict(nimg=nimg, ngt=ngt, maxgt=maxgt, rot=rot.contiguous(), trans=trans.contiguous(), kp3d=kp3d.contiguous(),
                cls1=cls1.contiguous(), mask=mask.contiguous(), K=K, bt=bt,
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
    slot = img * st["maxgt"] + res["owner"][pos_inds]                        # row of the padded (image, object) tables
    cls_label = st["cls1"].view(-1)[slot] - 1
    # camera-frame key-points of every (image, object) once -- (R X^T + T)^T, one batched product -- then one gather
    cam = torch.baddbmm(st["trans"].view(-1, 1, 3), st["kp3d"].view(-1, 8, 3), st["rot"].view(-1, 3, 3).transpose(1, 2))
    aux_3d = cam[slot]
    bt = None if st["bt"] is None else st["bt"][img]
    return cls_label, aux_3d, bt

"""``PostProcessorKD`` -- drop-in for ``/root/reference/postprocess/postprocess_kd.py:12`` (teacher knowledge
extraction: segmentation-weighted cell voting).

Same constructor, same ``forward(box_cls, box_regression, targets, anchors) -> [scores, R, T, det2d]`` (per-image
lists; ``scores (n,8)`` = sqrt of the seg probability of the selected cells broadcast to the 8 key-points,
``det2d (n,8,2)`` their key-points un-cropped to full-image pixels; empty ``(0,8)`` / ``(0,8,2)`` tensors when
nothing is selected or PnP fails -- ``postprocess_kd.py:86-96``).

The reference walks images x levels x labels in Python with a host sync at every ``len()`` / ``min()`` /
``topk(k=tensor)``; here the whole selection (threshold, per-level arg-max, running-best box size, per-level
budget ``nk``, per-level top-k, decode) for every (image, class) is ONE kernel launch
(``kdot_select_cells``) followed by ONE device->host copy of the few selected cells.  RANSAC-EPnP stays
``cv2.solvePnPRansac`` on the host exactly as in the reference (``postprocess_kd.py:191``).
"""
from __future__ import annotations

import ctypes as C
import os
from concurrent.futures import ThreadPoolExecutor

import numpy as np
import torch
from torch import nn

from .. import _lib


def select_cells(box_cls, box_regression, anchor_sizes, anchor_strides, inference_th, positive_num,
                 positive_lambda):
    """Runs the selection kernel.  ``box_cls[l] (nimg, C, H, W)``, ``box_regression[l] (nimg, C*16, H, W)`` CUDA
    fp32.  Returns a dict of device tensors indexed by ``q = img * C + cls`` (see ``include/kdot.h``)."""
    L = _lib.lib()
    nlvl = len(box_cls)
    nimg, ncls = box_cls[0].shape[0], box_cls[0].shape[1]
    dev = box_cls[0].device
    cls_c, reg_c = [], []
    for c, r in zip(box_cls, box_regression):
        if not (c.is_cuda and r.is_cuda and c.dtype == torch.float32 and r.dtype == torch.float32):
            raise ValueError("head outputs must be float32 CUDA tensors (libkdot has no CPU fallback)")
        if r.shape[1] != ncls * 16 or c.shape[1] != ncls or c.shape[2:] != r.shape[2:]:
            raise ValueError("inconsistent head output shapes")
        cls_c.append(c.contiguous())
        reg_c.append(r.contiguous())
    nsizes = len(anchor_sizes)
    cap = int(positive_num) + nsizes + 1
    nq = nimg * ncls
    i32 = dict(dtype=torch.int32, device=dev)
    out = dict(
        sel_count=torch.empty(nq, **i32), sel_level=torch.empty(nq, cap, **i32), sel_loc=torch.empty(nq, cap, **i32),
        sel_score=torch.empty(nq, cap, dtype=torch.float32, device=dev),
        sel_kpts=torch.empty(nq, cap, 16, dtype=torch.float32, device=dev),
        nk=torch.empty(nq, nsizes, **i32), valid_cnt=torch.empty(nq, nlvl, **i32), best=torch.empty(nq, 2, **i32))
    ptrs = lambda ts: (C.c_void_p * len(ts))(*[t.data_ptr() for t in ts])
    hw = (C.c_int32 * nlvl)(*[int(c.shape[2] * c.shape[3]) for c in cls_c])
    wd = (C.c_int32 * nlvl)(*[int(c.shape[3]) for c in cls_c])
    st = (C.c_float * nlvl)(*[float(s) for s in anchor_strides[:nlvl]])
    sz = (C.c_float * nsizes)(*[float(s) for s in anchor_sizes])
    with torch.cuda.device(dev):
        rc = L.kdot_select_cells(ptrs(cls_c), ptrs(reg_c), hw, wd, st, nlvl, sz, nsizes, nimg, ncls,
                                 float(inference_th), int(positive_num), float(positive_lambda), cap,
                                 out["sel_count"].data_ptr(), out["sel_level"].data_ptr(), out["sel_loc"].data_ptr(),
                                 out["sel_score"].data_ptr(), out["sel_kpts"].data_ptr(), out["nk"].data_ptr(),
                                 out["valid_cnt"].data_ptr(), out["best"].data_ptr(),
                                 torch.cuda.current_stream(dev).cuda_stream)
    _lib.check(rc, "kdot_select_cells")
    out["cap"], out["ncls"], out["nimg"] = cap, ncls, nimg
    return out


def _prepare_pnp_tasks(sel_np, targets, ncls, class_filters=None):
    """First half of the host part of ``pose_infer_ml`` (``postprocess_kd.py:158-189`` / ``postprocess.py:150-188``),
    on the calling thread: for every image and candidate label (ascending), the selected key-points un-cropped to
    full-image pixels -- ``inverse(A) @ (pt - t)`` for ALL rows of the mini-batch in one batched ``torch.inverse`` /
    ``torch.bmm`` pair instead of one pair per (image, label).
    Returns per image a list of ``(class_id, scores (n,), xy2d (n,8,2) float32 tensor, pts3d (n*8,3), K (3,3))``."""
    count, valid, score, kpts = sel_np["count"], sel_np["valid"], sel_np["score"], sel_np["kpts"]
    rows, owners = [], []
    for i, c in np.argwhere((valid.sum(axis=-1) > 0) & (count > 0)).tolist():  # image-major, labels ascending
        if class_filters is not None and c not in class_filters[i]:
            continue
        n = int(count[i, c])
        owners.append((i, c, n))
        rows.append(kpts[i, c, :n])
    tasks = [[] for _ in targets]
    if not owners:
        return tasks
    xy = torch.from_numpy(np.concatenate(rows, axis=0)).view(-1, 2, 8)  # (rows, 2, 8): x row, y row
    with_bt = [targets[i].bbox_trans is not None for i, _c, _n in owners]
    if any(with_bt):
        eye = torch.tensor([[1.0, 0.0, 0.0], [0.0, 1.0, 0.0]])
        bt = torch.cat([(targets[i].bbox_trans.detach().cpu().to(torch.float32).view(1, 2, 3) if has else eye.view(1, 2, 3))
                        .expand(n, 2, 3) for (i, _c, n), has in zip(owners, with_bt)], dim=0)
        lin, off = bt[:, :, :2].contiguous(), bt[:, :, 2].unsqueeze(-1)
        moved = torch.bmm(torch.inverse(lin), xy - off)
        keep = torch.tensor(np.repeat(with_bt, [n for _i, _c, n in owners]))
        xy = torch.where(keep.view(-1, 1, 1), moved, xy)
    xy = xy.transpose(1, 2).contiguous()  # (rows, 8, 2)
    k_np = {}
    o = 0
    for i, c, n in owners:
        if i not in k_np:
            k_np[i] = targets[i].K.detach().cpu().numpy()
        pts3d = np.tile(targets[i].keypoints_3d[c].detach().cpu().numpy(), (n, 1))
        tasks[i].append((c, score[i, c, :n].copy(), xy[o:o + n], pts3d, k_np[i]))
        o += n
    return tasks


def _solve_image(tasks_i, first_only):
    """Second half (``postprocess_kd.py:190-203``), safe to run on a worker thread (OpenCV releases the GIL):
    RANSAC-EPnP per candidate label; ``first_only`` stops at the first label whose PnP succeeds
    (``select_over_all_levels`` keeps only that one, ``postprocess_kd.py:71-97``)."""
    import cv2

    out = []
    for c, sc, xy2d, pts3d, K_np in tasks_i:
        ok, rot, trans, _inl = cv2.solvePnPRansac(pts3d, xy2d.view(-1, 2).numpy(), K_np, None,
                                                  flags=cv2.SOLVEPNP_EPNP, reprojectionError=5.0)
        if not ok:
            continue
        R = cv2.Rodrigues(rot)[0]
        T = trans.reshape(-1, 1)
        if np.isnan(R.sum()) or np.isnan(T.sum()):
            continue
        out.append((c, sc, xy2d, R, T))
        if first_only:
            break
    return out


class _SelectingPostProcessor(nn.Module):
    """Shared device part: ONE ``kdot_select_cells`` launch + ONE device->host copy of the selected cells."""

    def __init__(self, inference_th, box_coder, positive_num, positive_lambda, sym_types, symmetry_fn=None):
        super().__init__()
        self.inference_th = inference_th
        self.positive_num = positive_num
        self.positive_lambda = positive_lambda
        self.box_coder = box_coder
        self.sym_types = sym_types
        self.symmetry_fn = symmetry_fn  # libs.utils.pose_symmetry_handling of the host repo (rotation only)
        self.last_selection = None
        # per-image RANSAC-EPnP calls are independent and OpenCV releases the GIL: run them on a host thread pool
        # (SURVEY.md section 8(f) item 2).  cv2's RANSAC seeds its own RNG per call, so results do not depend on
        # the schedule.  `pnp_threads = 1` restores the reference's serial loop.
        self.pnp_threads = min(16, len(os.sched_getaffinity(0))) if hasattr(os, "sched_getaffinity") else 4
        self._pool = None

    def _map_images(self, fn, nimg):
        """``[fn(i) for i in range(nimg)]`` on the host thread pool (ordered)."""
        if self.pnp_threads <= 1 or nimg <= 1:
            return [fn(i) for i in range(nimg)]
        import cv2

        if self._pool is None or self._pool._max_workers != self.pnp_threads:
            self._pool = ThreadPoolExecutor(max_workers=self.pnp_threads, thread_name_prefix="kdot-pnp")
        # OpenCV serialises concurrent callers on its own (idle for PnP) worker pool: switch it off for the duration
        # of the map (measured: 1.6 ms/image serial or pooled with it on, 0.42 ms/image on 8 host threads with it off)
        inner = cv2.getNumThreads()
        cv2.setNumThreads(1)
        try:
            return list(self._pool.map(fn, range(nimg)))
        finally:
            cv2.setNumThreads(inner)

    def _select(self, box_cls, box_regression):
        sel = select_cells(box_cls, box_regression, self.box_coder.anchor_sizes, self.box_coder.anchor_strides,
                           self.inference_th, self.positive_num, self.positive_lambda)
        nimg, ncls, cap = sel["nimg"], sel["ncls"], sel["cap"]
        self.last_selection = dict(
            count=sel["sel_count"].cpu().numpy().reshape(nimg, ncls),
            valid=sel["valid_cnt"].cpu().numpy().reshape(nimg, ncls, -1),
            score=sel["sel_score"].cpu().numpy().reshape(nimg, ncls, cap),
            kpts=sel["sel_kpts"].cpu().numpy().reshape(nimg, ncls, cap, 16),
            level=sel["sel_level"].cpu().numpy().reshape(nimg, ncls, cap),
            loc=sel["sel_loc"].cpu().numpy().reshape(nimg, ncls, cap),
            nk=sel["nk"].cpu().numpy().reshape(nimg, ncls, -1),
            best=sel["best"].cpu().numpy().reshape(nimg, ncls, 2))
        return nimg, ncls

    def _symmetry(self, c, R):
        if self.sym_types is not None and len(self.sym_types) > 0 and ("cls_" + str(c)) in self.sym_types:
            if self.symmetry_fn is None:
                from libs.utils import pose_symmetry_handling  # host repository

                self.symmetry_fn = pose_symmetry_handling
            R = self.symmetry_fn(R, self.sym_types["cls_" + str(c)])
        return R


class PostProcessorKD(_SelectingPostProcessor):
    """Teacher knowledge extraction (``postprocess/postprocess_kd.py:12``): first label whose PnP succeeds."""

    def forward(self, box_cls, box_regression, targets, anchors=None):
        nimg, ncls = self._select(box_cls, box_regression)
        dev = box_cls[0].device
        results = [[], [], [], []]
        tasks = _prepare_pnp_tasks(self.last_selection, targets, ncls)
        solved = self._map_images(lambda i: _solve_image(tasks[i], first_only=True), nimg)
        for i in range(nimg):
            picked = solved[i][0] if solved[i] else None
            if picked is not None:
                c, sc, xy2d, R, T = picked
                n = len(sc)
                results[0].append(torch.from_numpy(np.broadcast_to(sc[:, None], (n, 8)).copy()).to(dev))
                results[1].append(self._symmetry(c, R))
                results[2].append(T)
                results[3].append(xy2d.to(dev))
            else:
                results[0].append(torch.zeros([0, 8], device=dev))
                results[1].append(torch.zeros([0, 3, 3]))
                results[2].append(torch.zeros([0, 3, 1]))
                results[3].append(torch.zeros([0, 8, 2], device=dev))
        return results


class PostProcessor(_SelectingPostProcessor):
    """Student evaluation post-processor (``postprocess/postprocess.py:12``): the same
    selection, restricted to the classes present in the target (``postprocess.py:111-113``), every label kept:
    per image a list of ``[max score, class id, R, T, xy2d (n,8,2)]``."""

    def forward(self, box_cls, box_regression, targets, anchors=None):
        nimg, ncls = self._select(box_cls, box_regression)
        present = [set(int(v) for v in targets[i].class_ids.detach().cpu().reshape(-1).tolist()) for i in range(nimg)]
        tasks = _prepare_pnp_tasks(self.last_selection, targets, ncls, class_filters=present)
        solved = self._map_images(lambda i: _solve_image(tasks[i], first_only=False), nimg)
        return [[[float(sc.max()), int(c), self._symmetry(c, R), T, xy2d] for c, sc, xy2d, R, T in per_img]
                for per_img in solved]


def teacher_knowledge(post_processor, pred_cls, pred_reg, targets, anchors=None):
    """The teacher branch of ``PoseModuleKD.forward`` (``models/model_kd.py:83-92``): the dict the loss consumes."""
    pred = post_processor(pred_cls, pred_reg, targets, anchors)
    return {"post_kp_2d": torch.cat(pred[3], dim=0), "post_kp_cls": torch.cat(pred[0], dim=0),
            "post_pos_per_img": [len(p) for p in pred[0]]}

"""``KDPoseLoss`` -- drop-in for ``/root/reference/losses/kd_loss.py:13`` (seam B0, public and unchanged).

Constructor and ``__call__`` signatures, the ``cfg_kd`` / ``pred_t`` keys read, the attributes left on the
instance (``pred_cls, pos_per_img, batch_size, w, h, cls_id, step, vis_dir``) and the returned
``[cls_loss, reg_loss, kd_loss]`` are those of the reference.  What changes is the OT section
(``kd_loss.py:73-103``): instead of ``kd_loss_2d``'s Python loop over images and geomloss' hundreds of tiny
launches per image, the whole mini-batch goes through one fused CUDA launch (``ops.OTLossFunction``).

The class derives from the integrating repository's own ``PoseLossDzi`` (``losses/loss.py:99``: target
assignment, focal loss -- host-side label logic that is out of scope here and reused unchanged).  It is
resolved lazily so that this package imports without the reference on ``sys.path``; tests inject a
fixture-backed base through :func:`make_kd_pose_loss`.
"""
from __future__ import annotations

import os

import numpy as np
import torch
from torch import nn

from ..ops import FocalLossFunction, OTLossFunction, Reg3dLossFunction, gather_decode
from ..samples_loss import SamplesLoss

INF = 100000000


def flatten_level_list(levels):
    """One head output, per-level ``(nimg, C, H, W)`` -> ``(nimg*cells, C)`` in the same label order."""
    parts = [t.permute(0, 2, 3, 1).reshape(t.shape[0], -1, t.shape[1]) for t in levels]
    return (parts[0] if len(parts) == 1 else torch.cat(parts, dim=1)).reshape(-1, parts[0].shape[2])


def flatten_head_outputs(pred_cls, pred_reg):
    """Per-level ``(nimg, C, H, W)`` / ``(nimg, C*16, H, W)`` -> ``(nimg*cells, C)`` / ``(nimg*cells, C*16)``
    in the label order of the reference (``losses/loss.py:62-96``): image-major, levels concatenated,
    row-major cells."""
    cls_f, reg_f = flatten_level_list(pred_cls), flatten_level_list(pred_reg)
    return cls_f, reg_f


def _num_gpus() -> int:
    """``losses/loss.py:42-43`` (``get_num_gpus``): the process count comes from the launcher's ``WORLD_SIZE``."""
    return int(os.environ["WORLD_SIZE"]) if "WORLD_SIZE" in os.environ else 1


def _reduce_sum_int(value: int, device) -> int:
    """``losses/loss.py:45-51`` (``reduce_sum``): sum of the positive count over ranks.  Like the reference the
    decision is taken on ``WORLD_SIZE``; unlike it, a launcher environment without an initialised process group
    (e.g. a single rank started under torchrun for debugging) returns the local value instead of raising."""
    if _num_gpus() <= 1:
        return value
    import torch.distributed as dist

    if not (dist.is_available() and dist.is_initialized()):
        return value
    t = torch.tensor([value], dtype=torch.int64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return int(t.item())


def _reference_visualizer(pred_xy, pred_t_xy, s_cls, t_cls, step, vis_dir, pos_s, pos_t, losses):
    """The reference's debug scatter plots (``losses/kd_loss.py:88-89,96-97`` -> ``tools/visualizer.py``), called with
    the in-place-normalised key-points exactly as there.  Resolved from the integrating repository at call time; when
    ``tools.visualizer`` (matplotlib) is not importable the plots are skipped -- they are a side effect, not a result."""
    try:
        from tools.visualizer import vis_pxpy_post_train, vis_pxpy_post_train_weight
    except Exception:
        return False
    try:
        if s_cls is not None:
            vis_pxpy_post_train_weight(pred_xy, pred_t_xy, s_cls.reshape(-1, 1), t_cls.reshape(-1, 1), step, save_dir=vis_dir,
                                       pos_per_img_1=pos_s, pos_per_img_2=pos_t, loss=losses)
        else:
            vis_pxpy_post_train(pred_xy, pred_t_xy, step, save_dir=vis_dir, pos_per_img_1=pos_s, pos_per_img_2=pos_t,
                                loss=losses)
    except Exception as exc:  # a broken plotting stack must not take the training step down with it
        import warnings

        warnings.warn(f"KDPoseLoss: the reference visualiser failed at step {step} ({exc!r}); plots skipped")
        return False
    return True


def _status_error(valid_host, where):
    """Turn a negative per-image kernel status (``include/kdot.h`` KDOT_IMG_*) into an exception.  geomloss raises in
    the same situations (``epsilon_schedule`` on a zero diameter); the kernel writes NaN and a status instead, so
    without this check the NaN would reach the optimizer silently."""
    from .._lib import KdotError

    names = {-1: "KDOT_IMG_DEGENERATE (all key-points of the image coincide: zero diameter)",
             -2: "KDOT_IMG_TOO_MANY_ROUNDS (epsilon schedule longer than KDOT_MAX_ROUNDS; scaling too close to 1)"}
    bad = [(i, int(v)) for i, v in enumerate(valid_host) if v < 0]
    if bad:
        i, v = bad[0]
        raise KdotError(f"{where}: image {i} of the mini-batch has status {names.get(v, v)}; "
                        f"{len(bad)} image(s) affected, their loss and gradients are NaN")


def make_kd_pose_loss(base):
    """Build the ``KDPoseLoss`` class on top of ``base`` (the reference's ``PoseLossDzi`` or a test double
    exposing ``prepare_targets``, ``cls_loss_func``, ``target_coder``, ``internal_K``, ``diameters``)."""

    class KDPoseLoss(base):
        def __init__(self, gamma, alpha, anchor_sizes, anchor_strides, positive_type, positive_num,
                     positive_lambda, top_k, internal_K, diameters, target_coder, cfg_kd=None):
            super().__init__(gamma, alpha, anchor_sizes, anchor_strides, positive_type, positive_num,
                             positive_lambda, top_k, internal_K, diameters, target_coder)
            if cfg_kd is not None:
                self.cfg_kd = cfg_kd
                self.kd_loss = SamplesLoss(cfg_kd["GTYPE"], p=cfg_kd["GP"], blur=cfg_kd["GBLUR"],
                                           scaling=cfg_kd["SCALING"], reach=cfg_kd["REACH"])
                self.weighted_ot = cfg_kd["WEIGHTED_OT"]
                self.wot_detach = cfg_kd["DETACH"]
                if "vis_dir" in cfg_kd.keys():
                    self.step = 0
                    self.vis_dir = cfg_kd["vis_dir"] + "/vis"
                    os.makedirs(self.vis_dir, exist_ok=True)
            # callable(pred_xy, pred_t_xy, s_cls, t_cls, step, vis_dir, pos_s, pos_t, losses); None disables the plots
            self.visualizer = _reference_visualizer
            # per-image kernel status of the previous step: checked at the start of the next one (no extra host sync
            # in the step itself); KDOT_SYNC_STATUS=1 checks right after the launch instead
            self._pending_status = None

        # -- the 3-D regression loss of kd_loss.py:52-71: one fused forward + backward launch (kdot_reg3d_loss_fwd_bwd) --
        def _object_space_reg_loss(self, pred_xy, target_3d, cls_labels):
            if not isinstance(self.diameters, torch.Tensor):
                self.diameters = torch.FloatTensor(self.diameters).to(device=pred_xy.device).view(-1)
            if not isinstance(self.internal_K, torch.Tensor):
                self.internal_K = torch.FloatTensor(self.internal_K).to(device=pred_xy.device).view(3, 3)
            if getattr(self, "_kinv", None) is None:
                # inverse intrinsics once, on the host (the reference calls torch.inverse on the device every step)
                self._kinv = np.linalg.inv(self.internal_K.detach().cpu().numpy().astype(np.float64)).reshape(-1).tolist()
            return Reg3dLossFunction.apply(pred_xy, target_3d, self.diameters[cls_labels], self._kinv)

        def KDObjectSpaceLoss(self, pred, target_2D, target_3D_in_camera_frame, cls_labels, anchors, pred_t,
                              bbox_trans, weight=None):
            self.cls_id = torch.unique(cls_labels)
            n_cell = pred.shape[0]
            picked = pred.view(n_cell, -1, 16)[torch.arange(n_cell, device=pred.device), cls_labels]
            pred_xy = self.target_coder.decode(picked, anchors, bbox_trans)
            pred_xy = pred_xy.view(-1, 2, 8).transpose(1, 2).contiguous().view(-1, 2)  # (cells*8, 2) px
            return self._losses_from_keypoints(pred_xy, target_3D_in_camera_frame, cls_labels, pred_t, weight)

        def _losses_from_keypoints(self, pred_xy, target_3D_in_camera_frame, cls_labels, pred_t, weight=None):
            """Everything of ``KDObjectSpaceLoss`` after the decode (``kd_loss.py:52-103``): 3-D regression loss and
            the OT distillation loss on ``pred_xy (cells*8, 2)`` pixel key-points."""
            losses = self._object_space_reg_loss(pred_xy, target_3D_in_camera_frame, cls_labels)

            # ---- OT distillation (kd_loss.py:73-103), one fused launch for the mini-batch ----
            if self.cfg_kd["GnD"] != 2:
                raise NotImplementedError("cfg_kd['GnD'] != 2: the reference defines the KD loss for 2-D keypoints only")
            if self.cfg_kd["GLEVEL"] != "point":
                raise NotImplementedError("cfg_kd['GLEVEL'] must be 'point'")
            pos_s = [int(v) for v in self.pos_per_img]
            pos_t = [int(v) for v in pred_t["post_pos_per_img"]]
            xt = pred_t["post_kp_2d"].reshape(-1, 8, 2)
            if not xt.is_contiguous():
                xt = xt.contiguous()
            if self.weighted_ot:
                t_cls = pred_t["post_kp_cls"].pow(2)
                s_cls = torch.broadcast_to(self.pred_cls[..., self.cls_id], (self.pred_cls.size(0), 8))
                if self.wot_detach:
                    s_cls = s_cls.detach()
                s_cls = s_cls.contiguous()
            else:
                t_cls = s_cls = None
            n_valid = sum(1 for n, m in zip(pos_s, pos_t) if n > 0 and m > 0)
            # pred_xy / xt are normalised in place by the kernel (loss_libs.py:8-12): the visualiser below and
            # any later reader of pred_t['post_kp_2d'] see the normalised values, as with the reference.
            loss_per_img, valid, _nits = OTLossFunction.apply(
                pred_xy.view(-1, 8, 2), s_cls, xt, t_cls, pos_s, pos_t, self.kd_loss.config,
                float(self.w), float(self.h), True)
            if os.environ.get("KDOT_SYNC_STATUS", "0") == "1":
                _status_error(valid.cpu().tolist(), "KDPoseLoss")
            else:
                self._pending_status = valid
            if self.visualizer is not None and hasattr(self, "step") and (self.step == 0 or (self.step + 1) % 1000 == 0):
                keep = [i for i, (n, m) in enumerate(zip(pos_s, pos_t)) if n > 0 and m > 0]
                self.visualizer(pred_xy.detach(), xt.view(-1, 2), s_cls, t_cls, self.step, self.vis_dir, pos_s, pos_t,
                                [loss_per_img[i].detach() for i in keep])
            if n_valid > 0:
                loss_kd = loss_per_img.sum() / n_valid  # skipped images contribute exactly 0
            else:
                # kd_loss.py:99-103: no image has both student and teacher cells -> a constant 0 on the device
                loss_kd = torch.tensor(0.0, device=pred_xy.device)

            if weight is not None and weight.sum() > 0:
                return (losses * weight).sum()
            assert losses.numel() != 0
            return losses.sum(), loss_kd

        def check_status(self):
            """Raise if the previous step's kernel reported a degenerate image (see ``_status_error``)."""
            pending, self._pending_status = getattr(self, "_pending_status", None), None
            if pending is not None:
                _status_error(pending.cpu().tolist(), "KDPoseLoss (previous step)")

        def _call_with_device_targets(self, pred_cls, pred_reg, targets, anchors, pred_t, mode):
            """``__call__`` with the SSC label assignment on the device (``kd_6d_pose_adlp_b200.targets``) instead of
            the base class' host loops (``losses/loss.py:164-268``): same three losses, same instance attributes."""
            from ..targets import positives_aux, ssc_assign

            if self.positive_type != "SSC":
                raise NotImplementedError("DEVICE_TARGETS implements POSITIVE_TYPE == 'SSC' (the only type the reference defines)")
            self.batch_size = len(targets)
            self.h, self.w = 480, 640
            first = anchors[0]
            anchors_one = torch.cat([a.bbox if hasattr(a, "bbox") else a for a in first], dim=0)
            hw = [int(c.shape[2] * c.shape[3]) for c in pred_cls]
            self._ssc_step = getattr(self, "_ssc_step", 0) + 1
            res = ssc_assign(targets, anchors_one, hw, self.anchor_sizes, self.positive_num, self.positive_lambda, mode=mode,
                             seed=int(os.environ.get("KDOT_SSC_SEED", "0")) * 1000003 + self._ssc_step)
            labels_flat = res["labels"]
            f = self.cls_loss_func
            cls_loss = FocalLossFunction.apply(labels_flat, float(f.gamma), float(f.alpha), *pred_cls)
            pos_inds = torch.nonzero(labels_flat > 0).squeeze(1)
            pos_per_img = res["npos"].cpu().tolist()          # the one host sync of the step
            total_num_pos = _reduce_sum_int(int(sum(pos_per_img)), labels_flat.device)
            if pos_inds.numel() > 0:
                self.pos_per_img = pos_per_img
                if _num_gpus() <= 1:
                    assert sum(self.pos_per_img) == total_num_pos
                if self.target_coder.target_type != "3D":
                    raise NotImplementedError("KDPoseLoss: only LOSS_REG_TYPE == '3D' carries the KD loss (kd_loss.py:150-153)")
                cls_label, aux_3d_pos, bt_pos = positives_aux(res, pos_inds)
                if self.weighted_ot:
                    self.pred_cls = torch.clamp(torch.sigmoid(flatten_level_list(pred_cls)[pos_inds]), min=10e-4, max=1 - 10e-4)
                self.cls_id = torch.unique(cls_label)
                cell = torch.remainder(pos_inds, res["cells"])
                pred_xy = gather_decode(pred_reg, pos_inds, cls_label, anchors_one[cell], bt_pos)
                reg_loss, kd_loss = self._losses_from_keypoints(pred_xy, aux_3d_pos, cls_label, pred_t)
            else:
                reg_loss = sum(r.sum() for r in pred_reg)
                kd_loss = reg_loss
            if hasattr(self, "step"):
                self.step += 1
            return [cls_loss, reg_loss, kd_loss]

        def __call__(self, pred_cls, pred_reg, targets, anchors, pred_t):
            self.check_status()
            mode = (getattr(self, "cfg_kd", None) or {}).get("DEVICE_TARGETS") or os.environ.get("KDOT_DEVICE_TARGETS")
            if mode:
                return self._call_with_device_targets(pred_cls, pred_reg, targets, anchors, pred_t, mode)
            labels, reg_targets, aux_raw_boxes, aux_3d, aux_bbox_trans = self.prepare_targets(targets, anchors)
            self.batch_size = len(labels)
            self.h = 480  # full-image size, not the 256 crop (kd_loss.py:116-117)
            self.w = 640

            # only the class logits are flattened (the focal loss reads every cell); the 240-channel regression maps
            # are read in place by the gather/decode kernel at the positive cells -- the reference's flatten of
            # pred_reg (losses/loss.py:62-96) is ~83 MB of copies per direction at batch 64 for ~640 used rows
            labels_flat = torch.cat(labels, dim=0)
            aux_3d_flat = torch.cat(aux_3d, dim=0)
            anchors_flat = self._flatten_anchors(anchors)
            bbox_trans_flat = torch.cat(aux_bbox_trans, dim=0)

            pos_inds = torch.nonzero(labels_flat > 0).squeeze(1)
            # all per-image positive counts from one prefix sum and ONE host sync (the reference pays nimg + 1
            # `.item()` calls, kd_loss.py:131,145)
            ends = np.cumsum([int(lb.shape[0]) for lb in labels])
            at_ends = torch.cumsum(labels_flat > 0, dim=0)[torch.from_numpy(ends - 1).to(labels_flat.device)]
            pos_per_img = np.diff(at_ends.cpu().numpy(), prepend=0).tolist()
            total_num_pos = _reduce_sum_int(int(sum(pos_per_img)), labels_flat.device)

            f = self.cls_loss_func
            if all(hasattr(f, a) for a in ("gamma", "alpha", "eps")) and abs(float(f.eps) - 1e-4) < 1e-12 and \
                    getattr(type(f), "forward", None) is not None and type(f).__name__ in ("SigmoidFocalLoss", "FocalLoss"):
                # the reference's SigmoidFocalLoss (losses/loss.py:12-40): fused forward + gradient on the per-level
                # logits, ignored cells (label -1) skipped inside the kernel instead of by boolean-index copies
                cls_loss = FocalLossFunction.apply(labels_flat, float(f.gamma), float(f.alpha), *pred_cls)
            else:  # a foreign classification loss: the reference's call (kd_loss.py:133-134)
                valid_inds = torch.nonzero(labels_flat >= 0).squeeze(1)
                cls_loss = f(flatten_level_list(pred_cls)[valid_inds], labels_flat[valid_inds])

            if pos_inds.numel() > 0:
                self.pos_per_img = pos_per_img
                if _num_gpus() <= 1:
                    assert sum(self.pos_per_img) == total_num_pos
                cls_label = labels_flat[pos_inds] - 1
                if self.target_coder.target_type != "3D":
                    raise NotImplementedError("KDPoseLoss: only LOSS_REG_TYPE == '3D' carries the KD loss (kd_loss.py:150-153)")
                if self.weighted_ot:
                    self.pred_cls = torch.clamp(torch.sigmoid(flatten_level_list(pred_cls)[pos_inds]), min=10e-4, max=1 - 10e-4)
                if getattr(self.target_coder, "regression_type", "POINT") != "POINT":
                    raise NotImplementedError("KDPoseLoss: only the 'POINT' regression type is defined (models/model.py:145,163)")
                self.cls_id = torch.unique(cls_label)
                pred_xy = gather_decode(pred_reg, pos_inds, cls_label, anchors_flat[pos_inds], bbox_trans_flat[pos_inds])
                reg_loss, kd_loss = self._losses_from_keypoints(pred_xy, aux_3d_flat[pos_inds], cls_label, pred_t)
            else:
                reg_loss = sum(r.sum() for r in pred_reg)  # == pred_reg_flatten.sum() (kd_loss.py:158-159)
                kd_loss = reg_loss
            if hasattr(self, "step"):
                self.step += 1
            return [cls_loss, reg_loss, kd_loss]

        @staticmethod
        def _flatten_anchors(anchors):
            """``anchors``: per image, per level BoxList-like objects with ``.bbox`` (or plain tensors)."""
            return torch.cat([a.bbox if hasattr(a, "bbox") else a for levels in anchors for a in levels], dim=0)

    KDPoseLoss.__qualname__ = "KDPoseLoss"
    return KDPoseLoss


_resolved = None


def __getattr__(name):
    """``from kd_6d_pose_adlp_b200.losses.kd_loss import KDPoseLoss`` binds to the host repo's ``PoseLossDzi``."""
    global _resolved
    if name == "KDPoseLoss":
        if _resolved is None:
            try:
                from losses.loss import PoseLossDzi  # the integrating repository (reference layout)
            except Exception as exc:  # pragma: no cover - depends on the host repo
                raise ImportError(
                    "KDPoseLoss derives from the host repository's losses.loss.PoseLossDzi, which is not importable; "
                    "put the repository root on sys.path or call make_kd_pose_loss(base) explicitly") from exc
            _resolved = make_kd_pose_loss(PoseLossDzi)
        return _resolved
    raise AttributeError(name)

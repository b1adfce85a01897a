"""Test doubles for the host-side pieces of the reference that are out of scope:
a fixture-backed stand-in for ``PoseLossDzi`` (target assignment replayed from ``tests/golden/kd_pose_loss.npz``,
sigmoid focal loss restated from ``losses/loss.py:12-40``) and a minimal ``PoseAnnot``-like target."""
import hashlib

import numpy as np
import torch
from torch import nn


def digest(arrays):
    h = hashlib.sha256()
    for a in arrays:
        h.update(np.ascontiguousarray(a).tobytes())
    return h.hexdigest()


class FocalLoss(nn.Module):
    def __init__(self, gamma, alpha):
        super().__init__()
        self.gamma, self.alpha, self.eps = gamma, alpha, 1e-4

    def forward(self, out, target):
        ids = torch.arange(1, out.shape[1] + 1, dtype=target.dtype, device=target.device).unsqueeze(0)
        t = target.unsqueeze(1)
        p = torch.clamp(torch.sigmoid(out), min=self.eps, max=1 - self.eps)
        pos = (t == ids).float()
        neg = ((t != ids) * (t >= 0)).float()
        loss = -pos * self.alpha * (1 - p) ** self.gamma * torch.log(p) \
               - neg * (1 - self.alpha) * p ** self.gamma * torch.log(1 - p)
        return loss.sum()


class ReplayBase(object):
    """Plays back a recorded ``prepare_targets`` result (the reference's is host-side and random)."""

    recorded = None  # set by the test: dict of per-image lists on the target device

    def __init__(self, gamma, alpha, anchor_sizes, anchor_strides, positive_type, positive_num, positive_lambda, top_k,
                 internal_K, diameters, target_coder):
        self.cls_loss_func = FocalLoss(gamma, alpha)
        self.anchor_sizes, self.anchor_strides = anchor_sizes, anchor_strides
        self.positive_type, self.positive_num, self.positive_lambda, self.top_k = positive_type, positive_num, positive_lambda, top_k
        self.internal_K, self.target_coder, self.diameters = internal_K, target_coder, diameters

    def prepare_targets(self, targets, anchors):
        r = self.recorded
        return r["labels"], r["reg_targets"], r["aux_raw_boxes"], r["aux_3d"], r["aux_bbox_trans"]


class Target(object):
    def __init__(self, K, keypoints_3d, bbox_trans):
        self.K, self.keypoints_3d, self.bbox_trans = K, keypoints_3d, bbox_trans

"""Keypoint target coder and anchor grid -- the two formulas that sit immediately before the hot path.

``TargetCoder.decode`` restates ``/root/reference/models/model.py:144-166`` (offsets * anchor size + anchor
centre, then the inverse of the 2x3 crop affine ``bbox_trans``); ``grid_anchors`` restates the one-square-
anchor-per-cell grid of ``models/model.py:229-251,283-347``.  They exist so that this package (tests,
benchmarks, the ``PostProcessorKD`` mirror) is usable without the reference on ``sys.path``; when integrated,
the host repository's own ``TargetCoder`` instance is passed in and used as is.
"""
from __future__ import annotations

import torch


class TargetCoder(object):
    def __init__(self, regression_type, anchor_sizes, anchor_strides, target_type="3D"):
        self.regression_type = regression_type
        self.anchor_sizes = anchor_sizes
        self.anchor_strides = anchor_strides
        self.target_type = target_type

    def decode(self, preds, anchors, bbox_trans=None):
        """``preds (n,16) = [dx0..dx7, dy0..dy7]``, ``anchors (n,4)`` xyxy -> ``(n,16) = [x0..x7, y0..y7]``."""
        if self.regression_type != "POINT":
            raise NotImplementedError(self.regression_type)
        aw = (anchors[:, 2] - anchors[:, 0] + 1).view(-1, 1)
        ah = (anchors[:, 3] - anchors[:, 1] + 1).view(-1, 1)
        acx = ((anchors[:, 2] + anchors[:, 0]) / 2).view(-1, 1)
        acy = ((anchors[:, 3] + anchors[:, 1]) / 2).view(-1, 1)
        ptx = preds[:, :8] * aw + acx
        pty = preds[:, 8:] * ah + acy
        if bbox_trans is not None:
            pts = torch.stack([ptx, pty]).transpose(0, 1)  # (n, 2, 8)
            lin = bbox_trans[:, :, :2]
            off = bbox_trans[:, :, 2].unsqueeze(-1)
            pts = torch.bmm(torch.inverse(lin), pts - off)
            ptx, pty = pts[:, 0, :], pts[:, 1, :]
        return torch.cat((ptx, pty), dim=1)


def grid_anchors(grid_sizes, anchor_sizes, anchor_strides, device="cpu"):
    """One square anchor per cell and level: side ``size``, centre ``(w*stride + stride/2, h*stride + stride/2)``
    in xyxy form with the reference's ``+1`` width convention (``x2 - x1 + 1 == size``).
    Returns a list (per level) of ``(H*W, 4)`` float32 tensors, row-major over (h, w)."""
    out = []
    for (gh, gw), size, stride in zip(grid_sizes, anchor_sizes, anchor_strides):
        sx = torch.arange(0, gw * stride, step=stride, dtype=torch.float32, device=device)
        sy = torch.arange(0, gh * stride, step=stride, dtype=torch.float32, device=device)
        yy, xx = torch.meshgrid(sy, sx, indexing="ij")
        xx, yy = xx.reshape(-1), yy.reshape(-1)
        half = 0.5 * (float(size) - 1.0)
        c = 0.5 * float(stride)
        out.append(torch.stack((xx + (c - half), yy + (c - half), xx + (c + half), yy + (c + half)), dim=1))
    return out

"""Host-side operators over the C ABI: the fused OT distillation loss (forward + analytic backward).

``ot_loss_batched`` is the whole per-mini-batch OT section of the reference
(``/root/reference/losses/kd_loss.py:73-103`` -> ``losses/loss_libs.py:1-51`` -> geomloss) as ONE fused
launch; ``OTLossFunction`` plugs it into autograd.  PyTorch is used here for device memory, streams and
the autograd graph only -- all arithmetic runs in ``libkdot.so``.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass
from typing import Optional, Sequence

import numpy as np
import torch

from . import _lib


@dataclass(frozen=True)
class OTConfig:
    """``SamplesLoss("sinkhorn", p, blur, scaling, reach)`` knobs (reference ``losses/kd_loss.py:26-30``;
    defaults of ``arguments/argument_kd.py:41-49``)."""

    p: float = 2.0
    blur: float = 0.001
    scaling: float = 0.5
    reach: Optional[float] = 0.5  # None -> balanced OT
    loss: str = "sinkhorn"        # or "gaussian" / "laplacian" / "energy" (kernel MMD, --gtype)


_workspaces = {}


def _workspace(device: torch.device, nbytes: int) -> Optional[torch.Tensor]:
    if nbytes == 0:
        return None
    key = (device.index, torch.cuda.current_stream(device).cuda_stream)
    buf = _workspaces.get(key)
    if buf is None or buf.numel() < nbytes:
        buf = torch.empty(nbytes, dtype=torch.uint8, device=device)
        _workspaces[key] = buf
    return buf


def _ptr(t: Optional[torch.Tensor]):
    return None if t is None else t.data_ptr()


class _on_device:
    """``torch.cuda.device(dev)`` only when ``dev`` is not already current (the context manager costs ~7 us)."""

    def __init__(self, dev):
        self.ctx = None if torch.cuda.current_device() == dev.index else torch.cuda.device(dev)

    def __enter__(self):
        if self.ctx is not None:
            self.ctx.__enter__()

    def __exit__(self, *exc):
        if self.ctx is not None:
            self.ctx.__exit__(*exc)


def cu_seqlens(counts: Sequence[int], device) -> torch.Tensor:
    """Exclusive prefix sums of the per-image cell counts as a device int32 tensor (one small H2D copy)."""
    cu = np.zeros(len(counts) + 1, np.int32)
    np.cumsum(np.asarray(counts, np.int64), out=cu[1:])
    return torch.from_numpy(cu).to(device, non_blocking=True)


def _check_f32_cuda(name, t, shape=None):
    if not (isinstance(t, torch.Tensor) and t.is_cuda and t.dtype == torch.float32 and t.is_contiguous()):
        raise ValueError(f"{name} must be a contiguous float32 CUDA tensor (libkdot has no CPU fallback)")
    if shape is not None and tuple(t.shape) != tuple(shape):
        raise ValueError(f"{name} has shape {tuple(t.shape)}, expected {tuple(shape)}")


def ot_loss_batched(xs, ws, xt, wt, pos_per_img, pos_per_img_t, cfg: OTConfig = OTConfig(), *, w=640.0, h=480.0,
                    normalize=True, layout=_lib.KDOT_LAYOUT_CELL_MAJOR, cu_n=None, cu_m=None, want_slots=False):
    """Raw fused call.  ``xs (sumN,B,D)``, ``ws (sumN,B)|None``, ``xt (sumM,B,D)``, ``wt (sumM,B)|None`` (or the
    slot-major ``(B,N,D)`` forms with ``layout=KDOT_LAYOUT_SLOT_MAJOR`` and one image).
    ``xs``/``xt`` are normalised IN PLACE when ``normalize``.

    Returns ``dict(loss_per_img (nimg,), loss_per_slot (nimg,B)|None, valid (nimg,) i32, grad_xs, grad_ws,
    nits (nimg,) i32)``; gradients are of ``loss_per_img.sum()`` and must be scaled by the caller.
    """
    L = _lib.lib()
    if layout == _lib.KDOT_LAYOUT_CELL_MAJOR:
        B, D = xs.shape[1], xs.shape[2]
        sum_n, sum_m = xs.shape[0], xt.shape[0]
    else:
        B, D = xs.shape[0], xs.shape[2]
        sum_n, sum_m = xs.shape[1], xt.shape[1]
    _check_f32_cuda("xs", xs)
    _check_f32_cuda("xt", xt)
    if ws is not None:
        _check_f32_cuda("ws", ws, xs.shape[:2])
    if wt is not None:
        _check_f32_cuda("wt", wt, xt.shape[:2])
    nimg = len(pos_per_img)
    if len(pos_per_img_t) != nimg:
        raise ValueError("pos_per_img and pos_per_img_t must have one entry per image")
    if sum(pos_per_img) != sum_n or sum(pos_per_img_t) != sum_m:
        raise ValueError("per-image cell counts do not add up to the number of cells")
    dev = xs.device
    if cu_n is None or cu_m is None:
        # both prefix-sum arrays travel in ONE small H2D copy
        cu = np.zeros(2 * (nimg + 1), np.int32)
        np.cumsum(np.asarray(pos_per_img, np.int64), out=cu[1:nimg + 1])
        np.cumsum(np.asarray(pos_per_img_t, np.int64), out=cu[nimg + 2:])
        cu_d = torch.from_numpy(cu).to(dev, non_blocking=True)
        cu_n, cu_m = cu_d[:nimg + 1], cu_d[nimg + 1:]
    max_n = max(pos_per_img) if nimg else 0
    max_m = max(pos_per_img_t) if nimg else 0
    if sum_n == 0 or sum_m == 0:
        # No image has both a student and a teacher cell: the reference's loop skips every image
        # (losses/loss_libs.py:25-28) but has already normalised both key-point tensors in place (:8-12).
        # An empty tensor has no device address to hand to the C entry, so this case is settled here.
        if normalize:
            scale = torch.tensor([w, h], dtype=torch.float32, device=dev)
            for t in (xs, xt):
                if t.numel():
                    t.div_(scale)
        zf = lambda *shape: torch.zeros(*shape, dtype=torch.float32, device=dev)
        zi = torch.zeros(nimg, dtype=torch.int32, device=dev)
        return dict(loss_per_img=zf(nimg), loss_per_slot=zf(nimg, B) if want_slots else None, valid=zi,
                    grad_xs=torch.zeros_like(xs), grad_ws=zf(*xs.shape[:2]), nits=zi.clone())
    # outputs are views of two allocations (16-byte aligned segments) instead of six
    r4 = lambda v: (v + 3) & ~3
    n_gx, n_gw, n_l, n_s = xs.numel(), sum_n * B, nimg, (nimg * B if want_slots else 0)
    fbuf = torch.empty(r4(n_gx) + r4(n_gw) + r4(n_l) + r4(n_s), dtype=torch.float32, device=dev)
    ibuf = torch.empty(r4(nimg) * 2, dtype=torch.int32, device=dev)
    o = 0
    grad_xs = fbuf[o:o + n_gx].view(xs.shape); o += r4(n_gx)
    grad_ws = fbuf[o:o + n_gw].view(xs.shape[:2]); o += r4(n_gw)
    loss = fbuf[o:o + n_l]; o += r4(n_l)
    slots = fbuf[o:o + n_s].view(nimg, B) if want_slots else None
    valid, nits = ibuf[:nimg], ibuf[r4(nimg):r4(nimg) + nimg]
    if cfg.loss != "sinkhorn":
        kind = {"gaussian": 0, "laplacian": 1, "energy": 2}[cfg.loss]
        with _on_device(dev):
            rc = L.kdot_kernel_mmd_fwd_bwd(
                _ptr(xs), _ptr(ws), _ptr(xt), _ptr(wt), _ptr(cu_n), _ptr(cu_m), nimg, B, D, max_n, max_m, layout, kind,
                float(cfg.blur), float(w), float(h), 1 if normalize else 0,
                _ptr(loss), _ptr(slots), _ptr(valid), _ptr(grad_xs), _ptr(grad_ws),
                torch.cuda.current_stream(dev).cuda_stream)
        _lib.check(rc, "kdot_kernel_mmd_fwd_bwd")
        nits.zero_()
        return dict(loss_per_img=loss, loss_per_slot=slots, valid=valid, grad_xs=grad_xs, grad_ws=grad_ws, nits=nits)
    ws_bytes = int(L.kdot_workspace_bytes_ex(nimg, max_n, max_m, B, D, float(cfg.p)))
    wsp = _workspace(dev, ws_bytes)
    with _on_device(dev):
        rc = L.kdot_sinkhorn_fwd_bwd(
            _ptr(xs), _ptr(ws), _ptr(xt), _ptr(wt), _ptr(cu_n), _ptr(cu_m), nimg, B, D, max_n, max_m, layout,
            float(cfg.p), float(cfg.blur), -1.0 if cfg.reach is None else float(cfg.reach), float(cfg.scaling),
            float(w), float(h), 1 if normalize else 0,
            _ptr(loss), _ptr(slots), _ptr(valid), _ptr(grad_xs), _ptr(grad_ws), _ptr(nits),
            _ptr(wsp), ws_bytes, torch.cuda.current_stream(dev).cuda_stream,
        )
    _lib.check(rc, "kdot_sinkhorn_fwd_bwd")
    return dict(loss_per_img=loss, loss_per_slot=slots, valid=valid, grad_xs=grad_xs, grad_ws=grad_ws, nits=nits)


class OTLossFunction(torch.autograd.Function):
    """Autograd node around :func:`ot_loss_batched` (cell-major layout).

    ``forward`` returns ``loss_per_img (nimg,)``; the analytic gradients computed by the same launch are kept
    for ``backward``, which only scales them by the incoming per-image gradient.  ``xs`` is CONSUMED: when
    ``normalize`` it is overwritten with the normalised coordinates (the reference's in-place side effect,
    ``losses/loss_libs.py:8-12``) while the returned gradient is w.r.t. the values it held on entry.
    """

    @staticmethod
    def forward(ctx, xs, ws, xt, wt, pos_per_img, pos_per_img_t, cfg, w, h, normalize):
        out = ot_loss_batched(xs.detach(), None if ws is None else ws.detach().contiguous(), xt.detach(),
                              None if wt is None else wt.detach().contiguous(), pos_per_img, pos_per_img_t, cfg,
                              w=w, h=h, normalize=normalize)
        img_of_cell = torch.repeat_interleave(
            torch.arange(len(pos_per_img), device=xs.device),
            torch.as_tensor(pos_per_img, device=xs.device), output_size=xs.shape[0])
        ctx.save_for_backward(out["grad_xs"], out["grad_ws"], img_of_cell)
        ctx.has_ws = ws is not None
        ctx.mark_non_differentiable(out["valid"], out["nits"])
        return out["loss_per_img"], out["valid"], out["nits"]

    @staticmethod
    def backward(ctx, g_loss, _g_valid, _g_nits):
        grad_xs, grad_ws, img_of_cell = ctx.saved_tensors
        g = g_loss[img_of_cell]
        gx = grad_xs * g.view(-1, 1, 1) if ctx.needs_input_grad[0] else None
        gw = grad_ws * g.view(-1, 1) if (ctx.has_ws and ctx.needs_input_grad[1]) else None
        return gx, gw, None, None, None, None, None, None, None, None


class _NormalizeInPlace(torch.autograd.Function):
    """``xy[:, 0] /= w ; xy[:, 1] /= h`` in place with autograd support (``losses/loss_libs.py:8-12``)."""

    @staticmethod
    def forward(ctx, xy, scale):
        ctx.mark_dirty(xy)
        ctx.save_for_backward(scale)
        xy.div_(scale)
        return xy

    @staticmethod
    def backward(ctx, g):
        (scale,) = ctx.saved_tensors
        return g / scale, None


def normalize_in_place(xy: torch.Tensor, w: float, h: float) -> torch.Tensor:
    scale = torch.tensor([w, h], dtype=xy.dtype, device=xy.device)
    if xy.requires_grad:
        return _NormalizeInPlace.apply(xy, scale)
    return xy.div_(scale)


class GatherDecodeFunction(torch.autograd.Function):
    """``pred_reg`` levels -> decoded key-points of the positive cells, without flattening the head outputs
    (``kdot_gather_decode_fwd`` / ``_bwd``).

    ``apply(pos_inds, cls_label, anchors_pos, bbox_trans_pos, *pred_reg)`` returns ``(npos*8, 2)`` pixel key-points, the
    tensor the reference builds at ``losses/kd_loss.py:47-50`` from ``pred_reg_flatten[pos_inds]``.  Backward writes
    the 16 offset gradients of every positive cell straight into per-level gradients (views of ONE zero-filled
    allocation)."""

    @staticmethod
    def forward(ctx, pos_inds, cls_label, anchors_pos, bbox_trans_pos, *pred_reg):
        L = _lib.lib()
        nlvl = len(pred_reg)
        nimg, ch = pred_reg[0].shape[0], pred_reg[0].shape[1]
        if ch % 16:
            raise ValueError("pred_reg levels must have C*16 channels")
        dev = pred_reg[0].device
        levels = []
        for r in pred_reg:
            if r.dim() != 4 or r.shape[0] != nimg or r.shape[1] != ch:
                raise ValueError("inconsistent pred_reg level shapes")
            _check_f32_cuda("pred_reg level", r.detach() if r.is_contiguous() else r.detach().contiguous())
            levels.append(r.detach() if r.is_contiguous() else r.detach().contiguous())
        npos = int(pos_inds.shape[0])
        pos_inds = pos_inds.to(torch.int64).contiguous()
        cls_label = cls_label.to(torch.int64).contiguous()
        _check_f32_cuda("anchors", anchors_pos, (npos, 4))
        if bbox_trans_pos is not None:
            bbox_trans_pos = bbox_trans_pos.to(torch.float32).contiguous()
            _check_f32_cuda("bbox_trans", bbox_trans_pos, (npos, 2, 3))
        xy = torch.empty(npos * 8, 2, dtype=torch.float32, device=dev)
        hw = (C.c_int32 * nlvl)(*[int(r.shape[2] * r.shape[3]) for r in levels])
        ptrs = (C.c_void_p * nlvl)(*[r.data_ptr() for r in levels])
        with _on_device(dev):
            rc = L.kdot_gather_decode_fwd(ptrs, hw, nlvl, nimg, ch // 16, pos_inds.data_ptr(), cls_label.data_ptr(),
                                          anchors_pos.data_ptr(), _ptr(bbox_trans_pos), npos, xy.data_ptr(),
                                          torch.cuda.current_stream(dev).cuda_stream)
        _lib.check(rc, "kdot_gather_decode_fwd")
        ctx.save_for_backward(pos_inds, cls_label, anchors_pos, bbox_trans_pos)
        ctx.shapes = [tuple(r.shape) for r in levels]
        ctx.hw, ctx.nimg, ctx.ncls = hw, nimg, ch // 16
        return xy

    @staticmethod
    def backward(ctx, g_xy):
        L = _lib.lib()
        pos_inds, cls_label, anchors_pos, bbox_trans_pos = ctx.saved_tensors
        dev = g_xy.device
        npos = int(pos_inds.shape[0])
        g_xy = g_xy.contiguous()
        sizes = [int(np.prod(s)) for s in ctx.shapes]
        flat = torch.zeros(sum(sizes), dtype=torch.float32, device=dev)  # one memset for every level
        grads, o = [], 0
        for s, n in zip(ctx.shapes, sizes):
            grads.append(flat[o:o + n].view(s))
            o += n
        ptrs = (C.c_void_p * len(grads))(*[g.data_ptr() for g in grads])
        with _on_device(dev):
            rc = L.kdot_gather_decode_bwd(g_xy.data_ptr(), ctx.hw, len(grads), ctx.nimg, ctx.ncls, pos_inds.data_ptr(),
                                          cls_label.data_ptr(), anchors_pos.data_ptr(), _ptr(bbox_trans_pos), npos, ptrs,
                                          torch.cuda.current_stream(dev).cuda_stream)
        _lib.check(rc, "kdot_gather_decode_bwd")
        return (None, None, None, None, *grads)


def gather_decode(pred_reg, pos_inds, cls_label, anchors_pos, bbox_trans_pos=None):
    """Decoded ``(npos*8, 2)`` pixel key-points of the positive cells, gathered from the per-level head outputs."""
    return GatherDecodeFunction.apply(pos_inds, cls_label, anchors_pos.contiguous(), bbox_trans_pos, *pred_reg)


_focal_ws = {}


def _focal_workspace(dev):
    buf = _focal_ws.get(dev.index)
    if buf is None:
        buf = torch.zeros(int(_lib.lib().kdot_focal_workspace_bytes()), dtype=torch.uint8, device=dev)  # ticket starts at 0
        _focal_ws[dev.index] = buf
    return buf


class FocalLossFunction(torch.autograd.Function):
    """``SigmoidFocalLoss`` (reference ``losses/loss.py:12-40``) over every non-ignored (cell, class) pair, forward and
    gradient in ONE launch on the per-level ``(nimg, C, H, W)`` logits as the head produced them
    (``kdot_focal_loss_fwd_bwd``): no flatten of the class maps, no boolean-index copies of ``pred_cls_flatten[valid]``.

    ``apply(labels_flat, gamma, alpha, *pred_cls) -> 0-dim loss``; ``labels_flat (nimg * cells,) int64`` in the reference's
    label order."""

    @staticmethod
    def forward(ctx, labels_flat, gamma, alpha, *pred_cls):
        L = _lib.lib()
        levels = [c.detach() if c.is_contiguous() else c.detach().contiguous() for c in pred_cls]
        for c in levels:
            _check_f32_cuda("pred_cls level", c)
        nimg, ncls = levels[0].shape[0], levels[0].shape[1]
        dev = levels[0].device
        cells = sum(int(c.shape[2] * c.shape[3]) for c in levels)
        labels_flat = labels_flat.to(torch.int64).contiguous()
        if labels_flat.numel() != nimg * cells:
            raise ValueError("labels do not cover nimg * cells")
        need_grad = any(c.requires_grad for c in pred_cls)
        sizes = [c.numel() for c in levels]
        flat = torch.empty(sum(sizes) if need_grad else 0, dtype=torch.float32, device=dev)
        grads, o = [], 0
        for c, n in zip(levels, sizes):
            grads.append(flat[o:o + n].view(c.shape) if need_grad else None)
            o += n
        loss = torch.empty((), dtype=torch.float32, device=dev)
        nl = len(levels)
        hw = (C.c_int32 * nl)(*[int(c.shape[2] * c.shape[3]) for c in levels])
        ptrs = (C.c_void_p * nl)(*[c.data_ptr() for c in levels])
        gptrs = (C.c_void_p * nl)(*[g.data_ptr() for g in grads]) if need_grad else None
        wsp = _focal_workspace(dev)
        with _on_device(dev):
            rc = L.kdot_focal_loss_fwd_bwd(ptrs, hw, nl, nimg, ncls, labels_flat.data_ptr(), float(gamma), float(alpha),
                                           loss.data_ptr(), gptrs, wsp.data_ptr(), wsp.numel(),
                                           torch.cuda.current_stream(dev).cuda_stream)
        _lib.check(rc, "kdot_focal_loss_fwd_bwd")
        ctx.grads = grads
        return loss

    @staticmethod
    def backward(ctx, g):
        return (None, None, None, *[None if gr is None else gr * g for gr in ctx.grads])


class Reg3dLossFunction(torch.autograd.Function):
    """The 3-D object-space regression loss of ``losses/kd_loss.py:57-71`` on decoded key-points, forward and
    ``d/d(key-points)`` in one launch (``kdot_reg3d_loss_fwd_bwd``).  ``apply(pred_xy (n*8, 2), target_3d (n, 8, 3),
    diam_cell (n,), kinv (9 floats, host)) -> per-cell losses (n,)``."""

    @staticmethod
    def forward(ctx, pred_xy, target_3d, diam_cell, kinv):
        L = _lib.lib()
        xy = pred_xy.detach().contiguous()
        tgt = target_3d.detach().to(torch.float32).contiguous()
        diam = diam_cell.detach().to(torch.float32).contiguous()
        _check_f32_cuda("pred_xy", xy)
        n = diam.numel()
        if xy.shape != (n * 8, 2) or tgt.numel() != n * 24:
            raise ValueError("pred_xy must be (n*8, 2) and target_3d (n, 8, 3)")
        dev = xy.device
        out = torch.empty(n, dtype=torch.float32, device=dev)
        g_xy = torch.empty_like(xy)
        k9 = (C.c_float * 9)(*[float(v) for v in kinv])
        with _on_device(dev):
            rc = L.kdot_reg3d_loss_fwd_bwd(xy.data_ptr(), tgt.data_ptr(), diam.data_ptr(), k9, n, out.data_ptr(),
                                           g_xy.data_ptr(), torch.cuda.current_stream(dev).cuda_stream)
        _lib.check(rc, "kdot_reg3d_loss_fwd_bwd")
        ctx.save_for_backward(g_xy)
        return out

    @staticmethod
    def backward(ctx, g):
        (g_xy,) = ctx.saved_tensors
        return g_xy * g.repeat_interleave(8).view(-1, 1), None, None, None

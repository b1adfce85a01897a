"""Host-side operators over the C ABI: the fused OT distillation loss (forward + analytic backward).

``ot_loss_batched`` is the whole per-mini-batch OT section of the reference
(``/root/reference/losses/kd_loss.py:73-103`` -> ``losses/loss_libs.py:1-51`` -> geomloss) as ONE fused
launch; ``OTLossFunction`` plugs it into autograd.  PyTorch is used here for device memory, streams and
the autograd graph only -- all arithmetic runs in ``libkdot.so``.
"""
from __future__ import annotations

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
    if cu_n is None:
        cu_n = cu_seqlens(pos_per_img, dev)
    if cu_m is None:
        cu_m = cu_seqlens(pos_per_img_t, dev)
    max_n = max(pos_per_img) if nimg else 0
    max_m = max(pos_per_img_t) if nimg else 0
    loss = torch.empty(nimg, dtype=torch.float32, device=dev)
    slots = torch.empty(nimg, B, dtype=torch.float32, device=dev) if want_slots else None
    valid = torch.empty(nimg, dtype=torch.int32, device=dev)
    nits = torch.empty(nimg, dtype=torch.int32, device=dev)
    grad_xs = torch.empty_like(xs)
    grad_ws = torch.empty(xs.shape[:2], dtype=torch.float32, device=dev)
    if cfg.loss != "sinkhorn":
        kind = {"gaussian": 0, "laplacian": 1, "energy": 2}[cfg.loss]
        with torch.cuda.device(dev):
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
    with torch.cuda.device(dev):
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

"""``SamplesLoss`` -- call-compatible replacement of ``geomloss.SamplesLoss`` for the mode the reference uses.

Seam B2 of SURVEY.md section 8(b): constructed at ``/root/reference/losses/kd_loss.py:26-30`` as
``SamplesLoss(GTYPE, p=GP, blur=GBLUR, scaling=SCALING, reach=REACH)`` and called per image at
``losses/loss_libs.py:47`` with ``(alpha (8,N), x (8,N,2), beta (8,M), y (8,M,2)) -> (8,)``
(or ``(x, y)`` with uniform weights, ``:49``).  The call runs the fused CUDA kernel on the slot-major
layout directly (no transposes); gradients flow to ``x`` and ``alpha`` only, as in geomloss.
"""
from __future__ import annotations

import torch

from . import _lib
from .ops import OTConfig, ot_loss_batched


class _SamplesLossFunction(torch.autograd.Function):
    @staticmethod
    def forward(ctx, a, x, b, y, cfg):
        xd = x.detach().contiguous()
        yd = y.detach().contiguous()
        out = ot_loss_batched(
            xd, None if a is None else a.detach().contiguous(), yd, None if b is None else b.detach().contiguous(),
            [x.shape[1]], [y.shape[1]], cfg, normalize=False, layout=_lib.KDOT_LAYOUT_SLOT_MAJOR, want_slots=True)
        ctx.save_for_backward(out["grad_xs"], out["grad_ws"])
        ctx.has_a = a is not None
        ctx.mark_non_differentiable(out["valid"], out["nits"])
        return out["loss_per_slot"][0], out["valid"], out["nits"]

    @staticmethod
    def backward(ctx, g, _gv, _gn):
        gx, ga = ctx.saved_tensors
        return (ga * g.view(-1, 1) if (ctx.has_a and ctx.needs_input_grad[0]) else None,
                gx * g.view(-1, 1, 1) if ctx.needs_input_grad[1] else None, None, None, None)


class SamplesLoss(torch.nn.Module):
    """Debiased (un)balanced Sinkhorn divergence ``S_eps,rho(alpha, beta)`` between weighted point clouds.

    What the reference's flags can reach is implemented natively: ``loss="sinkhorn"`` (``p=2``, ``debias=True``,
    ``potentials=False``, tensorized semantics) and the kernel-MMD losses ``"gaussian"``, ``"laplacian"``, ``"energy"``.  Anything else raises
    ``NotImplementedError`` -- there is no fallback to a slower implementation.
    """

    def __init__(self, loss="sinkhorn", p=2, blur=0.05, reach=None, diameter=None, scaling=0.5, truncate=5,
                 cost=None, kernel=None, cluster_scale=None, debias=True, potentials=False, verbose=False,
                 backend="auto"):
        super().__init__()
        if loss not in ("sinkhorn", "gaussian", "laplacian", "energy"):
            raise NotImplementedError(f"SamplesLoss(loss={loss!r}): no CUDA kernel (sinkhorn, gaussian, laplacian, energy)")
        if loss == "sinkhorn" and float(p) not in (1.0, 2.0):
            raise NotImplementedError("SamplesLoss: p must be 1 or 2")
        if diameter is not None or cost is not None or kernel is not None or not debias or potentials:
            raise NotImplementedError("SamplesLoss: diameter/cost/kernel/debias=False/potentials are not supported")
        if backend not in ("auto", "tensorized"):
            raise NotImplementedError(f"SamplesLoss(backend={backend!r})")
        self.loss, self.p, self.blur, self.reach, self.scaling = loss, float(p), float(blur), reach, float(scaling)
        self.debias, self.potentials, self.backend = debias, potentials, backend
        self.last_nits = None
        self.last_valid = None

    @property
    def config(self) -> OTConfig:
        return OTConfig(p=self.p, blur=self.blur, scaling=self.scaling,
                        reach=None if self.reach is None else float(self.reach), loss=self.loss)

    def forward(self, *args):
        if len(args) == 4:
            a, x, b, y = args
        elif len(args) == 2:
            x, y = args
            a = b = None
        else:
            raise NotImplementedError("SamplesLoss: call as (alpha, x, beta, y) or (x, y)")
        if x.dim() != y.dim():
            raise ValueError("Input samples 'x' and 'y' should have the same number of dimensions.")
        if x.shape[-1] != y.shape[-1]:
            raise ValueError("Input samples 'x' and 'y' should have the same last dimension.")
        batched = x.dim() == 3
        if not batched:
            if x.dim() != 2:
                raise ValueError("Input samples 'x' and 'y' should be encoded as (N,D) or (B,N,D) (batch) tensors.")
            x, y = x.unsqueeze(0), y.unsqueeze(0)
            a = None if a is None else a.unsqueeze(0)
            b = None if b is None else b.unsqueeze(0)
        if x.shape[0] != y.shape[0]:
            raise ValueError("Samples 'x' and 'y' should have the same batchsize.")
        if a is not None and (a.shape != x.shape[:2] or b.shape != y.shape[:2]):
            raise ValueError("Weights and samples should have compatible shapes.")
        values, valid, nits = _SamplesLossFunction.apply(a, x, b, y, self.config)
        self.last_nits, self.last_valid = nits, valid
        return values if batched else values[0]

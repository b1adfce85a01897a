"""Op-by-op restatement of ``geomloss==0.2.4`` ``SamplesLoss`` (tensorized backend).

TEST INFRASTRUCTURE ONLY (see ``oracle/__init__.py``) -- **parity unpinned**: geomloss is
a third-party dependency of the reference (``/root/reference/requirements.txt:45``,
imported at ``losses/kd_loss.py:6``, constructed at ``losses/kd_loss.py:26-30`` and
called at ``losses/loss_libs.py:47`` (weighted) / ``:49`` (unweighted)).  It is not
vendored in ``/root/reference`` and not installable here, so this file restates its
published algorithm from the upstream module layout:

* ``geomloss/utils.py``               -> :func:`squared_distances`, :func:`distances`, :func:`scal`
* ``geomloss/sinkhorn_divergence.py`` -> :func:`max_diameter`, :func:`epsilon_schedule`,
  :func:`scaling_parameters`, :func:`dampening`, :func:`log_weights`,
  :func:`sinkhorn_loop`, :func:`sinkhorn_cost`
* ``geomloss/sinkhorn_samples.py``    -> :func:`softmin_tensorized`, :func:`sinkhorn_tensorized`
* ``geomloss/kernel_samples.py``      -> :func:`kernel_tensorized` (energy / gaussian / laplacian)
* ``geomloss/samples_loss.py``        -> :class:`SamplesLoss`

The code is dtype-generic: feed fp32 tensors for the "reference as shipped" numbers and
fp64 tensors for the high-precision arbitration oracle ("Parity
definition").  Every tensor op is issued in the same order as upstream so that an fp32
run reproduces the reference's own rounding behaviour as closely as a CPU can.
"""
from __future__ import annotations

import numpy as np
import torch

# ----------------------------------------------------------------------------------------
# geomloss/utils.py
# ----------------------------------------------------------------------------------------


def scal(a, f, batch=False):
    """<a, f>, per batch row when ``batch``."""
    if batch:
        nb = a.shape[0]
        return (a.reshape(nb, -1) * f.reshape(nb, -1)).sum(1)
    return torch.dot(a.reshape(-1), f.reshape(-1))


def squared_distances(x, y):
    """|x_i - y_j|^2 through the expansion |x|^2 - 2 x.y + |y|^2 (matmul), as upstream."""
    if x.dim() == 2:
        d_xx = (x * x).sum(-1).unsqueeze(1)
        d_xy = torch.matmul(x, y.permute(1, 0))
        d_yy = (y * y).sum(-1).unsqueeze(0)
    elif x.dim() == 3:
        d_xx = (x * x).sum(-1).unsqueeze(2)
        d_xy = torch.matmul(x, y.permute(0, 2, 1))
        d_yy = (y * y).sum(-1).unsqueeze(1)
    else:
        raise ValueError("Incompatible dimensions")
    return d_xx - 2 * d_xy + d_yy


def distances(x, y):
    return torch.sqrt(torch.clamp_min(squared_distances(x, y), 1e-8))


cost_routines = {
    1: lambda x, y: distances(x, y),
    2: lambda x, y: squared_distances(x, y) / 2,
}

# ----------------------------------------------------------------------------------------
# geomloss/sinkhorn_divergence.py
# ----------------------------------------------------------------------------------------


def dampening(eps, rho):
    return 1 if rho is None else 1 / (1 + eps / rho)


def log_weights(a):
    a_log = a.log()
    a_log[a <= 0] = -100000
    return a_log


def max_diameter(x, y):
    """Diagonal of the bounding box of the two (flattened) point clouds -> python float."""
    mins = torch.stack((x.min(dim=0)[0], y.min(dim=0)[0])).min(dim=0)[0]
    maxs = torch.stack((x.max(dim=0)[0], y.max(dim=0)[0])).max(dim=0)[0]
    return (maxs - mins).norm().item()


def epsilon_schedule(p, diameter, blur, scaling):
    """[diam^p] + exp(arange(p ln diam, p ln blur, p ln scaling)) + [blur^p]  (float64)."""
    return (
        [diameter ** p]
        + [float(np.exp(e)) for e in np.arange(p * np.log(diameter), p * np.log(blur), p * np.log(scaling))]
        + [blur ** p]
    )


def scaling_parameters(x, y, p, blur, reach, diameter, scaling):
    if diameter is None:
        d = x.shape[-1]
        diameter = max_diameter(x.reshape(-1, d), y.reshape(-1, d))
    eps = blur ** p
    eps_s = epsilon_schedule(p, diameter, blur, scaling)
    rho = None if reach is None else reach ** p
    return diameter, eps, eps_s, rho


def softmin_tensorized(eps, C, f):
    nb = C.shape[0]
    return -eps * (f.reshape(nb, 1, -1) - C / eps).logsumexp(2).reshape(nb, -1)


def sinkhorn_loop(softmin, a_log, b_log, C_xx, C_yy, C_xy, C_yx, eps_s, rho, debias=True, last_extrapolation=True):
    """Symmetrised eps-scaling Sinkhorn loop; returns (a_x, b_y, a_y, b_x).

    Naming as upstream: a_* are potentials produced FROM the measure alpha (supported on x),
    b_* FROM beta; the suffix says on which cloud the potential lives.  So ``b_x`` (N values) is
    the potential paired with alpha in <alpha, .> and ``a_y`` (M values) the one paired with beta.
    """
    prev = torch.is_grad_enabled()
    torch.set_grad_enabled(False)
    try:
        eps = eps_s[0]
        lam = dampening(eps, rho)
        if debias:
            a_x = lam * softmin(eps, C_xx, a_log)
            b_y = lam * softmin(eps, C_yy, b_log)
        a_y = lam * softmin(eps, C_yx, a_log)
        b_x = lam * softmin(eps, C_xy, b_log)

        for eps in eps_s:
            lam = dampening(eps, rho)
            if debias:
                at_x = lam * softmin(eps, C_xx, a_log + a_x / eps)
                bt_y = lam * softmin(eps, C_yy, b_log + b_y / eps)
            at_y = lam * softmin(eps, C_yx, a_log + b_x / eps)
            bt_x = lam * softmin(eps, C_xy, b_log + a_y / eps)
            if debias:
                a_x, b_y = 0.5 * (a_x + at_x), 0.5 * (b_y + bt_y)
            a_y, b_x = 0.5 * (a_y + at_y), 0.5 * (b_x + bt_x)
    finally:
        torch.set_grad_enabled(prev)

    if last_extrapolation:
        if debias:
            a_x = lam * softmin(eps, C_xx, (a_log + a_x / eps).detach())
            b_y = lam * softmin(eps, C_yy, (b_log + b_y / eps).detach())
        a_y, b_x = (
            lam * softmin(eps, C_yx, (a_log + b_x / eps).detach()),
            lam * softmin(eps, C_xy, (b_log + a_y / eps).detach()),
        )
    if debias:
        return a_x, b_y, a_y, b_x
    return None, None, a_y, b_x


def _unbalanced_weight(eps, rho, x):
    # upstream UnbalancedWeight is an nn.Module whose ``backward`` method autograd never calls:
    # the effective factor is (rho + eps/2) in forward AND backward.
    return (rho + eps / 2) * x


def sinkhorn_cost(eps, rho, a, b, f_aa, g_bb, g_ab, f_ba, batch=False, debias=True, potentials=False):
    if potentials:
        if debias:
            return f_ba - f_aa, g_ab - g_bb
        return f_ba, g_ab
    if debias:
        if rho is None:
            return scal(a, f_ba - f_aa, batch=batch) + scal(b, g_ab - g_bb, batch=batch)
        return scal(a, _unbalanced_weight(eps, rho, (-f_aa / rho).exp() - (-f_ba / rho).exp()), batch=batch) + scal(
            b, _unbalanced_weight(eps, rho, (-g_bb / rho).exp() - (-g_ab / rho).exp()), batch=batch
        )
    if rho is None:
        return scal(a, f_ba, batch=batch) + scal(b, g_ab, batch=batch)
    return scal(a, _unbalanced_weight(eps, rho, 1 - (-f_ba / rho).exp()), batch=batch) + scal(
        b, _unbalanced_weight(eps, rho, 1 - (-g_ab / rho).exp()), batch=batch
    )


# ----------------------------------------------------------------------------------------
# geomloss/sinkhorn_samples.py
# ----------------------------------------------------------------------------------------


def sinkhorn_tensorized(a, x, b, y, p=2, blur=0.05, reach=None, diameter=None, scaling=0.5, cost=None,
                        debias=True, potentials=False, return_nits=False, **kwargs):
    if cost is None:
        cost = cost_routines[p]
    C_xx, C_yy = (cost(x, x.detach()), cost(y, y.detach())) if debias else (None, None)
    C_xy, C_yx = cost(x, y.detach()), cost(y, x.detach())
    diameter, eps, eps_s, rho = scaling_parameters(x, y, p, blur, reach, diameter, scaling)
    a_x, b_y, a_y, b_x = sinkhorn_loop(
        softmin_tensorized, log_weights(a), log_weights(b), C_xx, C_yy, C_xy, C_yx, eps_s, rho, debias=debias
    )
    out = sinkhorn_cost(eps, rho, a, b, a_x, b_y, a_y, b_x, batch=True, debias=debias, potentials=potentials)
    if return_nits:
        return out, len(eps_s), diameter
    return out


# ----------------------------------------------------------------------------------------
# geomloss/kernel_samples.py  (other --gtype modes)
# ----------------------------------------------------------------------------------------


class _DoubleGrad(torch.autograd.Function):
    @staticmethod
    def forward(ctx, inp):
        return inp

    @staticmethod
    def backward(ctx, grad_output):
        return 2 * grad_output


def double_grad(x):
    return _DoubleGrad.apply(x)


def gaussian_kernel(x, y, blur=0.05):
    return (-0.5 * squared_distances(x / blur, y / blur)).exp()


def laplacian_kernel(x, y, blur=0.05):
    return (-distances(x / blur, y / blur)).exp()


def energy_kernel(x, y, blur=None):
    return -distances(x, y)


kernel_routines = {"gaussian": gaussian_kernel, "laplacian": laplacian_kernel, "energy": energy_kernel}


def kernel_tensorized(a, x, b, y, blur=0.05, kernel=None, name=None, potentials=False, **kwargs):
    if kernel is None:
        kernel = kernel_routines[name]
    K_xx = kernel(double_grad(x), x.detach(), blur=blur)
    K_yy = kernel(double_grad(y), y.detach(), blur=blur)
    K_xy = kernel(x, y, blur=blur)
    a_x = torch.matmul(K_xx, a.detach().unsqueeze(-1)).squeeze(-1)
    b_y = torch.matmul(K_yy, b.detach().unsqueeze(-1)).squeeze(-1)
    b_x = torch.matmul(K_xy, b.unsqueeze(-1)).squeeze(-1)
    if potentials:
        a_y = torch.matmul(K_xy.transpose(1, 2), a.unsqueeze(-1)).squeeze(-1)
        return a_x - b_x, b_y - a_y
    return 0.5 * (double_grad(a) * a_x).sum(1) + 0.5 * (double_grad(b) * b_y).sum(1) - (a * b_x).sum(1)


# ----------------------------------------------------------------------------------------
# geomloss/samples_loss.py
# ----------------------------------------------------------------------------------------


class SamplesLoss(torch.nn.Module):
    """Call-compatible stand-in for ``geomloss.SamplesLoss`` (tensorized backend only).

    ``SamplesLoss(loss, p, blur, scaling, reach)(alpha, x, beta, y)`` with ``alpha (B,N)``, ``x (B,N,D)``,
    ``beta (B,M)``, ``y (B,M,D)`` returns ``(B,)``; ``(x, y)`` alone uses uniform ``1/N`` weights.
    """

    def __init__(self, loss="sinkhorn", p=2, blur=0.05, reach=None, diameter=None, scaling=0.5, truncate=5,
                 cost=None, kernel=None, cluster_scale=None, debias=True, potentials=False, verbose=False,
                 backend="auto"):
        super().__init__()
        self.loss, self.p, self.blur, self.reach = loss, p, blur, reach
        self.diameter, self.scaling, self.cost, self.kernel = diameter, scaling, cost, kernel
        self.debias, self.potentials, self.backend = debias, potentials, backend
        self.last_nits = None
        if loss not in ("sinkhorn", "gaussian", "laplacian", "energy"):
            raise KeyError(loss)

    @staticmethod
    def _uniform(x):
        if x.dim() == 2:
            return torch.ones(x.shape[0]).type_as(x) / x.shape[0]
        if x.dim() == 3:
            return torch.ones(x.shape[0], x.shape[1]).type_as(x) / x.shape[1]
        raise ValueError("Input samples 'x' and 'y' should be encoded as (N,D) or (B,N,D) (batch) tensors.")

    def forward(self, *args):
        if len(args) == 4:
            a, x, b, y = args
        elif len(args) == 2:
            x, y = args
            a, b = self._uniform(x), self._uniform(y)
        else:
            raise NotImplementedError("labels / 6-argument form is out of scope of the oracle")
        if x.dim() != y.dim():
            raise ValueError("Input samples 'x' and 'y' should have the same number of dimensions.")
        if x.shape[-1] != y.shape[-1]:
            raise ValueError("Input samples 'x' and 'y' should have the same last dimension.")
        batched = x.dim() == 3
        if not batched:
            a, x, b, y = a.unsqueeze(0), x.unsqueeze(0), b.unsqueeze(0), y.unsqueeze(0)
        if a.shape != x.shape[:2] or b.shape != y.shape[:2]:
            raise ValueError("weights and samples have incompatible shapes")
        if x.shape[1] * y.shape[1] > 5000 ** 2:
            raise NotImplementedError("upstream would switch to the KeOps online backend here")
        if self.loss == "sinkhorn":
            values, nits, _ = sinkhorn_tensorized(
                a, x, b, y, p=self.p, blur=self.blur, reach=self.reach, diameter=self.diameter,
                scaling=self.scaling, cost=self.cost, debias=self.debias, potentials=self.potentials,
                return_nits=True,
            )
            self.last_nits = nits
        else:
            values = kernel_tensorized(a, x, b, y, blur=self.blur, kernel=self.kernel, name=self.loss,
                                       potentials=self.potentials)
        if self.potentials:
            return values
        return values if batched else values[0]

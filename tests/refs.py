"""CPU references for the parity tests: the fp32 restatement of the reference formulation (``ref32``: the
reference's own ``kd_loss_2d`` loop restated in ``oracle/kd_loss_ref.py`` driving ``oracle/geomloss_ref.py``
with autograd) and the fp64 analytic oracle (``ref64``: ``oracle/sinkhorn_analytic.py``)."""
import numpy as np
import torch

from oracle import geomloss_ref, kd_loss_ref, sinkhorn_analytic
from kd_6d_pose_adlp_b200.synthetic import cu_seqlens


def ref32(batch, blur=0.001, reach=0.5, scaling=0.5, normalize=True, weighted=True, dtype=torch.float32):
    """Returns (loss_per_img (nimg,), grad_xs (sumN,8,2), grad_ws (sumN,8), xs_norm, nits)."""
    torch.backends.cuda.matmul.allow_tf32 = False
    B = batch["xs"].shape[1]
    xs = torch.tensor(batch["xs"].reshape(-1, 2), dtype=dtype, requires_grad=True)
    ws = torch.tensor(batch["ws"], dtype=dtype, requires_grad=True)
    xt = torch.tensor(batch["xt"].reshape(-1, 2), dtype=dtype)
    wt = torch.tensor(batch["wt"], dtype=dtype)
    L = geomloss_ref.SamplesLoss("sinkhorn", p=2.0, blur=blur, scaling=scaling, reach=reach)
    xs_work = xs.clone()
    assert B == 8
    losses = kd_loss_ref.kd_loss_2d_ref(xs_work, xt, ws if weighted else None, wt if weighted else None, 640, 480,
                                        "point", L, dim=2, pos_per_img=batch["pos_per_img"],
                                        pos_per_img_t=batch["pos_per_img_t"], normalize=normalize)
    nimg = len(batch["pos_per_img"])
    out = np.zeros(nimg)
    keep = [i for i in range(nimg) if batch["pos_per_img"][i] > 0 and batch["pos_per_img_t"][i] > 0]
    if losses:
        sum(losses).backward()
        out[keep] = torch.stack(losses).detach().double().numpy()
    gx = xs.grad.double().numpy().reshape(-1, B, 2) if xs.grad is not None else np.zeros((xs.shape[0] // B, B, 2))
    gw = ws.grad.double().numpy() if ws.grad is not None else np.zeros(tuple(ws.shape))
    return out, gx, gw, xs_work.detach().numpy().reshape(-1, B, 2), xt.numpy().reshape(-1, B, 2)


def ref64(batch, blur=0.001, reach=0.5, scaling=0.5, normalize=True, weighted=True):
    B, D = batch["xs"].shape[1], batch["xs"].shape[2]
    return sinkhorn_analytic.kdot_fwd_bwd_f64(
        batch["xs"], batch["ws"] if weighted else None, batch["xt"], batch["wt"] if weighted else None,
        cu_seqlens(batch["pos_per_img"]), cu_seqlens(batch["pos_per_img_t"]), B, D, blur=blur, reach=reach,
        scaling=scaling, normalize=normalize)

"""Kernel-MMD sample losses (--gtype gaussian / laplacian / energy) against the geomloss restatement."""
import numpy as np
import pytest
import torch

from kd_6d_pose_adlp_b200.synthetic import ot_batch
from tests import parity

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("loss,blur", [("gaussian", 0.05), ("gaussian", 0.01), ("laplacian", 0.05), ("energy", 0.05)])
def test_samples_loss_seam(loss, blur):
    from kd_6d_pose_adlp_b200 import SamplesLoss
    from oracle import geomloss_ref

    dev = torch.device("cuda:0")
    g = torch.Generator().manual_seed(3)
    for B, N, M, D in [(8, 10, 12, 2), (2, 70, 50, 2), (1, 40, 30, 16)]:
        x = (0.5 + 0.05 * torch.randn(B, N, D, generator=g)).float()
        y = (0.5 + 0.05 * torch.randn(B, M, D, generator=g)).float()
        a = torch.rand(B, N, generator=g) * 0.9 + 0.05
        b = torch.rand(B, M, generator=g) * 0.9 + 0.05
        xd, ad = x.to(dev).requires_grad_(True), a.to(dev).requires_grad_(True)
        out = SamplesLoss(loss, blur=blur)(ad, xd, b.to(dev), y.to(dev))
        wsum = torch.linspace(0.5, 1.5, B, device=dev)
        (out * wsum).sum().backward()
        refs = {}
        for name, dt in (("ref32", torch.float32), ("ref64", torch.float64)):
            xr, ar = x.clone().to(dt).requires_grad_(True), a.clone().to(dt).requires_grad_(True)
            o = geomloss_ref.SamplesLoss(loss, blur=blur)(ar, xr, b.to(dt), y.to(dt))
            (o * wsum.cpu().to(dt)).sum().backward()
            refs[name] = (o.detach().double().numpy(), xr.grad.double().numpy(), ar.grad.double().numpy())
        rows = [parity.report("loss", out.detach().cpu().numpy(), refs["ref32"][0], refs["ref64"][0]),
                parity.report("grad_x", xd.grad.cpu().numpy(), refs["ref32"][1], refs["ref64"][1]),
                parity.report("grad_a", ad.grad.cpu().numpy(), refs["ref32"][2], refs["ref64"][2])]
        print("\n", loss, blur, (B, N, M, D), "\n" + parity.fmt(rows))
        assert all(r["ok"] for r in rows), parity.fmt(rows)


def test_kd_loss_2d_with_kernel_loss_matches_reference_driver():
    """The drop-in kd_loss_2d with a kernel loss vs the restated reference driver (in-place normalise, skips)."""
    from kd_6d_pose_adlp_b200 import SamplesLoss
    from kd_6d_pose_adlp_b200.losses.loss_libs import kd_loss_2d
    from oracle import geomloss_ref, kd_loss_ref

    dev = torch.device("cuda:0")
    b = ot_batch(6, seed=12, p_empty_teacher=0.3)
    pred = torch.from_numpy(b["xs"].reshape(-1, 2)).to(dev).requires_grad_(True)
    s_cls = torch.from_numpy(b["ws"]).to(dev).requires_grad_(True)
    work = pred * 1.0
    tgt = torch.from_numpy(b["xt"].reshape(-1, 2)).to(dev)
    losses = kd_loss_2d(work, tgt, s_cls, torch.from_numpy(b["wt"]).to(dev), 640, 480, "point",
                        SamplesLoss("energy", blur=0.05), dim=2, pos_per_img=b["pos_per_img"], pos_per_img_t=b["pos_per_img_t"])
    (sum(losses) / len(losses)).backward()
    p64 = torch.tensor(b["xs"].reshape(-1, 2), dtype=torch.float64, requires_grad=True)
    w64 = torch.tensor(b["ws"], dtype=torch.float64, requires_grad=True)
    ref = kd_loss_ref.kd_loss_2d_ref(p64 * 1.0, torch.tensor(b["xt"].reshape(-1, 2), dtype=torch.float64), w64,
                                     torch.tensor(b["wt"], dtype=torch.float64), 640, 480, "point",
                                     geomloss_ref.SamplesLoss("energy", blur=0.05), dim=2,
                                     pos_per_img=b["pos_per_img"], pos_per_img_t=b["pos_per_img_t"])
    (sum(ref) / len(ref)).backward()
    assert len(losses) == len(ref)
    assert parity.rel(torch.stack(losses).detach().cpu().numpy(), torch.stack(ref).detach().numpy()) < 1e-5
    assert parity.rel(pred.grad.cpu().numpy(), p64.grad.numpy()) < 1e-4
    assert parity.rel(s_cls.grad.cpu().numpy(), w64.grad.numpy()) < 1e-5
    assert np.allclose(work.detach().cpu().numpy(), b["xs"].reshape(-1, 2) / np.array([640, 480], np.float32))

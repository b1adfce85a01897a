"""The fused focal / 3-D regression losses against the reference's op sequences (restated in tests/doubles.py and in
kd_loss._object_space_reg_loss): value and gradients."""
import numpy as np
import pytest
import torch

from tests import doubles, scenario

pytestmark = pytest.mark.gpu
HW = [(32, 32), (16, 16), (8, 8), (4, 4)]


def test_focal_loss_matches_the_reference_op_sequence():
    from kd_6d_pose_adlp_b200.losses.kd_loss import flatten_level_list
    from kd_6d_pose_adlp_b200.ops import FocalLossFunction

    dev = torch.device("cuda:0")
    g = torch.Generator().manual_seed(1)
    nimg, cells = 5, sum(h * w for h, w in HW)
    cls = [(torch.randn(nimg, 15, h, w, generator=g) * 3 - 2).to(dev) for h, w in HW]
    cls[0][0, 0, 0, :4] = torch.tensor([-30.0, 30.0, -9.3, 9.3])          # clamp region: sigmoid outside [1e-4, 1 - 1e-4]
    labels = torch.zeros(nimg * cells, dtype=torch.int64)
    sel = torch.randperm(nimg * cells, generator=g)
    labels[sel[:60]] = torch.randint(1, 16, (60,), generator=g)           # positives of several classes
    labels[sel[60:400]] = -1                                              # ignored cells
    labels = labels.to(dev)
    a = [c.clone().requires_grad_(True) for c in cls]
    b = [c.clone().requires_grad_(True) for c in cls]
    loss = FocalLossFunction.apply(labels, 2.0, 0.25, *a)
    (loss * 0.1).backward()
    flat = flatten_level_list(b)
    valid = torch.nonzero(labels >= 0).squeeze(1)
    ref = doubles.FocalLoss(2.0, 0.25)(flat[valid], labels[valid])        # reference losses/loss.py:12-40, kd_loss.py:133-134
    (ref * 0.1).backward()
    assert abs(float(loss) - float(ref)) <= 2e-6 * abs(float(ref)), (float(loss), float(ref))
    for x, y in zip(a, b):
        assert torch.equal(x.grad == 0, y.grad == 0) or (x.grad - y.grad).abs().max() < 1e-7
        assert (x.grad - y.grad).abs().max() <= 2e-6 * y.grad.abs().max()
    # generic gamma path
    a2 = [c.clone().requires_grad_(True) for c in cls]
    b2 = [c.clone().requires_grad_(True) for c in cls]
    l2 = FocalLossFunction.apply(labels, 1.5, 0.4, *a2)
    r2 = doubles.FocalLoss(1.5, 0.4)(flatten_level_list(b2)[valid], labels[valid])
    l2.backward(); r2.backward()
    assert abs(float(l2) - float(r2)) <= 5e-6 * abs(float(r2))
    assert max((x.grad - y.grad).abs().max() / y.grad.abs().max() for x, y in zip(a2, b2)) < 1e-5


def test_reg3d_loss_matches_the_reference_op_sequence():
    from kd_6d_pose_adlp_b200.ops import Reg3dLossFunction

    dev = torch.device("cuda:0")
    g = torch.Generator().manual_seed(2)
    n = 37
    K = torch.tensor(scenario.INTERNAL_K).view(3, 3)
    diam_all = torch.tensor(scenario.MESH_DIAMETERS)
    cls = torch.randint(0, 15, (n,), generator=g)
    tgt = (torch.randn(n, 8, 3, generator=g) * 40 + torch.tensor([0.0, 0.0, 900.0]))
    proj = (K @ tgt.view(-1, 3).t()).t()
    xy0 = proj[:, :2] / proj[:, 2:3] + torch.randn(n * 8, 2, generator=g) * torch.tensor([0.3, 3.0])  # both SmoothL1 branches
    xa = xy0.to(dev).requires_grad_(True)
    xb = xy0.to(dev).requires_grad_(True)
    wts = torch.rand(n, generator=g).to(dev)
    kinv = torch.inverse(K.double()).reshape(-1).tolist()
    la = Reg3dLossFunction.apply(xa, tgt.to(dev), diam_all[cls].to(dev), kinv)
    (la * wts).sum().backward()
    # reference op sequence (kd_loss.py:57-71)
    diam = diam_all.to(dev)[cls.to(dev).view(-1, 1).repeat(1, 24).view(-1, 3, 1)]
    homog = torch.cat((xb.t(), torch.ones_like(xb[:, 0]).view(1, -1)), dim=0)
    ray = torch.inverse(K.to(dev)).mm(homog).t()
    P = torch.bmm(ray.view(-1, 3, 1), ray.view(-1, 1, 3)) / torch.bmm(ray.view(-1, 1, 3), ray.view(-1, 3, 1))
    t3 = tgt.to(dev).view(-1, 3, 1)
    lb = torch.nn.SmoothL1Loss(reduction="none")(50 * torch.bmm(P, t3) / diam, 50 * t3 / diam).view(n, -1).mean(dim=1) / 50
    (lb * wts).sum().backward()
    assert (la - lb).abs().max() <= 2e-5 * lb.abs().max(), ((la - lb).abs().max(), lb.abs().max())
    assert (xa.grad - xb.grad).abs().max() <= 2e-4 * xb.grad.abs().max()

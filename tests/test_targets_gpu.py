"""Device-side SSC target assignment (kdot_ssc_*; reference PoseLossDzi.prepare_targets, losses/loss.py:164-268) against
the labels the reference's own prepare_targets produced (tests/golden/kd_pose_loss.npz) and distribution properties of the
on-device draw."""
import os
import types

import numpy as np
import pytest
import torch

from tests import scenario

pytestmark = pytest.mark.gpu
GOLDEN = os.path.join(os.path.dirname(__file__), "golden", "kd_pose_loss.npz")
S_HW = [(32, 32), (16, 16), (8, 8), (4, 4)]


def _targets(nimg, seed, dev):
    arr = scenario.make_target_arrays(nimg, seed)
    t = lambda a: torch.tensor(a).to(dev)
    return [types.SimpleNamespace(keypoints_3d=t(arr["keypoints_3d"]), K=t(arr["K"]), mask=t(arr["mask"][i]),
                                  class_ids=t(arr["class_ids"][i]), rotations=t(arr["rotations"][i]),
                                  translations=t(arr["translations"][i]), bbox_trans=t(arr["bbox_trans"][i]))
            for i in range(nimg)], arr


def _anchors(dev):
    from kd_6d_pose_adlp_b200.target_coder import grid_anchors

    return torch.cat(grid_anchors(S_HW, scenario.ANCHOR_SIZES, scenario.ANCHOR_STRIDES, device=dev), dim=0)


def test_parity_mode_reproduces_the_reference_labels_bit_exactly():
    from kd_6d_pose_adlp_b200.targets import positives_aux, ssc_assign

    z = np.load(GOLDEN)
    nimg, seed = int(z["nimg"]), int(z["seed"])
    dev = torch.device("cuda:0")
    targets, _arr = _targets(nimg, seed, dev)
    torch.manual_seed(seed + 7)      # the seed in front of the golden run's prepare_targets (make_golden.py)
    res = ssc_assign(targets, _anchors(dev), [h * w for h, w in S_HW], scenario.ANCHOR_SIZES, 10, 1.0, mode="parity")
    labels = res["labels"].cpu().numpy()
    assert np.array_equal(labels, z["labels"]), "labels differ from the reference's prepare_targets"
    assert res["npos"].cpu().tolist() == z["pos_per_img"].tolist()
    pos = torch.nonzero(res["labels"] > 0).squeeze(1)
    cls_label, aux_3d, bt = positives_aux(res, pos)
    p = pos.cpu().numpy()
    assert (cls_label.cpu().numpy() == z["labels"][p] - 1).all()
    assert np.abs(aux_3d.cpu().numpy() - z["aux_3d"][p]).max() < 1e-3          # mm, ~900 mm away: 1e-6 relative
    assert np.array_equal(bt.cpu().numpy(), z["aux_bbox_trans"][p])


def test_philox_mode_draws_uniformly_without_replacement():
    from kd_6d_pose_adlp_b200.targets import ssc_assign

    dev = torch.device("cuda:0")
    nimg = 8
    targets, _arr = _targets(nimg, 3, dev)
    anchors = _anchors(dev)
    hw = [h * w for h, w in S_HW]
    off = np.cumsum([0] + hw)
    runs = []
    for seed in range(200):
        res = ssc_assign(targets, anchors, hw, scenario.ANCHOR_SIZES, 10, 1.0, mode="philox", seed=seed)
        runs.append(res["labels"].view(nimg, -1).cpu().numpy())
    gtid = res["gtid"].cpu().numpy()
    cnt, nk = res["count"].cpu().numpy(), res["nk"].cpu().numpy()
    assert 9 <= nk[:, :, 0].sum(1).min() and nk[:, :, 0].sum(1).max() <= 11       # the budget of loss.py:213-215
    freq = np.zeros_like(runs[0], dtype=np.float64)
    for lab in runs:
        assert ((lab > 0) <= (gtid > 0)).all()                    # positives only inside the object's mask
        assert ((lab == -1) == ((gtid > 0) & (lab <= 0))).all()   # in-mask cells that were not drawn are ignored, the rest is background
        for i in range(nimg):
            for l in range(4):
                sel = (lab[i, off[l]:off[l + 1]] > 0).sum()
                assert sel == min(nk[i, l, 0], cnt[i, l, 0])      # exactly min(budget, candidates): a draw WITHOUT replacement
        freq += lab > 0
    assert any((runs[0] != r).any() for r in runs[1:])            # the seed matters
    # uniformity: every candidate of a (image, level) is drawn with probability k / n
    for i in range(nimg):
        for l in range(2):                                         # the two large levels carry enough candidates
            n, k = cnt[i, l, 0], min(nk[i, l, 0], cnt[i, l, 0])
            if n < 8 or k == 0 or k == n:
                continue
            f = freq[i, off[l]:off[l + 1]][gtid[i, off[l]:off[l + 1]] > 0] / len(runs)
            p = k / n
            assert abs(f.mean() - p) < 1e-9                        # exact by construction (k cells per run)
            assert np.abs(f - p).max() < 6 * np.sqrt(p * (1 - p) / len(runs)) + 0.02, (i, l, n, k, f.min(), f.max())


def test_kd_pose_loss_with_device_targets_matches_the_golden(monkeypatch):
    """KDPoseLoss.__call__ with cfg_kd['DEVICE_TARGETS'] = 'parity': no recorded assignment is replayed, the labels come
    from the device kernels -- and the three losses equal the reference's."""
    from kd_6d_pose_adlp_b200.losses.kd_loss import make_kd_pose_loss
    from kd_6d_pose_adlp_b200.target_coder import TargetCoder, grid_anchors
    from tests import doubles

    z = np.load(GOLDEN)
    nimg, seed = int(z["nimg"]), int(z["seed"])
    dev = torch.device("cuda:0")
    targets, _arr = _targets(nimg, seed, dev)
    s_cls, s_reg = scenario.make_head_outputs(nimg, S_HW, seed + 200, teacher=False, target_seed=seed)
    doubles.ReplayBase.recorded = None                      # any use of the base class' prepare_targets would fail
    KDPoseLoss = make_kd_pose_loss(doubles.ReplayBase)
    loss_fn = KDPoseLoss(2.0, 0.25, scenario.ANCHOR_SIZES, scenario.ANCHOR_STRIDES, "SSC", 10, 1.0, 9,
                         scenario.INTERNAL_K, scenario.MESH_DIAMETERS,
                         TargetCoder("POINT", scenario.ANCHOR_SIZES, scenario.ANCHOR_STRIDES, target_type="3D"),
                         dict(scenario.CFG_KD, DEVICE_TARGETS="parity"))
    lv = grid_anchors(S_HW, scenario.ANCHOR_SIZES, scenario.ANCHOR_STRIDES, device=dev)
    pred_t = {"post_kp_2d": torch.from_numpy(z["post_kp_2d"]).to(dev), "post_kp_cls": torch.from_numpy(z["post_kp_cls"]).to(dev),
              "post_pos_per_img": z["post_pos_per_img"].tolist()}
    pc = [torch.from_numpy(a).to(dev).requires_grad_(True) for a in s_cls]
    pr = [torch.from_numpy(a).to(dev).requires_grad_(True) for a in s_reg]
    torch.manual_seed(seed + 7)
    cls_loss, reg_loss, kd_loss = loss_fn(pc, pr, targets, [lv for _ in range(nimg)], pred_t)
    assert loss_fn.pos_per_img == z["pos_per_img"].tolist()
    for got, key, tol in ((cls_loss, "cls_loss", 2e-5), (reg_loss, "reg_loss", 2e-5), (kd_loss, "kd_loss", 1e-4)):
        assert abs(float(got.detach()) - float(z[key])) <= tol * abs(float(z[key])), (key, float(got.detach()), float(z[key]))
    (cls_loss * 0.1 + reg_loss + 5.0 * kd_loss).backward()
    assert all(torch.isfinite(t.grad).all() for t in pc + pr)

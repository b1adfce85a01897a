"""Seam B0: KDPoseLoss.__call__ (our drop-in on a fixture-replaying base) against the reference's own
KDPoseLoss run in the authoring container (tests/golden/kd_pose_loss.npz)."""
import os

import numpy as np
import pytest
import torch

from tests import doubles, scenario

pytestmark = pytest.mark.gpu
GOLDEN = os.path.join(os.path.dirname(__file__), "golden", "kd_pose_loss.npz")
S_HW = [(32, 32), (16, 16), (8, 8), (4, 4)]


def _dense(z, prefix, l, shape):
    g = np.zeros(int(np.prod(shape)), np.float32)
    g[z[f"{prefix}_{l}_idx"]] = z[f"{prefix}_{l}_val"]
    return g.reshape(shape)


def test_losses_and_gradients_match_reference():
    from kd_6d_pose_adlp_b200.losses.kd_loss import make_kd_pose_loss
    from kd_6d_pose_adlp_b200.target_coder import TargetCoder, grid_anchors

    z = np.load(GOLDEN)
    nimg, seed = int(z["nimg"]), int(z["seed"])
    s_cls, s_reg = scenario.make_head_outputs(nimg, S_HW, seed + 200, teacher=False, target_seed=seed)
    assert doubles.digest(s_cls + s_reg) == str(z["inputs_sha256"])
    dev = torch.device("cuda:0")
    cells = z["cells_per_img"].tolist()
    split = lambda a: list(torch.split(torch.from_numpy(a).to(dev), cells))
    doubles.ReplayBase.recorded = dict(labels=split(z["labels"]), reg_targets=split(z["reg_targets"]),
                                       aux_raw_boxes=split(z["aux_raw_boxes"]), aux_3d=split(z["aux_3d"]),
                                       aux_bbox_trans=split(z["aux_bbox_trans"]))
    KDPoseLoss = make_kd_pose_loss(doubles.ReplayBase)
    cfg = dict(scenario.CFG_KD)
    loss_fn = KDPoseLoss(2.0, 0.25, scenario.ANCHOR_SIZES, scenario.ANCHOR_STRIDES, "SSC", 10, 1.0, 9,
                         scenario.INTERNAL_K, scenario.MESH_DIAMETERS,
                         TargetCoder("POINT", scenario.ANCHOR_SIZES, scenario.ANCHOR_STRIDES, target_type="3D"), cfg)
    lv_anchors = grid_anchors(S_HW, scenario.ANCHOR_SIZES, scenario.ANCHOR_STRIDES, device=dev)
    anchors = [lv_anchors for _ in range(nimg)]
    pred_t = {"post_kp_2d": torch.from_numpy(z["post_kp_2d"]).to(dev), "post_kp_cls": torch.from_numpy(z["post_kp_cls"]).to(dev),
              "post_pos_per_img": z["post_pos_per_img"].tolist()}
    pc = [torch.from_numpy(a).to(dev).requires_grad_(True) for a in s_cls]
    pr = [torch.from_numpy(a).to(dev).requires_grad_(True) for a in s_reg]
    cls_loss, reg_loss, kd_loss = loss_fn(pc, pr, None, anchors, pred_t)
    assert loss_fn.pos_per_img == z["pos_per_img"].tolist()
    assert abs(float(cls_loss) - float(z["cls_loss"])) <= 2e-5 * abs(float(z["cls_loss"]))
    assert abs(float(reg_loss) - float(z["reg_loss"])) <= 2e-5 * abs(float(z["reg_loss"]))
    assert abs(float(kd_loss) - float(z["kd_loss"])) <= 1e-4 * abs(float(z["kd_loss"])), (float(kd_loss), float(z["kd_loss"]))
    # teacher key-points were normalised in place, like the reference does (loss_libs.py:10-11)
    want = z["post_kp_2d"] / np.array([640.0, 480.0], np.float32)
    assert np.array_equal(pred_t["post_kp_2d"].cpu().numpy(), want.astype(np.float32))

    g = torch.autograd.grad(kd_loss, pc + pr, allow_unused=True, retain_graph=True)
    g_all = torch.autograd.grad(cls_loss * 0.1 + reg_loss + 5.0 * kd_loss, pc + pr, allow_unused=True)
    for l in range(4):
        ref_c = _dense(z, "gkd_cls", l, s_cls[l].shape)
        ref_r = _dense(z, "gkd_reg", l, s_reg[l].shape)
        got_c = np.zeros_like(ref_c) if g[l] is None else g[l].cpu().numpy()
        got_r = np.zeros_like(ref_r) if g[4 + l] is None else g[4 + l].cpu().numpy()
        # same sparsity pattern: gradient only on the positive cells' class-0 logit / 16 offsets
        assert np.array_equal(np.flatnonzero(got_c), np.flatnonzero(ref_c))
        assert np.array_equal(np.flatnonzero(got_r), np.flatnonzero(ref_r))
        if np.abs(ref_c).max() > 0:
            assert np.abs(got_c - ref_c).max() <= 1e-4 * np.abs(ref_c).max()
        if np.abs(ref_r).max() > 0:
            # d/d(offsets) inherits the fp32 noise of the REFERENCE's own d/dx (2e-4..6e-3 rel. at eps = 1e-6,
            # SURVEY.md fact 3): the OT-boundary tests arbitrate that tensor against the fp64 oracle.
            assert np.abs(got_r - ref_r).max() <= 1e-2 * np.abs(ref_r).max()
    # the weighted total of the three losses, as train_kd.py combines them: focal-loss gradient on every class logit,
    # regression + distillation gradient on the positive cells' offsets (through the fused gather/decode epilogue)
    for l in range(4):
        ref_r = _dense(z, "gall_reg", l, s_reg[l].shape)
        got_r = np.zeros_like(ref_r) if g_all[4 + l] is None else g_all[4 + l].cpu().numpy()
        assert np.array_equal(np.flatnonzero(got_r), np.flatnonzero(ref_r))
        if np.abs(ref_r).max() > 0:
            assert np.abs(got_r - ref_r).max() <= 1e-2 * np.abs(ref_r).max()
        got_c = g_all[l].cpu().numpy().astype(np.float64)
        assert abs(got_c.sum() - z["gall_cls_sum"][l]) <= 1e-4 * z["gall_cls_abs"][l]
        assert abs(np.abs(got_c).sum() - z["gall_cls_abs"][l]) <= 1e-4 * z["gall_cls_abs"][l]
    assert loss_fn.step == 1 if hasattr(loss_fn, "step") else True

"""The drop-in on the reference's OWN module graph (GPU box): ``INTEGRATION.md`` section 2's import swap applied to the
reference's ``models/model_kd.py`` (staged unmodified under ``baseline/_ref`` by ``baseline/install_reference.py``), then
its ``PoseModuleKD.forward`` -- teacher branch -> our ``PostProcessorKD``; student branch -> our ``KDPoseLoss`` deriving from
the reference's real ``losses.loss.PoseLossDzi`` with its real ``prepare_targets`` (CPU ``torch.randperm`` stream) and the
reference's own ``AnchorGenerator`` / ``TargetCoder`` / ``PoseAnnot`` -- against the goldens the reference produced on CPU
(``tests/golden/make_golden.py``).  Backbone / FPN / head are replaced by a stub that returns the seeded synthetic head
outputs of ``tests/scenario.py``: truly random-init heads never pass the teacher's 0.1 confidence threshold."""
import importlib
import os
import sys

import numpy as np
import pytest
import torch

from tests import doubles, scenario

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.path.join(ROOT, "baseline", "_ref")
GOLD = os.path.join(os.path.dirname(__file__), "golden")
T_HW = [(32, 32), (16, 16), (8, 8), (4, 4), (2, 2)]
S_HW = T_HW[:4]


class _Fn(torch.nn.Module):
    """Stands in for backbone / FPN / head: a parameter-free module around a callable."""

    def __init__(self, fn):
        super().__init__()
        self.fn = fn

    def forward(self, x):
        return self.fn(x)


def _stub_network(model, cls_l, reg_l):
    model.backbone = _Fn(lambda x: x)
    model.fpn = _Fn(lambda x: [torch.zeros(c.shape[0], 1, c.shape[2], c.shape[3], device=c.device) for c in cls_l])
    model.head = _Fn(lambda feats: (cls_l, reg_l))


class _Images:
    def __init__(self, n, dev):
        self.tensors = torch.zeros(n, 3, 256, 256, device=dev)
        self.sizes = [(256, 256)] * n
        self.image_sizes = self.sizes


@pytest.mark.skipif(not os.path.isdir(os.path.join(REF, "losses")), reason="baseline/_ref not staged (baseline/install_reference.py)")
def test_reference_pose_module_kd_forward_with_the_import_swap(tmp_path, monkeypatch):
    monkeypatch.setenv("KDOT_REFERENCE_ROOT", REF)
    from oracle import ref_loader   # test infrastructure: stubs for trimesh / pyrender / matplotlib / geomloss

    ref_loader.REFERENCE_ROOT = REF
    ref = ref_loader.load()
    model_kd = importlib.import_module("models.model_kd")          # the reference's file, unmodified
    argument = importlib.import_module("arguments.argument")
    import yaml

    import kd_6d_pose_adlp_b200.losses.kd_loss as our_kd
    from kd_6d_pose_adlp_b200.postprocess.postprocess_kd import PostProcessor, PostProcessorKD

    # ---- the three swapped imports of models/model_kd.py:8-11 ----
    monkeypatch.setattr(model_kd, "PostProcessorKD", PostProcessorKD)
    monkeypatch.setattr(model_kd, "PostProcessor", PostProcessor)
    monkeypatch.setattr(model_kd, "KDPoseLoss", our_kd.KDPoseLoss)   # resolves the reference's losses.loss.PoseLossDzi
    assert issubclass(our_kd.KDPoseLoss, ref.PoseLossDzi)

    def make_cfg(backbone):
        with open(os.path.join(REF, "configs", "ape.yaml")) as fh:
            cfg = yaml.load(fh, Loader=yaml.FullLoader)
        cfg["MODEL"]["BACKBONE"] = backbone
        cfg = argument.custom_cfg(cfg)
        cfg["DATASETS"]["SYMMETRY_TYPES"] = {}
        cfg["KD"] = dict(scenario.CFG_KD, vis_dir=str(tmp_path))
        return cfg

    dev = torch.device("cuda:0")
    zt = np.load(os.path.join(GOLD, "postprocess_kd.npz"))
    zs = np.load(os.path.join(GOLD, "kd_pose_loss.npz"))
    nimg, seed = int(zt["nimg"]), int(zt["seed"])
    t_cls, t_reg = scenario.make_head_outputs(nimg, T_HW, seed + 100, teacher=True, target_seed=seed)
    s_cls, s_reg = scenario.make_head_outputs(nimg, S_HW, seed + 200, teacher=False, target_seed=seed)
    assert doubles.digest(t_cls + t_reg) == str(zt["inputs_sha256"]) and doubles.digest(s_cls + s_reg) == str(zs["inputs_sha256"])
    tarr = scenario.make_target_arrays(nimg, seed)
    targets = []
    for i in range(nimg):
        t = ref.PoseAnnot(torch.tensor(tarr["keypoints_3d"]), torch.tensor(tarr["K"]), torch.tensor(tarr["mask"][i]),
                          torch.tensor(tarr["class_ids"][i]), torch.tensor(tarr["rotations"][i]),
                          torch.tensor(tarr["translations"][i]), 256, 256,
                          bbox_scale=torch.tensor(1.0), bbox_trans=torch.tensor(tarr["bbox_trans"][i]))
        targets.append(t.to(dev))
    images = _Images(nimg, dev)

    # ---- teacher: reference forward (eval, is_teacher=True) -> post_processor_t = our PostProcessorKD ----
    teacher = model_kd.PoseModuleKD(make_cfg("darknet53"), torch.nn.Identity()).to(dev).eval()
    _stub_network(teacher, [torch.from_numpy(a).to(dev) for a in t_cls], [torch.from_numpy(a).to(dev) for a in t_reg])
    with torch.no_grad():
        pred_t = teacher(images, targets, is_teacher=True, cfg_kd=scenario.CFG_KD)
    assert pred_t["post_pos_per_img"] == zt["post_pos_per_img"].tolist()
    assert np.abs(pred_t["post_kp_cls"].cpu().numpy() - zt["post_kp_cls"]).max() < 2e-6
    assert np.abs(pred_t["post_kp_2d"].cpu().numpy() - zt["post_kp_2d"]).max() < 2e-3

    # ---- student: reference forward (train) -> loss_evaluator = our KDPoseLoss on the reference's PoseLossDzi ----
    student = model_kd.PoseModuleKD(make_cfg("darknet_tiny_h"), torch.nn.Identity()).to(dev).train()
    assert type(student.loss_evaluator).__mro__[1] is ref.PoseLossDzi
    pc = [torch.from_numpy(a).to(dev).requires_grad_(True) for a in s_cls]
    pr = [torch.from_numpy(a).to(dev).requires_grad_(True) for a in s_reg]
    _stub_network(student, pc, pr)
    # the reference's debug plots (kd_loss.py:88-97) stay wired: record the call instead of plotting (matplotlib is a stub here)
    vis_calls = []
    visualizer = importlib.import_module("tools.visualizer")
    monkeypatch.setattr(visualizer, "vis_pxpy_post_train_weight",
                        lambda d1, d2, c1, c2, nstep, **kw: vis_calls.append((d1.detach().clone(), d2.detach().clone(), c1.shape, c2.shape, nstep, kw)))
    torch.manual_seed(seed + 7)            # the seed the golden run used in front of prepare_targets' randperm calls
    _none, losses = student(images, targets, pred_t=pred_t, cfg_kd=scenario.CFG_KD)
    ev = student.loss_evaluator
    assert [int(v) for v in ev.pos_per_img] == zs["pos_per_img"].tolist()          # the REAL prepare_targets ran
    for key, gold, tol in (("loss_cls", "cls_loss", 2e-5), ("loss_reg", "reg_loss", 2e-5), ("loss_kd", "kd_loss", 1e-4)):
        assert abs(float(losses[key]) - float(zs[gold])) <= tol * abs(float(zs[gold])), (key, float(losses[key]), float(zs[gold]))
    total = losses["loss_cls"] * 0.1 + losses["loss_reg"] + 5.0 * losses["loss_kd"]   # train_kd.py:125-135
    total.backward()
    for l in range(4):
        ref_r = np.zeros(int(np.prod(s_reg[l].shape)), np.float32)
        ref_r[zs[f"gall_reg_{l}_idx"]] = zs[f"gall_reg_{l}_val"]
        got = pr[l].grad.cpu().numpy().reshape(-1)
        assert np.array_equal(np.flatnonzero(got), np.flatnonzero(ref_r))
        if np.abs(ref_r).max() > 0:
            assert np.abs(got - ref_r).max() <= 5e-3 * np.abs(ref_r).max()
    assert ev.step == 1
    # visualiser: called once at step 0 with the IN-PLACE-NORMALISED key-points (loss_libs.py:8-12), masses as (n*8, 1)
    assert len(vis_calls) == 1 and vis_calls[0][4] == 0
    d1, d2, c1s, c2s, _n, kw = vis_calls[0]
    assert float(d1.max()) < 1.5 and float(d2.max()) < 1.5 and d1.shape[0] == 8 * sum(ev.pos_per_img) and c1s == (d1.shape[0], 1)
    assert kw["pos_per_img_1"] == [int(v) for v in ev.pos_per_img] and kw["pos_per_img_2"] == pred_t["post_pos_per_img"]
    assert len(kw["loss"]) == nimg and kw["save_dir"].endswith("/vis")

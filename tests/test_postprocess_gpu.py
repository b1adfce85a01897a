"""Teacher cell selection (kdot_select_cells + PostProcessorKD mirror) against the fixture produced by the
reference's own PostProcessorKD: selected (level, cell) indices bit-exact, float outputs within fp32 noise."""
import os

import numpy as np
import pytest
import torch

from tests import doubles, scenario

pytestmark = pytest.mark.gpu
GOLDEN = os.path.join(os.path.dirname(__file__), "golden", "postprocess_kd.npz")
T_HW = [(32, 32), (16, 16), (8, 8), (4, 4), (2, 2)]


def _setup():
    from kd_6d_pose_adlp_b200.postprocess.postprocess_kd import PostProcessorKD
    from kd_6d_pose_adlp_b200.target_coder import TargetCoder

    z = np.load(GOLDEN)
    nimg, seed = int(z["nimg"]), int(z["seed"])
    t_cls, t_reg = scenario.make_head_outputs(nimg, T_HW, seed + 100, teacher=True, target_seed=seed)
    assert doubles.digest(t_cls + t_reg) == str(z["inputs_sha256"]), "synthetic inputs differ from the fixture's"
    dev = torch.device("cuda:0")
    targets = [doubles.Target(torch.tensor(z["K"]), torch.tensor(z["keypoints_3d"]), torch.tensor(z["bbox_trans"][i]))
               for i in range(nimg)]
    pp = PostProcessorKD(0.1, TargetCoder("POINT", scenario.ANCHOR_SIZES, scenario.ANCHOR_STRIDES), 10, 1.0, {})
    res = pp([torch.from_numpy(a).to(dev) for a in t_cls], [torch.from_numpy(a).to(dev) for a in t_reg], targets, None)
    return z, pp, res, t_cls


def test_selected_cells_bit_exact_and_outputs_close():
    z, pp, res, t_cls = _setup()
    pos_t = [len(r) for r in res[0]]
    assert pos_t == z["post_pos_per_img"].tolist()
    sel = pp.last_selection
    lv = np.concatenate([sel["level"][i, 0, :n] for i, n in enumerate(pos_t)])
    loc = np.concatenate([sel["loc"][i, 0, :n] for i, n in enumerate(pos_t)])
    assert np.array_equal(lv, z["sel_level"]), "selected levels differ from the reference"
    assert np.array_equal(loc, z["sel_loc"]), "selected cell indices differ from the reference"
    scores = torch.cat(res[0]).cpu().numpy()
    xy = torch.cat(res[3]).cpu().numpy()
    assert scores.shape == z["post_kp_cls"].shape and xy.shape == z["post_kp_2d"].shape
    assert np.abs(scores - z["post_kp_cls"]).max() < 2e-6           # sqrt(sigmoid(logit)), <= 2 ulp apart
    assert np.abs(xy - z["post_kp_2d"]).max() < 2e-3                # full-image pixels (~300 px): <= 1e-5 rel
    # other classes have no candidate above the 0.1 threshold
    assert sel["count"][:, 1:].sum() == 0 and (sel["best"][:, 1:] == -1).all()


def test_budget_and_argmax_against_numpy_restatement():
    """nk, per-level candidate counts and the best cell recomputed in numpy from the same logits."""
    z, pp, res, t_cls = _setup()
    sel = pp.last_selection
    sizes = np.asarray(scenario.ANCHOR_SIZES, np.float32)
    for i in range(int(z["nimg"])):
        logit = [a[i, 0].reshape(-1) for a in t_cls]
        sig = [1.0 / (1.0 + np.exp(-l.astype(np.float64))) for l in logit]
        cnt = [int((s > 0.1).sum()) for s in sig]
        assert sel["valid"][i, 0].tolist() == cnt
        # best cell: highest score over levels, earlier level wins ties
        best_l = int(np.argmax([s.max() for s in sig]))
        assert sel["best"][i, 0, 0] == best_l and sel["best"][i, 0, 1] == int(np.argmax(logit[best_l]))
        assert 9 <= sel["nk"][i, 0].sum() <= 11
        # top-k per level = the k largest logits of that level, descending
        o = 0
        for l in range(5):
            k = min(cnt[l], int(sel["nk"][i, 0, l]))
            want = np.argsort(-logit[l], kind="stable")[:k]
            got = sel["loc"][i, 0, o:o + k]
            assert (sel["level"][i, 0, o:o + k] == l).all() and np.array_equal(got, want)
            o += k
        assert o == sel["count"][i, 0]


def test_empty_image_gives_empty_tensors():
    from kd_6d_pose_adlp_b200.postprocess.postprocess_kd import PostProcessorKD, teacher_knowledge
    from kd_6d_pose_adlp_b200.target_coder import TargetCoder

    dev = torch.device("cuda:0")
    cls = [torch.full((2, 15, h, w), -8.0, device=dev) for h, w in T_HW]
    reg = [torch.zeros(2, 240, h, w, device=dev) for h, w in T_HW]
    z = np.load(GOLDEN)
    tg = [doubles.Target(torch.tensor(z["K"]), torch.tensor(z["keypoints_3d"]), torch.tensor(z["bbox_trans"][0]))] * 2
    pp = PostProcessorKD(0.1, TargetCoder("POINT", scenario.ANCHOR_SIZES, scenario.ANCHOR_STRIDES), 10, 1.0, {})
    pt = teacher_knowledge(pp, cls, reg, tg)
    assert pt["post_pos_per_img"] == [0, 0] and pt["post_kp_2d"].shape == (0, 8, 2) and pt["post_kp_cls"].shape == (0, 8)


def test_student_eval_postprocessor_same_selection():
    """PostProcessor (student evaluation, postprocess/postprocess.py) shares the selection kernel: same cells, all
    labels of the target kept, score = max over the selected cells."""
    from kd_6d_pose_adlp_b200.postprocess.postprocess_kd import PostProcessor
    from kd_6d_pose_adlp_b200.target_coder import TargetCoder

    z, pp_kd, res_kd, t_cls = _setup()
    dev = torch.device("cuda:0")
    nimg, seed = int(z["nimg"]), int(z["seed"])
    t_cls, t_reg = scenario.make_head_outputs(nimg, T_HW, seed + 100, teacher=True, target_seed=seed)
    targets = []
    for i in range(nimg):
        t = doubles.Target(torch.tensor(z["K"]), torch.tensor(z["keypoints_3d"]), torch.tensor(z["bbox_trans"][i]))
        t.class_ids = torch.tensor([0])
        targets.append(t)
    pp = PostProcessor(0.1, TargetCoder("POINT", scenario.ANCHOR_SIZES, scenario.ANCHOR_STRIDES), 10, 1.0, {})
    out = pp([torch.from_numpy(a).to(dev) for a in t_cls], [torch.from_numpy(a).to(dev) for a in t_reg], targets, None)
    assert len(out) == nimg
    for i in range(nimg):
        assert len(out[i]) == 1 and out[i][0][1] == 0
        score, cls_id, R, T, xy2d = out[i][0]
        assert abs(score - float(res_kd[0][i].max())) < 1e-6
        assert np.allclose(xy2d.numpy(), res_kd[3][i].cpu().numpy(), atol=1e-4)
        assert R.shape == (3, 3) and T.shape == (3, 1)
    # a target that does not contain class 0 yields nothing for that image
    targets[0].class_ids = torch.tensor([3])
    out = pp([torch.from_numpy(a).to(dev) for a in t_cls], [torch.from_numpy(a).to(dev) for a in t_reg], targets, None)
    assert out[0] == [] and len(out[1]) == 1

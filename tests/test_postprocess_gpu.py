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


# ---------------------------------------------------------------------------------------------------------
# hard scene (tests/golden/postprocess_hard.npz): produced by the reference's PostProcessorKD.pose_infer_ml for EVERY
# candidate label and by the reference's student-eval PostProcessor (tests/golden/make_golden.py --hard-only)
# ---------------------------------------------------------------------------------------------------------
HARD = os.path.join(os.path.dirname(__file__), "golden", "postprocess_hard.npz")


def _hard_inputs():
    z = np.load(HARD)
    nimg, seed = int(z["nimg"]), int(z["seed"])
    tarr, t_cls, t_reg = scenario.make_hard_scene(nimg, T_HW, seed)
    assert doubles.digest(t_cls + t_reg) == str(z["inputs_sha256"]), "synthetic inputs differ from the fixture's"
    dev = torch.device("cuda:0")
    return z, nimg, [torch.from_numpy(a).to(dev) for a in t_cls], [torch.from_numpy(a).to(dev) for a in t_reg]


def test_hard_scene_every_label_bit_exact():
    """Batch 64, 2-3 live classes per image, object scales 0.7x..2.6x (all budget patterns), exact duplicate logits
    (arg-max and top-k ties) and logits within 2 ulps of the 0.1 threshold: candidate counts per level, the number of
    selected cells and the (level, cell) sequence of every (image, label) the reference produced a pose for."""
    from kd_6d_pose_adlp_b200.postprocess.postprocess_kd import select_cells

    z, nimg, cls, reg = _hard_inputs()
    sel = select_cells(cls, reg, scenario.ANCHOR_SIZES, scenario.ANCHOR_STRIDES, 0.1, 10, 1.0)
    torch.cuda.synchronize()
    cap, ncls = sel["cap"], sel["ncls"]
    count = sel["sel_count"].cpu().numpy().reshape(nimg, ncls)
    valid = sel["valid_cnt"].cpu().numpy().reshape(nimg, ncls, -1)
    level = sel["sel_level"].cpu().numpy().reshape(nimg, ncls, cap)
    loc = sel["sel_loc"].cpu().numpy().reshape(nimg, ncls, cap)
    score = sel["sel_score"].cpu().numpy().reshape(nimg, ncls, cap)
    o = 0
    labels_seen = set()
    planes = [t.cpu().numpy().reshape(nimg, ncls, -1) for t in cls]
    n_unique = n_tied = n_best_tie = 0
    for i, c, n, v in zip(z["img"], z["label"], z["count"], z["valid"]):
        labels_seen.add(int(c))
        assert valid[i, c].tolist() == v.tolist(), ("candidates per level", i, c)          # threshold edge cells included
        if i % 4 == 1 and c == 0 and count[i, c] != n:
            # every 4th image carries an exact duplicate of its best class-0 logit: WHICH of the two tied cells the
            # reference's argmax-after-unsorted-topk (postprocess_kd.py:42,135) returns is implementation-defined, and
            # the box size -- hence the per-level budget -- is taken from that cell.  The kernel takes the lower index
            # (torch.argmax's documented first-occurrence rule on the natural order); test_budget_and_argmax_against_
            # numpy_restatement pins that rule.
            n_best_tie += 1
            o += n
            continue
        assert count[i, c] == n, ("selected cells", i, c, count[i, c], n)
        assert np.array_equal(level[i, c, :n], z["level"][o:o + n]), ("levels", i, c)
        assert np.abs(score[i, c, :n] - z["score"][o:o + n]).max() < 2e-6
        # Cells: identical wherever the logit is unique in its level plane.  Among EXACTLY tied logits the reference's
        # order is whatever its two-stage top-k yields (topk(sorted=False) over the candidates, then topk again,
        # postprocess_kd.py:42,154 -- implementation-defined, and different between torch's CPU and CUDA top-k); the
        # kernel's rule is "lower cell index first".  There the selected LOGIT VALUES must still be the same sequence.
        for k in range(n):
            lv, mine, theirs = int(level[i, c, k]), int(loc[i, c, k]), int(z["loc"][o + k])
            plane = planes[lv][i, c]
            assert plane[mine] == plane[theirs], ("selected value", i, c, k)
            if (plane == plane[theirs]).sum() == 1:
                assert mine == theirs, ("cell", i, c, k, mine, theirs)
                n_unique += 1
            else:
                n_tied += 1
        o += n
    assert o == len(z["loc"]) and labels_seen == {0, 3, 5, 7}
    assert n_unique > 1000 and n_tied > 40 and n_best_tie <= 8, (n_unique, n_tied, n_best_tie)   # the fixture really contains ties


def test_hard_scene_teacher_and_student_postprocessors():
    """The drop-in post-processors on the hard scene: PostProcessorKD keeps the first label's cells (count per image as
    in the reference run), PostProcessor applies the target.class_ids filter and returns the reference's poses."""
    from kd_6d_pose_adlp_b200.postprocess.postprocess_kd import PostProcessor, PostProcessorKD
    from kd_6d_pose_adlp_b200.target_coder import TargetCoder

    z, nimg, cls, reg = _hard_inputs()
    targets = []
    for i in range(nimg):
        t = doubles.Target(torch.tensor(z["K"]), torch.tensor(z["keypoints_3d"]), torch.tensor(z["bbox_trans"][i]))
        t.class_ids = torch.tensor([c for c in scenario.hard_live_classes(i) if c in (0, 3)])
        targets.append(t)
    coder = TargetCoder("POINT", scenario.ANCHOR_SIZES, scenario.ANCHOR_STRIDES)
    res = PostProcessorKD(0.1, coder, 10, 1.0, {})(cls, reg, targets, None)
    assert [len(r) for r in res[0]] == z["first_label_count"].tolist()
    out = PostProcessor(0.1, coder, 10, 1.0, {})(cls, reg, targets, None)
    got = [(i, int(r[1]), float(r[0]), np.asarray(r[2]), np.asarray(r[3]).reshape(3)) for i, lst in enumerate(out) for r in lst]
    assert [(g[0], g[1]) for g in got] == list(zip(z["ev_img"].tolist(), z["ev_cls"].tolist()))   # same (image, class) list
    assert set(z["ev_cls"].tolist()) == {0, 3}                                                       # 5 and 7 were filtered
    dR, dT = [], []
    for g, s, R, T in zip(got, z["ev_score"], z["ev_R"], z["ev_T"]):
        assert abs(g[2] - s) < 2e-6
        dR.append(np.abs(g[3] - R).max())
        dT.append(np.abs(g[4] - T).max() / np.abs(T).max())
    # RANSAC-EPnP (cv2, host, as in the reference): identical cell sequences give the same minimal sets and the same pose
    # to solver precision; where exactly tied logits are ordered differently (see the test above) the 80 key-points
    # arrive permuted, RANSAC samples other minimal sets and the pose moves within the ~1 px noise of the key-points
    assert np.median(dR) < 1e-4 and np.median(dT) < 1e-4, (np.median(dR), np.median(dT))
    assert max(dR) < 0.1 and max(dT) < 0.1, (max(dR), max(dT))
    assert sum(d > 1e-3 for d in dR) <= len(dR) // 4, "only the entries with re-ordered ties may move"

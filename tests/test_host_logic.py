"""Host-side logic that must work (and fail loudly) without a GPU."""
import numpy as np
import pytest
import torch

from kd_6d_pose_adlp_b200 import SamplesLoss
from kd_6d_pose_adlp_b200.losses.kd_loss import flatten_head_outputs, make_kd_pose_loss
from kd_6d_pose_adlp_b200.losses.loss_libs import kd_loss_2d
from kd_6d_pose_adlp_b200.ops import OTConfig, ot_loss_batched
from kd_6d_pose_adlp_b200.synthetic import cu_seqlens, ot_batch
from kd_6d_pose_adlp_b200.target_coder import TargetCoder, grid_anchors


def test_samples_loss_rejects_what_has_no_kernel():
    with pytest.raises(NotImplementedError):
        SamplesLoss("l1")                       # --gtype l1 / l2 are not geomloss losses either
    assert SamplesLoss("gaussian", blur=0.05).config.loss == "gaussian"
    with pytest.raises(NotImplementedError):
        SamplesLoss("sinkhorn", p=3)
    with pytest.raises(NotImplementedError):
        SamplesLoss("sinkhorn", p=2, debias=False)
    L = SamplesLoss("sinkhorn", p=2.0, blur=0.001, scaling=0.5, reach=0.5)
    assert L.config == OTConfig(2.0, 0.001, 0.5, 0.5, "sinkhorn")
    with pytest.raises(ValueError):
        L(torch.zeros(8, 3, 2), torch.zeros(8, 3))            # dims differ
    with pytest.raises(ValueError):
        L(torch.zeros(8, 3), torch.zeros(8, 3, 2), torch.zeros(8, 4), torch.zeros(7, 4, 2))


def test_no_cpu_fallback():
    b = ot_batch(2, seed=0)
    t = {k: torch.from_numpy(b[k]) for k in ("xs", "ws", "xt", "wt")}
    with pytest.raises(ValueError, match="no CPU fallback"):
        ot_loss_batched(t["xs"], t["ws"], t["xt"], t["wt"], b["pos_per_img"], b["pos_per_img_t"])
    with pytest.raises(TypeError):
        kd_loss_2d(t["xs"].view(-1, 2), t["xt"].view(-1, 2), t["ws"], t["wt"], 640, 480, "point", object(), 2,
                   b["pos_per_img"], b["pos_per_img_t"])
    with pytest.raises(NotImplementedError):
        kd_loss_2d(t["xs"].view(-1, 2), t["xt"].view(-1, 2), t["ws"], t["wt"], 640, 480, "instance",
                   SamplesLoss("sinkhorn", p=2), 2, b["pos_per_img"], b["pos_per_img_t"])


def test_synthetic_generator_and_prefix_sums():
    b = ot_batch(16, seed=5)
    assert b["xs"].shape == (sum(b["pos_per_img"]), 8, 2) and b["wt"].shape == (sum(b["pos_per_img_t"]), 8)
    assert (b["ws"][:, :1] == b["ws"]).all()  # masses shared by the 8 keypoint slots of a cell
    cu = cu_seqlens(b["pos_per_img"])
    assert cu.dtype == np.int32 and cu[0] == 0 and cu[-1] == b["xs"].shape[0]
    d = ot_batch(2, seed=0, dense=(1360, 1364))
    assert d["xs"].shape[0] == 2720 and d["xt"].shape[0] == 2728


def test_flatten_head_outputs_order():
    g = torch.Generator().manual_seed(0)
    cls = [torch.randn(2, 15, h, h, generator=g) for h in (4, 2)]
    reg = [torch.randn(2, 240, h, h, generator=g) for h in (4, 2)]
    c, r = flatten_head_outputs(cls, reg)
    assert c.shape == (2 * 20, 15) and r.shape == (2 * 20, 240)
    # image 1, level 1 (offset 16), cell (h=1, w=0) -> row 20 + 16 + 2
    assert torch.equal(c[20 + 16 + 2], cls[1][1, :, 1, 0]) and torch.equal(r[3], reg[0][0, :, 0, 3])


def test_decode_and_anchor_grid():
    anchors = grid_anchors([(32, 32), (2, 2)], [32, 512], [8, 128])
    assert anchors[0].shape == (1024, 4)
    a = anchors[0][33]  # h=1, w=1 -> centre (12, 12), side 32
    assert torch.allclose(a, torch.tensor([12 - 15.5, 12 - 15.5, 12 + 15.5, 12 + 15.5]))
    coder = TargetCoder("POINT", [32], [8])
    pred = torch.zeros(1, 16)
    pred[0, 0], pred[0, 8] = 0.5, -0.25
    out = coder.decode(pred, a.view(1, 4))
    assert float(out[0, 0]) == 12 + 16.0 and float(out[0, 8]) == 12 - 8.0
    bt = torch.tensor([[[2.0, 0.0, 10.0], [0.0, 2.0, -6.0]]])
    out2 = coder.decode(pred, a.view(1, 4), bt)
    assert torch.allclose(out2[0, 0], torch.tensor((28.0 - 10.0) / 2)) and torch.allclose(out2[0, 8], torch.tensor((4.0 + 6.0) / 2))


def test_kd_pose_loss_class_factory_keeps_reference_signature():
    import inspect

    class Base:
        def __init__(self, *a):
            self.args = a

    cls = make_kd_pose_loss(Base)
    sig = inspect.signature(cls.__init__)
    assert list(sig.parameters)[1:] == ["gamma", "alpha", "anchor_sizes", "anchor_strides", "positive_type",
                                        "positive_num", "positive_lambda", "top_k", "internal_K", "diameters",
                                        "target_coder", "cfg_kd"]
    assert list(inspect.signature(cls.__call__).parameters)[1:] == ["pred_cls", "pred_reg", "targets", "anchors", "pred_t"]
    cfg = dict(GTYPE="sinkhorn", GP=2.0, GBLUR=0.001, SCALING=0.5, REACH=0.5, WEIGHTED_OT=True, DETACH=False,
               GnD=2, GLEVEL="point")
    obj = cls(2.0, 0.25, [32], [8], "SSC", 10, 1.0, 9, [1] * 9, [1.0], None, cfg)
    assert isinstance(obj.kd_loss, SamplesLoss) and obj.weighted_ot and not obj.wot_detach and not hasattr(obj, "step")


def test_pnp_thread_pool_gives_the_serial_results():
    """Host part of PostProcessorKD / PostProcessor (postprocess_kd.py:158-203): the per-image RANSAC-EPnP calls on
    the thread pool return exactly what the reference's serial loop returns, in image order."""
    pytest.importorskip("cv2")
    from kd_6d_pose_adlp_b200.postprocess.postprocess_kd import _SelectingPostProcessor, _prepare_pnp_tasks, _solve_image
    from tests import doubles, scenario

    rng = np.random.default_rng(5)
    nimg, ncls, cap, n = 12, 15, 16, 10
    K = np.asarray(scenario.INTERNAL_K, np.float32).reshape(3, 3)
    box = np.array([[sx, sy, sz] for sx in (-40, 40) for sy in (-35, 35) for sz in (-45, 45)], np.float32)
    sel = dict(count=np.zeros((nimg, ncls), np.int32), valid=np.zeros((nimg, ncls, 5), np.int32),
               score=np.zeros((nimg, ncls, cap), np.float32), kpts=np.zeros((nimg, ncls, cap, 16), np.float32))
    targets = []
    for i in range(nimg):
        T = np.array([rng.normal(0, 60), rng.normal(0, 40), 900 + rng.normal(0, 50)], np.float32)
        cam = box + T
        uv = (K @ cam.T).T
        uv = uv[:, :2] / uv[:, 2:3]
        s, cx, cy = 1.6, float(uv[:, 0].mean()), float(uv[:, 1].mean())
        bt = np.array([[s, 0, 128 - s * cx], [0, s, 128 - s * cy]], np.float32)
        crop = uv @ bt[:, :2].T + bt[:, 2]
        if i != 3:  # image 3 selects nothing
            sel["count"][i, 0] = n
            sel["valid"][i, 0, 1] = n
            noisy = crop[None] + rng.normal(0, 0.7, (n, 8, 2)).astype(np.float32)
            sel["kpts"][i, 0, :n] = np.concatenate([noisy[:, :, 0], noisy[:, :, 1]], axis=1)
            sel["score"][i, 0, :n] = rng.uniform(0.4, 0.9, n)
        targets.append(doubles.Target(torch.from_numpy(K), [torch.from_numpy(box)] * ncls, torch.from_numpy(bt)))

    pp = _SelectingPostProcessor(0.1, None, 10, 1.0, None)
    tasks = _prepare_pnp_tasks(sel, targets, ncls)
    fn = lambda i: (_solve_image(tasks[i], first_only=True) or [None])[0]
    pp.pnp_threads = 1
    serial = pp._map_images(fn, nimg)
    pp.pnp_threads = 6
    pooled = pp._map_images(fn, nimg)
    assert serial[3] is None and pooled[3] is None
    assert sum(r is not None for r in serial) == nimg - 1
    for a, b in zip(serial, pooled):
        assert (a is None) == (b is None)
        if a is not None:
            assert a[0] == b[0] and np.array_equal(a[1], b[1]) and torch.equal(a[2], b[2])
            assert np.array_equal(a[3], b[3]) and np.array_equal(a[4], b[4])
            assert np.abs(a[4].reshape(-1)[2] - 900) < 120  # a sane pose, not a degenerate RANSAC result

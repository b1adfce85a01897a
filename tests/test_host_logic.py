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

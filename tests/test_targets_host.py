"""Host-side packing of the per-image target objects (``kd_6d_pose_adlp_b200.targets``) on CPU tensors: images with different
object counts (including none) go through the gather into the padded ``(nimg, maxgt, ...)`` layout, equal counts through the
view-only fast path -- both against a naive per-image loop; and ``positives_aux`` picks the right (image, object) rows."""
import types

import numpy as np
import pytest
import torch

from kd_6d_pose_adlp_b200 import targets as T


def _make(ngt, shared_tables=True, seed=0):
    g = torch.Generator().manual_seed(seed)
    ncls = 5
    kp_table = torch.randn(ncls, 8, 3, generator=g)
    K = torch.tensor([[572.4, 0.0, 325.3], [0.0, 573.6, 242.0], [0.0, 0.0, 1.0]])
    out = []
    for n in ngt:
        out.append(types.SimpleNamespace(
            class_ids=torch.randint(0, ncls, (n,), generator=g),
            rotations=torch.randn(n, 3, 3, generator=g), translations=torch.randn(n, 3, generator=g),
            keypoints_3d=kp_table if shared_tables else kp_table.clone(), K=K if shared_tables else K.clone(),
            mask=torch.randint(0, n + 1, (16, 16), generator=g).float(), bbox_trans=torch.randn(2, 3, generator=g)))
    return out


@pytest.mark.parametrize("ngt,shared", [([2, 0, 3, 1], True), ([2, 2, 2], True), ([1, 1, 1, 1], False), ([0, 0], True),
                                        ([3, 1, 0, 2, 3], False)])
def test_stack_targets_matches_a_per_image_loop(ngt, shared):
    tg = _make(ngt, shared)
    st = T._stack_targets(tg, torch.device("cpu"))
    maxgt = max(1, max(ngt))
    assert st["maxgt"] == maxgt and st["ngt"] == ngt
    assert st["num_gt"].tolist() == ngt
    for name, shape in (("rot", (len(ngt), maxgt, 3, 3)), ("trans", (len(ngt), maxgt, 3)), ("kp3d", (len(ngt), maxgt, 8, 3)),
                        ("cls1", (len(ngt), maxgt))):
        assert tuple(st[name].shape) == shape and st[name].is_contiguous()
    for i, t in enumerate(tg):
        n = ngt[i]
        assert torch.equal(st["rot"][i, :n], t.rotations.float())
        assert torch.equal(st["trans"][i, :n], t.translations.float())
        assert torch.equal(st["cls1"][i, :n], t.class_ids + 1)
        assert torch.equal(st["kp3d"][i, :n], t.keypoints_3d[t.class_ids].float())
        for arr in (st["rot"], st["trans"], st["kp3d"], st["cls1"]):      # padding rows are zero (cls1 = 0: "no object")
            assert not arr[i, n:].any()
        assert torch.equal(st["mask"][i], t.mask)
        assert torch.equal(st["K"][i], t.K.float())
        assert torch.equal(st["bt"][i], t.bbox_trans)


def test_more_than_eight_objects_is_rejected():
    with pytest.raises(ValueError):
        T._stack_targets(_make([9]), torch.device("cpu"))


def test_positives_aux_reads_the_owner_rows():
    ngt = [2, 0, 3, 1]
    tg = _make(ngt, seed=3)
    st = T._stack_targets(tg, torch.device("cpu"))
    cells = 10
    # positives: (image, object, cell) triples in label order
    trip = [(0, 1, 4), (0, 0, 7), (2, 2, 0), (2, 0, 9), (3, 0, 5)]
    pos_inds = torch.tensor([i * cells + c for i, _g, c in trip])
    owner = torch.zeros(len(ngt) * cells, dtype=torch.int32)
    for i, g, c in trip:
        owner[i * cells + c] = g
    res = dict(st=st, cells=cells, owner=owner)
    cls_label, aux_3d, bt = T.positives_aux(res, pos_inds)
    for k, (i, g, _c) in enumerate(trip):
        t = tg[i]
        assert int(cls_label[k]) == int(t.class_ids[g])
        want = t.keypoints_3d[t.class_ids[g]].float() @ t.rotations[g].float().T + t.translations[g].float()
        np.testing.assert_allclose(aux_3d[k].numpy(), want.numpy(), rtol=1e-6, atol=1e-6)
        assert torch.equal(bt[k], t.bbox_trans)

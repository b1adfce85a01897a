"""Size-independent properties of the fused loss on the GPU, at BASELINE.json's full sizes
(batch 64 / 512 sparse, dense 1360 x 1364 cells): what any correct implementation of the algorithm must satisfy."""
import numpy as np
import pytest
import torch

from kd_6d_pose_adlp_b200.synthetic import ot_batch

pytestmark = pytest.mark.gpu


def run(batch, **kw):
    from kd_6d_pose_adlp_b200.ops import OTConfig, ot_loss_batched

    dev = torch.device("cuda:0")
    t = {k: torch.from_numpy(np.ascontiguousarray(batch[k])).to(dev) for k in ("xs", "ws", "xt", "wt")}
    out = ot_loss_batched(t["xs"], t["ws"], t["xt"], t["wt"], batch["pos_per_img"], batch["pos_per_img_t"],
                          OTConfig(**kw.pop("cfg", {})), **kw)
    torch.cuda.synchronize()
    return {k: v.cpu().numpy() for k, v in out.items() if v is not None}


def take(batch, idx):
    """Sub-batch made of the images `idx` (in that order)."""
    cn = np.concatenate([[0], np.cumsum(batch["pos_per_img"])])
    cm = np.concatenate([[0], np.cumsum(batch["pos_per_img_t"])])
    sel_s = np.concatenate([np.arange(cn[i], cn[i + 1]) for i in idx]).astype(int)
    sel_t = np.concatenate([np.arange(cm[i], cm[i + 1]) for i in idx]).astype(int)
    return dict(xs=batch["xs"][sel_s], ws=batch["ws"][sel_s], xt=batch["xt"][sel_t], wt=batch["wt"][sel_t],
                pos_per_img=[batch["pos_per_img"][i] for i in idx], pos_per_img_t=[batch["pos_per_img_t"][i] for i in idx]), sel_s


def test_batch_composition_invariance_bit_exact_batch512():
    """Config 5 size (batch 512): an image's loss / gradients do not depend on what else is in the batch or where."""
    b = ot_batch(512, seed=77)
    full = run(b)
    assert (full["valid"] >= 0).all() and np.isfinite(full["loss_per_img"]).all()
    rng = np.random.default_rng(0)
    idx = rng.permutation(512)[:64].tolist()
    sub, sel_s = take(b, idx)
    part = run(sub)
    np.testing.assert_array_equal(part["loss_per_img"], full["loss_per_img"][idx])
    np.testing.assert_array_equal(part["nits"], full["nits"][idx])
    np.testing.assert_array_equal(part["grad_xs"], full["grad_xs"][sel_s])
    np.testing.assert_array_equal(part["grad_ws"], full["grad_ws"][sel_s])
    # run-to-run determinism
    again = run(b)
    for k in ("loss_per_img", "grad_xs", "grad_ws"):
        np.testing.assert_array_equal(again[k], full[k])


def test_self_divergence_is_zero_and_grad_vanishes():
    """F(alpha, x; alpha, x) = 0 (debiased divergence), at the sparse and at the dense size."""
    for b in (ot_batch(64, seed=5, p_empty_teacher=0.0), ot_batch(2, seed=6, dense=(1360, 1360))):
        b = dict(b, xt=b["xs"].copy(), wt=b["ws"].copy(), pos_per_img_t=list(b["pos_per_img"]))
        out = run(b)
        scale = np.abs(b["ws"]).sum() / len(b["pos_per_img"]) * 1e-3   # loss scale of a typical non-trivial problem
        assert np.abs(out["loss_per_img"]).max() < 1e-4 * max(scale, 1e-3)
        assert np.abs(out["grad_ws"]).max() < 1e-6


def test_cell_permutation_invariance_dense():
    """Permuting the student / teacher cells of a dense image permutes the gradients and keeps the loss."""
    b = ot_batch(2, seed=8, dense=(1360, 1364), sigma=0.1)
    out = run(b)
    rng = np.random.default_rng(1)
    ps = np.concatenate([rng.permutation(1360), 1360 + rng.permutation(1360)])
    pt = np.concatenate([rng.permutation(1364), 1364 + rng.permutation(1364)])
    bp = dict(b, xs=b["xs"][ps], ws=b["ws"][ps], xt=b["xt"][pt], wt=b["wt"][pt])
    outp = run(bp)
    assert np.array_equal(out["nits"], outp["nits"])
    assert np.abs(outp["loss_per_img"] - out["loss_per_img"]).max() <= 2e-6 * np.abs(out["loss_per_img"]).max()
    assert np.abs(outp["grad_ws"] - out["grad_ws"][ps]).max() <= 2e-5 * np.abs(out["grad_ws"]).max()
    assert np.abs(outp["grad_xs"] - out["grad_xs"][ps]).max() <= 5e-3 * np.abs(out["grad_xs"]).max()


def test_translation_invariance():
    """Shifting both clouds by the same vector changes nothing (cost by differences; diameter unchanged)."""
    b = ot_batch(64, seed=9, in_pixels=False)
    out = run(b, normalize=False)
    shift = np.array([0.125, -0.0625], np.float32)  # exactly representable: the shifted inputs are exact
    bs = dict(b, xs=b["xs"] + shift, xt=b["xt"] + shift)
    outs = run(bs, normalize=False)
    assert np.array_equal(out["nits"], outs["nits"])
    assert np.abs(outs["loss_per_img"] - out["loss_per_img"]).max() <= 1e-5 * np.abs(out["loss_per_img"]).max()
    assert np.abs(outs["grad_xs"] - out["grad_xs"]).max() <= 5e-3 * np.abs(out["grad_xs"]).max()


def test_gradient_is_descent_direction_dense():
    """A small step along -grad_xs decreases the dense loss (first-order check of the analytic backward at full size)."""
    b = ot_batch(1, seed=10, dense=(1360, 1364), sigma=0.1, in_pixels=False)
    out = run(b, normalize=False, cfg=dict(blur=0.05))
    g = out["grad_xs"]
    step = 1e-3 / np.abs(g).max()
    b2 = dict(b, xs=(b["xs"] - step * g).astype(np.float32))
    out2 = run(b2, normalize=False, cfg=dict(blur=0.05))
    pred = -step * float((g.astype(np.float64) ** 2).sum())
    got = float(out2["loss_per_img"][0]) - float(out["loss_per_img"][0])
    assert got < 0 and 0.3 < got / pred < 3.0, (got, pred)

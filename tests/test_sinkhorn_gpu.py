"""GPU parity of the fused OT kernels (through the C ABI) against the CPU oracles."""
import numpy as np
import pytest
import torch

from kd_6d_pose_adlp_b200.synthetic import ot_batch
from tests import parity, refs

pytestmark = pytest.mark.gpu


def run_gpu(batch, blur=0.001, reach=0.5, scaling=0.5, normalize=True, weighted=True):
    from kd_6d_pose_adlp_b200.ops import OTConfig, ot_loss_batched

    dev = torch.device("cuda:0")
    xs = torch.from_numpy(batch["xs"]).to(dev)
    xt = torch.from_numpy(batch["xt"]).to(dev)
    ws = torch.from_numpy(batch["ws"]).to(dev) if weighted else None
    wt = torch.from_numpy(batch["wt"]).to(dev) if weighted else None
    out = ot_loss_batched(xs, ws, xt, wt, batch["pos_per_img"], batch["pos_per_img_t"],
                          OTConfig(p=2.0, blur=blur, scaling=scaling, reach=reach), normalize=normalize)
    torch.cuda.synchronize()
    res = {k: (v.cpu().numpy() if v is not None else None) for k, v in out.items()}
    res["xs_norm"] = xs.cpu().numpy()
    res["xt_norm"] = xt.cpu().numpy()
    return res


def check(batch, **kw):
    g = run_gpu(batch, **kw)
    l32, gx32, gw32, xsn32, xtn32 = refs.ref32(batch, **kw)
    o64 = refs.ref64(batch, **kw)
    keep = o64["valid"] == 1
    # integer / bit-exact parts
    np.testing.assert_array_equal(g["valid"], o64["valid"])
    np.testing.assert_array_equal(g["nits"], o64["nits"])
    if kw.get("normalize", True):
        np.testing.assert_array_equal(g["xs_norm"], xsn32)   # in-place side effect, bit-exact
        np.testing.assert_array_equal(g["xt_norm"], xtn32)
    assert np.all(g["loss_per_img"][~keep] == 0)
    rows = [parity.report("loss_per_img", g["loss_per_img"], l32, o64["loss_per_img"]),
            parity.report("grad_xs", g["grad_xs"], gx32, o64["grad_xs"])]
    if kw.get("weighted", True):
        rows.append(parity.report("grad_ws", g["grad_ws"], gw32, o64["grad_ws"]))
    print("\n" + parity.fmt(rows))
    assert all(r["ok"] for r in rows), parity.fmt(rows)
    return rows


@pytest.mark.parametrize("sigma", [0.05, 0.005, 0.15])
@pytest.mark.parametrize("nimg", [8, 64])
def test_small_kernel_ape_shape(nimg, sigma):
    check(ot_batch(nimg, seed=nimg, sigma=sigma))


def test_small_kernel_up_to_64_points():
    check(ot_batch(5, seed=3, n_range=(20, 32), m_range=(20, 32)))


def test_balanced_and_unweighted():
    b = ot_batch(6, seed=11)
    check(b, reach=None)
    check(b, weighted=False)
    check(b, normalize=False) if False else None


@pytest.mark.parametrize("scaling,blur", [(0.7, 0.001), (0.9, 0.01), (0.5, 0.05)])
def test_schedules(scaling, blur):
    check(ot_batch(4, seed=5), scaling=scaling, blur=blur)


def test_tiled_kernel_medium():
    check(ot_batch(3, seed=7, n_range=(40, 90), m_range=(50, 120), p_empty_teacher=0.0))


def test_tiled_kernel_mixed_with_empty():
    b = ot_batch(6, seed=9, n_range=(1, 70), m_range=(1, 70), p_empty_teacher=0.3)
    check(b)


def test_tiled_kernel_dense_one_image():
    check(ot_batch(1, seed=1, dense=(1360, 1364), sigma=0.1))


@pytest.mark.parametrize("kind", ["small_zero_copy", "small_zero_copy_nowb", "small_memcpy", "tiled"])
def test_host_buffer_entry_point_matches_device_path(kind, monkeypatch):
    """kdot_sinkhorn_fwd_bwd_host (host NumPy buffers in/out) is bit-identical to the device-pointer call."""
    import ctypes

    from kd_6d_pose_adlp_b200 import _lib

    monkeypatch.setenv("KDOT_HOST_ZERO_COPY", "1" if kind.startswith("small_zero_copy") else "0")
    wb = 0 if kind.endswith("nowb") else 1
    batch = ot_batch(9, seed=21) if kind != "tiled" else ot_batch(3, seed=22, n_range=(40, 70), m_range=(40, 70))
    g = run_gpu(batch)
    L = _lib.lib()
    nimg = len(batch["pos_per_img"])
    sn, sm = batch["xs"].shape[0], batch["xt"].shape[0]
    ctx = L.kdot_host_ctx_create(0, nimg, sn, sm, 8, 2)
    assert ctx, L.kdot_last_error()
    xs, xt = batch["xs"].copy(), batch["xt"].copy()
    pn, pm = np.asarray(batch["pos_per_img"], np.int32), np.asarray(batch["pos_per_img_t"], np.int32)
    loss, valid, nits = np.empty(nimg, np.float32), np.empty(nimg, np.int32), np.empty(nimg, np.int32)
    gx, gw = np.empty_like(xs), np.empty_like(batch["ws"])
    p = lambda a: a.ctypes.data
    rc = L.kdot_sinkhorn_fwd_bwd_host(ctx, p(xs), p(batch["ws"]), p(xt), p(batch["wt"]), p(pn), p(pm), nimg, 2.0, 0.001, 0.5,
                                      0.5, 640.0, 480.0, 1, wb, p(loss), p(valid), p(gx), p(gw), p(nits))
    assert rc == 0, L.kdot_last_error()
    h2d, d2h = ctypes.c_size_t(0), ctypes.c_size_t(0)
    L.kdot_host_ctx_last_traffic(ctx, ctypes.byref(h2d), ctypes.byref(d2h))
    L.kdot_host_ctx_destroy(ctx)
    assert h2d.value >= xs.nbytes + xt.nbytes and d2h.value >= gx.nbytes
    np.testing.assert_array_equal(loss, g["loss_per_img"])
    np.testing.assert_array_equal(valid, g["valid"])
    np.testing.assert_array_equal(nits, g["nits"])
    np.testing.assert_array_equal(gx, g["grad_xs"])
    np.testing.assert_array_equal(gw, g["grad_ws"])
    if wb:
        np.testing.assert_array_equal(xs, g["xs_norm"])      # write_back_normalized
        np.testing.assert_array_equal(xt, g["xt_norm"])
    else:
        np.testing.assert_array_equal(xs, batch["xs"])


def test_bad_arguments_fail_loudly():
    from kd_6d_pose_adlp_b200 import _lib
    from kd_6d_pose_adlp_b200.ops import OTConfig, ot_loss_batched

    dev = torch.device("cuda:0")
    b = ot_batch(2, seed=0)
    t = {k: torch.from_numpy(b[k]).to(dev) for k in ("xs", "ws", "xt", "wt")}
    with pytest.raises(_lib.KdotError, match="p == 2"):
        ot_loss_batched(t["xs"], t["ws"], t["xt"], t["wt"], b["pos_per_img"], b["pos_per_img_t"], OTConfig(p=1.0))
    with pytest.raises(ValueError):
        ot_loss_batched(t["xs"], t["ws"], t["xt"], t["wt"], b["pos_per_img"][:-1] + [999], b["pos_per_img_t"])
    # clouds too large for the shared-memory plan are refused, not silently mis-computed
    big = ot_batch(1, seed=0, dense=(4000, 4000))
    tb = {k: torch.from_numpy(big[k]).to(dev) for k in ("xs", "ws", "xt", "wt")}
    with pytest.raises(_lib.KdotError, match="shared memory"):
        ot_loss_batched(tb["xs"], tb["ws"], tb["xt"], tb["wt"], big["pos_per_img"], big["pos_per_img_t"])


@pytest.mark.parametrize("scaling", [0.95, 0.975])
def test_long_schedules_take_the_log_exp_route(scaling):
    """> 48 schedule entries (BASELINE.json configs[4]: 100-200 Sinkhorn rounds): device log/exp schedule route."""
    rows = check(ot_batch(3, seed=13), scaling=scaling)
    b = ot_batch(2, seed=14, n_range=(40, 50), m_range=(40, 50), p_empty_teacher=0.0)
    check(b, scaling=scaling)


def test_samples_loss_seam_b2_with_autograd():
    """SamplesLoss(alpha, x, beta, y) -> (B,) on the (B,N,D) layout, gradients to x and alpha (loss_libs.py:47)."""
    from kd_6d_pose_adlp_b200 import SamplesLoss
    from oracle import geomloss_ref

    dev = torch.device("cuda:0")
    g = torch.Generator().manual_seed(5)
    for B, N, M in [(8, 10, 12), (1, 7, 5), (3, 50, 40)]:
        x = (0.5 + 0.05 * torch.randn(B, N, 2, generator=g)).float()
        y = (0.5 + 0.05 * torch.randn(B, M, 2, generator=g)).float()
        a = torch.rand(B, N, generator=g) * 0.9 + 0.05
        b = torch.rand(B, M, generator=g) * 0.9 + 0.05
        xd, ad = x.to(dev).requires_grad_(True), a.to(dev).requires_grad_(True)
        out = SamplesLoss("sinkhorn", p=2, blur=0.001, scaling=0.5, reach=0.5)(ad, xd, b.to(dev), y.to(dev))
        assert out.shape == (B,)
        wsum = torch.linspace(0.5, 1.5, B, device=dev)
        (out * wsum).sum().backward()
        # fp64 oracle through autograd
        x64, a64 = x.double().requires_grad_(True), a.double().requires_grad_(True)
        ref = geomloss_ref.SamplesLoss("sinkhorn", p=2, blur=0.001, scaling=0.5, reach=0.5)(a64, x64, b.double(), y.double())
        (ref * wsum.cpu().double()).sum().backward()
        assert parity.rel(out.detach().cpu().numpy(), ref.detach().numpy()) < 2e-6
        assert parity.rel(ad.grad.cpu().numpy(), a64.grad.numpy()) < 2e-5
        assert parity.rel(xd.grad.cpu().numpy(), x64.grad.numpy()) < 5e-3
        # unweighted (x, y) call: uniform 1/N, 1/M masses (loss_libs.py:49)
        out_u = SamplesLoss("sinkhorn", p=2, blur=0.01, scaling=0.5, reach=None)(x.to(dev), y.to(dev))
        ref_u = geomloss_ref.SamplesLoss("sinkhorn", p=2, blur=0.01, scaling=0.5, reach=None)(x.double(), y.double())
        assert parity.rel(out_u.cpu().numpy(), ref_u.numpy()) < 2e-6


def test_zero_mass_cells_and_status_codes():
    from kd_6d_pose_adlp_b200 import _lib

    b = ot_batch(4, seed=17, p_empty_teacher=0.0)
    b["ws"][::3] = 0.0           # zero-mass student cells: log-weight -100000 like geomloss
    check(b)
    # an image whose points all coincide has zero diameter: geomloss would raise from np.arange; we flag it
    d = ot_batch(2, seed=18, p_empty_teacher=0.0)
    n0 = d["pos_per_img"][0]
    m0 = d["pos_per_img_t"][0]
    d["xs"][:n0] = 100.0
    d["xt"][:m0] = 100.0
    g = run_gpu(d)
    assert g["valid"][0] == _lib.KDOT_IMG_DEGENERATE and np.isnan(g["loss_per_img"][0])
    assert g["valid"][1] == _lib.KDOT_IMG_OK and np.isfinite(g["loss_per_img"][1])
    assert np.isnan(g["grad_xs"][:n0]).all() and np.isfinite(g["grad_xs"][n0:]).all()

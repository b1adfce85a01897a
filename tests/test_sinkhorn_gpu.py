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


def check(batch, tol=parity.TOL, **kw):
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
            parity.report("grad_xs", g["grad_xs"], gx32, o64["grad_xs"], tol=tol)]
    if kw.get("weighted", True):
        rows.append(parity.report("grad_ws", g["grad_ws"], gw32, o64["grad_ws"]))
    print("\n" + parity.fmt(rows))
    assert all(r["ok"] for r in rows), parity.fmt(rows)
    return rows


@pytest.mark.parametrize("sigma", [0.05, 0.005, 0.15])
@pytest.mark.parametrize("nimg", [8, 64])
def test_small_kernel_ape_shape(nimg, sigma):
    check(ot_batch(nimg, seed=nimg, sigma=sigma))


@pytest.mark.parametrize("B", [1, 2, 4, 16])
@pytest.mark.parametrize("nimg", [3, 200])
def test_small_kernel_slot_counts(B, nimg):
    """Slot counts other than the 8 key-points of the reference (B = 16: two CTAs per image at least; nimg = 200: more
    CTAs than SMs, the rolled variant): every warp derives the image-wide bounding box from the raw points of all B
    slots, and the in-place normalisation waits for the whole cluster -- diameter (hence `nits`), normalised
    coordinates and gradients have to match for every split."""
    from kd_6d_pose_adlp_b200.synthetic import cu_seqlens
    from oracle import sinkhorn_analytic

    b = ot_batch(nimg, seed=10 * B + nimg, B=B, p_empty_teacher=0.2)
    g = run_gpu(b)
    o = sinkhorn_analytic.kdot_fwd_bwd_f64(b["xs"], b["ws"], b["xt"], b["wt"], cu_seqlens(b["pos_per_img"]),
                                           cu_seqlens(b["pos_per_img_t"]), B, 2)
    np.testing.assert_array_equal(g["valid"], o["valid"])
    np.testing.assert_array_equal(g["nits"], o["nits"])
    np.testing.assert_array_equal(g["xs_norm"], o["xs_norm"])   # in-place side effect, bit-exact
    np.testing.assert_array_equal(g["xt_norm"], o["xt_norm"])
    for name in ("loss_per_img", "grad_xs", "grad_ws"):
        assert parity.rel(g[name], o[name]) <= parity.TOL, (name, parity.rel(g[name], o[name]))


def test_small_kernel_up_to_64_points():
    check(ot_batch(5, seed=3, n_range=(20, 32), m_range=(20, 32)))


def test_balanced_and_unweighted():
    b = ot_batch(6, seed=11)
    check(b, reach=None)
    check(b, weighted=False)
    check(b, normalize=False)   # SamplesLoss called on already-normalised coordinates (seam B2)


@pytest.mark.parametrize("scaling,blur", [(0.7, 0.001), (0.9, 0.01), (0.5, 0.05)])
def test_schedules(scaling, blur):
    check(ot_batch(4, seed=5), scaling=scaling, blur=blur)


def test_tiled_kernel_medium():
    check(ot_batch(3, seed=7, n_range=(40, 90), m_range=(50, 120), p_empty_teacher=0.0))


def test_tiled_kernel_mixed_with_empty():
    b = ot_batch(6, seed=9, n_range=(1, 70), m_range=(1, 70), p_empty_teacher=0.3)
    check(b)


def test_dense_all_cells_one_image():
    """Every cell of the darknet_tiny / darknet53 grids (BASELINE.json configs[2], variant 3b): streaming kernel."""
    check(ot_batch(1, seed=1, dense=(1360, 1364), sigma=0.1), tol=parity.TOL_STREAM)


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
    with pytest.raises(_lib.KdotError, match="p must be 1 or 2"):
        ot_loss_batched(t["xs"], t["ws"], t["xt"], t["wt"], b["pos_per_img"], b["pos_per_img_t"], OTConfig(p=3.0))
    with pytest.raises(ValueError):
        ot_loss_batched(t["xs"], t["ws"], t["xt"], t["wt"], b["pos_per_img"][:-1] + [999], b["pos_per_img_t"])


@pytest.mark.parametrize("path", ["stream", "tiled"])
def test_large_path_kernels_match_oracle_d2(path, monkeypatch):
    """Both large-cloud kernels (streaming cooperative = default, tiled = KDOT_FORCE_PATH=tiled) on D = 2 problems."""
    monkeypatch.setenv("KDOT_FORCE_PATH", path)
    tol = parity.TOL_STREAM if path == "stream" else parity.TOL
    check(ot_batch(3, seed=7, n_range=(40, 90), m_range=(50, 120), p_empty_teacher=0.0), tol=tol)
    check(ot_batch(6, seed=9, n_range=(1, 70), m_range=(1, 70), p_empty_teacher=0.3), tol=tol)
    check(ot_batch(4, seed=31, n_range=(33, 40), m_range=(33, 40)), tol=tol, reach=None, scaling=0.7)


@pytest.mark.parametrize("n_range,m_range", [((16, 16), (16, 16)), ((16, 17), (16, 17)), ((3, 40), (2, 24)),
                                             ((128, 128), (128, 128)), ((120, 129), (126, 128))])
def test_default_dispatch_around_the_kernel_boundaries(n_range, m_range, monkeypatch):
    """No path forced: 32 | 33 points (register kernel -> CTA-resident tiled kernel) and 256 | 257 points
    (tiled -> streaming) give oracle-matching results on either side; the batch maximum decides for every image."""
    monkeypatch.delenv("KDOT_FORCE_PATH", raising=False)
    tol = parity.TOL_STREAM if n_range[1] + m_range[1] > 256 else parity.TOL
    check(ot_batch(5, seed=sum(n_range) + 3 * sum(m_range), n_range=n_range, m_range=m_range, p_empty_teacher=0.2), tol=tol)


@pytest.mark.parametrize("dense,sigma,B", [((600, 640), 0.005, 2), ((900, 300), 0.3, 1), ((4200, 150), 0.1, 1)])
def test_stream_kernel_sorted_staging_and_exact_tile_skipping(dense, sigma, B):
    """The streaming kernel stages D = 2 clouds of up to 4096 points in Morton order and skips, in the cold rounds,
    column tiles whose every exponential would flush to zero.  Tight clusters (most tiles skipped), wide clouds (few)
    and a cloud beyond the sorting limit (identity order, nothing to skip) all have to match the fp64 oracle."""
    from kd_6d_pose_adlp_b200.ops import OTConfig, ot_loss_batched
    from kd_6d_pose_adlp_b200.synthetic import cu_seqlens
    from oracle import sinkhorn_analytic

    b = ot_batch(2, seed=dense[0], dense=dense, B=B, sigma=sigma)
    dev = torch.device("cuda:0")
    t = {k: torch.from_numpy(b[k]).to(dev) for k in ("xs", "ws", "xt", "wt")}
    out = ot_loss_batched(t["xs"], t["ws"], t["xt"], t["wt"], b["pos_per_img"], b["pos_per_img_t"], OTConfig())
    torch.cuda.synchronize()
    o = sinkhorn_analytic.kdot_fwd_bwd_f64(b["xs"], b["ws"], b["xt"], b["wt"], cu_seqlens(b["pos_per_img"]),
                                           cu_seqlens(b["pos_per_img_t"]), B, 2)
    assert np.array_equal(out["nits"].cpu().numpy(), o["nits"])
    e_l, e_w = parity.rel(out["loss_per_img"].cpu().numpy(), o["loss_per_img"]), parity.rel(out["grad_ws"].cpu().numpy(), o["grad_ws"])
    e_x, q_x = parity.rel(out["grad_xs"].cpu().numpy(), o["grad_xs"]), parity.rel_quantile(out["grad_xs"].cpu().numpy(), o["grad_xs"])
    print(f"\nstream {dense} sigma {sigma}: loss {e_l:.2e} d/dalpha {e_w:.2e} d/dx {e_x:.2e} (99.9 % quantile {q_x:.2e})")
    assert e_l < 2e-6 and e_w < 5e-5
    assert e_x <= parity.TOL_STREAM and q_x <= parity.TOL / 2
    # run-to-run reproducibility of the sorted path (deterministic ranks, fixed-order reductions)
    t2 = {k: torch.from_numpy(b[k]).to(dev) for k in ("xs", "ws", "xt", "wt")}
    out2 = ot_loss_batched(t2["xs"], t2["ws"], t2["xt"], t2["wt"], b["pos_per_img"], b["pos_per_img_t"], OTConfig())
    assert torch.equal(out["grad_xs"], out2["grad_xs"]) and torch.equal(out["loss_per_img"], out2["loss_per_img"])


def test_stream_kernel_clouds_beyond_shared_memory():
    """N = M = 3500 cells in one slot: too large for the tiled kernel's shared-memory plan -> streaming kernel."""
    from kd_6d_pose_adlp_b200.ops import OTConfig, ot_loss_batched
    from oracle import sinkhorn_analytic
    from kd_6d_pose_adlp_b200.synthetic import cu_seqlens

    b = ot_batch(1, seed=41, dense=(3500, 3400), B=1, sigma=0.1)
    dev = torch.device("cuda:0")
    t = {k: torch.from_numpy(b[k]).to(dev) for k in ("xs", "ws", "xt", "wt")}
    out = ot_loss_batched(t["xs"], t["ws"], t["xt"], t["wt"], b["pos_per_img"], b["pos_per_img_t"], OTConfig())
    torch.cuda.synchronize()
    o = sinkhorn_analytic.kdot_fwd_bwd_f64(b["xs"], b["ws"], b["xt"], b["wt"], cu_seqlens(b["pos_per_img"]),
                                           cu_seqlens(b["pos_per_img_t"]), 1, 2)
    assert np.array_equal(out["nits"].cpu().numpy(), o["nits"])
    e_x, q_x = parity.rel(out["grad_xs"].cpu().numpy(), o["grad_xs"]), parity.rel_quantile(out["grad_xs"].cpu().numpy(), o["grad_xs"])
    print(f"\nstream 3500x3400: d/dx {e_x:.2e} (99.9 % quantile {q_x:.2e})")
    assert parity.rel(out["loss_per_img"].cpu().numpy(), o["loss_per_img"]) < 2e-6
    assert parity.rel(out["grad_ws"].cpu().numpy(), o["grad_ws"]) < 5e-5
    assert e_x <= parity.TOL_STREAM and q_x <= parity.TOL / 2


def test_zebra_full_size_one_image():
    """BASELINE.json configs[3] at FULL size (one of its 8 images): N = M = 4096 cells, D = 16 code probabilities,
    blur 0.05, one slot -- the generic-dimension streaming kernel against the float64 oracle (4 x 4096^2 float64 cost
    matrices, ~10 rounds: some tens of seconds of numpy)."""
    from kd_6d_pose_adlp_b200.ops import OTConfig, ot_loss_batched
    from kd_6d_pose_adlp_b200.synthetic import cu_seqlens
    from oracle import sinkhorn_analytic

    b = ot_batch(1, seed=16, dense=(4096, 4096), B=1, D=16)
    dev = torch.device("cuda:0")
    t = {k: torch.from_numpy(b[k]).to(dev) for k in ("xs", "ws", "xt", "wt")}
    out = ot_loss_batched(t["xs"], t["ws"], t["xt"], t["wt"], b["pos_per_img"], b["pos_per_img_t"],
                          OTConfig(blur=0.05), normalize=False)
    torch.cuda.synchronize()
    o = sinkhorn_analytic.kdot_fwd_bwd_f64(b["xs"], b["ws"], b["xt"], b["wt"], cu_seqlens(b["pos_per_img"]),
                                           cu_seqlens(b["pos_per_img_t"]), 1, 16, blur=0.05, normalize=False)
    assert np.array_equal(out["nits"].cpu().numpy(), o["nits"])
    e_l, e_w = parity.rel(out["loss_per_img"].cpu().numpy(), o["loss_per_img"]), parity.rel(out["grad_ws"].cpu().numpy(), o["grad_ws"])
    e_x = parity.rel(out["grad_xs"].cpu().numpy(), o["grad_xs"])
    print(f"\nzebra 4096 x 4096, D = 16: loss {e_l:.2e} d/dalpha {e_w:.2e} d/dx {e_x:.2e}")
    assert e_l <= parity.TOL and e_w <= parity.TOL and e_x <= parity.TOL


@pytest.mark.parametrize("D,blur", [(16, 0.05), (16, 0.01), (3, 0.01), (8, 0.001)])
def test_stream_kernel_generic_dimension(D, blur):
    """ZebraPose-style per-cell code distributions (BASELINE.json configs[3]): D-dimensional points, one slot."""
    from kd_6d_pose_adlp_b200 import SamplesLoss
    from oracle import geomloss_ref

    dev = torch.device("cuda:0")
    g = torch.Generator().manual_seed(D)
    B, N, M = 2, 150, 170
    x = torch.sigmoid(torch.randn(B, N, D, generator=g)).float()
    y = torch.sigmoid(torch.randn(B, M, D, generator=g)).float()
    a = torch.sigmoid(torch.randn(B, N, generator=g)).float()
    bb = torch.sigmoid(torch.randn(B, M, generator=g)).float()
    xd, ad = x.to(dev).requires_grad_(True), a.to(dev).requires_grad_(True)
    L = SamplesLoss("sinkhorn", p=2, blur=blur, scaling=0.5, reach=0.5)
    out = L(ad, xd, bb.to(dev), y.to(dev))
    out.sum().backward()
    res = {}
    for name, dt in (("ref32", torch.float32), ("ref64", torch.float64)):
        xr, ar = x.detach().clone().to(dt).requires_grad_(True), a.detach().clone().to(dt).requires_grad_(True)
        Lr = geomloss_ref.SamplesLoss("sinkhorn", p=2, blur=blur, scaling=0.5, reach=0.5)
        o = Lr(ar, xr, bb.to(dt), y.to(dt))
        o.sum().backward()
        res[name] = (o.detach().double().numpy(), xr.grad.double().numpy(), ar.grad.double().numpy(), Lr.last_nits)
    assert int(L.last_nits.cpu()[0]) == res["ref64"][3]
    rows = [parity.report("loss", out.detach().cpu().numpy(), res["ref32"][0], res["ref64"][0]),
            parity.report("grad_x", xd.grad.cpu().numpy(), res["ref32"][1], res["ref64"][1]),
            parity.report("grad_a", ad.grad.cpu().numpy(), res["ref32"][2], res["ref64"][2])]
    print("\n" + parity.fmt(rows))
    assert all(r["ok"] for r in rows), parity.fmt(rows)


@pytest.mark.parametrize("blur,reach", [(0.01, 0.5), (0.05, None), (0.001, 0.5)])
def test_p1_distance_cost(blur, reach):
    """SamplesLoss("sinkhorn", p=1): cost |x - y| (geomloss distances with its 1e-8 clamp), streaming kernel."""
    from kd_6d_pose_adlp_b200 import SamplesLoss
    from oracle import geomloss_ref

    dev = torch.device("cuda:0")
    g = torch.Generator().manual_seed(11)
    for B, N, M in [(8, 10, 12), (2, 80, 60)]:
        x = (0.5 + 0.05 * torch.randn(B, N, 2, generator=g)).float()
        y = (0.5 + 0.05 * torch.randn(B, M, 2, generator=g)).float()
        a = torch.rand(B, N, generator=g) * 0.9 + 0.05
        b = torch.rand(B, M, generator=g) * 0.9 + 0.05
        xd, ad = x.to(dev).requires_grad_(True), a.to(dev).requires_grad_(True)
        L = SamplesLoss("sinkhorn", p=1, blur=blur, scaling=0.5, reach=reach)
        out = L(ad, xd, b.to(dev), y.to(dev))
        out.sum().backward()
        refs = {}
        for name, dt in (("ref32", torch.float32), ("ref64", torch.float64)):
            xr, ar = x.clone().to(dt).requires_grad_(True), a.clone().to(dt).requires_grad_(True)
            Lr = geomloss_ref.SamplesLoss("sinkhorn", p=1, blur=blur, scaling=0.5, reach=reach)
            o = Lr(ar, xr, b.to(dt), y.to(dt))
            o.sum().backward()
            refs[name] = (o.detach().double().numpy(), xr.grad.double().numpy(), ar.grad.double().numpy(), Lr.last_nits)
        assert int(L.last_nits.cpu()[0]) == refs["ref64"][3]
        rows = [parity.report("loss", out.detach().cpu().numpy(), refs["ref32"][0], refs["ref64"][0]),
                parity.report("grad_x", xd.grad.cpu().numpy(), refs["ref32"][1], refs["ref64"][1]),
                parity.report("grad_a", ad.grad.cpu().numpy(), refs["ref32"][2], refs["ref64"][2])]
        print("\n" + parity.fmt(rows))
        assert all(r["ok"] for r in rows), parity.fmt(rows)


@pytest.mark.parametrize("kernel_points", [6, 60, 300])   # register / CTA-resident / streaming kernel
def test_schedule_length_at_exact_ties(kernel_points):
    """scaling = 0.5, blur = 2^-10 and a bounding-box diagonal of exactly 1, 1/2, 1/4: diam^2 * 0.25^k hits blur^2
    EXACTLY, so the length of geomloss' epsilon schedule -- numpy's len(arange(2 ln diam, 2 ln blur, 2 ln 0.5)) + 2 -- is
    decided by rounding.  nits is an integer output: it has to equal the oracle's np.arange-based value bit for bit."""
    from kd_6d_pose_adlp_b200.ops import OTConfig, ot_loss_batched
    from kd_6d_pose_adlp_b200.synthetic import cu_seqlens
    from oracle import sinkhorn_analytic

    rng = np.random.default_rng(kernel_points)
    n = m = kernel_points
    extents = [(0.25, 1.25), (0.25, 0.75), (0.5, 0.75), (0.125, 0.1875)]   # diagonals 1, 1/2, 1/4, 1/16 (y is constant)
    xs = np.empty((len(extents) * n, 8, 2), np.float32)
    xt = np.empty((len(extents) * m, 8, 2), np.float32)
    for i, (lo, hi) in enumerate(extents):
        for arr, cnt in ((xs, n), (xt, m)):
            blk = arr[i * cnt:(i + 1) * cnt]
            blk[..., 0] = rng.uniform(lo, hi, (cnt, 8)).astype(np.float32)
            blk[..., 1] = 0.5
        xs[i * n, :, 0] = lo          # the box is spanned exactly
        xt[i * m, :, 0] = hi
    ws = rng.uniform(0.1, 0.9, (len(extents) * n, 8)).astype(np.float32)
    wt = rng.uniform(0.1, 0.9, (len(extents) * m, 8)).astype(np.float32)
    pos_n, pos_m = [n] * len(extents), [m] * len(extents)
    blur = 2.0 ** -10
    dev = torch.device("cuda:0")
    t = lambda a: torch.from_numpy(a.copy()).to(dev)
    out = ot_loss_batched(t(xs), t(ws), t(xt), t(wt), pos_n, pos_m, OTConfig(blur=blur, scaling=0.5), normalize=False)
    torch.cuda.synchronize()
    o = sinkhorn_analytic.kdot_fwd_bwd_f64(xs, ws, xt, wt, cu_seqlens(pos_n), cu_seqlens(pos_m), 8, 2, blur=blur, scaling=0.5,
                                           normalize=False)
    # the oracle's lengths come from np.arange itself; the expected values spell out what "tie" means here
    for (lo, hi), nits in zip(extents, o["nits"]):
        diam = np.float32(hi) - np.float32(lo)
        want = len(np.arange(2 * np.log(float(diam)), 2 * np.log(blur), 2 * np.log(0.5))) + 2
        assert nits == want
    assert np.array_equal(out["nits"].cpu().numpy(), o["nits"]), (out["nits"].cpu().numpy(), o["nits"])
    assert parity.rel(out["loss_per_img"].cpu().numpy(), o["loss_per_img"]) < 1e-5


def _adversarial_clouds():
    """Clouds built to break an over-eager skip rule: tight clusters with far outliers (a tile's box is huge while its
    points are clumped), masses of 1e-3 next to 0.999 (log-weights 10 units apart inside a tile), N >> M and M >> N."""
    rng = np.random.default_rng(123)
    cases = []
    for name, (n, m), blur in (("clusters+outliers", (700, 650), 1e-3), ("clusters+outliers", (700, 650), 1e-2),
                               ("n>>m", (3000, 40), 1e-3), ("m>>n", (40, 2500), 5e-2), ("masses", (900, 900), 1e-3)):
        B = 2
        def cloud(k):
            centres = rng.uniform(0.2, 0.8, (6, 1, 2))
            pts = (centres[rng.integers(0, 6, k)] + 0.004 * rng.standard_normal((k, B, 2)))
            far = rng.random(k) < 0.03                      # 3 % outliers anywhere in the frame
            pts[far] = rng.uniform(0.0, 1.0, (int(far.sum()), B, 2))
            return (pts * np.array([640.0, 480.0])).astype(np.float32)
        xs, xt = cloud(n), cloud(m)
        if name == "masses":
            ws = np.where(rng.random((n, 1)) < 0.5, 1e-3, 0.999).astype(np.float32).repeat(B, 1)
            wt = np.where(rng.random((m, 1)) < 0.5, 1e-3, 0.999).astype(np.float32).repeat(B, 1)
        else:
            ws = rng.uniform(0.05, 0.95, (n, 1)).astype(np.float32).repeat(B, 1)
            wt = rng.uniform(0.05, 0.95, (m, 1)).astype(np.float32).repeat(B, 1)
        cases.append((f"{name} {n}x{m} blur {blur}", xs, ws, xt, wt, blur))
    return cases


def test_tile_skipping_is_bit_exact():
    """The streaming kernel's tile skipping drops only terms whose exponential is exactly +0: the shipped library and
    the -DKDOT_NO_TILE_SKIP build of the same sources (libkdot_noskip.so, built by __graft_entry__.build(): same Morton
    order, seeds and arithmetic, every tile evaluated) must agree BIT FOR BIT on loss, d/dx and d/dalpha."""
    from kd_6d_pose_adlp_b200 import _lib
    from kd_6d_pose_adlp_b200.build import NOSKIP_LIB_PATH

    dev = torch.device("cuda:0")
    libs = {"skip": _lib.lib(), "noskip": _lib.load(NOSKIP_LIB_PATH)}
    for name, xs, ws, xt, wt, blur in _adversarial_clouds():
        n, m, B = xs.shape[0], xt.shape[0], xs.shape[1]
        res = {}
        for tag, L in libs.items():
            t = lambda a: torch.from_numpy(a.copy()).to(dev)
            dxs, dws, dxt, dwt = t(xs), t(ws), t(xt), t(wt)
            cu = torch.tensor([0, n, 0, m], dtype=torch.int32, device=dev)
            loss = torch.empty(1, device=dev); valid = torch.empty(1, dtype=torch.int32, device=dev)
            nits = torch.empty(1, dtype=torch.int32, device=dev)
            gx, gw = torch.empty_like(dxs), torch.empty_like(dws)
            nb = int(L.kdot_workspace_bytes(1, n, m, B, 2))
            wsp = torch.empty(max(nb, 16), dtype=torch.uint8, device=dev)
            rc = L.kdot_sinkhorn_fwd_bwd(dxs.data_ptr(), dws.data_ptr(), dxt.data_ptr(), dwt.data_ptr(), cu[:2].data_ptr(),
                                         cu[2:].data_ptr(), 1, B, 2, n, m, 0, 2.0, blur, 0.5, 0.5, 640.0, 480.0, 1,
                                         loss.data_ptr(), None, valid.data_ptr(), gx.data_ptr(), gw.data_ptr(), nits.data_ptr(),
                                         wsp.data_ptr(), nb, torch.cuda.current_stream(dev).cuda_stream)
            assert rc == 0, L.kdot_last_error()
            torch.cuda.synchronize()
            res[tag] = (loss.cpu().numpy(), gx.cpu().numpy(), gw.cpu().numpy(), int(nits.cpu()[0]))
        a, b = res["skip"], res["noskip"]
        assert a[3] == b[3] and np.isfinite(a[0]).all() and np.isfinite(a[1]).all(), name
        for k, what in enumerate(("loss", "d/dx", "d/dalpha")):
            assert np.array_equal(a[k].view(np.uint32), b[k].view(np.uint32)), f"{name}: {what} differs between the skipping and the non-skipping build"

"""the eight details of the geomloss-0.2.4 restatement that decide the (schedule-dependent) value
of the loss and would have to be re-verified first if real geomloss sources ever became available -- as an executable
checklist.  Each test instruments ``oracle/geomloss_ref.py`` (test infrastructure) and states the behaviour the CUDA
kernels reproduce; should any of them ever be found to differ from upstream geomloss, this file names the assumption that
broke.  geomloss itself is absent (requirements.txt:45, not installable offline): parity remains UNPINNED."""
import numpy as np
import pytest
import torch

from oracle import geomloss_ref as G


def _problem(B=2, N=5, M=6, seed=0, dtype=torch.float64):
    g = torch.Generator().manual_seed(seed)
    x = (0.5 + 0.05 * torch.randn(B, N, 2, generator=g)).to(dtype)
    y = (0.5 + 0.05 * torch.randn(B, M, 2, generator=g)).to(dtype)
    a = (torch.rand(B, N, generator=g) * 0.9 + 0.05).to(dtype)
    b = (torch.rand(B, M, generator=g) * 0.9 + 0.05).to(dtype)
    return a, x, b, y


def _trace(monkeypatch, **kw):
    """Run sinkhorn_tensorized and record every softmin call as (eps, rows, cols, grad_enabled, h)."""
    calls = []
    real = G.softmin_tensorized

    def spy(eps, C, f):
        calls.append(dict(eps=float(eps), rows=C.shape[1], cols=C.shape[2], grad=torch.is_grad_enabled(), h=f.detach().clone()))
        return real(eps, C, f)

    monkeypatch.setattr(G, "softmin_tensorized", spy)
    a, x, b, y = _problem()
    x = x.clone().requires_grad_(True)
    out, nits, diam = G.sinkhorn_tensorized(a, x, b, y, p=2, blur=0.001, reach=0.5, scaling=0.5, return_nits=True, **kw)
    return calls, out, nits, diam, (a, x, b, y)


def test_i_coarsest_temperature_is_visited_three_times(monkeypatch):
    """(i) init at eps_s[0], the loop starts again at eps_s[0], and np.arange re-emits diam**p as its first element."""
    calls, _out, nits, diam, _ = _trace(monkeypatch)
    eps_s = G.epsilon_schedule(2, diam, 0.001, 0.5)
    assert len(eps_s) == nits and eps_s[0] == diam ** 2 and abs(eps_s[1] - diam ** 2) < 1e-12 * diam ** 2
    per_round = [calls[i]["eps"] for i in range(0, len(calls), 4)]
    assert len(calls) == 4 * (nits + 2)                         # init + nits loop rounds + last extrapolation, 4 softmins each
    assert per_round[0] == per_round[1] == eps_s[0] and abs(per_round[2] - eps_s[0]) < 1e-12 * eps_s[0]
    assert per_round[1:-1] == eps_s and per_round[-1] == eps_s[-1] == 0.001 ** 2


def test_ii_cross_pairing_and_iii_all_four_averaged(monkeypatch):
    """(ii) at_y is computed from b_x and bt_x from a_y; (iii) all four potentials take the 1/2-averaged update."""
    calls, _out, nits, _diam, (a, x, b, y) = _trace(monkeypatch)
    N, M = x.shape[1], y.shape[1]
    la, lb = G.log_weights(a), G.log_weights(b)
    C = {k: G.cost_routines[2](u, v) for k, (u, v) in dict(xx=(x, x), yy=(y, y), yx=(y, x), xy=(x, y)).items()}
    # replay round 1 (the first loop round) by hand from the init round's outputs
    e0 = calls[0]["eps"]
    lam = G.dampening(e0, 0.25)
    sm = lambda Cm, h: -e0 * (h[:, None, :] - Cm / e0).logsumexp(2)
    a_x, b_y, a_y, b_x = lam * sm(C["xx"], la), lam * sm(C["yy"], lb), lam * sm(C["yx"], la), lam * sm(C["xy"], lb)
    # order of the four calls inside a round: (xx, a_x) (yy, b_y) (yx, b_x -> at_y) (xy, a_y -> bt_x)
    r1 = calls[4:8]
    assert [(c["rows"], c["cols"]) for c in r1] == [(N, N), (M, M), (M, N), (N, M)]
    assert torch.allclose(r1[2]["h"], (la + b_x / e0).detach(), rtol=0, atol=1e-12)   # at_y <- b_x   (cross)
    assert torch.allclose(r1[3]["h"], (lb + a_y / e0).detach(), rtol=0, atol=1e-12)   # bt_x <- a_y   (cross)
    assert torch.allclose(r1[0]["h"], (la + a_x / e0).detach(), rtol=0, atol=1e-12)
    # round 2 consumes the AVERAGED potentials of round 1 -- for all four
    at = [lam * sm(C["xx"], r1[0]["h"]), lam * sm(C["yy"], r1[1]["h"]), lam * sm(C["yx"], r1[2]["h"]), lam * sm(C["xy"], r1[3]["h"])]
    avg = [0.5 * (a_x + at[0]), 0.5 * (b_y + at[1]), 0.5 * (a_y + at[2]), 0.5 * (b_x + at[3])]
    e1 = calls[8]["eps"]
    r2 = calls[8:12]
    want = [la + avg[0] / e1, lb + avg[1] / e1, la + avg[3] / e1, lb + avg[2] / e1]
    for c, w in zip(r2, want):
        assert torch.allclose(c["h"], w.detach(), rtol=0, atol=1e-9)


def test_iv_last_extrapolation_is_plain_detached_and_from_the_same_old_potentials(monkeypatch):
    calls, out, nits, _diam, (a, x, b, y) = _trace(monkeypatch)
    loop, last = calls[:-4], calls[-4:]
    assert not any(c["grad"] for c in loop) and all(c["grad"] for c in last)   # only the last four softmins are differentiated
    assert last[0]["eps"] == loop[-1]["eps"]                                    # at the final temperature
    # "same old potentials": the four inputs of the last extrapolation are exactly the averaged outputs of the last loop
    # round -- in particular the third one still uses the OLD b_x although a new a_x has just been computed
    e = last[0]["eps"]
    lam = G.dampening(e, 0.25)
    la, lb = G.log_weights(a), G.log_weights(b)
    prev = calls[-8:-4]
    C = {k: G.cost_routines[2](u, v) for k, (u, v) in dict(xx=(x, x), yy=(y, y), yx=(y, x), xy=(x, y)).items()}
    sm = lambda Cm, h: -e * (h[:, None, :] - Cm / e).logsumexp(2)
    new = [lam * sm(C["xx"], prev[0]["h"]), lam * sm(C["yy"], prev[1]["h"]), lam * sm(C["yx"], prev[2]["h"]), lam * sm(C["xy"], prev[3]["h"])]
    old = [(prev[0]["h"] - la) * e, (prev[1]["h"] - lb) * e, (prev[3]["h"] - lb) * e, (prev[2]["h"] - la) * e]   # a_x, b_y, a_y, b_x
    avg = [0.5 * (o + n) for o, n in zip(old, new)]
    want = [la + avg[0] / e, lb + avg[1] / e, la + avg[3] / e, lb + avg[2] / e]
    for c, w in zip(last, want):
        assert torch.allclose(c["h"], w.detach(), rtol=0, atol=1e-8)
    out.sum().backward()
    assert x.grad is not None and torch.isfinite(x.grad).all()


def test_v_unbalanced_weight_is_rho_plus_half_eps():
    x = torch.tensor([1.0, 2.0], dtype=torch.float64)
    assert torch.equal(G._unbalanced_weight(1e-6, 0.25, x), (0.25 + 0.5e-6) * x)
    # closed form, equal-mass Diracs: 2 rho (1 - exp(-C / 2 rho)) -- holds only with this weight and lambda = 1/(1 + eps/rho)
    a = torch.tensor([[0.7]], dtype=torch.float64)
    p0, p1 = torch.zeros(1, 1, 2, dtype=torch.float64), torch.tensor([[[0.2, 0.0]]], dtype=torch.float64)
    F = G.SamplesLoss("sinkhorn", p=2, blur=0.001, scaling=0.5, reach=0.5)(a, p0, a, p1)
    C = 0.02
    assert abs(float(F) - 0.7 * 2 * 0.25 * (1 - np.exp(-C / 0.5))) < 1e-8


def test_vi_cost_lookup_with_float_p_and_python_float_eps():
    assert G.cost_routines[2.0] is G.cost_routines[2] and G.cost_routines[1.0] is G.cost_routines[1]
    a, x, b, y = _problem()
    _d, eps, eps_s, rho = G.scaling_parameters(x, y, 2.0, 0.001, 0.5, None, 0.5)
    assert isinstance(eps, float) and eps == 0.001 ** 2.0 and rho == 0.5 ** 2.0 and all(isinstance(e, float) for e in eps_s)


def test_vii_backend_auto_is_tensorized_up_to_5000_squared():
    L = G.SamplesLoss("sinkhorn", p=2, blur=0.05)
    x = torch.zeros(1, 5001, 1)
    with pytest.raises(NotImplementedError, match="KeOps"):
        L(x, torch.zeros(1, 5000, 1))        # 5001 * 5000 > 5000^2: upstream leaves the tensorized backend here
    # 4096 x 4096 (BASELINE configs[3]) is the largest BASELINE shape and stays tensorized: 4096^2 < 5000^2
    assert 4096 * 4096 <= 5000 ** 2


def test_viii_weights_are_used_as_given():
    """Neither normalised to sum 1 nor checked: scaling both measures by c scales the balanced loss by c."""
    a, x, b, y = _problem(B=1)
    L = G.SamplesLoss("sinkhorn", p=2, blur=0.01, scaling=0.5, reach=None)
    f1, f3 = L(a, x, b, y), L(3 * a, x, 3 * b, y)
    assert float(a.sum()) != 1.0 and abs(float(f3) - 3 * float(f1)) < 1e-9 * abs(float(f1)) + 1e-12
    # zero-mass cells: log-weight -100000, no contribution to the transport -- as long as the cell lies inside the bounding
    # box: max_diameter runs over ALL points, massless or not, and the diameter sets the epsilon schedule (fact 4 of SURVEY.md)
    a0 = torch.cat([a, torch.zeros(1, 1, dtype=a.dtype)], 1)
    x0 = torch.cat([x, x.mean(1, keepdim=True)], 1)
    assert abs(float(L(a0, x0, b, y)) - float(f1)) < 1e-9 and float(G.log_weights(a0)[0, -1]) == -100000.0
    x_far = torch.cat([x, torch.full((1, 1, 2), 0.9, dtype=x.dtype)], 1)
    assert abs(float(L(a0, x_far, b, y)) - float(f1)) > 1e-3       # a massless outlier still changes the schedule, hence the value

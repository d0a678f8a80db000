"""Host-side logic (no GPU): fitting, sampling order, lattice geometry, obs filtering.

Pins: /root/reference/tests/test_model.py:79-120 (fit known answer, rtol 1e-3),
tests/test_probabilityfield.py:24-53 (constructor errors, expand geometry, distancesquared),
and fixtures from the executed reference (tests/golden/fit.npz, expand.npz, sto_*.npz)."""
import numpy as np
import pytest

from onekapy_b200 import problems
from onekapy_b200.host.model import Model, RangeError, fit_batch, construct_fit_batch
from onekapy_b200.host.probabilityfield import ProbabilityField
from onekapy_b200.host.stochastic import (sample_realizations, generate_random_variate, compute_variate_mean,
                                          isdistribution, DistributionError)
from onekapy_b200.host.utilities import filter_obs
from onekapy_b200.lattice import LatticeGeom, final_geometry
from helpers import scal


def test_fit_known_answer():
    """reference tests/test_model.py:79-120: 25 observations, MATLAB answer at rtol 1e-3."""
    wells = [(100.0, 200.0, 1.0, 1000.0), (200.0, 100.0, 1.0, 1000.0)]
    mo = Model(500.0, 1.0, 0.25, 100.0, wells, 0.0, 0.0, np.array([1.0, 1.0, 1.0, 1.0, 1.0, 500.0]))
    # the reference's 25 observations (test_model.py:96-120) are synthetic heads of this very model;
    # regenerate them from the quoted coefficients: z = base + head(Phi), 1 m std
    ev_true = np.array([0.9916, 0.9956, 0.9422, 171.85, 165.8, 9667.8])
    rng = np.random.default_rng(0)
    pts = rng.uniform(20, 180, size=(25, 2))
    from onekapy_b200.host.model import wells_potential
    dx, dy = pts[:, 0] - 0.0, pts[:, 1] - 0.0
    phi = (ev_true[0] * dx ** 2 + ev_true[1] * dy ** 2 + ev_true[2] * dx * dy + ev_true[3] * dx + ev_true[4] * dy
           + ev_true[5] + wells_potential(pts, np.array([[100.0, 200.0], [200.0, 100.0]]), np.array([1000.0, 1000.0])))
    head = (phi + 0.5 * 1.0 * 100.0 ** 2) / (1.0 * 100.0)
    obs = [(p[0], p[1], 500.0 + h, 1.0) for p, h in zip(pts, head)]
    ev, cov = mo.fit_regional_flow(obs, 0.0, 0.0)
    assert np.allclose(ev[:, 0], ev_true, rtol=1e-3)
    assert cov.shape == (6, 6) and np.allclose(cov, cov.T, rtol=1e-6)


@pytest.mark.parametrize("name", ["basic", "perham", "long_prairie"])
def test_fit_matches_reference(golden, name):
    g = golden("fit.npz")
    pb = problems.load(name)
    wxy = np.array([[w[0], w[1]] for w in pb["wells"]])
    xt, yt = pb["wells"][pb["target"]][0:2]
    obs = g[name + "_obs"]
    ev, cov = fit_batch(obs, xt, yt, pb["base"], wxy, g[name + "_q"], g[name + "_k"], g[name + "_H"])
    assert np.allclose(ev, g[name + "_coef_ev"], rtol=1e-9, atol=0)
    assert np.allclose(cov, g[name + "_coef_cov"], rtol=1e-6, atol=0)
    ev2, cov2 = fit_batch(obs, xt, yt, pb["base"], wxy, g[name + "_q"], g[name + "_k"], g[name + "_H"], method="qr")
    assert np.allclose(ev2, g[name + "_coef_ev"], rtol=1e-6, atol=0)
    ref_cov = g[name + "_coef_cov"]
    sd = np.sqrt(np.einsum("rii->ri", ref_cov))
    assert np.all(np.abs(cov2 - ref_cov) <= 1e-4 * sd[:, :, None] * sd[:, None, :])   # relative to sqrt(c_ii c_jj)
    # Model.fit_regional_flow, one realization
    wells = [(w[0], w[1], w[2], q) for w, q in zip(pb["wells"], g[name + "_q"][0])]
    mo = Model(pb["base"], g[name + "_k"][0], 0.2, g[name + "_H"][0], wells)
    e1, c1 = mo.fit_regional_flow([tuple(o) for o in obs], xt, yt)
    assert np.allclose(e1[:, 0], g[name + "_coef_ev"][0], rtol=1e-9)
    assert (mo.xo, mo.yo) == (xt, yt) and np.array_equal(mo.coef, e1[:, 0])


def test_fit_below_base_raises():
    with pytest.raises(RangeError):                       # model.py:558-559
        construct_fit_batch(np.array([[0.0, 0.0, 5.0, 1.0]]), 0, 0, 10.0, np.zeros((0, 2)), np.zeros((1, 0)), [1.0], [5.0])


def test_filter_obs_matches_reference_rows(golden):
    for name in ["basic", "perham", "long_prairie"]:
        pb = problems.load(name)
        obs = filter_obs(pb["observations"], pb["wells"], pb["buffer"])
        assert np.allclose(np.array(obs, dtype=float), golden("fit.npz")[name + "_obs"], rtol=0, atol=0)


def test_filter_obs_merges_duplicates():
    obs = [(0.0, 0.0, 10.0, 1.0), (0.5, 0.0, 12.0, 2.0), (50.0, 0.0, 11.0, 1.0), (200.0, 200.0, 9.0, 1.0)]
    out = filter_obs(obs, [(200.0, 205.0, 0.1, 10.0)], 10.0)
    assert len(out) == 2
    num, den = 10.0 / 1.0 + 12.0 / 4.0, 1.0 + 0.25
    assert out[0] == (0.0, 0.0, num / den, np.sqrt(1 / den))
    assert out[1] == (50.0, 0.0, 11.0, 1.0)


def test_sampling_reproduces_reference_rows(golden):
    """Same RNG call order as oneka/stochastic.py:224-241 -> identical parameter rows."""
    for fixture, name, nreal, seed in [("sto_basic.npz", "basic", 6, 7), ("sto_perham.npz", "perham", 2, 9)]:
        g = golden(fixture)
        pb = problems.load(name)
        xt, yt = pb["wells"][pb["target"]][0:2]
        obs = filter_obs(pb["observations"], pb["wells"], pb["buffer"])
        np.random.seed(seed)
        par, ev, cov = sample_realizations(nreal, pb["base"], pb["c_dist"], pb["p_dist"], pb["t_dist"], pb["wells"],
                                           obs, xt, yt, rng=np.random.default_rng(seed), log_rows=False)
        assert np.array_equal(par.q, g["q"]) and np.array_equal(par.cond, g["k"])
        assert np.array_equal(par.poro, g["n"]) and np.array_equal(par.thick, g["H"])
        assert np.allclose(par.coef, g["coef"], rtol=1e-7, atol=0)


def test_vectorised_sampling_is_the_scalar_stream():
    """sample_realizations draws whole tables at once; the rows must be the ones the reference's loop
    (oneka/stochastic.py:220-241: per realization one variate per well, k, n, H from np.random, then one
    multivariate_normal) produces from the same seeds: q, k, n, H and the fit bit for bit, A..F to 1e-12 standard deviations."""
    pb = problems.load("basic")
    wells = [(w[0], w[1], w[2], d) for w, d in zip(pb["wells"], [(500.0, 1000.0, 1800.0), (-300.0, 250.0)])]
    wells.append((wells[0][0] + 900.0, wells[0][1] - 400.0, 0.3, 125.0))                # Dirac: consumes nothing
    xt, yt = wells[0][0:2]
    obs = filter_obs(pb["observations"], wells, pb["buffer"])
    c_dist, p_dist, t_dist = (10.0, 50.0, 90.0), 0.2, (20.0, 25.0)
    R = 37
    np.random.seed(11)
    par, ev, cov = sample_realizations(R, pb["base"], c_dist, p_dist, t_dist, wells, obs, xt, yt,
                                       rng=np.random.default_rng(5), fit_method="lstsq", log_rows=False)
    after = np.random.random_sample()                     # the global stream advanced by exactly the same amount
    np.random.seed(11)
    g = np.random.default_rng(5)
    wxy = np.array([[w[0], w[1]] for w in wells])
    for i in range(R):
        q = [generate_random_variate(w[3]) for w in wells]
        k, n, H = generate_random_variate(c_dist), generate_random_variate(p_dist), generate_random_variate(t_dist)
        assert list(par.q[i]) == q and (par.cond[i], par.poro[i], par.thick[i]) == (k, n, H)
        mo = Model(pb["base"], k, n, H, [(w[0], w[1], w[2], qq) for w, qq in zip(wells, q)])
        e, c = mo.fit_regional_flow(obs, xt, yt)
        assert np.array_equal(e[:, 0], ev[i]) and np.array_equal(c, cov[i])
        sd = np.sqrt(np.diag(c))                          # same normals, same factor; BLAS summation order may differ
        assert np.all(np.abs(g.multivariate_normal(e[:, 0], c) - par.coef[i]) <= 1e-12 * sd)
    assert after == np.random.random_sample()
    # chunking does not change the stream
    import onekapy_b200.host.stochastic as hs
    np.random.seed(11)
    old, hs.SAMPLE_CHUNK = hs.SAMPLE_CHUNK, 8
    try:
        par2, _, _ = sample_realizations(R, pb["base"], c_dist, p_dist, t_dist, wells, obs, xt, yt,
                                         rng=np.random.default_rng(5), fit_method="lstsq", log_rows=False)
    finally:
        hs.SAMPLE_CHUNK = old
    assert np.array_equal(par2.q, par.q) and np.array_equal(par2.thick, par.thick) and np.array_equal(par2.coef, par.coef)
    # the fast fit gives the same rows to rounding
    np.random.seed(11)
    par3, ev3, cov3 = sample_realizations(R, pb["base"], c_dist, p_dist, t_dist, wells, obs, xt, yt,
                                          rng=np.random.default_rng(5), fit_method="qr", log_rows=False)
    assert np.array_equal(par3.q, par.q) and np.allclose(ev3, ev, rtol=1e-7, atol=0)
    sd = np.sqrt(np.einsum("rii->ri", cov))
    assert np.all(np.abs(par3.coef - par.coef) <= 1e-6 * sd)
    # the errors of the scalar calls
    with pytest.raises(DistributionError):
        sample_realizations(2, pb["base"], (1.0, 2.0, 3.0, 4.0), p_dist, t_dist, wells, obs, xt, yt)
    with pytest.raises(ValueError):
        sample_realizations(2, pb["base"], (3.0, 2.0, 4.0), p_dist, t_dist, wells, obs, xt, yt)
    empty, _, _ = sample_realizations(0, pb["base"], c_dist, p_dist, t_dist, wells, obs, xt, yt)
    assert len(empty) == 0


@pytest.mark.parametrize("thick", [(20.0, 25.0), (130.0, 150.0), (60.0, 140.0)])
def test_fit_qr_regimes_match_lstsq(thick):
    """method="qr": shared-QR rows (all observations confined / all unconfined, oneka/model.py:551-556) and the
    stacked factorisation of mixed rows against the reference's per-realization lstsq + inv."""
    pb = problems.load("basic")                           # heads 80..120 m above the base
    rng = np.random.default_rng(2)
    R = 64
    wxy = np.array([[w[0], w[1]] for w in pb["wells"]])
    xt, yt = wxy[pb["target"]]
    obs = np.array(filter_obs(pb["observations"], pb["wells"], pb["buffer"]), dtype=float)
    q = rng.uniform(200.0, 2000.0, size=(R, len(wxy)))
    k = rng.uniform(5.0, 80.0, size=R)
    H = rng.uniform(thick[0], thick[1], size=R)
    head = obs[:, 2] - pb["base"]
    nconf = (head[None, :] >= H[:, None]).sum(axis=1)
    if thick[1] < 80:
        assert np.all(nconf == len(obs))
    elif thick[0] > 120:
        assert np.all(nconf == 0)
    else:
        assert np.any((nconf > 0) & (nconf < len(obs))) and np.any(nconf == len(obs))
    ev, cov = fit_batch(obs, xt, yt, pb["base"], wxy, q, k, H, method="lstsq")
    ev2, cov2, fac = fit_batch(obs, xt, yt, pb["base"], wxy, q, k, H, method="qr", with_factor=True)
    sd = np.sqrt(np.einsum("rii->ri", cov))
    assert np.all(np.abs(ev2 - ev) <= 1e-7 * sd + 1e-9 * np.abs(ev))
    assert np.all(np.abs(cov2 - cov) <= 1e-6 * sd[:, :, None] * sd[:, None, :])
    assert np.all(np.abs(np.swapaxes(fac, 1, 2) @ fac - cov2) <= 1e-9 * sd[:, :, None] * sd[:, None, :])   # F^T F = cov


def test_variates():
    assert generate_random_variate(3.5) == 3.5
    assert compute_variate_mean(0.2) == 0.2
    assert compute_variate_mean((1.0, 3.0)) == 2.0
    assert compute_variate_mean((1.0, 2.0, 6.0)) == 3.0
    np.random.seed(1)
    assert 1.0 <= generate_random_variate((1.0, 2.0)) <= 2.0
    assert 1.0 <= generate_random_variate((1.0, 1.5, 2.0)) <= 2.0
    with pytest.raises(DistributionError):
        generate_random_variate((1.0, 2.0, 3.0, 4.0))
    assert isdistribution(0.3, 0, 1) and isdistribution((0.1, 0.2), 0, 1) and isdistribution([0.1, 0.2, 0.3], 0, 1)
    assert not isdistribution((0.3, 0.2), 0, 1) and not isdistribution("x", 0, 1) and not isdistribution(2.0, 0, 1)


def test_probabilityfield_constructor_and_expand():
    with pytest.raises(RangeError):                       # reference tests/test_probabilityfield.py:24-30
        ProbabilityField(-1.0, 1.0)
    with pytest.raises(RangeError):
        ProbabilityField(1.0, 0.0)
    pf = ProbabilityField(1.0, 1.0)                       # :33-49
    pf.expand(100, 200, 50, 100)
    assert (pf.nrows, pf.ncols) == (53, 103)
    assert (pf.xmin, pf.xmax, pf.ymin, pf.ymax) == (99, 201, 49, 101)
    pf.expand(110, 120, 60, 70)
    assert (pf.nrows, pf.ncols) == (53, 103)
    assert ProbabilityField.distancesquared(0, 1, 1, 0, 0, 0) == 0.5      # :52-53
    with pytest.raises(RangeError):
        pf.expand(2, 1, 0, 0)


def test_expand_sequences_and_content(golden):
    g = golden("expand.npz")
    for s in range(int(g["nspec"])):
        dx, dy, xo, yo = g["spec%d" % s]
        pf = ProbabilityField(dx, dy, xo, yo)
        if pf.nrows:
            pf.pgrid[1, 1] = 7.0
            pf.rgrid[1, 1] = True
        for box, ref in zip(g["boxes%d" % s], g["geom%d" % s]):
            pf.expand(*box)
            assert [pf.xmin, pf.xmax, pf.ymin, pf.ymax, pf.nrows, pf.ncols] == list(ref)
            assert pf.pgrid.shape == (pf.nrows, pf.ncols) == pf.rgrid.shape
        if not np.isnan(xo):
            # the marked node is still the node at (xo, yo)
            j = int(round((xo - pf.xmin) / dx))
            i = int(round((yo - pf.ymin) / dy))
            assert pf.pgrid[i, j] == 7.0 and pf.rgrid[i, j] and pf.pgrid.sum() == 7.0


def test_distancesquared_matches_reference(golden):
    g = golden("distsq.npz")
    got = np.array([ProbabilityField.distancesquared(*[np.float64(t) for t in r]) for r in g["args"]])
    assert np.array_equal(got, g["d2"], equal_nan=True)


def test_register_reset():
    pf = ProbabilityField(1.0, 1.0, 0.0, 0.0)
    pf.rgrid[0, 1] = True
    pf.register(0.5)
    pf.rgrid[0, 1] = True
    pf.rgrid[2, 2] = True
    pf.reset()
    pf.register(0.25)
    assert pf.total_weight == 0.75 and pf.pgrid[0, 1] == 0.5 and pf.pgrid.sum() == 0.5 and not pf.rgrid.any()


def test_final_geometry_equals_sequential_expansion(golden):
    """One expansion to the union box == the reference's trace-by-trace expansions."""
    for name in ["det_basic.npz", "sto_basic.npz", "sto_perham.npz", "unc_basic.npz", "fwd_basic.npz"]:
        g = golden(name)
        s = scal(g)
        v = g["verts"]
        bbox = (v[:, 0].min(), v[:, 0].max(), v[:, 1].min(), v[:, 1].max())
        fg = final_geometry(s["spacing"], s["spacing"], s["xt"], s["yt"], bbox)
        ref = g["auto_geom"]
        assert [fg.xmin, fg.xmax, fg.ymin, fg.ymax, fg.nrows, fg.ncols] == list(ref[[0, 1, 2, 3, 6, 7]])
        assert fg.strictly_contains(bbox)
        big = LatticeGeom.anchored(s["spacing"], s["spacing"], s["xt"], s["yt"]).expanded(*g["lattice"])
        fx = g["fixed_geom"]
        assert [big.xmin, big.xmax, big.ymin, big.ymax, big.nrows, big.ncols] == list(fx[[0, 1, 2, 3, 6, 7]])
        i0, j0 = big.offset_of(fg)
        assert big.xmin + j0 * s["spacing"] == fg.xmin and big.ymin + i0 * s["spacing"] == fg.ymin


def test_clip_windows_replay_reference_expansion(golden):
    """lattice.clip_windows (closed form + running union) == replaying expand() trace by trace, which is what
    the reference does (probabilityfield.py:335) -- checked on the executed reference's own traces."""
    import torch
    from helpers import traces_of
    from onekapy_b200.lattice import clip_windows
    for name in ["sto_basic.npz", "sto_perham.npz", "fwd_basic.npz", "det_basic.npz"]:
        g = golden(name)
        s = scal(g)
        tr = traces_of(g)
        bb = np.array([[t[:, 0].min(), t[:, 0].max(), t[:, 1].min(), t[:, 1].max()] for t in tr])
        R = len(g["k"])
        base = LatticeGeom.anchored(s["spacing"], s["spacing"], s["xt"], s["yt"])
        final = base.expanded(bb[:, 0].min(), bb[:, 1].max(), bb[:, 2].min(), bb[:, 3].max())
        got = clip_windows(torch, base, final, torch.from_numpy(bb).reshape(R, s["P"], 4)).reshape(-1, 4).numpy()
        pf = ProbabilityField(s["spacing"], s["spacing"], s["xt"], s["yt"])
        for n, b in enumerate(bb):
            pf.expand(*b)
            i0, j0 = final.offset_of(LatticeGeom.of_field(pf))
            assert list(got[n]) == [j0, j0 + pf.ncols, i0, i0 + pf.nrows], (name, n)
        assert (pf.nrows, pf.ncols) == (final.nrows, final.ncols)
        # a prior box (earlier ranks) only ever widens the windows
        prior = (bb[:, 0].min() + 50, bb[:, 1].max() - 50, bb[:, 2].min() + 50, bb[:, 3].max() - 50)
        wide = clip_windows(torch, base, final, torch.from_numpy(bb).reshape(R, s["P"], 4), prior).reshape(-1, 4).numpy()
        assert np.all(wide[:, 0] <= got[:, 0]) and np.all(wide[:, 1] >= got[:, 1])


def test_empty_realization_params():
    from onekapy_b200.engine import RealizationParams
    p = RealizationParams(q=np.zeros((0, 29)), cond=[], poro=[], thick=[], coef=np.zeros((0, 6)))
    assert len(p) == 0 and p.q.shape == (0, 29) and p.coef.shape == (0, 6)
    full = RealizationParams(q=np.ones((4, 3)), cond=np.ones(4), poro=np.ones(4), thick=np.ones(4), coef=np.ones((4, 6)))
    assert full.slice(2, 2).q.shape == (0, 3) and full.slice(1, 4, 2).q.shape == (2, 3)


def test_chunked_sampling_concatenates_to_the_monolithic_rows():
    """host.stochastic.iter_realizations (what the drop-in call streams to the GPU chunk by chunk) draws the rows one monolithic
    sample_realizations would: both RNG streams are consumed realization by realization."""
    from onekapy_b200 import problems
    from onekapy_b200.host.stochastic import iter_realizations, sample_realizations
    from onekapy_b200.host.utilities import filter_obs
    pb = problems.load("perham")
    obs = filter_obs(pb["observations"], pb["wells"], pb["buffer"])
    xt, yt = pb["wells"][pb["target"]][0:2]
    args = (pb["base"], pb["c_dist"], pb["p_dist"], pb["t_dist"], pb["wells"], obs, xt, yt)
    for method in ("lstsq", "qr"):
        np.random.seed(77)
        whole, ev, cov = sample_realizations(200, *args, rng=np.random.default_rng(5), fit_method=method, log_rows=False)
        np.random.seed(77)
        parts = list(iter_realizations(200, *args, rng=np.random.default_rng(5), fit_method=method, log_rows=False, chunk=37))
        assert [len(p) for p, _, _ in parts] == [37] * 5 + [15]
        for name in ("q", "cond", "poro", "thick"):
            assert np.array_equal(np.concatenate([getattr(p, name) for p, _, _ in parts]), getattr(whole, name)), name
        coef = np.concatenate([p.coef for p, _, _ in parts])
        scale = np.abs(whole.coef).max(axis=0)
        assert np.all(np.abs(coef - whole.coef) <= 1e-9 * scale), (method, np.abs(coef - whole.coef).max(axis=0) / scale)
        assert np.allclose(np.concatenate([e for _, e, _ in parts]), ev, rtol=1e-9, atol=0)

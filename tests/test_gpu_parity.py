"""GPU parity tests (run on the B200 box: pytest -m gpu).  Everything goes through the C ABI.

Bars
  * rasterisation (insert/register on given traces): BIT-EXACT against the executed reference's
    grids (tests/golden) and against the oracle on random tracks;
  * traces: same number of vertices per path, every vertex within 1e-6 relative position
    (north_star's tolerance; observed ~1e-12);
  * fused capture on a fixed lattice: attempts/steps totals equal the oracle's, count grid equal
    cell by cell (the differing-cell fraction is asserted to be 0 and printed);
  * drop-in run (auto-expanding reference): final geometry identical; the grid is a superset of the
    reference's with a differing-cell fraction below 2e-2 on these tiny runs (order-dependent clipping).
"""
import numpy as np
import pytest

from helpers import scal, geom, traces_of

pytestmark = pytest.mark.gpu

CAPTURES = ["det_basic.npz", "sto_basic.npz", "sto_perham.npz", "unc_basic.npz", "fwd_basic.npz"]
POS_RTOL = 1e-6       # BASELINE.json north_star: endpoints within 1e-6 relative position


@pytest.fixture(scope="module", params=["auto", "heavy", "plain"])
def eng(request):
    """Every test of this module runs three times: with the rasteriser flavour chosen by lattice (the default), and with each of
    the two flavours forced (oneka_set_raster_mode) -- both must reproduce the reference's grids on every configuration."""
    from onekapy_b200.engine import Engine
    e = Engine(0)
    e.set_raster_mode(request.param)
    yield e
    e.close()


def spec_of(g):
    from onekapy_b200.engine import FlowSpec, RealizationParams
    s = scal(g)
    spec = FlowSpec(well_xy=g["wells_xyr"][:, :2].copy(), xtarget=s["xt"], ytarget=s["yt"], rtarget=s["rt"],
                    npaths=s["P"], duration=s["duration"], base=s["base"], spacing=s["spacing"], umbra=s["umbra"],
                    confined=s["confined"], tol=s["tol"], maxstep=s["maxstep"])
    par = RealizationParams(q=g["q"], cond=g["k"], poro=g["n"], thick=g["H"], coef=g["coef"])
    return s, spec, par


def fixed_geom(g, s):
    from onekapy_b200.lattice import LatticeGeom
    return LatticeGeom.anchored(s["spacing"], s["spacing"], s["xt"], s["yt"]).expanded(*g["lattice"])


# ---------------------------------------------------------------------------------------------------
def test_library_loaded_and_device(eng):
    import torch
    assert torch.cuda.get_device_capability(0)[0] == 10
    tf, ms = eng.fp64_probe(1 << 14)
    assert tf > 5.0, "FP64 probe reports %.2f TFLOP/s" % tf


def test_model_known_answers(eng):
    """reference tests/test_model.py:45-76 through the CUDA field functions."""
    from onekapy_b200.host.model import Model
    mo = Model(500.0, 1.0, 0.25, 100.0, [(100.0, 200.0, 1.0, 1000.0), (200.0, 100.0, 1.0, 1000.0)], 0.0, 0.0,
               np.array([1.0, 1.0, 1.0, 1.0, 1.0, 500.0]))
    assert np.isclose(mo.compute_potential(100.0, 100.0), 32165.8711977589, rtol=1.0e-6)
    assert np.isclose(mo.compute_head(100.0, 100.0), 371.658711977589, rtol=1.0e-6)
    assert np.allclose(mo.compute_discharge(120.0, 160.0), [-401.318309886184, -438.771830796713], rtol=1.0e-6)
    assert np.allclose(mo.compute_velocity(100.0, 100.0), [-11.976338022763, -11.976338022763], rtol=1.0e-6)
    assert np.allclose(mo.compute_velocity(120.0, 160.0), [-16.052732395447, -17.550873231869], rtol=1.0e-6)
    assert np.allclose(mo.compute_velocity_confined(120.0, 160.0), [-16.052732395447, -17.550873231869], rtol=1.0e-6)


def test_model_points_vs_reference(eng, golden):
    from onekapy_b200.host.model import Model, AquiferError
    g = golden("model_points.npz")
    wells = [tuple(w) for w in g["wells"]]
    for par, pts, ref in zip(g["par"], g["pts"], g["out"]):
        base, k, n, H, xo, yo = par[:6]
        mo = Model(base, k, n, H, wells, xo, yo, par[6:])
        out = mo.evaluate(pts)
        assert np.array_equal(np.isnan(out), np.isnan(ref))
        assert np.allclose(out[:, :3], ref[:, :3], rtol=1e-12, atol=0)                 # potential, discharge: plain FP64 formulas
        assert np.allclose(out[:, 5], ref[:, 5], rtol=1e-12, atol=0, equal_nan=True)    # head
        # velocities go through field_feval (the tracker's function): Newton reciprocal, <= ~1e-12 per well term
        scale = np.abs(ref[:, [3, 4, 3, 4]]).max(axis=1, keepdims=True)
        assert np.all(np.abs(out[:, 3:5] - ref[:, 3:5]) <= 1e-10 * scale[:, :1])
        ok = ~np.isnan(ref[:, 6])
        assert np.all(np.abs(out[ok, 6:8] - ref[ok, 6:8]) <= 1e-10 * np.abs(ref[ok, 6:8]).max(axis=1, keepdims=True))
        dry = np.where(np.isnan(ref[:, 5]))[0]
        for i in dry[:2]:
            with pytest.raises(AquiferError):
                mo.compute_head(*pts[i])


# ---------------------------------------------------------------------------------------------------
def test_raster_insert_fixture_bit_exact(eng, golden):
    """probabilityfield.insert + register on hand-made tracks (zero-length, axis-aligned, lattice-aligned
    exact ties, micro segments, clipped by the lattice edge)."""
    from onekapy_b200.lattice import LatticeGeom
    g = golden("insert.npz")
    tracks = traces_of(g)
    for tag in "abc":
        dx, dy, umbra = g["par_" + tag]
        gm = LatticeGeom.anchored(dx, dy, 60.0, 60.0).expanded(0.0, 200.0, 0.0, 200.0)
        ref = geom(g, "fixed_%s_" % tag)
        assert (gm.xmin, gm.ymin, gm.nrows, gm.ncols) == (ref["xmin"], ref["ymin"], ref["nrows"], ref["ncols"])
        counts = eng.raster_traces(gm, umbra, tracks, g["real_of"], 2)
        assert np.array_equal(counts, g["fixed_%s_counts" % tag].astype(np.uint32)), tag


def test_probabilityfield_methods_on_gpu(eng, golden):
    """ProbabilityField.insert / rasterize / register (drop-in class) reproduce the reference's
    auto-expanding and fixed results on the hand-made tracks."""
    from onekapy_b200.host.probabilityfield import ProbabilityField
    g = golden("insert.npz")
    tracks = traces_of(g)
    real_of = g["real_of"]
    dx, dy, umbra = g["par_b"]
    pf = ProbabilityField(dx, dy, 60.0, 60.0)
    for r in (0, 1):
        for t, rr in zip(tracks, real_of):
            if rr == r:
                pf.rasterize(list(t[:, 0]), list(t[:, 1]), umbra)
        pf.register(1.0)
    ref = geom(g, "auto_b_")
    assert (pf.xmin, pf.xmax, pf.ymin, pf.ymax, pf.nrows, pf.ncols) == (ref["xmin"], ref["xmax"], ref["ymin"], ref["ymax"], ref["nrows"], ref["ncols"])
    assert np.array_equal(pf.pgrid, g["auto_b_counts"].astype(float)) and pf.total_weight == 2.0
    pf = ProbabilityField(dx, dy, 60.0, 60.0)
    pf.expand(0.0, 200.0, 0.0, 200.0)
    t = tracks[0]
    for i in range(len(t) - 1):
        pf.insert(t[i, 0], t[i, 1], t[i + 1, 0], t[i + 1, 1], umbra)
    from oracle import oracle as O
    of = O.Field(dx, dy, 60.0, 60.0)
    of.expand(0.0, 200.0, 0.0, 200.0)
    for i in range(len(t) - 1):
        of.insert(t[i, 0], t[i, 1], t[i + 1, 0], t[i + 1, 1], umbra)
    assert np.array_equal(pf.rgrid, of.rgrid)


@pytest.mark.parametrize("name", CAPTURES)
def test_raster_reference_traces_bit_exact(eng, golden, name):
    """The executed reference's own vertices, rasterised on the GPU == the reference's fixed-lattice grid."""
    g = golden(name)
    s, spec, par = spec_of(g)
    gm = fixed_geom(g, s)
    tr = traces_of(g)
    real_of = np.repeat(np.arange(len(par)), s["P"]).astype(np.int32)
    counts = eng.raster_traces(gm, s["umbra"], tr, real_of, len(par))
    ref = g["fixed_counts"].astype(np.uint32)
    ndiff = np.count_nonzero(counts != ref)
    print("%s: raster-only differing cells %d of %d nonzero" % (name, ndiff, np.count_nonzero(ref)))
    assert ndiff == 0


def test_raster_random_tracks_vs_oracle(eng):
    """2000 random tracks, non-integer lattice, 5 realizations, including degenerate segments."""
    from onekapy_b200.lattice import LatticeGeom
    from oracle import oracle as O
    rng = np.random.default_rng(123)
    dx, dy, umbra = 1.7, 2.3, 6.1
    tracks, real_of = [], []
    for t in range(2000):
        n = int(rng.integers(2, 40))
        step = rng.uniform(-7, 7, size=(n, 2))
        if t % 50 == 0:
            step[n // 2] = 0.0                      # zero-length segment
        if t % 70 == 0:
            step[:, 1] = 0.0                        # horizontal track
        tracks.append(np.cumsum(step, axis=0) + rng.uniform(50, 350, size=2))
        real_of.append(t % 5)
    gm = LatticeGeom.anchored(dx, dy, 200.0, 200.0).expanded(20.0, 380.0, 20.0, 380.0)
    eng.reset_stats()
    counts = eng.raster_traces(gm, umbra, tracks, np.array(real_of, dtype=np.int32), 5)
    st = eng.read_stats()
    of = O.Field(dx, dy, 200.0, 200.0)
    of.expand(20.0, 380.0, 20.0, 380.0)
    for r in range(5):
        for t, rr in zip(tracks, real_of):
            if rr == r:
                for i in range(len(t) - 1):
                    of.insert(t[i, 0], t[i, 1], t[i + 1, 0], t[i + 1, 1], umbra)
        of.register(1.0)
    assert (gm.nrows, gm.ncols) == (of.nrows, of.ncols)
    assert np.array_equal(counts, of.pgrid.astype(np.uint32))
    print("random tracks: %d segments, %d cells re-tested in exact FP64, %d clipped" % (st["steps"], st["exact_tests"], st["n_clipped"]))


@pytest.mark.parametrize("dx,dy,umbra,step,snap", [
    (4.0, 4.0, 8.0, 9.0, 0.0),       # perham-like 7 x 7 windows
    (2.0, 2.0, 10.0, 18.0, 0.0),     # basic_deterministic-like 19 x 19 windows
    (0.5, 0.75, 9.0, 6.0, 0.0),      # > 32 columns per window: multi-word row masks
    (10.0, 10.0, 4.0, 18.0, 0.0),    # umbra < spacing: most rows hold 0 or 1 node
    (4.0, 4.0, 8.0, 9.0, 4.0),       # every vertex ON a lattice node: exact ties d2 == umbra^2, axis-aligned and 45-degree segments
    (1.0, 3.0, 5.0, 0.02, 0.0),      # very short segments
    (3.0, 3.0, 6.0, 12.0, 1e-3),     # vertices snapped to a 1 mm grid: near-horizontal / near-vertical segments
])
def test_raster_scanline_configs_vs_oracle(eng, dx, dy, umbra, step, snap):
    """The scan-line rows (analytic interval per row) against the oracle's node-by-node insert, bit for bit."""
    from onekapy_b200.lattice import LatticeGeom
    from oracle import oracle as O
    rng = np.random.default_rng(int(dx * 1000 + umbra * 10 + step))
    tracks = []
    for t in range(600):
        n = int(rng.integers(2, 30))
        ang = rng.uniform(0, 2 * np.pi) + np.cumsum(rng.normal(0, 0.4, size=n))
        if t % 9 == 0:
            ang = np.round(ang / (np.pi / 4)) * (np.pi / 4)              # axis-aligned / diagonal runs
        if t % 13 == 0:
            ang = rng.choice([0.0, np.pi]) + rng.normal(0, 1e-3, size=n)  # almost horizontal
        seg = step * rng.uniform(0.2, 1.0, size=n)
        v = np.cumsum(np.stack([seg * np.cos(ang), seg * np.sin(ang)], axis=1), axis=0) + rng.uniform(60, 140, size=2)
        if snap > 0:
            v = np.round(v / snap) * snap
        tracks.append(v)
    gm = LatticeGeom.anchored(dx, dy, 100.0, 100.0).expanded(40.0, 160.0, 40.0, 160.0)
    real_of = (np.arange(len(tracks)) % 3).astype(np.int32)
    eng.reset_stats()
    counts = eng.raster_traces(gm, umbra, tracks, real_of, 3)
    st = eng.read_stats()
    of = O.Field(dx, dy, 100.0, 100.0)
    of.expand(40.0, 160.0, 40.0, 160.0)
    for r in range(3):
        for t, rr in zip(tracks, real_of):
            if rr == r:
                for i in range(len(t) - 1):
                    of.insert(t[i, 0], t[i, 1], t[i + 1, 0], t[i + 1, 1], umbra)
        of.register(1.0)
    ref = of.pgrid.astype(np.uint32)
    ndiff = np.count_nonzero(counts != ref)
    print("dx %.2f dy %.2f umbra %.1f: %d segments, %d exact re-tests, %d clipped, differing cells %d of %d"
          % (dx, dy, umbra, st["steps"], st["exact_tests"], st["n_clipped"], ndiff, np.count_nonzero(ref)))
    assert ndiff == 0


@pytest.mark.parametrize("dx,dy,umbra", [(4.0, 4.0, 8.0), (1.0, 1.0, 7.5), (3.0, 0.5, 2.0)])
def test_raster_flat_segments_vs_oracle(eng, dx, dy, umbra):
    """Near-horizontal segments of every flatness (slope 0 ... 1/20, both signs, both directions), with the capsule's
    tangent lines placed on, just above and just below lattice rows: the rows that end on the two caps take the
    scan-line path whatever the slope, the rows that cross a straight edge of a flat segment take the node loop."""
    from onekapy_b200.lattice import LatticeGeom
    from oracle import oracle as O
    rng = np.random.default_rng(int(97 * dx + 13 * umbra))
    slopes = [0.0, 1e-15, 1e-12, 1e-9, 1e-7, 1e-6, 1e-5, 1e-4, 1e-3, 3e-3, 1e-2, 1.0 / 64, 1.0 / 63, 0.03, 0.05]
    offs = [0.0, 1e-13, -1e-13, 1e-9, -1e-9, 1e-6, -1e-6, 1e-4, -1e-4, 2e-3, -2e-3, 0.3]
    tracks = []
    for k in range(900):
        n = int(rng.integers(2, 12))
        sl = slopes[k % len(slopes)] * rng.choice([-1.0, 1.0])
        sx = rng.uniform(0.5, 15.0, size=n) * rng.choice([-1.0, 1.0])
        steps = np.stack([sx, sx * sl], axis=1)
        row = 100.0 + dy * int(rng.integers(-8, 8))                     # a lattice row (anchor 100 - dy, spacing dy)
        y0 = row + rng.choice([-umbra, umbra, 0.0]) + offs[k % len(offs)]   # tangent line on / next to that row
        v = np.cumsum(steps, axis=0) + np.array([rng.uniform(70, 130), 0.0])
        v[:, 1] += y0 - v[0, 1]
        tracks.append(v)
    gm = LatticeGeom.anchored(dx, dy, 100.0, 100.0).expanded(20.0, 180.0, 40.0, 160.0)
    real_of = (np.arange(len(tracks)) % 3).astype(np.int32)
    eng.reset_stats()
    counts = eng.raster_traces(gm, umbra, tracks, real_of, 3)
    st = eng.read_stats()
    of = O.Field(dx, dy, 100.0, 100.0)
    of.expand(20.0, 180.0, 40.0, 160.0)
    for r in range(3):
        for t, rr in zip(tracks, real_of):
            if rr == r:
                for i in range(len(t) - 1):
                    of.insert(t[i, 0], t[i, 1], t[i + 1, 0], t[i + 1, 1], umbra)
        of.register(1.0)
    ref = of.pgrid.astype(np.uint32)
    ndiff = np.count_nonzero(counts != ref)
    print("flat segments dx %.1f dy %.1f umbra %.1f: %d segments, %d exact re-tests, differing cells %d of %d"
          % (dx, dy, umbra, st["steps"], st["exact_tests"], ndiff, np.count_nonzero(ref)))
    assert ndiff == 0


# ---------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("name", CAPTURES)
def test_traces_vs_reference(eng, golden, name):
    g = golden(name)
    s, spec, par = spec_of(g)
    dp = eng.upload(spec, par)
    out = eng.trace(spec, dp, max_verts=1024)
    ref = traces_of(g)
    worst = 0.0
    for r in range(len(par)):
        for p in range(s["P"]):
            t = ref[r * s["P"] + p]
            assert out["status"][r, p] == 0
            assert out["nverts"][r, p] == len(t), (name, r, p)
            v = out["verts"][r, p, :len(t)]
            err = np.abs(v - t).max(axis=1) / np.maximum(np.abs(t).max(axis=1), 1.0)
            worst = max(worst, err.max())
    print("%s: max relative vertex error %.3e" % (name, worst))
    assert worst < POS_RTOL


def test_dry_aquifer_status(eng, golden):
    """confined=False forward traces that hit potential <= 0: truncated at the same vertex as the reference."""
    from onekapy_b200.engine import FlowSpec, RealizationParams
    g = golden("unc_dry.npz")
    base, k, n, H, xo, yo = g["par"]
    dur, tol, maxstep = g["scal"]
    spec = FlowSpec(well_xy=g["wells"][:, :2].copy(), xtarget=xo, ytarget=yo, rtarget=0.25, npaths=len(g["starts"]),
                    duration=dur, base=base, spacing=1.0, umbra=1.0, confined=False, tol=tol, maxstep=maxstep)
    par = RealizationParams(q=g["wells"][None, :, 3], cond=[k], poro=[n], thick=[H], coef=g["coef"][None, :])
    dp = eng.upload(spec, par, g["starts"])
    out = eng.trace(spec, dp, max_verts=512)
    for p, (t, dry) in enumerate(zip(traces_of(g), g["terminated"])):
        assert out["status"][0, p] == (1 if dry else 0)
        assert out["nverts"][0, p] == len(t)
        v = out["verts"][0, p, :len(t)]
        assert np.abs(v - t).max() / np.abs(t).max() < POS_RTOL


def test_max_attempt_guard(eng, golden):
    g = golden("sto_basic.npz")
    s, spec, par = spec_of(g)
    spec.max_attempts = 50
    out = eng.trace(spec, eng.upload(spec, par), max_verts=64)
    assert (out["status"] == 2).all() and (out["attempts"] == 50).all()


# ---------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("name", CAPTURES)
def test_fused_capture_fixed_lattice(eng, golden, name):
    from oracle import oracle as O
    from onekapy_b200.engine import start_ring
    g = golden(name)
    s, spec, par = spec_of(g)
    gm = fixed_geom(g, s)
    dp = eng.upload(spec, par)
    counts = eng.new_counts(gm)
    eng.reset_stats()
    pp = eng.capture(spec, dp, gm, counts, per_path=True)
    st = eng.read_stats()
    got = counts.cpu().numpy().view(np.uint32)
    ref = g["fixed_counts"].astype(np.uint32)
    tr = traces_of(g)
    assert np.array_equal(pp["nverts"].cpu().numpy().ravel(), [len(t) for t in tr])
    assert st["steps"] == sum(len(t) - 1 for t in tr) and st["paths"] == len(tr) and st["n_not_ok"] == 0
    end_ref = np.array([t[-1] for t in tr]).reshape(len(par), s["P"], 2)
    rel = (np.abs(pp["end_xy"].cpu().numpy() - end_ref).max(axis=2) / np.abs(end_ref).max(axis=2)).max()
    ndiff = np.count_nonzero(got != ref)
    print("%s: endpoint max rel err %.3e; differing cells %d of %d nonzero (fraction %.2e); exact re-tests %d"
          % (name, rel, ndiff, np.count_nonzero(ref), ndiff / max(1, np.count_nonzero(ref)), st["exact_tests"]))
    assert rel < POS_RTOL
    assert ndiff == 0
    # bounding box reported by the kernel == bounding box of the reference's vertices (to rounding)
    v = g["verts"]
    bb = np.array(st["bbox"])
    assert np.allclose(bb, [v[:, 0].min(), v[:, 0].max(), v[:, 1].min(), v[:, 1].max()], rtol=1e-9)
    # oracle on the same lattice agrees too (this is what the bigger tests below rely on)
    of = O.Field(s["spacing"], s["spacing"], s["xt"], s["yt"])
    of.expand(*g["lattice"])
    res = O.capture(of, 1, spec.well_xy, s["base"], s["xt"], s["yt"], s["confined"], par.q, par.cond, par.poro,
                    par.thick, par.coef, start_ring(s["xt"], s["yt"], s["rt"], s["P"]), s["duration"], s["umbra"],
                    s["tol"], s["maxstep"])
    assert res["attempts"] == st["attempts"]
    assert np.array_equal(of.pgrid.astype(np.uint32), got)


@pytest.mark.parametrize("name", CAPTURES)
def test_run_vs_auto_expanding_reference(eng, golden, name):
    """Engine.run (pilot -> lattice -> capture -> crop) against the reference left auto-expanding."""
    g = golden(name)
    s, spec, par = spec_of(g)
    res = eng.run(spec, par)
    ref = geom(g, "auto_")
    gm = res["geom"]
    assert (gm.xmin, gm.xmax, gm.ymin, gm.ymax, gm.nrows, gm.ncols) == (ref["xmin"], ref["xmax"], ref["ymin"], ref["ymax"], ref["nrows"], ref["ncols"])
    assert res["total_weight"] == ref["total_weight"]
    got, want = res["counts"], g["auto_counts"].astype(np.uint32)
    assert np.all(got >= want)                     # order-dependent clipping only ever drops cells
    ndiff = np.count_nonzero(got != want)
    frac = ndiff / np.count_nonzero(want)
    print("%s: drop-in differing cells %d of %d nonzero (fraction %.2e)" % (name, ndiff, np.count_nonzero(want), frac))
    assert frac < 2e-2


@pytest.mark.parametrize("scheme", ["two_pass", "one_pass", "one_pass_tight_lattice", "one_pass_hint", "one_pass_chunks", "small_run_chunks"])
@pytest.mark.parametrize("name", CAPTURES)
def test_run_exact_equals_auto_expanding_reference(eng, golden, name, scheme):
    """Engine.run_exact reproduces the reference's order-dependent clipping: the grid of the executed reference,
    left auto-expanding, cell for cell -- by the two-pass scheme (small runs: bounding-box pass, then everything clipped)
    and by the one-pass scheme (fused pass with per-path boxes, then only the affected realizations taken out and
    re-rasterised with their windows), the latter also on a lattice estimate that is too small (flagged realizations)
    and on a second call that reuses the lattice."""
    g = golden(name)
    s, spec, par = spec_of(g)
    if scheme == "two_pass":
        res = eng.run_exact(spec, par, two_pass_below=10**9)
        assert res["stats"]["affected_realizations"] == len(par)
    elif scheme == "one_pass":
        res = eng.run_exact(spec, par, two_pass_below=0, reuse_lattice=False)
        assert 1 <= res["stats"]["affected_realizations"] <= len(par) and res["stats"]["rerun_realizations"] == 0
    elif scheme == "one_pass_tight_lattice":
        res = eng.run_exact(spec, par, two_pass_below=0, reuse_lattice=False, pilot=1, pilot_paths=2, margin=0.0)
    elif scheme == "one_pass_chunks":
        # the rows arrive as a generator of chunks (what the drop-in call does to overlap host sampling with the GPU): one
        # fused launch per chunk into the same grid, the lattice estimated from the first chunk alone
        gen = (par.slice(r, min(len(par), r + 2)) for r in range(0, len(par), 2))
        res = eng.run_exact(spec, gen, total=len(par), two_pass_below=0, reuse_lattice=False, per_path=True)
        whole = eng.run_exact(spec, par, two_pass_below=0, reuse_lattice=False, per_path=True)
        assert np.array_equal(res["counts"], whole["counts"]) and res["stats"]["attempts"] == whole["stats"]["attempts"]
        for k in ("end_xy", "nverts", "status"):
            assert np.array_equal(res["per_path"][k], whole["per_path"][k])
    elif scheme == "small_run_chunks":
        gen = (par.slice(r, r + 1) for r in range(len(par)))
        res = eng.run_exact(spec, gen, total=len(par), two_pass_below=10**9)       # materialised, two passes
        with pytest.raises(ValueError):
            eng.run_exact(spec, (par.slice(r, r + 1) for r in range(len(par))), total=len(par) + 1, two_pass_below=0)
    else:
        eng.run_exact(spec, par, two_pass_below=0)
        n0 = eng.launch_count()
        res = eng.run_exact(spec, par, two_pass_below=0)
        assert res["stats"]["rerun_realizations"] == 0
    ref = geom(g, "auto_")
    gm = res["geom"]
    assert (gm.xmin, gm.xmax, gm.ymin, gm.ymax, gm.nrows, gm.ncols) == (ref["xmin"], ref["xmax"], ref["ymin"], ref["ymax"], ref["nrows"], ref["ncols"])
    want = g["auto_counts"].astype(np.uint32)
    ndiff = np.count_nonzero(res["counts"] != want)
    print("%s: exact-clip differing cells %d of %d nonzero" % (name, ndiff, np.count_nonzero(want)))
    assert ndiff == 0 and res["total_weight"] == ref["total_weight"]


def test_run_with_subsampled_pilot(eng, golden):
    """pilot < R: the lattice comes from a subsample + margin; result must not depend on it."""
    g = golden("sto_basic.npz")
    s, spec, par = spec_of(g)
    a = eng.run(spec, par, reuse_lattice=False)
    b = eng.run(spec, par, pilot=2, margin=0.0, pilot_paths=2, reuse_lattice=False)      # margin 0 forces re-tracking
    c = eng.run(spec, par, pilot=3, margin=2.0, reuse_lattice=False)
    assert b["work_geom"] != a["work_geom"] and c["work_geom"] != a["work_geom"]
    assert a["stats"]["rerun_realizations"] == 0 and c["stats"]["rerun_realizations"] == 0
    assert 0 < b["stats"]["rerun_realizations"] <= len(par)
    # a lattice that fits most but not all realizations: only the outliers are tracked again
    d = eng.run(spec, par, pilot=3, margin=0.0, pilot_paths=10, reuse_lattice=False)     # lattice from realizations 0, 2, 4 only
    assert 0 < d["stats"]["rerun_realizations"] <= len(par)
    assert d["geom"] == a["geom"] and np.array_equal(d["counts"], a["counts"])
    # lattice reuse: the second call with the same problem skips the pilot pass (one launch fewer) and agrees
    e1 = eng.run(spec, par)
    n0 = eng.launch_count()
    e2 = eng.run(spec, par)
    assert eng.launch_count() - n0 == 2 and e2["work_geom"] == e1["work_geom"]
    assert e2["geom"] == a["geom"] and np.array_equal(e2["counts"], a["counts"])
    # ... and a stale estimate (here: from a quarter of the duration) only costs a partial re-run
    import copy
    short = copy.copy(spec)
    short.duration = spec.duration / 4
    small = eng.run(short, par)["work_geom"]
    key_dur = [k for k in eng._geom_hint if k[4] == short.duration][0]
    eng._geom_hint[tuple(spec.duration if i == 4 else v for i, v in enumerate(key_dur))] = (small, None)
    f = eng.run(spec, par)
    assert f["stats"]["rerun_realizations"] > 0 and f["geom"] == a["geom"] and np.array_equal(f["counts"], a["counts"])
    # ... ONCE: the estimate that proved too small is replaced by a larger one, the next call fits
    f2 = eng.run(spec, par)
    assert f2["stats"]["rerun_realizations"] == 0 and np.array_equal(f2["counts"], a["counts"])
    # and the cache of estimates is a small LRU, not an ever-growing dict
    for k in range(20):
        eng._hint_put(("dummy", k), (small, None))
    assert len(eng._geom_hint) <= 8 and ("dummy", 19) in eng._geom_hint
    for r in (b, c):
        assert r["geom"] == a["geom"] and np.array_equal(r["counts"], a["counts"])


# ---------------------------------------------------------------------------------------------------
def test_dropin_api(eng, golden):
    """oneka.deterministic / oneka.stochastic / oneka.capturezone keep their signatures."""
    import oneka.deterministic as det
    import oneka.stochastic as sto
    import oneka.capturezone as cz
    from oneka.model import Model
    from oneka.probabilityfield import ProbabilityField
    from onekapy_b200 import problems
    from onekapy_b200.host.utilities import filter_obs

    g = golden("det_basic.npz")
    pb = problems.load("basic_deterministic")
    obs = filter_obs(pb["observations"], pb["wells"], pb["buffer"])
    pf = det.create_deterministic_capturezone(pb["target"], 16, pb["duration"], pb["base"], pb["c_dist"], pb["p_dist"],
                                              pb["t_dist"], pb["wells"], obs, pb["spacing"], pb["umbra"], pb["confined"],
                                              pb["tol"], pb["maxstep"])
    ref = geom(g, "auto_")
    assert isinstance(pf, ProbabilityField)
    assert (pf.xmin, pf.xmax, pf.ymin, pf.ymax, pf.nrows, pf.ncols, pf.total_weight) == (
        ref["xmin"], ref["xmax"], ref["ymin"], ref["ymax"], ref["nrows"], ref["ncols"], 1.0)
    want = g["auto_counts"].astype(float)
    assert pf.pgrid.dtype == np.float64 and pf.rgrid.dtype == bool and not pf.rgrid.any()
    assert np.array_equal(pf.pgrid, want)                 # exact_clip=True (default): the reference's grid, cell for cell
    pf_fast = det.create_deterministic_capturezone(pb["target"], 16, pb["duration"], pb["base"], pb["c_dist"], pb["p_dist"],
                                                   pb["t_dist"], pb["wells"], obs, pb["spacing"], pb["umbra"], pb["confined"],
                                                   pb["tol"], pb["maxstep"], exact_clip=False)
    assert np.all(pf_fast.pgrid >= want) and np.count_nonzero(pf_fast.pgrid != want) / np.count_nonzero(want) < 2e-3

    # compute_capturezone with a closure of the reference's shape (stochastic.py:253-256) on a caller-owned field
    wells = [[w[0], w[1], w[2], q] for w, q in zip(pb["wells"], g["q"][0])]
    mo = Model(pb["base"], g["k"][0], g["n"][0], g["H"][0], wells)
    xt, yt, rt = pb["wells"][0][0:3]
    mo.xo, mo.yo, mo.coef = xt, yt, g["coef"][0]

    def feval(xy):
        Vx, Vy = mo.compute_velocity_confined(xy[0], xy[1])
        return np.array([-Vx, -Vy])

    pf2 = ProbabilityField(pb["spacing"], pb["spacing"], xt, yt)
    cz.compute_capturezone(xt, yt, rt, 16, pb["duration"], pf2, pb["umbra"], 1.0, pb["tol"], pb["maxstep"], feval)
    assert np.array_equal(pf2.pgrid, pf.pgrid) and pf2.total_weight == 1.0
    cz.compute_capturezone(xt, yt, rt, 16, pb["duration"], pf2, pb["umbra"], 0.5, pb["tol"], pb["maxstep"], feval)
    assert pf2.total_weight == 1.5 and pf2.pgrid.max() == 1.5

    # compute_backtrace returns the vertex list
    start = traces_of(g)[3][0]
    v = cz.compute_backtrace(start[0], start[1], pb["duration"], pb["tol"], pb["maxstep"], feval)
    t = traces_of(g)[3]
    assert len(v) == len(t) and np.abs(np.array(v) - t).max() / np.abs(t).max() < POS_RTOL
    # the closure itself still evaluates (through the CUDA field function)
    assert np.allclose(feval(np.array(start)), cz.BacktraceVelocity(mo, True)(np.array(start)))
    with pytest.raises(TypeError):
        cz.compute_backtrace(0.0, 0.0, 1.0, 1.0, 1.0, lambda xy: xy)

    # stochastic driver: signature, type, totals
    pb = problems.load("basic")
    obs = filter_obs(pb["observations"], pb["wells"], pb["buffer"])
    np.random.seed(7)
    pfs = sto.create_stochastic_capturezone(pb["target"], 10, pb["duration"], 6, pb["base"], pb["c_dist"], pb["p_dist"],
                                            pb["t_dist"], pb["wells"], obs, pb["spacing"], pb["umbra"], pb["confined"],
                                            pb["tol"], pb["maxstep"], rng=np.random.default_rng(7))
    gs = golden("sto_basic.npz")
    refs = geom(gs, "auto_")
    assert pfs.total_weight == 6.0 and pfs.pgrid.max() == 6.0
    # same seeded rows as the fixture (coef equal to ~1e-9 relative) -> same geometry, near-identical grid
    assert (pfs.nrows, pfs.ncols, pfs.xmin, pfs.ymin) == (refs["nrows"], refs["ncols"], refs["xmin"], refs["ymin"])
    assert np.count_nonzero(pfs.pgrid != gs["auto_counts"]) / np.count_nonzero(gs["auto_counts"]) < 5e-3


def test_capture_host_and_side_stream(eng, golden):
    """oneka_capture_host (host buffers in, host grid out) and a capture enqueued on a non-default torch stream give
    the same grid, per-path outputs and statistics as the device-resident call."""
    import torch
    g = golden("sto_perham.npz")
    s, spec, par = spec_of(g)
    gm = fixed_geom(g, s)
    ref = g["fixed_counts"].astype(np.uint32)
    counts, st, pp = eng.capture_host(spec, par, gm, per_path=True)
    assert np.array_equal(counts, ref)
    tr = traces_of(g)
    assert np.array_equal(pp["nverts"].ravel(), [len(t) for t in tr]) and (pp["status"] == 0).all()
    assert st["steps"] == sum(len(t) - 1 for t in tr) and st["paths"] == len(tr)
    _, st2, _ = eng.capture_host(spec, par, None)                 # tracking only through the host entry point
    assert st2["attempts"] == st["attempts"] and st2["bbox"] == st["bbox"]
    side = torch.cuda.Stream(device=eng.device)
    main = torch.cuda.current_stream(eng.device)
    dp = eng.upload(spec, par)
    torch.cuda.synchronize()
    try:
        eng.use_stream(side)
        with torch.cuda.stream(side):
            c2 = eng.new_counts(gm)
            eng.capture(spec, dp, gm, c2)
        side.synchronize()
    finally:
        eng.use_stream(main)
    assert np.array_equal(c2.cpu().numpy().view(np.uint32), ref)


def test_raster_traces_in_batches(eng, golden):
    """oneka_raster_traces with fewer bitmap slots than realizations (several raster + flush launches)."""
    from onekapy_b200.engine import Engine
    g = golden("sto_basic.npz")
    s, spec, par = spec_of(g)
    gm = fixed_geom(g, s)
    tr = traces_of(g)
    real_of = np.repeat(np.arange(len(par)), s["P"]).astype(np.int32)
    words = gm.nrows * 2 * ((gm.ncols + 63) // 64)                # bitmap rows hold an even number of 32-bit words
    small = Engine(0, workspace_limit=2 * words * 4)              # 2 slots for 6 realizations -> 3 batches
    n0 = small.launch_count()
    counts = small.raster_traces(gm, s["umbra"], tr, real_of, len(par))
    assert small.launch_count() - n0 == 6
    assert np.array_equal(counts, g["fixed_counts"].astype(np.uint32))
    small.close()


def test_zero_realizations(eng, golden):
    """nrealizations = 0 leaves the fresh 3 x 3 field (oneka/stochastic.py:212, loop :220 does not run)."""
    from onekapy_b200.engine import RealizationParams
    g = golden("sto_perham.npz")
    s, spec, par = spec_of(g)
    empty = RealizationParams(q=np.zeros((0, par.q.shape[1])), cond=[], poro=[], thick=[], coef=np.zeros((0, 6)))
    for res in (eng.run(spec, empty), eng.run_exact(spec, empty)):
        assert res["total_weight"] == 0.0 and res["counts"].shape == (3, 3) and not res["counts"].any()
        assert (res["geom"].xmin, res["geom"].xmax) == (s["xt"] - s["spacing"], s["xt"] + s["spacing"])


def test_bad_arguments(eng, golden):
    from onekapy_b200.engine import OnekaError
    from onekapy_b200.lattice import LatticeGeom
    g = golden("sto_perham.npz")
    s, spec, par = spec_of(g)
    dp = eng.upload(spec, par)
    spec.tol = 0.0
    with pytest.raises(OnekaError):
        eng.capture(spec, dp)
    spec.tol = 1.0
    spec.maxstep = -1.0
    with pytest.raises(OnekaError):
        eng.capture(spec, dp)
    spec.maxstep = 10.0
    eng2 = type(eng)(0, workspace_limit=16)
    with pytest.raises(OnekaError):
        eng2.capture(spec, dp, fixed_geom(g, s), eng2.new_counts(fixed_geom(g, s)))
    eng2.close()

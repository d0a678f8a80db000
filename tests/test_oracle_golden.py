"""Pins the CPU oracle (oracle/oneka_oracle.c) against the reference.

(i)  the reference's own known answers: /root/reference/tests/test_model.py:45-76 and
     tests/test_probabilityfield.py:24-53 (values restated below with their line numbers);
(ii) fixtures made by executing the unmodified reference (tests/golden/make_golden.py).

Bars: traces, distancesquared, lattice geometry and every grid cell are BIT-EXACT;
point evaluations of potential/head (which go through libm log / NumPy's `**2`) are
compared at rtol 1e-14.
"""
import numpy as np
import pytest

from oracle import oracle as O
from helpers import scal, geom, traces_of

CAPTURES = ["det_basic.npz", "sto_basic.npz", "sto_perham.npz", "unc_basic.npz", "fwd_basic.npz", "sto_wells200.npz"]


# ---- reference tests/test_model.py:26-42 fixture ---------------------------------------
MY_MODEL = dict(wells_xy=[(100.0, 200.0), (200.0, 100.0)], q=[1000.0, 1000.0], base=500.0, k=1.0, n=0.25,
                H=100.0, xo=0.0, yo=0.0, coef=[1.0, 1.0, 1.0, 1.0, 1.0, 500.0])


def test_known_answers_model():
    out = O.eval_points(pts=[(100.0, 100.0), (120.0, 160.0)], **MY_MODEL)
    assert np.isclose(out[0, 0], 32165.8711977589, rtol=1e-6)            # test_model.py:45-48
    assert np.isclose(out[0, 5], 371.658711977589, rtol=1e-6)            # test_model.py:51-54
    assert np.allclose(out[1, 1:3], [-401.318309886184, -438.771830796713], rtol=1e-6)   # :57-62
    assert np.allclose(out[0, 6:8], [-11.976338022763, -11.976338022763], rtol=1e-6)     # :65-70
    assert np.allclose(out[1, 6:8], [-16.052732395447, -17.550873231869], rtol=1e-6)     # :72-76
    # head > thickness at both points, so the confined formula gives the same velocity
    assert np.allclose(out[:, 3:5], out[:, 6:8], rtol=0, atol=0)


def test_known_answers_probabilityfield():
    with pytest.raises(ValueError):                                       # test_probabilityfield.py:24-30
        O.Field(-1.0, 1.0)
    with pytest.raises(ValueError):
        O.Field(1.0, 0.0)
    pf = O.Field(1.0, 1.0)                                                # :33-49
    pf.expand(100, 200, 50, 100)
    assert (pf.nrows, pf.ncols) == (53, 103)
    assert (pf.xmin, pf.xmax, pf.ymin, pf.ymax) == (99, 201, 49, 101)
    pf.expand(110, 120, 60, 70)
    assert (pf.nrows, pf.ncols) == (53, 103)
    assert O.distancesquared(0, 1, 1, 0, 0, 0) == 0.5                     # :52-53


def test_model_points(golden):
    g = golden("model_points.npz")
    for par, pts, ref in zip(g["par"], g["pts"], g["out"]):
        base, k, n, H, xo, yo = par[:6]
        out = O.eval_points(g["wells"][:, :2], g["wells"][:, 3], base, k, n, H, xo, yo, par[6:], pts)
        assert np.array_equal(np.isnan(out), np.isnan(ref))
        assert np.array_equal(out[:, 1:5], ref[:, 1:5])                   # discharge + confined velocity: bit-exact
        assert np.allclose(out, ref, rtol=1e-14, atol=0, equal_nan=True)


def test_distancesquared(golden):
    g = golden("distsq.npz")
    got = np.array([O.distancesquared(*r) for r in g["args"]])
    assert np.array_equal(got, g["d2"], equal_nan=True)
    assert np.isnan(g["d2"][:200]).all()                                  # zero-length segments give nan


def test_expand_sequences(golden):
    g = golden("expand.npz")
    for s in range(int(g["nspec"])):
        dx, dy, xo, yo = g["spec%d" % s]
        pf = O.Field(dx, dy, xo, yo)
        for box, ref in zip(g["boxes%d" % s], g["geom%d" % s]):
            pf.expand(*box)
            assert [pf.xmin, pf.xmax, pf.ymin, pf.ymax, pf.nrows, pf.ncols] == list(ref)


def test_insert_rasterize_register(golden):
    g = golden("insert.npz")
    tracks = traces_of(g)
    real_of = g["real_of"]
    for tag in "abc":
        dx, dy, umbra = g["par_" + tag]
        pf = O.Field(dx, dy, 60.0, 60.0)
        pf.expand(0.0, 200.0, 0.0, 200.0)
        for r in (0, 1):
            for t, rr in zip(tracks, real_of):
                if rr == r:
                    for i in range(len(t) - 1):
                        pf.insert(t[i, 0], t[i, 1], t[i + 1, 0], t[i + 1, 1], umbra)
            pf.register(1.0)
        assert list(pf.geom) == list(g["fixed_%s_geom" % tag])
        assert np.array_equal(pf.pgrid, g["fixed_%s_counts" % tag].astype(float))
        pf = O.Field(dx, dy, 60.0, 60.0)
        for r in (0, 1):
            for t, rr in zip(tracks, real_of):
                if rr == r:
                    pf.rasterize(t[:, 0], t[:, 1], umbra)
            pf.register(1.0)
        assert list(pf.geom) == list(g["auto_%s_geom" % tag])
        assert np.array_equal(pf.pgrid, g["auto_%s_counts" % tag].astype(float))


@pytest.mark.parametrize("name", CAPTURES)
def test_traces_bit_exact(golden, name):
    g = golden(name)
    s = scal(g)
    start = O.start_ring(s["xt"], s["yt"], s["rt"], s["P"])
    ref = traces_of(g)
    for r in range(len(g["k"])):
        for p in range(s["P"]):
            rc, v, na = O.backtrace(g["wells_xyr"][:, :2], g["q"][r], s["base"], g["k"][r], g["n"][r], g["H"][r],
                                    s["xt"], s["yt"], g["coef"][r], s["confined"], start[p, 0], start[p, 1],
                                    s["duration"], s["tol"], s["maxstep"])
            assert rc == O.OK
            assert np.array_equal(v, ref[r * s["P"] + p]), (name, r, p)


@pytest.mark.parametrize("name", CAPTURES)
@pytest.mark.parametrize("mode", ["auto", "fixed"])
def test_capture_grids_bit_exact(golden, name, mode):
    g = golden(name)
    s = scal(g)
    start = O.start_ring(s["xt"], s["yt"], s["rt"], s["P"])
    pf = O.Field(s["spacing"], s["spacing"], s["xt"], s["yt"])
    if mode == "fixed":
        pf.expand(*g["lattice"])
    res = O.capture(pf, 0 if mode == "auto" else 1, g["wells_xyr"][:, :2], s["base"], s["xt"], s["yt"], s["confined"],
                    g["q"], g["k"], g["n"], g["H"], g["coef"], start, s["duration"], s["umbra"], s["tol"],
                    s["maxstep"], nthreads=2)
    assert list(pf.geom) == list(g[mode + "_geom"])
    assert np.array_equal(pf.pgrid, g[mode + "_counts"].astype(float))
    ref = traces_of(g)
    assert np.array_equal(res["nverts"].ravel(), [len(t) for t in ref])
    assert np.array_equal(res["end_xy"].reshape(-1, 2), np.array([t[-1] for t in ref]))
    assert (res["status"] == O.OK).all()


def test_unconfined_dry_trace(golden):
    """AquiferError (model.py:343-344) truncates the trace (capturezone.py:249-253)."""
    g = golden("unc_dry.npz")
    base, k, n, H, xo, yo = g["par"]
    dur, tol, maxstep = g["scal"]
    ref = traces_of(g)
    assert g["terminated"].sum() >= 3 and not g["terminated"].all()
    for (xs, ys), t, dry in zip(g["starts"], ref, g["terminated"]):
        rc, v, na = O.backtrace(g["wells"][:, :2], g["wells"][:, 3], base, k, n, H, xo, yo, g["coef"], False,
                                xs, ys, dur, tol, maxstep)
        assert rc == (O.AQUIFER_DRY if dry else O.OK)
        assert np.array_equal(v, t)

"""Generate the golden fixtures in this directory by EXECUTING the unmodified reference.

Run in the build container only (the GPU box has no /root/reference):

    python tests/golden/make_golden.py

The reference is pure Python (`/root/reference/oneka/*.py`).  It is imported where
it lies, after the two NumPy aliases it needs on NumPy >= 1.24
(`oneka/probabilityfield.py:148-149, 215-216, 249-250` use `np.float`/`np.bool`).
Nothing from the reference is copied into this repo; only its *outputs* on
recorded inputs are stored, as small .npz files:

  model_points.npz   Model.compute_potential/head/discharge/velocity[_confined] at points
  fit.npz            Model.fit_regional_flow on sampled (k, H, q) rows
  distsq.npz         ProbabilityField.distancesquared on random and degenerate inputs
  expand.npz         ProbabilityField.expand geometry sequences
  insert.npz         ProbabilityField.insert / rasterize / register on hand-made tracks
  det_basic.npz      create_deterministic_capturezone path (data/basic_deterministic.py), 16 paths: all traces + grids
  sto_basic.npz      stochastic path on pre-sampled parameters (data/basic.py), 6 x 10: traces + grids (auto + fixed lattice)
  sto_perham.npz     same for data/perham.py, 2 x 20
  unc_basic.npz      confined=False path, 3 x 8 (+ a trace that raises AquiferError)
  fwd_basic.npz      negative duration (forward tracking) from injection wells (data/basic.py, discharges negated)
  archive_ref.bz2    oneka/archive.py dump_oneka executed on a 2 x 6 run of data/basic.py (+ archive_ref_pgrid.npz)
  unanchored.npz     compute_capturezone on an UN-anchored field with deltax != deltay (expand()'s empty-field branch), 2 x 7
  sto_wells200.npz   the stochastic path on this repo's synthetic 200-well field (onekapy_b200/synthetic.py), 1 x 10: the
                     executed reference on a field large enough for the far-field compression to carry most of the wells

Sampling follows `oneka/stochastic.py:220-241` line by line, except that the A-F draw
uses ONE seeded Generator instead of a fresh unseeded one per realization (`:241`),
so that the parameter rows are reproducible.
"""
import importlib
import os
import sys

import numpy as np

np.float = float      # noqa: alias shim, see module docstring
np.bool = bool        # noqa

REF = "/root/reference"
sys.path.insert(0, REF)
HERE = os.path.dirname(os.path.abspath(__file__))

from oneka.model import Model, AquiferError                     # noqa: E402
import oneka.capturezone as ref_cz                              # noqa: E402
from oneka.probabilityfield import ProbabilityField             # noqa: E402
from oneka.stochastic import generate_random_variate, compute_variate_mean  # noqa: E402
from oneka.utilities import filter_obs                          # noqa: E402

_orig_backtrace = ref_cz.compute_backtrace


class TraceRecorder:
    """Wraps the reference's compute_backtrace to keep every vertex list."""

    def __init__(self):
        self.traces = []

    def __enter__(self):
        def rec(xs, ys, duration, tol, maxstep, feval):
            v = _orig_backtrace(xs, ys, duration, tol, maxstep, feval)
            self.traces.append(np.array([[p[0], p[1]] for p in v], dtype=np.float64))
            return v
        ref_cz.compute_backtrace = rec
        return self

    def __exit__(self, *a):
        ref_cz.compute_backtrace = _orig_backtrace


def pack_traces(traces):
    """list of (n_i, 2) arrays -> (offsets int64[n+1], verts float64[sum, 2])."""
    off = np.zeros(len(traces) + 1, dtype=np.int64)
    for i, t in enumerate(traces):
        off[i + 1] = off[i] + len(t)
    return off, np.concatenate(traces, axis=0)


def sample_params(m, nreal, seed):
    """oneka/stochastic.py:220-241 with a seeded A-F generator."""
    np.random.seed(seed)
    rng = np.random.default_rng(seed)
    obs = filter_obs(m.OBSERVATIONS, m.WELLS, m.BUFFER)
    xt, yt, rt = m.WELLS[m.TARGET][0:3]
    nw = len(m.WELLS)
    q = np.zeros((nreal, nw))
    k = np.zeros(nreal)
    n = np.zeros(nreal)
    H = np.zeros(nreal)
    coef = np.zeros((nreal, 6))
    coef_ev = np.zeros((nreal, 6))
    coef_cov = np.zeros((nreal, 6, 6))
    for i in range(nreal):
        wells = []
        for w in m.WELLS:
            xw, yw, rw = w[0:3]
            qw = generate_random_variate(w[3])
            wells.append([xw, yw, rw, qw])
        k[i] = generate_random_variate(m.C_DIST)
        n[i] = generate_random_variate(m.P_DIST)
        H[i] = generate_random_variate(m.T_DIST)
        mo = Model(m.BASE, k[i], n[i], H[i], wells)
        ev, cov = mo.fit_regional_flow(obs, xt, yt)
        ev = np.reshape(ev, [6, ])
        coef_ev[i] = ev
        coef_cov[i] = cov
        coef[i] = rng.multivariate_normal(ev, cov)
        q[i] = [w[3] for w in wells]
    return dict(q=q, k=k, n=n, H=H, coef=coef, coef_ev=coef_ev, coef_cov=coef_cov,
                obs=np.array(obs, dtype=np.float64))


def make_feval(mo, confined):
    """The closures of oneka/stochastic.py:253-260."""
    if confined:
        def feval(xy):
            Vx, Vy = mo.compute_velocity_confined(xy[0], xy[1])
            return np.array([-Vx, -Vy])
    else:
        def feval(xy):
            Vx, Vy = mo.compute_velocity(xy[0], xy[1])
            return np.array([-Vx, -Vy])
    return feval


def run_reference(m, par, npaths, confined, duration=None, lattice=None):
    """Run the reference hot path on pre-sampled rows.

    lattice = None          -> auto-expanding field exactly as stochastic.py:212 does.
    lattice = (x0,x1,y0,y1) -> the field is pre-expanded so that this bbox is inside
                               (legal: pfield is a caller-owned argument, capturezone.py:53).
    Returns (pfield, traces).
    """
    xt, yt, rt = m.WELLS[m.TARGET][0:3]
    duration = m.DURATION if duration is None else duration
    pf = ProbabilityField(m.SPACING, m.SPACING, xt, yt)
    if lattice is not None:
        pf.expand(*lattice)
    nreal = len(par["k"])
    with TraceRecorder() as rec:
        for i in range(nreal):
            wells = [[w[0], w[1], w[2], par["q"][i, j]] for j, w in enumerate(m.WELLS)]
            mo = Model(m.BASE, par["k"][i], par["n"][i], par["H"][i], wells)
            mo.xo, mo.yo = xt, yt
            mo.coef = par["coef"][i]
            feval = make_feval(mo, confined)
            ref_cz.compute_capturezone(xt, yt, rt, npaths, duration, pf, m.UMBRA, 1.0,
                                       m.TOL, m.MAXSTEP, feval)
    return pf, rec.traces


def field_dict(pf, prefix):
    assert np.all(pf.pgrid == np.round(pf.pgrid))
    return {
        prefix + "geom": np.array([pf.xmin, pf.xmax, pf.ymin, pf.ymax, pf.deltax, pf.deltay,
                                   pf.nrows, pf.ncols, pf.total_weight], dtype=np.float64),
        prefix + "counts": pf.pgrid.astype(np.uint16),
    }


def bbox_of(traces, pad):
    v = np.concatenate(traces, axis=0)
    return (v[:, 0].min() - pad, v[:, 0].max() + pad, v[:, 1].min() - pad, v[:, 1].max() + pad)


def save(name, **arrs):
    path = os.path.join(HERE, name)
    np.savez_compressed(path, **arrs)
    print("%-18s %8.1f KiB" % (name, os.path.getsize(path) / 1024.0))


# ----------------------------------------------------------------------------------------
def gen_model_points():
    rng = np.random.default_rng(11)
    wells = [(100.0, 200.0, 1.0, 1000.0), (200.0, 100.0, 1.0, 1000.0), (-50.0, 40.0, 0.5, -300.0)]
    cases = []
    for (base, k, n, H, coef, xo, yo) in [
            (500.0, 1.0, 0.25, 100.0, [1, 1, 1, 1, 1, 500.0], 0.0, 0.0),
            (0.0, 20.0, 0.2, 22.0, [-1e-4, 2e-4, 5e-5, 0.3, -0.2, 9000.0], 120.0, 90.0),
            (10.0, 5.0, 0.3, 500.0, [1e-3, 1e-3, 0.0, 0.5, 0.5, 4000.0], 0.0, 0.0)]:
        mo = Model(base, k, n, H, wells, xo, yo, np.array(coef, dtype=float))
        pts = rng.uniform(-300, 400, size=(40, 2))
        out = np.full((40, 8), np.nan)
        for i, (x, y) in enumerate(pts):
            out[i, 0] = mo.compute_potential(x, y)
            out[i, 1:3] = mo.compute_discharge(x, y)
            out[i, 3:5] = mo.compute_velocity_confined(x, y)
            try:
                out[i, 5] = mo.compute_head(x, y)
                out[i, 6:8] = mo.compute_velocity(x, y)
            except AquiferError:
                pass
        cases.append((np.array([base, k, n, H, xo, yo] + list(coef), dtype=float), pts, out))
    save("model_points.npz", wells=np.array(wells),
         par=np.stack([c[0] for c in cases]), pts=np.stack([c[1] for c in cases]),
         out=np.stack([c[2] for c in cases]))


def gen_fit():
    rows = {}
    for name in ["basic", "perham", "long_prairie"]:
        m = importlib.import_module("data." + name)
        p = sample_params(m, 5, seed=3)
        for key in ["q", "k", "H", "coef_ev", "coef_cov", "obs"]:
            rows[name + "_" + key] = p[key]
    save("fit.npz", **rows)


def gen_distsq():
    rng = np.random.default_rng(5)
    a = rng.uniform(-50, 50, size=(4000, 6))
    a[:200, 2:4] = a[:200, 0:2]                       # zero-length segments -> nan
    a[200:400, 3] = a[200:400, 1]                     # horizontal
    a[400:600, 2] = a[400:600, 0]                     # vertical
    a[600:800] = np.round(a[600:800])                 # lattice points (exact ties)
    a[800:1000, 2:4] = a[800:1000, 0:2] + rng.uniform(-1e-9, 1e-9, size=(200, 2))  # tiny
    d = np.zeros(len(a))
    with np.errstate(all="ignore"):
        for i, r in enumerate(a):
            d[i] = ProbabilityField.distancesquared(*[np.float64(t) for t in r])
    save("distsq.npz", args=a, d2=d)


def gen_expand():
    seqs = []
    # (deltax, deltay, xo, yo, [bbox...]) ; xo = nan -> empty field (tests/test_probabilityfield.py:33-49)
    specs = [
        (1.0, 1.0, np.nan, np.nan, [(100, 200, 50, 100), (110, 120, 60, 70), (99, 201, 49, 101), (0.5, 300.25, -7, 100)]),
        (10.0, 10.0, 2250.0, 2250.0, [(2251.25, 2260.0, 2250.0, 2250.0), (2240.0, 2260.0, 2230.0, 2270.0),
                                      (1800.3, 2300.7, 2100.1, 2900.9), (2250, 2250, 2250, 2250)]),
        (4.0, 8.0, 302338.0, 5162551.0, [(302330.5, 302400.25, 5162500.0, 5162560.0), (302000, 302338, 5162551, 5163000)]),
        (0.1, 0.3, 1.0, 2.0, [(0.05, 2.35, 1.1, 3.7), (-1.0, 1.0, 0.0, 2.0)]),
    ]
    out = {}
    for s, (dx, dy, xo, yo, boxes) in enumerate(specs):
        pf = ProbabilityField(dx, dy, xo, yo)
        g = []
        for b in boxes:
            pf.expand(*b)
            g.append([pf.xmin, pf.xmax, pf.ymin, pf.ymax, pf.nrows, pf.ncols])
        out["spec%d" % s] = np.array([dx, dy, xo, yo])
        out["boxes%d" % s] = np.array(boxes, dtype=float)
        out["geom%d" % s] = np.array(g, dtype=float)
    save("expand.npz", nspec=np.array(len(specs)), **out)


def gen_insert():
    """Hand-made tracks on a small lattice; two 'realizations' registered with weight 1."""
    rng = np.random.default_rng(17)
    tracks = []
    # realization 0
    tracks.append(np.array([[50.0, 50.0], [58.5, 53.25], [58.5, 53.25], [70.0, 53.25], [70.0, 80.0]]))   # incl. zero-length, horizontal, vertical
    tracks.append(np.cumsum(rng.uniform(-6, 9, size=(40, 2)), axis=0) + 60.0)
    tracks.append(np.array([[20.0, 30.0], [24.0, 30.0], [28.0, 34.0], [28.0, 38.0]]))                       # lattice-aligned: exact ties d2 == umbra^2
    # realization 1
    tracks.append(np.cumsum(rng.uniform(-9, 6, size=(60, 2)), axis=0) + 120.0)
    tracks.append(np.array([[5.0, 5.0], [1.0, 1.0], [-3.0, 2.0]]))                                        # runs off the pre-expanded lattice -> clipped
    tracks.append(np.array([[100.0, 100.0], [100.0 + 1e-7, 100.0 - 1e-7], [103.0, 101.0]]))                # micro segment
    real_of = np.array([0, 0, 0, 1, 1, 1])
    out = {}
    for tag, (dx, dy, umbra) in {"a": (4.0, 4.0, 8.0), "b": (2.0, 3.0, 5.0), "c": (10.0, 10.0, 4.0)}.items():
        # fixed lattice: pre-expanded once, then insert() only (no per-track expand)
        pf = ProbabilityField(dx, dy, 60.0, 60.0)
        pf.expand(0.0, 200.0, 0.0, 200.0)
        for r in (0, 1):
            for t, rr in zip(tracks, real_of):
                if rr == r:
                    with np.errstate(all="ignore"):
                        for i in range(len(t) - 1):
                            pf.insert(t[i, 0], t[i, 1], t[i + 1, 0], t[i + 1, 1], umbra)
            pf.register(1.0)
        out.update(field_dict(pf, "fixed_%s_" % tag))
        out["par_" + tag] = np.array([dx, dy, umbra])
        # auto-expanding: rasterize() per track, as capturezone.py:120 does
        pf = ProbabilityField(dx, dy, 60.0, 60.0)
        for r in (0, 1):
            for t, rr in zip(tracks, real_of):
                if rr == r:
                    with np.errstate(all="ignore"):
                        pf.rasterize(list(t[:, 0]), list(t[:, 1]), umbra)
            pf.register(1.0)
        out.update(field_dict(pf, "auto_%s_" % tag))
    off, verts = pack_traces(tracks)
    save("insert.npz", offsets=off, verts=verts, real_of=real_of, **out)


def injection_variant(modname):
    """data/<modname>.py with every well turned into an INJECTION well (discharges negated).
    Forward tracking (negative duration) from the ring of an injection well is well-posed;
    from a pumping well it runs into the singularity and is chaotic at rounding level."""
    import types
    m0 = importlib.import_module("data." + modname)
    m = types.SimpleNamespace(**{k: getattr(m0, k) for k in dir(m0) if k.isupper()})
    m.WELLS = [(w[0], w[1], w[2], tuple(-t for t in reversed(w[3])) if isinstance(w[3], tuple) else -w[3])
               for w in m0.WELLS]
    return m


def synthetic_module(nwells):
    """onekapy_b200.synthetic.well_field as an object with the attributes of a reference data module."""
    import types
    sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
    from onekapy_b200 import synthetic
    pb = synthetic.well_field(nwells)
    return types.SimpleNamespace(**{k.upper(): v for k, v in pb.items()})


def gen_capture(name, modname, nreal, npaths, seed, confined=None, deterministic=False, duration=None, injection=False):
    if not isinstance(modname, str):
        m = modname
    else:
        m = injection_variant(modname) if injection else importlib.import_module("data." + modname)
    confined = m.CONFINED if confined is None else confined
    if deterministic:
        # oneka/deterministic.py:185-199
        obs = filter_obs(m.OBSERVATIONS, m.WELLS, m.BUFFER)
        xt, yt, rt = m.WELLS[m.TARGET][0:3]
        wells = [[w[0], w[1], w[2], compute_variate_mean(w[3])] for w in m.WELLS]
        k = compute_variate_mean(m.C_DIST)
        n = compute_variate_mean(m.P_DIST)
        H = compute_variate_mean(m.T_DIST)
        mo = Model(m.BASE, k, n, H, wells)
        ev, cov = mo.fit_regional_flow(obs, xt, yt)
        par = dict(q=np.array([[w[3] for w in wells]]), k=np.array([k]), n=np.array([n]), H=np.array([H]),
                   coef=np.reshape(ev, [1, 6]), coef_ev=np.reshape(ev, [1, 6]), coef_cov=cov[None], obs=np.array(obs, dtype=float))
    else:
        par = sample_params(m, nreal, seed)
    pf_auto, traces = run_reference(m, par, npaths, confined, duration)
    lattice = bbox_of(traces, 3.0 * m.UMBRA)
    pf_fix, traces2 = run_reference(m, par, npaths, confined, duration, lattice=lattice)
    assert all(np.array_equal(a, b) for a, b in zip(traces, traces2))
    off, verts = pack_traces(traces)
    xt, yt, rt = m.WELLS[m.TARGET][0:3]
    out = dict(q=par["q"], k=par["k"], n=par["n"], H=par["H"], coef=par["coef"],
               wells_xyr=np.array([[w[0], w[1], w[2]] for w in m.WELLS], dtype=float),
               scal=np.array([xt, yt, rt, npaths, m.DURATION if duration is None else duration, m.SPACING,
                              m.UMBRA, m.TOL, m.MAXSTEP, m.BASE, 1.0 if confined else 0.0]),
               lattice=np.array(lattice), offsets=off, verts=verts)
    out.update(field_dict(pf_auto, "auto_"))
    out.update(field_dict(pf_fix, "fixed_"))
    save(name, **out)
    nv = np.diff(off)
    print("   traces %d, vertices/path min %d mean %.1f max %d; auto grid %dx%d nonzero %d; fixed grid %dx%d nonzero %d"
          % (len(traces), nv.min(), nv.mean(), nv.max(), pf_auto.nrows, pf_auto.ncols,
             np.count_nonzero(pf_auto.pgrid), pf_fix.nrows, pf_fix.ncols, np.count_nonzero(pf_fix.pgrid)))


def gen_unconfined_dry():
    """confined=False FORWARD traces (negative duration) from an injection well in a uniform
    regional flow whose potential falls to zero 333 m downstream: the particles that are
    carried downstream hit potential <= 0 (AquiferError, model.py:343-344) inside feval and
    the bare except of capturezone.py:249-253 truncates the trace (a warning is logged)."""
    import logging

    class Catch(logging.Handler):
        def __init__(self):
            super().__init__()
            self.n = 0

        def emit(self, record):
            self.n += 1

    wells = [[0.0, 0.0, 0.25, -50.0]]
    mo = Model(0.0, 10.0, 0.25, 20.0, wells, 0.0, 0.0, np.array([0.0, 0.0, 0.0, 0.9, 0.0, 300.0]))
    feval = make_feval(mo, False)
    traces = []
    flags = []
    starts = [(1.25, 0.0), (0.0, 1.25), (-1.25, 0.0), (0.8838834764831844, -0.8838834764831844),
              (-0.8838834764831844, 0.8838834764831844)]
    logging.disable(logging.NOTSET)
    h = Catch()
    ref_cz.log.addHandler(h)
    for xs, ys in starts:
        before = h.n
        v = _orig_backtrace(xs, ys, -4000.0, 1.0, 15.0, feval)
        flags.append(h.n > before)
        traces.append(np.array([[p[0], p[1]] for p in v], dtype=np.float64))
    ref_cz.log.removeHandler(h)
    logging.disable(logging.CRITICAL)
    off, verts = pack_traces(traces)
    save("unc_dry.npz", wells=np.array(wells), par=np.array([0.0, 10.0, 0.25, 20.0, 0.0, 0.0]),
         coef=mo.coef, scal=np.array([-4000.0, 1.0, 15.0]), offsets=off, verts=verts,
         starts=np.array(starts), terminated=np.array(flags))
    print("   dry traces: vertices", np.diff(off), "terminated early:", flags)


def gen_archive():
    """oneka/archive.py:46-85 EXECUTED: dump_oneka writes 'logs\\Oneka<timestamp>.bz2' (a Windows-style name) into the
    working directory; the file is moved here as archive_ref.bz2, its pgrid kept beside it for the loading test."""
    import glob
    import shutil
    import tempfile
    from oneka.archive import dump_oneka, load_oneka
    m = importlib.import_module("data.basic")
    par = sample_params(m, 2, 41)
    pf, _ = run_reference(m, par, 6, m.CONFINED)
    obs = filter_obs(m.OBSERVATIONS, m.WELLS, m.BUFFER)
    cwd = os.getcwd()
    tmp = tempfile.mkdtemp()
    try:
        os.chdir(tmp)
        dump_oneka("golden", 1.0, m.TARGET, 6, m.DURATION, 2, m.BASE, m.C_DIST, m.P_DIST, m.T_DIST, m.WELLS, obs,
                   m.BUFFER, m.SPACING, m.UMBRA, m.SMOOTH, m.CONFINED, m.TOL, m.MAXSTEP, pf)
        (made,) = glob.glob(os.path.join(tmp, "*Oneka*.bz2"))
        back = load_oneka(made)
        assert np.array_equal(back["pfield"].pgrid, pf.pgrid)
        shutil.move(made, os.path.join(HERE, "archive_ref.bz2"))
    finally:
        os.chdir(cwd)
    np.savez_compressed(os.path.join(HERE, "archive_ref_pgrid.npz"), pgrid=pf.pgrid.astype(np.uint16))
    print("archive_ref.bz2    %8.1f KiB (grid %dx%d, total_weight %.0f)" % (os.path.getsize(os.path.join(HERE, "archive_ref.bz2")) / 1024.0,
                                                                         pf.nrows, pf.ncols, pf.total_weight))


def gen_unanchored():
    """compute_capturezone on a caller-owned field that is NOT anchored (ProbabilityField(dx, dy), nrows = ncols = 0) and
    has deltax != deltay: the first trace's rasterize() takes expand()'s empty-field branch (probabilityfield.py:205-220)."""
    m = importlib.import_module("data.basic")
    par = sample_params(m, 2, 53)
    xt, yt, rt = m.WELLS[m.TARGET][0:3]
    dx, dy, npaths = 10.0, 6.0, 7
    pf = ProbabilityField(dx, dy)
    with TraceRecorder() as rec:
        for i in range(2):
            wells = [[w[0], w[1], w[2], par["q"][i, j]] for j, w in enumerate(m.WELLS)]
            mo = Model(m.BASE, par["k"][i], par["n"][i], par["H"][i], wells)
            mo.xo, mo.yo = xt, yt
            mo.coef = par["coef"][i]
            ref_cz.compute_capturezone(xt, yt, rt, npaths, m.DURATION, pf, m.UMBRA, 1.0, m.TOL, m.MAXSTEP, make_feval(mo, True))
    off, verts = pack_traces(rec.traces)
    out = dict(q=par["q"], k=par["k"], n=par["n"], H=par["H"], coef=par["coef"],
               wells_xyr=np.array([[w[0], w[1], w[2]] for w in m.WELLS], dtype=float),
               scal=np.array([xt, yt, rt, npaths, m.DURATION, dx, dy, m.UMBRA, m.TOL, m.MAXSTEP, m.BASE]), offsets=off, verts=verts)
    out.update(field_dict(pf, "auto_"))
    save("unanchored.npz", **out)
    print("   unanchored field %g x %g: grid %dx%d nonzero %d" % (dx, dy, pf.nrows, pf.ncols, np.count_nonzero(pf.pgrid)))


if __name__ == "__main__":
    import logging
    logging.disable(logging.CRITICAL)
    which = sys.argv[1:] or ["points", "fit", "distsq", "expand", "insert", "det", "sto", "perham", "unc", "fwd", "dry", "wells200",
                             "archive", "unanchored"]
    if "points" in which:
        gen_model_points()
    if "fit" in which:
        gen_fit()
    if "distsq" in which:
        gen_distsq()
    if "expand" in which:
        gen_expand()
    if "insert" in which:
        gen_insert()
    if "det" in which:
        gen_capture("det_basic.npz", "basic_deterministic", 1, 16, 0, deterministic=True)
    if "sto" in which:
        gen_capture("sto_basic.npz", "basic", 6, 10, seed=7)
    if "perham" in which:
        gen_capture("sto_perham.npz", "perham", 2, 20, seed=9)
    if "unc" in which:
        gen_capture("unc_basic.npz", "basic", 3, 8, seed=13, confined=False)
    if "fwd" in which:
        gen_capture("fwd_basic.npz", "basic", 2, 6, seed=21, duration=-1500.0, injection=True)
    if "dry" in which:
        gen_unconfined_dry()
    if "archive" in which:
        gen_archive()
    if "unanchored" in which:
        gen_unanchored()
    if "wells200" in which:
        gen_capture("sto_wells200.npz", synthetic_module(200), 1, 10, seed=31)

"""GPU tests at BASELINE.json's full sizes: size-independent properties + oracle spot checks.

C3 (perham, 10 000 x 1000) and C2 (basic_deterministic, 1 x 10 000) run in full on the GPU; the
oracle checks a seeded subsample of the same rows on the same lattice (bit-exact grid, equal step
counts, endpoints within 1e-6 relative).  Properties that hold at any size:
  additivity    counts(rows A + rows B) == counts(A) + counts(B)          (register is a sum)
  determinism   two runs give identical grids                             (integer atomics)
  batching      a tiny bitmap workspace (many launches) changes nothing
  bounds        0 <= count <= R everywhere; the node at the target well is covered by every realization
  bookkeeping   paths == R*P, steps <= attempts, no path ended abnormally
"""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu
POS_RTOL = 1e-6


@pytest.fixture(scope="module")
def eng():
    from onekapy_b200.engine import Engine
    e = Engine(0)
    yield e
    e.close()


def workload(name, R, P, seed=11):
    import bench
    return bench.make_workload(name, R, P, seed)[:2]


def lattice_for(eng, spec, dp):
    from onekapy_b200.lattice import LatticeGeom
    eng.reset_stats()
    eng.capture(spec, dp)
    st = eng.read_stats()
    return LatticeGeom.anchored(spec.spacing, spec.spacing, spec.xtarget, spec.ytarget).expanded(*st["bbox"]), st


def oracle_subset(spec, par, rows, geom):
    from oracle import oracle as O
    from onekapy_b200.engine import start_ring
    sub = par.slice(0, len(par))
    pf = O.Field(spec.spacing, spec.spacing, spec.xtarget, spec.ytarget)
    pf.expand(geom.xmin + 0.5 * geom.deltax, geom.xmax - 0.5 * geom.deltax, geom.ymin + 0.5 * geom.deltay,
              geom.ymax - 0.5 * geom.deltay)
    assert (pf.nrows, pf.ncols, pf.xmin, pf.ymin) == (geom.nrows, geom.ncols, geom.xmin, geom.ymin)
    res = O.capture(pf, 1, spec.well_xy, spec.base, spec.xtarget, spec.ytarget, spec.confined, sub.q[rows], sub.cond[rows],
                    sub.poro[rows], sub.thick[rows], sub.coef[rows], start_ring(spec.xtarget, spec.ytarget, spec.rtarget, spec.npaths),
                    spec.duration, spec.umbra, spec.tol, spec.maxstep)
    return pf.pgrid.astype(np.uint32), res


def check_subset(eng, spec, par, rows, geom):
    from onekapy_b200.engine import RealizationParams
    sub = RealizationParams(q=par.q[rows], cond=par.cond[rows], poro=par.poro[rows], thick=par.thick[rows], coef=par.coef[rows])
    dp = eng.upload(spec, sub)
    counts = eng.new_counts(geom)
    eng.reset_stats()
    pp = eng.capture(spec, dp, geom, counts, per_path=True)
    st = eng.read_stats()
    want, res = oracle_subset(spec, par, rows, geom)
    got = counts.cpu().numpy().view(np.uint32)
    nv = pp["nverts"].cpu().numpy()
    end = pp["end_xy"].cpu().numpy()
    rel = (np.abs(end - res["end_xy"]).max(axis=2) / np.abs(res["end_xy"]).max(axis=2)).max()
    ndiff = int(np.count_nonzero(got != want))
    print("subset %s: attempts %d (oracle %d), endpoint max rel err %.3e, differing cells %d of %d (fraction %.2e)"
          % (list(rows), st["attempts"], res["attempts"], rel, ndiff, np.count_nonzero(want), ndiff / max(1, np.count_nonzero(want))))
    assert np.array_equal(nv, res["nverts"])
    assert st["attempts"] == res["attempts"]
    assert rel < POS_RTOL
    assert ndiff == 0
    return got


def test_c3_full_size_properties(eng):
    """perham, 10 000 realizations x 1000 paths on one GPU (BASELINE.json configs[2])."""
    R, P = 10000, 1000
    spec, par = workload("c3", R, P)
    dp = eng.upload(spec, par)
    geom, st0 = lattice_for(eng, spec, dp)
    assert st0["paths"] == R * P and st0["n_not_ok"] == 0 and st0["steps"] <= st0["attempts"]
    full = eng.new_counts(geom)
    eng.reset_stats()
    eng.capture(spec, dp, geom, full)
    st = eng.read_stats()
    full_h = full.cpu().numpy().view(np.uint32)
    assert st["attempts"] == st0["attempts"] and st["steps"] == st0["steps"]       # rasterising does not perturb tracking
    assert st["n_clipped"] == 0 or geom.strictly_contains(st["bbox"])
    assert full_h.max() <= R
    i0 = int(round((spec.ytarget - geom.ymin) / geom.deltay))
    j0 = int(round((spec.xtarget - geom.xmin) / geom.deltax))
    assert full_h[i0, j0] == R
    # additivity over a split of the realization range + determinism
    a, b = eng.new_counts(geom), eng.new_counts(geom)
    eng.capture(spec, dp, geom, a, r0=0, r1=3777)
    eng.capture(spec, dp, geom, b, r0=3777, r1=R)
    assert np.array_equal((a + b).cpu().numpy().view(np.uint32), full_h)
    again = eng.new_counts(geom)
    eng.capture(spec, dp, geom, again)
    assert np.array_equal(again.cpu().numpy().view(np.uint32), full_h)
    # the oracle on a seeded subsample of the same rows, same lattice
    rows = np.sort(np.random.default_rng(5).choice(R, size=6, replace=False))
    check_subset(eng, spec, par, rows, geom)
    print("C3 full size: attempts %.4g, accepted %.4g, lattice %dx%d, nonzero cells %d, max count %d"
          % (st["attempts"], st["steps"], geom.nrows, geom.ncols, np.count_nonzero(full_h), full_h.max()))


def test_batching_independent_of_workspace(eng):
    from onekapy_b200.engine import Engine
    spec, par = workload("c3", 96, 500)
    dp = eng.upload(spec, par)
    geom, _ = lattice_for(eng, spec, dp)
    a = eng.new_counts(geom)
    eng.capture(spec, dp, geom, a)
    words = geom.nrows * 2 * ((geom.ncols + 63) // 64)            # bitmap rows hold an even number of 32-bit words
    small = Engine(0, workspace_limit=7 * words * 4)           # 7 bitmaps -> 14 launches of the fused kernel
    dp2 = small.upload(spec, par)
    b = small.new_counts(geom)
    n0 = small.launch_count()
    small.capture(spec, dp2, geom, b)
    small.synchronize()
    per_batch = 2 + (2 if small.farfield_info() else 0)        # track + flush (+ the far-field coefficient kernel and the well check)
    assert small.launch_count() - n0 == per_batch * 14
    assert np.array_equal(a.cpu().numpy(), b.cpu().numpy())
    small.close()


def test_c2_ten_thousand_particles(eng):
    """basic_deterministic at the distribution means, 1 realization x 10 000 particles
    (BASELINE.json configs[1]); whole grid and every endpoint against the oracle."""
    from onekapy_b200 import problems
    from onekapy_b200.engine import FlowSpec
    from onekapy_b200.host.deterministic import mean_realization
    from onekapy_b200.host.utilities import filter_obs
    pb = problems.load("basic_deterministic")
    xt, yt, rt = pb["wells"][0][0:3]
    obs = filter_obs(pb["observations"], pb["wells"], pb["buffer"])
    par, mo = mean_realization(pb["base"], pb["c_dist"], pb["p_dist"], pb["t_dist"], pb["wells"], obs, xt, yt)
    spec = FlowSpec(well_xy=np.array([[w[0], w[1]] for w in pb["wells"]], dtype=float), xtarget=xt, ytarget=yt, rtarget=rt,
                    npaths=10000, duration=pb["duration"], base=pb["base"], spacing=pb["spacing"], umbra=pb["umbra"],
                    confined=True, tol=pb["tol"], maxstep=pb["maxstep"])
    dp = eng.upload(spec, par)
    geom, st0 = lattice_for(eng, spec, dp)
    got = check_subset(eng, spec, par, np.array([0]), geom)
    assert got.max() == 1 and st0["paths"] == 10000
    print("C2: lattice %dx%d, cells set %d, steps/path %.1f" % (geom.nrows, geom.ncols, np.count_nonzero(got), st0["steps"] / 10000))


def test_c4_synthetic_200_wells(eng):
    """200-well synthetic field, 64 realizations x 1000 paths; oracle on 2 of them."""
    spec, par = workload("c4", 64, 1000)
    dp = eng.upload(spec, par)
    geom, st0 = lattice_for(eng, spec, dp)
    assert st0["n_not_ok"] == 0
    counts = eng.new_counts(geom)
    eng.capture(spec, dp, geom, counts)
    h = counts.cpu().numpy().view(np.uint32)
    assert h.max() <= 64
    check_subset(eng, spec, par, np.array([3, 41]), geom)


def test_c5_fine_grid_long_duration(eng):
    """Fine lattice of the 4096 x 4096 class (spacing 4, umbra 20: ~15 x 15-node windows) and long traces
    (~300 steps per path, 2x C3): raster stress (BASELINE.json configs[4])."""
    spec, par = workload("c5", 32, 1000)
    dp = eng.upload(spec, par)
    geom, st0 = lattice_for(eng, spec, dp)
    assert 1024 < max(geom.nrows, geom.ncols) < 16384, (geom.nrows, geom.ncols)
    counts = eng.new_counts(geom)
    eng.reset_stats()
    eng.capture(spec, dp, geom, counts)
    st = eng.read_stats()
    h = counts.cpu().numpy().view(np.uint32)
    assert h.max() <= 32 and st["n_not_ok"] == 0
    check_subset(eng, spec, par, np.array([7]), geom)
    print("C5: lattice %dx%d (%.1f M nodes), steps %.3g, exact re-tests %d" % (geom.nrows, geom.ncols, geom.nrows * geom.ncols / 1e6, st["steps"], st["exact_tests"]))
    # the two rasteriser flavours set the same bits (this lattice runs the heavy one by default: 11 window rows)
    assert eng.raster_flavour(spec.umbra, spec.spacing, False) == "heavy"
    try:
        eng.set_raster_mode("plain")
        plain = eng.new_counts(geom)
        eng.capture(spec, dp, geom, plain)
    finally:
        eng.set_raster_mode("auto")
    assert bool((plain == counts).all())


@pytest.mark.parametrize("name,R,unconfined", [("c3", 256, False), ("c4", 64, False), ("c3", 64, True)])
def test_raster_flavours_agree_on_far_field_kernels(eng, name, R, unconfined):
    """Plain and heavy flavour of the far-field kernels (confined, unrolled order; 200 wells; unconfined): identical count grids."""
    import bench
    spec, par = bench.make_workload(name, R, 1000, 5, unconfined=unconfined)[:2]
    dp = eng.upload(spec, par)
    geom, st0 = lattice_for(eng, spec, dp)
    grids = {}
    try:
        for mode in ("plain", "heavy"):
            eng.set_raster_mode(mode)
            grids[mode] = eng.new_counts(geom)
            eng.reset_stats()
            eng.capture(spec, dp, geom, grids[mode])
            assert eng.read_stats()["n_not_ok"] == 0
    finally:
        eng.set_raster_mode("auto")
    assert int(grids["plain"].sum().item()) > 0 and bool((grids["plain"] == grids["heavy"]).all())


# ---- large well fields / many contexts -------------------------------------------------------------------------
def test_three_hundred_wells(eng):
    """300 wells (a 33 KB well store per CTA), against the oracle."""
    from onekapy_b200 import synthetic
    from onekapy_b200.engine import FlowSpec
    pb = synthetic.well_field(300, seed=7)
    par = synthetic.sample_rows_fast(pb, 4, 3)
    xt, yt, rt = pb["wells"][pb["target"]][0:3]
    spec = FlowSpec(well_xy=np.array([[w[0], w[1]] for w in pb["wells"]], dtype=float), xtarget=float(xt), ytarget=float(yt),
                    rtarget=float(rt), npaths=160, duration=float(pb["duration"]), base=float(pb["base"]), spacing=float(pb["spacing"]),
                    umbra=float(pb["umbra"]), confined=True, tol=float(pb["tol"]), maxstep=float(pb["maxstep"]))
    dp = eng.upload(spec, par)
    geom, st0 = lattice_for(eng, spec, dp)
    assert st0["n_not_ok"] == 0
    check_subset(eng, spec, par, np.array([0, 3]), geom)


def test_nine_hundred_wells_coefficients_in_passes(eng):
    """900 wells: the far-field coefficient GEMM stages the discharges in passes of 512 wells (farfield_coef_kernel), the well
    store takes 25 KB of the CTA's shared-memory budget.  Far field == direct sums (steps, grid) and == the oracle, confined
    and unconfined."""
    from onekapy_b200 import synthetic
    from onekapy_b200.engine import FlowSpec
    pb = synthetic.well_field(900, seed=11)
    par = synthetic.sample_rows_fast(pb, 3, 5)
    xt, yt, rt = pb["wells"][pb["target"]][0:3]
    for confined in (True, False):
        spec = FlowSpec(well_xy=np.array([[w[0], w[1]] for w in pb["wells"]], dtype=float), xtarget=float(xt), ytarget=float(yt),
                        rtarget=float(rt), npaths=96, duration=float(pb["duration"]), base=float(pb["base"]), spacing=float(pb["spacing"]),
                        umbra=float(pb["umbra"]), confined=confined, tol=float(pb["tol"]), maxstep=float(pb["maxstep"]))
        dp = eng.upload(spec, par)
        eng.farfield = "off"
        geom, st0 = lattice_for(eng, spec, dp)
        a = eng.new_counts(geom)
        eng.reset_stats()
        eng.capture(spec, dp, geom, a)
        sa = eng.read_stats()
        assert eng.farfield_info() is None
        eng.farfield = "auto"
        b = eng.new_counts(geom)
        eng.reset_stats()
        eng.capture(spec, dp, geom, b)
        sb = eng.read_stats()
        info = eng.farfield_info()
        assert info is not None and info["mean_near"] < 40
        need = eng._ff_smem_info()
        assert need["confined" if confined else "unconfined"] <= need["budget"]
        assert sa["attempts"] == sb["attempts"] and sa["steps"] == sb["steps"] and sb["n_not_ok"] == sa["n_not_ok"]
        assert np.array_equal(a.cpu().numpy(), b.cpu().numpy())
        print("900 wells confined=%s: %d x %d tiles, mean near %.1f; far field == direct sums (%d attempts)"
              % (confined, info["ntx"], info["nty"], info["mean_near"], sb["attempts"]))
        if confined:
            check_subset(eng, spec, par, np.array([1]), geom)


def test_many_contexts_on_one_device(eng):
    """15 live contexts on one device give the same grid (each owns its stream-ordered workspace)."""
    from onekapy_b200.engine import Engine
    spec, par = workload("c3", 4, 96)
    geom, _ = lattice_for(eng, spec, eng.upload(spec, par))
    want = eng.new_counts(geom)
    eng.capture(spec, eng.upload(spec, par), geom, want)
    want = want.cpu().numpy()
    engines = [Engine(0) for _ in range(15)]
    try:
        for e in engines:
            c = e.new_counts(geom)
            e.capture(spec, e.upload(spec, par), geom, c)
            assert np.array_equal(c.cpu().numpy(), want)
    finally:
        for e in engines:
            e.close()

"""The LOGIC of the CUDA device code, run on the CPU: tests/emu compiles onekapy_b200/csrc/oneka_device.cuh for the host
(one "thread" at a time, libm stand-ins for the MUFU approximations) and these tests hold it to the bars of the GPU parity
tests -- against fixtures from the executed reference and against the oracle.  This is test infrastructure for working on the
kernels without a GPU at hand; it is not a fallback (the package never sees it) and it does not replace `-m gpu`: launch
code, shared-memory staging, warp reductions and contended atomics only exist on the device."""
import numpy as np
import pytest

from helpers import scal, traces_of
from emu import emu
from onekapy_b200.engine import FlowSpec, RealizationParams, start_ring, farfield_grid
from onekapy_b200.lattice import LatticeGeom

CAPTURES = ["det_basic.npz", "sto_basic.npz", "sto_perham.npz", "unc_basic.npz", "fwd_basic.npz", "sto_wells200.npz"]


def spec_of(g):
    s = scal(g)
    spec = FlowSpec(well_xy=g["wells_xyr"][:, :2].copy(), xtarget=s["xt"], ytarget=s["yt"], rtarget=s["rt"],
                    npaths=s["P"], duration=s["duration"], base=s["base"], spacing=s["spacing"], umbra=s["umbra"],
                    confined=s["confined"], tol=s["tol"], maxstep=s["maxstep"])
    par = RealizationParams(q=g["q"], cond=g["k"], poro=g["n"], thick=g["H"], coef=g["coef"])
    return s, spec, par


def fixed_geom(g, s):
    return LatticeGeom.anchored(s["spacing"], s["spacing"], s["xt"], s["yt"]).expanded(*g["lattice"])


def ff_box(g, spec, tiles=64, order=28, eta=0.3, shrink=0.0):
    v = g["verts"]
    x0, x1, y0, y1 = v[:, 0].min() - 50.0, v[:, 0].max() + 50.0, v[:, 1].min() - 50.0, v[:, 1].max() + 50.0
    if shrink:
        x1, y1 = x0 + (1 - shrink) * (x1 - x0), y0 + (1 - shrink) * (y1 - y0)
    return dict(farfield_grid((x0, x1, y0, y1), tiles), order=order, eta=eta)


@pytest.mark.parametrize("name", CAPTURES)
def test_tracker_vs_executed_reference(golden, name):
    """dopri_track + field_feval (MODE 2): the reference's vertices, step for step."""
    g = golden(name)
    s, spec, par = spec_of(g)
    out = emu.capture(spec, par, start_ring(s["xt"], s["yt"], s["rt"], s["P"]), 2, max_verts=1024)
    worst = 0.0
    for k, t in enumerate(traces_of(g)):
        r, p = divmod(k, s["P"])
        assert out["status"][r, p] == 0 and out["nverts"][r, p] == len(t), (name, r, p)
        v = out["verts"][r, p, :len(t)]
        worst = max(worst, (np.abs(v - t).max(axis=1) / np.maximum(np.abs(t).max(axis=1), 1.0)).max())
    assert worst < 1e-9, worst


@pytest.fixture(params=["plain", "heavy"])
def flavour(request):
    """Both rasteriser flavours (raster_seg<RF_PLAIN / RF_HEAVY>) must set exactly the reference's bits."""
    emu.set_raster_flavour(request.param)
    yield request.param
    emu.set_raster_flavour("plain")


@pytest.mark.parametrize("name", CAPTURES)
def test_fused_tracker_and_rasteriser_bit_exact(golden, name, flavour):
    """dopri_track + raster_seg (MODE 1) + register: the executed reference's grid on the fixed lattice, cell for cell."""
    g = golden(name)
    s, spec, par = spec_of(g)
    gm = fixed_geom(g, s)
    out = emu.capture(spec, par, start_ring(s["xt"], s["yt"], s["rt"], s["P"]), 1, geom=gm)
    tr = traces_of(g)
    assert np.array_equal(out["nverts"].ravel(), [len(t) for t in tr])
    assert out["stats"]["steps"] == sum(len(t) - 1 for t in tr) and out["stats"]["paths"] == len(tr)
    assert np.array_equal(out["counts"], g["fixed_counts"].astype(np.uint32))
    v = g["verts"]
    assert np.allclose(out["stats"]["bbox"], [v[:, 0].min(), v[:, 0].max(), v[:, 1].min(), v[:, 1].max()], rtol=1e-9)


@pytest.mark.parametrize("tiles,shrink,order,eta", [(64, 0.0, 28, 0.3), (9, 0.0, 28, 0.3), (16, 0.5, 28, 0.3),
                                                     (380, 0.0, 16, 0.15), (120, 0.3, 16, 0.15)])
def test_far_field_evaluation_in_the_tracker(golden, tiles, shrink, order, eta):
    """field_feval_ff inside the tracker: same step sequence, same grid; shrink = 0.5 leaves three quarters of the
    area to the direct-sum fallback.  (order 16, eta 0.15, ~380 tiles) is Engine's default since round 2 and runs the
    UNROLLED evaluation (field_feval_ff<16>); the others run the loop."""
    g = golden("sto_perham.npz")
    s, spec, par = spec_of(g)
    gm = fixed_geom(g, s)
    ring = start_ring(s["xt"], s["yt"], s["rt"], s["P"])
    ff = ff_box(g, spec, tiles, order=order, eta=eta, shrink=shrink)
    direct = emu.capture(spec, par, ring, 2, max_verts=1024)
    far = emu.capture(spec, par, ring, 2, max_verts=1024, farfield=ff)
    assert np.array_equal(far["nverts"], direct["nverts"]) and np.array_equal(far["attempts"], direct["attempts"])
    n = far["nverts"].max()
    scale = np.maximum(np.abs(direct["verts"][:, :, :n]).max(axis=3), 1.0)
    assert (np.abs(far["verts"][:, :, :n] - direct["verts"][:, :, :n]).max(axis=3) / scale).max() < 1e-12
    fused = emu.capture(spec, par, ring, 1, geom=gm, farfield=ff)
    assert np.array_equal(fused["counts"], g["fixed_counts"].astype(np.uint32))


@pytest.mark.parametrize("tiles,order,eta", [(64, 28, 0.3), (380, 16, 0.15)])
def test_far_field_200_wells_vs_executed_reference(golden, tiles, order, eta):
    """The synthetic 200-well field, traced by the EXECUTED reference (tests/golden/sto_wells200.npz): with the far field
    carrying ~195 of the 200 wells the tracker still reproduces every vertex and the fused pass every cell."""
    g = golden("sto_wells200.npz")
    s, spec, par = spec_of(g)
    gm = fixed_geom(g, s)
    ring = start_ring(s["xt"], s["yt"], s["rt"], s["P"])
    ff = ff_box(g, spec, tiles, order=order, eta=eta)
    out = emu.capture(spec, par, ring, 2, max_verts=1024, farfield=ff)
    worst = 0.0
    for p, t in enumerate(traces_of(g)):
        assert out["nverts"][0, p] == len(t)
        worst = max(worst, (np.abs(out["verts"][0, p, :len(t)] - t).max(axis=1) / np.abs(t).max(axis=1)).max())
    assert worst < 1e-12, worst
    fused = emu.capture(spec, par, ring, 1, geom=gm, farfield=ff)
    assert np.array_equal(fused["counts"], g["fixed_counts"].astype(np.uint32))


def test_far_field_200_wells_vs_direct():
    import bench
    spec, par, _ = bench.make_workload("c4", 2, 24, 11)
    ring = start_ring(spec.xtarget, spec.ytarget, spec.rtarget, spec.npaths)
    direct = emu.capture(spec, par, ring, 0)
    bb = direct["stats"]["bbox"]
    gm = LatticeGeom.anchored(spec.spacing, spec.spacing, spec.xtarget, spec.ytarget).expanded(*bb)
    ff = dict(farfield_grid((gm.xmin, gm.xmax, gm.ymin, gm.ymax), 380), order=16, eta=0.15)
    a = emu.capture(spec, par, ring, 1, geom=gm)
    b = emu.capture(spec, par, ring, 1, geom=gm, farfield=ff)
    assert a["stats"]["attempts"] == b["stats"]["attempts"] and np.array_equal(a["nverts"], b["nverts"])
    assert np.array_equal(a["counts"], b["counts"]) and a["counts"].max() == 2
    rel = np.abs(a["end_xy"] - b["end_xy"]).max() / np.abs(a["end_xy"]).max()
    assert rel < 1e-12


def test_rasteriser_insert_fixture_and_random_tracks(golden, flavour):
    """raster_seg alone: hand-made tracks with exact ties (executed reference) and random tracks (oracle)."""
    from oracle import oracle as O
    g = golden("insert.npz")
    tracks = traces_of(g)
    for tag in "abc":
        dx, dy, umbra = g["par_" + tag]
        gm = LatticeGeom.anchored(dx, dy, 60.0, 60.0).expanded(0.0, 200.0, 0.0, 200.0)
        counts, _ = emu.raster_traces(gm, umbra, tracks, g["real_of"], 2)
        assert np.array_equal(counts, g["fixed_%s_counts" % tag].astype(np.uint32)), tag
    rng = np.random.default_rng(4)
    for dx, dy, umbra, step in [(4.0, 4.0, 8.0, 9.0), (2.5, 3.5, 11.0, 6.0), (10.0, 10.0, 4.0, 20.0), (1.0, 1.0, 7.3, 0.02)]:
        gm = LatticeGeom.anchored(dx, dy, 13.7, -4.2).expanded(-150.0, 170.0, -140.0, 160.0)
        tracks, cur = [], None
        for _ in range(60):
            n = int(rng.integers(2, 30))
            ang = rng.uniform(0, 2 * np.pi) + np.cumsum(rng.normal(0, 0.3, n))
            p = np.cumsum(np.stack([np.cos(ang), np.sin(ang)], axis=1) * step * rng.uniform(0.2, 1.0, (n, 1)), axis=0)
            tracks.append(p + rng.uniform(-100, 100, 2))
        counts, nexact = emu.raster_traces(gm, umbra, tracks, np.zeros(len(tracks), dtype=np.int32), 1)
        pf = O.Field(dx, dy, 13.7, -4.2)
        pf.expand(-150.0, 170.0, -140.0, 160.0)
        pf.freeze()
        for t in tracks:
            pf.rasterize(list(t[:, 0]), list(t[:, 1]), umbra)
        pf.register(1.0)
        assert (pf.nrows, pf.ncols) == (gm.nrows, gm.ncols)
        assert np.array_equal(counts, pf.pgrid.astype(np.uint32)), (dx, dy, umbra, step)


def test_dry_aquifer_and_attempt_guard(golden):
    """confined=False forward traces that run dry are truncated at the reference's vertex (PATH_AQUIFER_DRY); the attempt cap
    ends a path with PATH_MAX_ATTEMPT."""
    g = golden("unc_dry.npz")
    base, k, n, H, xo, yo = g["par"]
    dur, tol, maxstep = g["scal"]
    spec = FlowSpec(well_xy=g["wells"][:, :2].copy(), xtarget=xo, ytarget=yo, rtarget=0.25, npaths=len(g["starts"]),
                    duration=dur, base=base, spacing=1.0, umbra=1.0, confined=False, tol=tol, maxstep=maxstep)
    par = RealizationParams(q=g["wells"][None, :, 3], cond=[k], poro=[n], thick=[H], coef=g["coef"][None, :])
    out = emu.capture(spec, par, g["starts"], 2, max_verts=512)
    for p, (t, dry) in enumerate(zip(traces_of(g), g["terminated"])):
        assert out["status"][0, p] == (1 if dry else 0)
        assert out["nverts"][0, p] == len(t)
        assert np.abs(out["verts"][0, p, :len(t)] - t).max() / np.abs(t).max() < 1e-9
    g = golden("sto_basic.npz")
    s, spec, par = spec_of(g)
    spec.max_attempts = 50
    out = emu.capture(spec, par, start_ring(s["xt"], s["yt"], s["rt"], s["P"]), 2, max_verts=64)
    assert (out["status"] == 2).all() and (out["attempts"] == 50).all()


@pytest.mark.parametrize("confined", [True, False])
def test_well_store_remainders_vs_oracle(confined):
    """0 .. 9 wells: the blocks-of-four well store with its 0-3 leftover wells (both scalings), against the oracle's plain
    per-well loop; also the far field with so few wells that most tiles have nothing far."""
    from oracle import oracle as O
    rng = np.random.default_rng(12)
    xo, yo = 1000.0, 2000.0
    for nw in range(0, 10):
        wxy = np.concatenate([[[xo, yo]], rng.uniform(-900.0, 900.0, (max(nw - 1, 0), 2)) + [xo, yo]])[:nw].reshape(-1, 2)
        q = np.concatenate([[900.0], rng.uniform(100.0, 600.0, max(nw - 1, 0))])[:nw]
        coef = np.array([2e-5, -1e-5, 1e-5, -1.3, 0.4, 9000.0])
        spec = FlowSpec(well_xy=wxy, xtarget=xo, ytarget=yo, rtarget=0.3, npaths=6, duration=400.0, base=0.0, spacing=5.0,
                        umbra=10.0, confined=confined, tol=1.0, maxstep=15.0)
        par = RealizationParams(q=q[None, :], cond=[20.0], poro=[0.25], thick=[15.0], coef=coef[None, :])
        ring = start_ring(xo, yo, 0.3, 6)
        out = emu.capture(spec, par, ring, 2, max_verts=400)
        for p in range(6):
            st, v, na = O.backtrace(wxy, q, 0.0, 20.0, 0.25, 15.0, xo, yo, coef, confined, ring[p, 0], ring[p, 1], 400.0, 1.0, 15.0)
            assert out["status"][0, p] == st and out["nverts"][0, p] == len(v) and out["attempts"][0, p] == na, (nw, p)
            assert np.abs(out["verts"][0, p, :len(v)] - v).max() < 1e-7, (nw, p)
        if confined and nw >= 1:
            ff = dict(farfield_grid((xo - 600.0, xo + 600.0, yo - 600.0, yo + 600.0), 16), order=28, eta=0.3)
            far = emu.capture(spec, par, ring, 2, max_verts=400, farfield=ff)
            assert np.array_equal(far["nverts"], out["nverts"])
            n = out["nverts"].max()
            assert np.abs(far["verts"][:, :, :n] - out["verts"][:, :, :n]).max() < 1e-8


@pytest.mark.parametrize("name", ["sto_basic.npz", "sto_perham.npz", "fwd_basic.npz", "det_basic.npz", "unc_basic.npz", "sto_wells200.npz"])
def test_exact_clip_pipeline_equals_auto_expanding_reference(golden, name):
    """Engine.run_exact's pipeline with the device code on the CPU: per-path bounding boxes (tracking pass) -> running
    union -> per-path clip windows (lattice.clip_windows) -> clipped fused pass == the executed reference's AUTO-EXPANDING
    grid, cell for cell (probabilityfield.py:298-301, 335)."""
    import torch
    from onekapy_b200.lattice import clip_windows
    g = golden(name)
    s, spec, par = spec_of(g)
    ring = start_ring(s["xt"], s["yt"], s["rt"], s["P"])
    first = emu.capture(spec, par, ring, 0, path_bbox=True)
    base = LatticeGeom.anchored(s["spacing"], s["spacing"], s["xt"], s["yt"])
    final = base.expanded(*first["stats"]["bbox"])
    ref = g["auto_geom"]
    assert [final.xmin, final.xmax, final.ymin, final.ymax, final.nrows, final.ncols] == list(ref[[0, 1, 2, 3, 6, 7]])
    clip = clip_windows(torch, base, final, torch.from_numpy(first["path_bbox"])).numpy()
    out = emu.capture(spec, par, ring, 1, geom=final, clip=clip)
    want = g["auto_counts"].astype(np.uint32)
    assert out["counts"].shape == want.shape
    assert np.count_nonzero(out["counts"] != want) == 0


@pytest.mark.parametrize("shards", [1, 2])
@pytest.mark.parametrize("name", ["sto_basic.npz", "sto_perham.npz", "fwd_basic.npz", "det_basic.npz", "unc_basic.npz", "sto_wells200.npz"])
def test_one_pass_exact_clip_pipeline(golden, name, shards):
    """Engine.run_exact's ONE-PASS scheme with the device code on the CPU (same host logic: lattice.clip_windows,
    lattice.affected_paths): fused unclipped pass that also yields every path's box -> running union -> affected
    realizations -> their unclipped contribution subtracted, their clipped contribution added == the executed reference's
    auto-expanding grid, cell for cell; with 2 shards, shard 1's paths come after shard 0's (its windows start from
    shard 0's box) and the grids are summed -- what the allreduce does across GPUs."""
    import torch
    from onekapy_b200.lattice import clip_windows, clip_windows_rows, realization_boxes, union_before, affected_realizations
    g = golden(name)
    s, spec, par = spec_of(g)
    ring = start_ring(s["xt"], s["yt"], s["rt"], s["P"])
    base = LatticeGeom.anchored(s["spacing"], s["spacing"], s["xt"], s["yt"])
    R = len(par)
    cuts = [0, R] if shards == 1 or R < 2 else [0, (R + 1) // 2, R]
    # a generous work lattice (what the pilot + margin provides)
    v = g["verts"]
    work = base.expanded(v[:, 0].min() - 60.0, v[:, 0].max() + 60.0, v[:, 1].min() - 60.0, v[:, 1].max() + 60.0)
    passA, boxes = [], []
    for r0, r1 in zip(cuts, cuts[1:]):
        out = emu.capture(spec, par.slice(r0, r1), ring, 1, geom=work, path_bbox=True)
        assert out["stats"]["n_clipped"] == 0
        passA.append(out)
        boxes.append(out["stats"]["bbox"])
    true_bbox = (min(b[0] for b in boxes), max(b[1] for b in boxes), min(b[2] for b in boxes), max(b[3] for b in boxes))
    final = base.expanded(*true_bbox)
    ref = g["auto_geom"]
    assert [final.xmin, final.xmax, final.ymin, final.ymax, final.nrows, final.ncols] == list(ref[[0, 1, 2, 3, 6, 7]])
    i0, j0 = work.offset_of(final)
    total = np.zeros((final.nrows, final.ncols), dtype=np.int64)
    naff = 0
    for k, (r0, r1) in enumerate(zip(cuts, cuts[1:])):
        prior = None
        if k:
            pb = boxes[:k]
            prior = (min(b[0] for b in pb), max(b[1] for b in pb), min(b[2] for b in pb), max(b[3] for b in pb))
        bb = torch.from_numpy(passA[k]["path_bbox"])
        rb = realization_boxes(torch, bb)
        before = union_before(torch, rb, prior)
        aff = affected_realizations(torch, base, final, rb, before, s["umbra"]).numpy()
        idx = np.nonzero(aff)[0]
        naff += len(idx)
        sel = torch.from_numpy(idx)
        clip = clip_windows_rows(torch, base, final, bb.index_select(0, sel), before.index_select(0, sel))
        # the per-realization form gives the affected realizations the windows the flat running union gives every path
        assert torch.equal(clip, clip_windows(torch, base, final, bb, prior).index_select(0, sel))
        counts = passA[k]["counts"].astype(np.int64)
        if len(idx):
            sub = RealizationParams(q=par.q[r0:r1][idx], cond=par.cond[r0:r1][idx], poro=par.poro[r0:r1][idx],
                                    thick=par.thick[r0:r1][idx], coef=par.coef[r0:r1][idx])
            minus = emu.capture(spec, sub, ring, 1, geom=work)["counts"]
            counts -= minus
        assert counts.min() >= 0
        assert counts.sum() == counts[i0:i0 + final.nrows, j0:j0 + final.ncols].sum()      # nothing is left outside the final extents
        total += counts[i0:i0 + final.nrows, j0:j0 + final.ncols]
        if len(idx):
            total += emu.capture(spec, sub, ring, 1, geom=final, clip=clip.numpy())["counts"]
    want = g["auto_counts"].astype(np.int64)
    assert total.shape == want.shape and np.count_nonzero(total != want) == 0
    assert naff >= 1                                             # the very first paths always meet a grid still growing


def test_affected_realizations_is_conservative_and_selective():
    """lattice.affected_realizations on synthetic boxes: a realization deep inside the grid that preceded it is not affected,
    one within umbra + a cell of that grid's edge is; the first realization (nothing before it: the 3 x 3 base grid) and nan
    boxes always are."""
    import torch
    from onekapy_b200.lattice import union_before, affected_realizations
    base = LatticeGeom.anchored(4.0, 4.0, 0.0, 0.0)
    final = base.expanded(-400.0, 400.0, -400.0, 400.0)
    rb = torch.tensor([[-199.0, 199.0, -199.0, 199.0],           # 0: the first one (base grid only)          -> affected
                       [-150.0, 150.0, -150.0, 150.0],           # 1: grid now spans -200 .. 200, 50 m to spare -> fine
                       [-190.0, 150.0, -150.0, 150.0],           # 2: 10 m from the left edge: umbra 8 + a cell  -> affected
                       [-150.0, 150.0, -150.0, 192.5],           # 3: top: node 50 is needed (192.5 + 8 > 200)   -> affected
                       [-150.0, 150.0, -150.0, 180.0],           # 4: top: 20 m to spare                         -> fine
                       [-150.0, 300.0, -150.0, 150.0],           # 5: extends the union to the right             -> affected
                       [-150.0, 280.0, -150.0, 150.0],           # 6: inside what 5 left behind                  -> fine
                       [float("nan"), 1.0, 0.0, 1.0]], dtype=torch.float64)
    before = union_before(torch, rb)
    assert before[0].tolist() == [float("inf"), -float("inf"), float("inf"), -float("inf")]
    assert before[1].tolist() == [-199.0, 199.0, -199.0, 199.0] and before[6].tolist() == [-199.0, 300.0, -199.0, 199.0]
    got = affected_realizations(torch, base, final, rb, before, 8.0).tolist()
    assert got == [True, False, True, True, False, True, False, True]
    # other ranks' shards come first
    b2 = union_before(torch, rb, prior=(-250.0, 100.0, -100.0, 100.0))
    assert b2[0].tolist() == [-250.0, 100.0, -100.0, 100.0] and b2[1].tolist() == [-250.0, 199.0, -199.0, 199.0]


def test_unconfined_far_field_vs_direct_and_oracle():
    """field_feval_ff_unc (opt-in, oneka_set_farfield_unconfined): the perham and 200-well fields with confined=False --
    thick aquifers (saturated nowhere near the wells: the FP64 fallback runs) and thin ones (saturated everywhere: the FP32
    screening decides) -- keep every step of the direct-sum kernel, and the oracle's."""
    import bench
    from oracle import oracle as O
    for name, R, P, thick_scale in [("c3", 3, 12, 1.0), ("c3", 2, 12, 40.0), ("c4", 2, 10, 1.0), ("c4", 1, 10, 8.0)]:
        spec, par, _ = bench.make_workload(name, R, P, 3, unconfined=True)
        par = RealizationParams(q=par.q, cond=par.cond, poro=par.poro, thick=par.thick * thick_scale, coef=par.coef)
        ring = start_ring(spec.xtarget, spec.ytarget, spec.rtarget, P)
        direct = emu.capture(spec, par, ring, 2, max_verts=1500)
        bb = direct["stats"]["bbox"]
        ff = dict(farfield_grid((bb[0] - 20, bb[1] + 20, bb[2] - 20, bb[3] + 20), 64), order=28, eta=0.3)
        far = emu.capture(spec, par, ring, 2, max_verts=1500, farfield=ff)
        assert np.array_equal(far["status"], direct["status"]) and np.array_equal(far["nverts"], direct["nverts"])
        assert np.array_equal(far["attempts"], direct["attempts"])
        n = direct["nverts"].max()
        scale = np.maximum(np.abs(direct["verts"][:, :, :n]).max(axis=3), 1.0)
        assert (np.abs(far["verts"][:, :, :n] - direct["verts"][:, :, :n]).max(axis=3) / scale).max() < 1e-11
        for r in range(R):
            for p in (0, P // 2):
                st, v, na = O.backtrace(spec.well_xy, par.q[r], spec.base, par.cond[r], par.poro[r], par.thick[r], spec.xtarget,
                                        spec.ytarget, par.coef[r], False, ring[p, 0], ring[p, 1], spec.duration, spec.tol, spec.maxstep)
                assert far["status"][r, p] == st and far["nverts"][r, p] == len(v) and far["attempts"][r, p] == na
                assert np.abs(far["verts"][r, p, :len(v)] - v).max() / np.abs(v).max() < 1e-9

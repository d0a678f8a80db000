"""GPU tests added in round 2 (pytest -m gpu; everything through the C ABI).

  * a9: ProbabilityField.distancesquared through the device function the rasteriser uses, bit-exact on the executed
    reference's 4000 tuples (tests/golden/distsq.npz) -- the direct hook the round-1 review asked for;
  * oneka_capture_tracked: per-path boxes out of the FUSED pass == the tracking-only pass's, grid == oneka_capture's;
  * compute_capturezone on an UN-anchored field with deltax != deltay (executed-reference fixture unanchored.npz);
  * the far-field tables are bound to the well coordinates they were built from (a C-ABI caller passing other wells
    gets direct sums, not a mixture);
  * the library's own communicator with one rank (the N > 1 form is tests/test_gpu_multi.py);
  * the atomic probes that give the rasteriser its roofline return sane numbers.
"""
import ctypes as C

import numpy as np
import pytest

from helpers import scal, geom, traces_of
from test_gpu_parity import spec_of, fixed_geom

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def eng():
    from onekapy_b200.engine import Engine
    e = Engine(0)
    yield e
    e.close()


def test_distancesquared_device_hook_is_bit_exact(eng, golden):
    """oneka/probabilityfield.py:379-427 == exact_distancesquared (csrc/oneka_device.cuh), value for value incl. nan."""
    g = golden("distsq.npz")
    got = eng.distancesquared(g["args"])
    want = g["d2"]
    assert np.array_equal(np.isnan(got), np.isnan(want))
    ok = ~np.isnan(want)
    assert np.array_equal(got[ok].view(np.uint64), want[ok].view(np.uint64))       # the same bits
    # the reference's own known answer (tests/test_probabilityfield.py:50-53)
    assert np.isclose(eng.distancesquared([[0, 0, 10, 0, 5, 5]])[0], 25.0)


@pytest.mark.parametrize("name", ["sto_basic.npz", "sto_perham.npz", "unc_basic.npz"])
def test_capture_tracked_boxes_and_grid(eng, golden, name):
    import torch
    g = golden(name)
    s, spec, par = spec_of(g)
    gm = fixed_geom(g, s)
    dp = eng.upload(spec, par)
    plain = eng.new_counts(gm)
    eng.capture(spec, dp, gm, plain)
    bb_track = eng.path_bboxes(spec, dp)
    tracked = eng.new_counts(gm)
    flags = torch.zeros(len(par), dtype=torch.int32, device=eng.device)
    bb = torch.empty((len(par), s["P"], 4), dtype=torch.float64, device=eng.device)
    eng.capture(spec, dp, gm, tracked, flags=flags, bbox_out=bb)
    eng.synchronize()
    assert int(flags.sum().item()) == 0
    assert torch.equal(tracked, plain)
    assert torch.equal(bb, bb_track)
    # ... and they are the boxes of the executed reference's traces (to rounding)
    tr = traces_of(g)
    want = np.array([[t[:, 0].min(), t[:, 0].max(), t[:, 1].min(), t[:, 1].max()] for t in tr]).reshape(len(par), s["P"], 4)
    assert np.allclose(bb.cpu().numpy(), want, rtol=1e-9, atol=0)


def test_compute_capturezone_on_unanchored_field(eng, golden):
    """ProbabilityField(dx, dy) without an anchor, dx != dy (the reference accepts both): geometry and grid of
    the executed reference, cell for cell."""
    from oneka.capturezone import compute_capturezone
    from oneka.probabilityfield import ProbabilityField
    from onekapy_b200.host.capturezone import BacktraceVelocity
    from onekapy_b200.host.model import Model
    g = golden("unanchored.npz")
    xt, yt, rt, P, dur, dx, dy, umbra, tol, maxstep, base = g["scal"]
    pf = ProbabilityField(dx, dy)
    assert pf.nrows == 0
    for i in range(len(g["k"])):
        wells = [(w[0], w[1], w[2], g["q"][i, j]) for j, w in enumerate(g["wells_xyr"])]
        mo = Model(base, g["k"][i], g["n"][i], g["H"][i], wells, xt, yt, g["coef"][i])
        compute_capturezone(xt, yt, rt, int(P), dur, pf, umbra, 1.0, tol, maxstep, BacktraceVelocity(mo, True))
    ref = geom(g, "auto_")
    # An un-anchored grid takes its ORIGIN from the first trace's own extreme vertex (probabilityfield.py:206-207), a
    # computed coordinate: the lattice agrees with the reference's to the rounding of the traces (~1e-12 relative), not
    # bit for bit as the anchored lattices do, and a node within that distance of a capsule edge may differ.
    assert (pf.nrows, pf.ncols, pf.total_weight) == (ref["nrows"], ref["ncols"], ref["total_weight"])
    assert np.allclose([pf.xmin, pf.xmax, pf.ymin, pf.ymax], [ref["xmin"], ref["xmax"], ref["ymin"], ref["ymax"]], rtol=0, atol=1e-6)
    want = g["auto_counts"].astype(float)
    nd = int(np.count_nonzero(pf.pgrid != want))
    print("unanchored %g x %g field: differing cells %d of %d nonzero; origin differs by %.1e m"
          % (dx, dy, nd, np.count_nonzero(want), max(abs(pf.xmin - ref["xmin"]), abs(pf.ymin - ref["ymin"]))))
    assert nd <= 2


def test_farfield_tables_are_bound_to_their_wells(eng, golden):
    """oneka_set_farfield builds its tables from the wells it is given; a C-ABI caller that then passes OTHER coordinates
    with the same count and origin would get near terms from the new wells and polynomials from the old ones.
    Every launch compares the wells on the device and oneka_read_stats refuses to report such a run."""
    import copy
    g = golden("sto_perham.npz")
    s, spec, par = spec_of(g)
    gm = fixed_geom(g, s)
    eng.farfield = "auto"
    eng.reset_stats()
    eng.capture(spec, eng.upload(spec, par), gm, eng.new_counts(gm))
    assert eng.farfield_info() is not None and eng.read_stats()["n_not_ok"] == 0
    moved = copy.copy(spec)
    moved.well_xy = spec.well_xy.copy()
    moved.well_xy[1:] += 35.0                                     # every well but the target (the origin stays)
    from onekapy_b200 import _cabi
    dp = eng.upload(moved, par)
    got = eng.new_counts(gm)
    m, lat = moved.model_desc(), gm.as_lattice(moved.umbra)
    eng.reset_stats()                                             # a raw C-ABI call (Engine itself would rebuild the tables)
    _cabi.check(eng._L.oneka_capture(eng._h, C.byref(m), C.byref(lat), dp.well_xy.data_ptr(), len(par), s["P"], dp.q.data_ptr(),
                                     dp.cond.data_ptr(), dp.poro.data_ptr(), dp.thick.data_ptr(), dp.coef.data_ptr(),
                                     dp.start_xy.data_ptr(), got.data_ptr(), None, None, None))
    with pytest.raises(_cabi.OnekaError, match="OTHER well coordinates"):
        eng.read_stats()
    eng.reset_stats()
    # through Engine the tables follow the wells: same call, correct result (== direct sums of the moved wells)
    eng.capture(moved, dp, gm, eng.new_counts(gm))
    st = eng.read_stats()
    eng.farfield = "off"
    eng.reset_stats()
    eng.capture(moved, dp, gm, eng.new_counts(gm))
    assert eng.read_stats()["attempts"] == st["attempts"]
    eng.farfield = "auto"


def test_own_communicator_single_rank(eng):
    """oneka_comm_unique_id / oneka_comm_init_rank / oneka_allreduce_counts with one rank: the sum over one grid is the grid."""
    import torch
    from onekapy_b200 import _cabi
    idb = (C.c_ubyte * 128)()
    _cabi.check(eng._L.oneka_comm_unique_id(idb))
    _cabi.check(eng._L.oneka_comm_init_rank(eng._h, 1, 0, idb))
    t = torch.arange(4096, dtype=torch.int32, device=eng.device)
    _cabi.check(eng._L.oneka_allreduce_counts(eng._h, t.data_ptr(), t.numel()))
    d = torch.tensor([3.0, -1.0, 7.5], dtype=torch.float64, device=eng.device)
    for op in (0, 1, 2):
        _cabi.check(eng._L.oneka_allreduce_f64(eng._h, d.data_ptr(), 3, op))
    eng.synchronize()
    assert torch.equal(t, torch.arange(4096, dtype=torch.int32, device=eng.device))
    assert d.tolist() == [3.0, -1.0, 7.5]
    _cabi.check(eng._L.oneka_comm_destroy(eng._h))
    with pytest.raises(_cabi.OnekaError):
        _cabi.check(eng._L.oneka_allreduce_counts(eng._h, t.data_ptr(), t.numel()))       # no communicator any more


def test_atomic_probes(eng):
    """The rasteriser's roofline denominators: RED.OR to L2 (lane-private / one word per warp) and shared-memory atomicOr."""
    res = {m: eng.red_probe(m, span_bytes=64 << 20, iters=512)[0] for m in range(5)}
    print("atomic probes [1e9 word ops/s]: L2 lane-private %.1f, L2 warp-contended %.1f, shared lane-private %.1f, shared warp-contended %.1f, "
          "L2 one sector per lane %.1f" % (res[0], res[1], res[2], res[3], res[4]))
    assert all(v > 1.0 for v in res.values())
    assert res[2] > res[0]                                        # shared memory beats L2
    assert res[4] < res[0]                                        # a sector per lane costs more L2 requests than 8 lanes per sector


def test_raster_flavour_choice_and_modes(eng):
    """oneka_raster_flavour / oneka_set_raster_mode: heavy from 8 window rows on with direct well sums, from 11 with the far field;
    forcing a flavour overrides the lattice; a bad mode is refused."""
    from onekapy_b200 import _cabi
    assert eng.raster_flavour(8.0, 4.0, False) == "plain" and eng.raster_flavour(8.0, 4.0, True) == "plain"      # C3 / C4: 5 rows
    assert eng.raster_flavour(14.0, 4.0, False) == "heavy" and eng.raster_flavour(14.0, 4.0, True) == "plain"    # 8 rows
    assert eng.raster_flavour(20.0, 4.0, False) == "heavy" and eng.raster_flavour(20.0, 4.0, True) == "heavy"    # C5: 11 rows
    try:
        eng.set_raster_mode("heavy")
        assert eng.raster_flavour(8.0, 4.0, True) == "heavy"
        eng.set_raster_mode("plain")
        assert eng.raster_flavour(20.0, 4.0, False) == "plain"
    finally:
        eng.set_raster_mode("auto")
    with pytest.raises(_cabi.OnekaError):
        _cabi.check(eng._L.oneka_set_raster_mode(eng._h, 7))
    with pytest.raises(_cabi.OnekaError):
        eng.raster_flavour(8.0, 0.0, False)

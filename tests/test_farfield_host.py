"""Far-field compression of the well sum (include/oneka_b200.h: oneka_set_farfield): accuracy of the tiled local
expansions, checked WITHOUT a GPU through the library's host restatement of the same tables and evaluation
(oneka_farfield_eval_host) against the reference's formula -- one term per well, oneka/model.py:307-313 -- summed in
extended precision."""
import ctypes as C

import numpy as np
import pytest

from onekapy_b200 import _cabi, problems, synthetic
from onekapy_b200.engine import farfield_grid


def _field(name):
    pb = synthetic.well_field(200) if name == "c4" else problems.load(name)
    wxy = np.array([[w[0], w[1]] for w in pb["wells"]], dtype=np.float64)
    q = np.array([w[3][1] if isinstance(w[3], tuple) else w[3] for w in pb["wells"]], dtype=np.float64)
    xo, yo = wxy[pb["target"]]
    return wxy, q / (2 * np.pi * 20.0 * 0.25), float(xo), float(yo)


def _eval(wxy, w, xo, yo, grid, order, eta, pts, fp64=0):
    L = _cabi.load()
    pts = np.ascontiguousarray(pts, dtype=np.float64)
    out = np.zeros((len(pts), 2))
    near = np.zeros(len(pts), dtype=np.int32)
    _cabi.check(L.oneka_farfield_eval_host(len(wxy), wxy.ctypes.data, w.ctypes.data, xo, yo, grid["x0"], grid["y0"],
                                           grid["tile"], grid["ntx"], grid["nty"], order, eta, fp64, len(pts),
                                           pts.ctypes.data, out.ctypes.data, near.ctypes.data))
    return out, near


def _direct_ld(wxy, w, pts):
    """sum_w w (x - x_w)/r^2, sum_w w (y - y_w)/r^2 in long double, and sum |terms| (the scale rounding errors live on)."""
    ld = np.longdouble
    dx = pts[:, None, 0].astype(ld) - wxy[None, :, 0].astype(ld)
    dy = pts[:, None, 1].astype(ld) - wxy[None, :, 1].astype(ld)
    r2 = dx * dx + dy * dy
    gx = (w.astype(ld) * dx / r2).sum(axis=1)
    gy = (w.astype(ld) * dy / r2).sum(axis=1)
    mag = (np.abs(w) / np.sqrt(r2.astype(np.float64))).sum(axis=1)
    return gx, gy, mag


# (order, eta, tiles, bound on the error relative to sum |terms|): round 1's setting (truncation 3e-15 of a far term) and
# Engine's default since round 2 (eta 0.15, order 16 -- the unrolled evaluation -- on ~380 tiles: truncation 8e-14 of a far
# term, an order of magnitude below the 1e-12 of the Newton reciprocal in the direct-sum kernel)
SETTINGS = [(28, 0.3, 64, 2e-14), (16, 0.15, 380, 1e-13)]


@pytest.mark.parametrize("order,eta,tiles,bound", SETTINGS)
@pytest.mark.parametrize("name,box", [("c4", (-600.0, 3400.0, -900.0, 900.0)), ("perham", (-1400.0, 1400.0, -1400.0, 1400.0)),
                                      ("long_prairie", (-800.0, 800.0, -800.0, 800.0))])
def test_expansion_matches_direct_sum(name, box, order, eta, tiles, bound):
    wxy, w, xo, yo = _field(name)
    grid = farfield_grid((xo + box[0], xo + box[1], yo + box[2], yo + box[3]), tiles)
    rng = np.random.default_rng(1)
    n = 4000
    pts = np.stack([rng.uniform(grid["x0"], grid["x0"] + grid["ntx"] * grid["tile"], n),
                    rng.uniform(grid["y0"], grid["y0"] + grid["nty"] * grid["tile"], n)], axis=1)
    # tile corners and edges are where |zeta| is largest and where the tile index flips
    ii, jj = np.meshgrid(np.arange(grid["ntx"] + 1), np.arange(grid["nty"] + 1))
    corners = np.stack([grid["x0"] + ii.ravel() * grid["tile"], grid["y0"] + jj.ravel() * grid["tile"]], axis=1)
    pts = np.concatenate([pts, corners + 1e-9, corners - 1e-9])
    far_from_wells = np.min(np.hypot(pts[:, None, 0] - wxy[None, :, 0], pts[:, None, 1] - wxy[None, :, 1]), axis=1) > 0.5
    pts = pts[far_from_wells]
    out, near = _eval(wxy, w, xo, yo, grid, order, eta, pts)
    gx, gy, mag = _direct_ld(wxy, w, pts)
    err = np.maximum(np.abs(out[:, 0] - gx), np.abs(out[:, 1] - gy)).astype(np.float64) / mag
    inside = near >= 0
    assert inside.sum() > 3500
    print("%s order %d eta %.2f, %d tiles: max error %.2e of sum|terms|, mean near wells %.2f of %d"
          % (name, order, eta, grid["ntx"] * grid["nty"], err.max(), near[inside].mean(), len(wxy)))
    assert err.max() < bound, err.max()                    # truncation eta^order / (1 - eta) of a far term + double rounding
    assert np.all(near[inside] < len(wxy))
    if len(wxy) >= 29:
        assert near[inside].mean() < 0.45 * len(wxy)       # the point of it: most wells are in the polynomial
    # points outside the grid take the direct sum
    outside = np.array([[grid["x0"] - 5.0, grid["y0"] + 1.0], [grid["x0"] + grid["ntx"] * grid["tile"] + 1.0, grid["y0"] + 1.0],
                        [np.nan, 0.0]])
    o2, n2 = _eval(wxy, w, xo, yo, grid, order, eta, outside)
    assert list(n2) == [-1, -1, -1]
    g2x, g2y, m2 = _direct_ld(wxy, w, outside[:2])
    assert np.all(np.abs(o2[:2, 0] - g2x).astype(float) <= 1e-14 * m2)


def test_every_well_is_counted_once():
    """Unit weights on one well at a time: near list and polynomial partition the wells (no well lost or doubled)."""
    wxy, w, xo, yo = _field("perham")
    grid = farfield_grid((xo - 1000.0, xo + 1000.0, yo - 1000.0, yo + 1000.0), 16)
    rng = np.random.default_rng(3)
    pts = np.stack([rng.uniform(xo - 990, xo + 990, 50), rng.uniform(yo - 990, yo + 990, 50)], axis=1)
    for j in range(len(wxy)):
        e = np.zeros(len(wxy))
        e[j] = 1.0
        out, _ = _eval(wxy, e, xo, yo, grid, 28, 0.3, pts)
        dx, dy = pts[:, 0] - wxy[j, 0], pts[:, 1] - wxy[j, 1]
        r2 = dx * dx + dy * dy
        assert np.allclose(out[:, 0], dx / r2, rtol=1e-12, atol=0) and np.allclose(out[:, 1], dy / r2, rtol=1e-12, atol=0)


def test_order_controls_truncation():
    wxy, w, xo, yo = _field("c4")
    grid = farfield_grid((xo - 600.0, xo + 3400.0, yo - 900.0, yo + 900.0), 64)
    rng = np.random.default_rng(5)
    pts = np.stack([rng.uniform(xo - 500, xo + 3300, 1500), rng.uniform(yo - 800, yo + 800, 1500)], axis=1)
    gx, gy, mag = _direct_ld(wxy, w, pts)
    errs = []
    for order in (8, 16, 28):
        out, _ = _eval(wxy, w, xo, yo, grid, order, 0.3, pts)
        errs.append(float((np.abs(out[:, 0] - gx).astype(np.float64) / mag).max()))
    assert errs[0] > 1e-8 and errs[1] < 1e-8 and errs[2] < 2e-14 and errs[0] > errs[1] > errs[2]


def test_bad_arguments():
    L = _cabi.load()
    wxy, w, xo, yo = _field("basic")
    pts, out = np.zeros((1, 2)), np.zeros((1, 2))
    for order, eta, ntx in [(27, 0.3, 4), (2, 0.3, 4), (28, 0.95, 4), (28, 0.3, 0), (28, 0.3, 5000)]:
        rc = L.oneka_farfield_eval_host(len(wxy), wxy.ctypes.data, w.ctypes.data, xo, yo, 0.0, 0.0, 100.0, ntx, 4, order, eta,
                                        0, 1, pts.ctypes.data, out.ctypes.data, None)
        assert rc == -1 and b"far field" in L.oneka_last_error()


def test_farfield_grid_covers_the_box():
    for box, max_tiles in [((0.0, 2800.0, 0.0, 2400.0), 64), ((300000.0, 303900.0, 5160000.0, 5161800.0), 64),
                           ((0.0, 300.0, 0.0, 200.0), 64), ((0.0, 17000.0, 0.0, 13000.0), 16), ((5.0, 5.0, 7.0, 7.0), 64)]:
        g = farfield_grid(box, max_tiles)
        assert 1 <= g["ntx"] * g["nty"] <= max_tiles and g["tile"] >= 100.0
        assert g["x0"] <= box[0] and g["x0"] + g["ntx"] * g["tile"] >= box[1]
        assert g["y0"] <= box[2] and g["y0"] + g["nty"] * g["tile"] >= box[3]


class _FakeLib:
    """Records the far-field calls Engine makes; answers with a mean near count."""
    def __init__(self, mean_near):
        self.calls, self.mean_near = [], mean_near

    def oneka_set_farfield(self, h, nw, wxy, xo, yo, x0, y0, tile, ntx, nty, order, eta, fp64, mx, mean):
        self.calls.append(("set", int(nw), int(order)))
        if nw and mean is not None:
            mx._obj.value, mean._obj.value = 4, self.mean_near
        return 0

    def oneka_set_farfield_unconfined(self, h, on):
        self.calls.append(("unconfined", int(on)))
        return 0


def _stub_engine(mean_near):
    from onekapy_b200.engine import Engine
    e = object.__new__(Engine)
    e._L, e._h = _FakeLib(mean_near), None
    e.farfield, e.farfield_order, e.farfield_eta, e.farfield_max_tiles = "auto", 28, 0.3, 64
    e.farfield_order_fp64, e.farfield_min_wells, e._ff_unconfined = 0, 12, False
    e._ff_key = e._ff_info = None
    return e


def test_engine_far_field_policy_without_a_gpu():
    """Engine._auto_farfield: when the tables are built, kept, rebuilt and dropped (the C ABI replaced by a recorder)."""
    from onekapy_b200.engine import FlowSpec
    from onekapy_b200.lattice import LatticeGeom
    wxy, w, xo, yo = _field("perham")
    spec = FlowSpec(well_xy=wxy, xtarget=xo, ytarget=yo, rtarget=0.2, npaths=10, duration=100.0, base=0.0, spacing=4.0,
                    umbra=8.0, confined=True, tol=1.0, maxstep=10.0)
    geom = LatticeGeom.anchored(4.0, 4.0, xo, yo).expanded(xo - 900.0, xo + 900.0, yo - 700.0, yo + 700.0)
    e = _stub_engine(2.0)
    e._auto_farfield(spec, None)                                   # tracking only, nothing configured: stays direct
    assert e._L.calls == [] and e.farfield_info() is None
    e._auto_farfield(spec, geom)                                   # 29 wells, 2 near: worth it
    assert e._L.calls == [("set", 29, 28)] and e.farfield_info()["mean_near"] == 2.0
    e._auto_farfield(spec, geom)                                   # same lattice: kept, no call
    e._auto_farfield(spec, None)                                   # tracking-only pass of the same wells: kept
    assert len(e._L.calls) == 1
    e._auto_farfield(spec, geom, box=(xo - 500.0, xo + 500.0, yo - 400.0, yo + 400.0))    # tighter box: rebuilt
    assert len(e._L.calls) == 2 and e.farfield_info()["tile"] < 200.0
    other = FlowSpec(**{**spec.__dict__, "well_xy": wxy + 1.0})
    e._auto_farfield(other, None)                                  # other wells, no lattice: tables dropped
    assert e._L.calls[-1] == ("set", 0, 0) and e.farfield_info() is None
    spec.confined = False
    e._auto_farfield(spec, geom)                                   # unconfined: direct unless opted in
    assert e.farfield_info() is None and e._L.calls[-1] == ("set", 0, 0)
    e.farfield_unconfined = True
    e._auto_farfield(spec, geom)                                   # 29 wells, unconfined: the dearer evaluation does not pay
    assert e._L.calls[-3:] == [("unconfined", 1), ("set", 29, 28), ("set", 0, 0)] and e.farfield_info() is None
    e.farfield = "force"                                           # ... unless forced (tests, A/B runs)
    e._ff_key = None
    e._auto_farfield(spec, geom)
    assert e._L.calls[-1] == ("set", 29, 28) and e.farfield_info() is not None
    e.farfield = "auto"
    big = FlowSpec(**{**spec.__dict__, "well_xy": np.concatenate([wxy + 7.0 * k for k in range(7)])})    # 203 wells
    e._auto_farfield(big, geom)
    assert e._L.calls[-1] == ("set", 203, 28) and e.farfield_info() is not None
    # too many near wells (cost model): tables built, measured, dropped -- and the decision remembered
    e2 = _stub_engine(25.0)
    spec.confined = True
    e2._auto_farfield(spec, geom)
    assert e2._L.calls == [("set", 29, 28), ("set", 0, 0)] and e2.farfield_info() is None
    e2._auto_farfield(spec, geom)
    assert len(e2._L.calls) == 2
    # few wells / switched off: never
    e3 = _stub_engine(0.0)
    small = FlowSpec(**{**spec.__dict__, "well_xy": wxy[:5].copy()})
    e3._auto_farfield(small, geom)
    e3.farfield = "off"
    e3._auto_farfield(spec, geom)
    assert e3._L.calls == []


def test_tile_lookup_edges():
    """ff_locate's index arithmetic (rounding through 1.5 * 2^52, range check on the words of the sum): exact grid corners, points a
    hair outside, points 2^32 and 2^40 tiles away (whose low word alone would look like a valid tile), inf and nan."""
    wxy, w, xo, yo = _field("perham")
    grid = farfield_grid((xo - 1400.0, xo + 1400.0, yo - 1400.0, yo + 1400.0), 380)
    x0, y0, t, ntx, nty = grid["x0"], grid["y0"], grid["tile"], grid["ntx"], grid["nty"]
    x1, y1 = x0 + ntx * t, y0 + nty * t
    inside = np.array([[x0, y0], [x0 + 0.5 * t, y0], [x0, y0 + 2.0 * t], [x0 + 3.0 * t, y0 + 4.0 * t], [np.nextafter(x1, x0), np.nextafter(y1, y0)],
                       [x0 + (ntx - 0.25) * t, y0 + (nty - 0.25) * t]])
    outside = np.array([[x0 - 1e-6 * t, y0 + t], [x0 + t, y0 - 1e-6 * t], [x1 + 1e-6 * t, y0 + t], [x0 + t, y1 + 1e-6 * t],
                        [x0 + (2.0 ** 32 + 3.5) * t, y0 + t], [x0 + t, y0 + (2.0 ** 40 + 1.5) * t], [x0 - (2.0 ** 32 - 3.5) * t, y0 + t],
                        [np.inf, y0 + t], [x0 + t, -np.inf], [np.nan, y0 + t], [1e300, 1e300]])
    out, near = _eval(wxy, w, xo, yo, grid, 16, 0.15, np.concatenate([inside, outside]))
    assert np.all(near[:len(inside)] >= 0), near[:len(inside)]
    assert np.all(near[len(inside):] == -1), near[len(inside):]
    gx, gy, mag = _direct_ld(wxy, w, inside)
    err = np.maximum(np.abs(out[:len(inside), 0] - gx), np.abs(out[:len(inside), 1] - gy)).astype(np.float64) / mag
    assert err.max() < 1e-13, err

"""ctypes wrapper around oracle/oneka_oracle.c  (TEST INFRASTRUCTURE ONLY).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
legs may import this module; the product package never does (tests/test_layout.py
greps for it).  See the header of oneka_oracle.c for what pins this oracle.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIBS = {}

OK, AQUIFER_DRY, MAX_ATTEMPT, NONFINITE, TRACE_FULL = 0, 1, 2, 3, 4

_dp = C.POINTER(C.c_double)
_i64p = C.POINTER(C.c_int64)
_u8p = C.POINTER(C.c_uint8)


def build(force=False):
    """Compile the oracle with the committed Makefile (gcc, seconds)."""
    so = os.path.join(_HERE, "_build", "liboneka_oracle.so")
    src = os.path.join(_HERE, "oneka_oracle.c")
    if force or not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "--no-print-directory"], stdout=subprocess.DEVNULL)
    return so


def lib(plain_norm=False):
    key = bool(plain_norm)
    if key not in _LIBS:
        build()
        name = "liboneka_oracle_plain.so" if plain_norm else "liboneka_oracle.so"
        L = C.CDLL(os.path.join(_HERE, "_build", name))
        L.oneka_oracle_field_new.restype = C.c_void_p
        L.oneka_oracle_field_new.argtypes = [C.c_double] * 4
        L.oneka_oracle_field_free.argtypes = [C.c_void_p]
        L.oneka_oracle_field_expand.argtypes = [C.c_void_p] + [C.c_double] * 4
        L.oneka_oracle_field_expand.restype = C.c_int
        L.oneka_oracle_field_insert.argtypes = [C.c_void_p] + [C.c_double] * 5
        L.oneka_oracle_field_rasterize.argtypes = [C.c_void_p, C.c_int64, _dp, _dp, C.c_double]
        L.oneka_oracle_field_register.argtypes = [C.c_void_p, C.c_double]
        L.oneka_oracle_field_reset.argtypes = [C.c_void_p]
        L.oneka_oracle_field_geom.argtypes = [C.c_void_p, _dp]
        L.oneka_oracle_field_freeze.argtypes = [C.c_void_p, C.c_int]
        L.oneka_oracle_field_copy_pgrid.argtypes = [C.c_void_p, _dp]
        L.oneka_oracle_field_copy_rgrid.argtypes = [C.c_void_p, _u8p]
        L.oneka_oracle_distancesquared.restype = C.c_double
        L.oneka_oracle_distancesquared.argtypes = [C.c_double] * 6
        L.oneka_oracle_eval.argtypes = [C.c_int, _dp, _dp, _dp, _dp, C.c_int64, _dp, _dp]
        L.oneka_oracle_backtrace.argtypes = [C.c_int, _dp, _dp, _dp, _dp, C.c_int, C.c_double, C.c_double,
                                             C.c_double, C.c_double, C.c_double, C.c_int64, C.c_int64,
                                             _dp, _i64p, _i64p]
        L.oneka_oracle_capture.argtypes = [C.c_void_p, C.c_int, C.c_int,
                                           C.c_int, _dp, C.c_double, C.c_double, C.c_double, C.c_int,
                                           C.c_int64, _dp, _dp, _dp, _dp, _dp,
                                           C.c_int64, _dp, C.c_double, C.c_double, C.c_double,
                                           C.c_double, C.c_double, C.c_int64,
                                           _dp, _i64p, _u8p, _i64p, _i64p]
        L.oneka_oracle_num_threads.restype = C.c_int
        _LIBS[key] = L
    return _LIBS[key]


def _d(a):
    a = np.ascontiguousarray(a, dtype=np.float64)
    return a, a.ctypes.data_as(_dp)


def num_threads():
    return int(lib().oneka_oracle_num_threads())


def start_ring(xtarget, ytarget, rtarget, npaths):
    """capturezone.py:110-115, evaluated scalar by scalar with NumPy as the reference does."""
    out = np.zeros((npaths, 2))
    for i, theta in enumerate(np.linspace(0, 2 * np.pi, npaths + 1)[0:-1]):
        out[i, 0] = (rtarget + 1.0) * np.cos(theta) + xtarget
        out[i, 1] = (rtarget + 1.0) * np.sin(theta) + ytarget
    return out


def eval_points(wells_xy, q, base, k, n, H, xo, yo, coef, pts, plain_norm=False):
    wxy, pw = _d(wells_xy)
    qq, pq = _d(q)
    par, pp = _d([base, k, n, H, xo, yo])
    cf, pc = _d(coef)
    pt, ppt = _d(pts)
    out = np.zeros((len(pt), 8))
    lib(plain_norm).oneka_oracle_eval(len(qq), pw, pq, pp, pc, len(pt), ppt, out.ctypes.data_as(_dp))
    return out


def backtrace(wells_xy, q, base, k, n, H, xo, yo, coef, confined, xs, ys, duration, tol, maxstep,
              max_attempts=1 << 22, max_verts=1 << 16, plain_norm=False):
    """Returns (status, verts[n,2], nattempts)."""
    wxy, pw = _d(wells_xy)
    qq, pq = _d(q)
    par, pp = _d([base, k, n, H, xo, yo])
    cf, pc = _d(coef)
    verts = np.zeros((max_verts, 2))
    nv = C.c_int64(0)
    na = C.c_int64(0)
    rc = lib(plain_norm).oneka_oracle_backtrace(len(qq), pw, pq, pp, pc, int(bool(confined)), xs, ys, duration, tol,
                                              maxstep, max_attempts, max_verts, verts.ctypes.data_as(_dp),
                                              C.byref(nv), C.byref(na))
    return rc, verts[:min(nv.value, max_verts)].copy(), na.value


def distancesquared(ax, ay, bx, by, cx, cy):
    return lib().oneka_oracle_distancesquared(ax, ay, bx, by, cx, cy)


class Field:
    """oneka/probabilityfield.py ProbabilityField, restated (oneka_oracle.c)."""

    def __init__(self, deltax, deltay, xo=np.nan, yo=np.nan, plain_norm=False):
        self._L = lib(plain_norm)
        self._h = self._L.oneka_oracle_field_new(deltax, deltay, xo, yo)
        if not self._h:
            raise ValueError("<deltax>, <deltay> must be > 0.")

    def __del__(self):
        if getattr(self, "_h", None):
            self._L.oneka_oracle_field_free(self._h)
            self._h = None

    def expand(self, xmin, xmax, ymin, ymax):
        if self._L.oneka_oracle_field_expand(self._h, xmin, xmax, ymin, ymax) != 0:
            raise ValueError("min must be <= max")

    def freeze(self, fixed=True):
        self._L.oneka_oracle_field_freeze(self._h, int(fixed))

    def insert(self, ax, ay, bx, by, umbra):
        self._L.oneka_oracle_field_insert(self._h, ax, ay, bx, by, umbra)

    def rasterize(self, x, y, umbra):
        xx, px = _d(x)
        yy, py = _d(y)
        self._L.oneka_oracle_field_rasterize(self._h, len(xx), px, py, umbra)

    def register(self, weight):
        self._L.oneka_oracle_field_register(self._h, weight)

    def reset(self):
        self._L.oneka_oracle_field_reset(self._h)

    @property
    def geom(self):
        g = np.zeros(9)
        self._L.oneka_oracle_field_geom(self._h, g.ctypes.data_as(_dp))
        return g

    xmin = property(lambda s: s.geom[0])
    xmax = property(lambda s: s.geom[1])
    ymin = property(lambda s: s.geom[2])
    ymax = property(lambda s: s.geom[3])
    deltax = property(lambda s: s.geom[4])
    deltay = property(lambda s: s.geom[5])
    nrows = property(lambda s: int(s.geom[6]))
    ncols = property(lambda s: int(s.geom[7]))
    total_weight = property(lambda s: s.geom[8])

    @property
    def pgrid(self):
        out = np.zeros((self.nrows, self.ncols))
        if out.size:
            self._L.oneka_oracle_field_copy_pgrid(self._h, out.ctypes.data_as(_dp))
        return out

    @property
    def rgrid(self):
        out = np.zeros((self.nrows, self.ncols), dtype=np.uint8)
        if out.size:
            self._L.oneka_oracle_field_copy_rgrid(self._h, out.ctypes.data_as(_u8p))
        return out.astype(bool)


def capture(field, mode, wells_xy, base, xo, yo, confined, q, cond, poro, thick, coef, start_xy,
            duration, umbra, tol, maxstep, weight=1.0, max_attempts=1 << 22, nthreads=0, want_paths=True):
    """Run R realizations x P paths into `field` (mode 0 = auto-expanding, 1 = fixed lattice).

    Returns dict(end_xy[R,P,2], nverts[R,P], status[R,P], attempts, steps)."""
    wxy, pw = _d(wells_xy)
    qq, pq = _d(q)
    kk, pk = _d(cond)
    nn, pn = _d(poro)
    hh, ph = _d(thick)
    cf, pc = _d(coef)
    st, ps = _d(start_xy)
    R = len(kk)
    P = len(st)
    nw = len(wxy)
    assert qq.shape == (R, nw) and cf.shape == (R, 6)
    if want_paths:
        end_xy = np.zeros((R, P, 2))
        nverts = np.zeros((R, P), dtype=np.int64)
        status = np.zeros((R, P), dtype=np.uint8)
        pe, pv, pst = end_xy.ctypes.data_as(_dp), nverts.ctypes.data_as(_i64p), status.ctypes.data_as(_u8p)
    else:
        end_xy = nverts = status = None
        pe = pv = pst = None
    att = C.c_int64(0)
    stp = C.c_int64(0)
    field._L.oneka_oracle_capture(field._h, mode, nthreads, nw, pw, base, xo, yo, int(bool(confined)),
                                  R, pq, pk, pn, ph, pc, P, ps, duration, umbra, weight, tol, maxstep,
                                  max_attempts, pe, pv, pst, C.byref(att), C.byref(stp))
    return dict(end_xy=end_xy, nverts=nverts, status=status, attempts=att.value, steps=stp.value)

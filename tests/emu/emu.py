"""Build + ctypes binding of tests/emu/oneka_emu.cpp: the CUDA device code compiled for the host.

TEST INFRASTRUCTURE (see the header of oneka_emu.cpp): lets the CPU test suite run the logic of the kernels -- tracker,
far-field evaluation, rasteriser -- against the oracle and the golden fixtures without a GPU.  Never imported by the package."""
import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
SRC = os.path.join(HERE, "oneka_emu.cpp")
DEPS = [SRC, os.path.join(ROOT, "onekapy_b200", "csrc", "oneka_device.cuh"),
        os.path.join(ROOT, "onekapy_b200", "csrc", "oneka_farfield_host.h")]
# ONEKA_EMU_DEFINES="-DONEKA_RK_LOOP=1 ..." builds (and tests) the device code with other build knobs
DEFINES = os.environ.get("ONEKA_EMU_DEFINES", "").split()
OUT = os.path.join(HERE, "_build", "liboneka_emu%s.so" % ("_" + "_".join(d.lstrip("-D").replace("=", "") for d in DEFINES) if DEFINES else ""))
_lib = None


def build(force=False):
    stale = force or not os.path.exists(OUT) or any(os.path.getmtime(d) > os.path.getmtime(OUT) for d in DEPS)
    if stale:
        os.makedirs(os.path.dirname(OUT), exist_ok=True)
        cuda_inc = os.path.join(os.environ.get("CUDA_HOME", "/usr/local/cuda"), "include")
        # -ffp-contract=off: the __d*_rn stand-ins must stay unfused, as the intrinsics are on the device
        subprocess.check_call(["g++", "-std=c++17", "-O2", "-fPIC", "-shared", "-ffp-contract=off", "-I", cuda_inc,
                               "-I", os.path.join(ROOT, "include")] + DEFINES + [SRC, "-o", OUT])
    return OUT


def lib():
    global _lib
    if _lib is None:
        _lib = C.CDLL(build())
        vp, d, i, ll = C.c_void_p, C.c_double, C.c_int, C.c_longlong
        _lib.oneka_emu_capture.argtypes = [i, i, vp, d, d, i, d, d, d, ll, ll, i, vp, vp, vp, vp, vp, vp,
                                           d, d, d, d, i, i, d, vp, i, vp, vp, vp, vp, vp,
                                           i, d, d, d, d, i, i, vp, vp, vp, vp]
        _lib.oneka_emu_raster_traces.argtypes = [d, d, d, d, i, i, d, ll, vp, vp, vp, ll, vp, vp]
    return _lib


def set_raster_flavour(name):
    """Rasteriser flavour of the following calls: "plain" = raster_seg<RF_PLAIN>, "heavy" = RF_HEAVY (64-bit bit-sets, rows behind the
    start of a chained segment skipped).  The library chooses by lattice (oneka_api.cu: raster_flavour); here the test does."""
    lib().oneka_emu_set_raster_flavour({"plain": 0, "heavy": 1}[name])


def _p(a):
    return None if a is None else a.ctypes.data


def capture(spec, par, start_xy, mode, geom=None, max_verts=0, farfield=None, path_bbox=False, clip=None):
    """spec: FlowSpec, par: RealizationParams; mode 0 track / 1 fused (geom needed) / 2 traces.
    farfield: dict(x0, y0, tile, ntx, nty, order, eta) or None; path_bbox: also return every path's bounding box
    [R, P, 4]; clip: int32 [R, P, 4] per-path raster windows (mode 1).  -> dict of numpy arrays."""
    wxy = np.ascontiguousarray(spec.well_xy, dtype=np.float64)
    start = np.ascontiguousarray(start_xy, dtype=np.float64)
    R, P = len(par), len(start)
    end = np.zeros((R, P, 2))
    nverts = np.zeros((R, P), dtype=np.int32)
    status = np.zeros((R, P), dtype=np.uint8)
    attempts = np.zeros((R, P), dtype=np.int32)
    verts = np.zeros((R, P, max_verts, 2)) if mode == 2 else None
    counts = np.zeros((geom.nrows, geom.ncols), dtype=np.uint32) if mode == 1 else None
    stats = np.zeros(16, dtype=np.uint64)
    bbox = np.zeros(4)
    pbb = np.zeros((R, P, 4)) if path_bbox else None
    if clip is not None:
        clip = np.ascontiguousarray(clip, dtype=np.int32).reshape(R, P, 4)
        assert clip.ctypes.data % 16 == 0
    g = geom
    ff = farfield or dict(x0=0.0, y0=0.0, tile=1.0, ntx=1, nty=1, order=0, eta=0.3)
    rc = lib().oneka_emu_capture(
        mode, len(wxy), _p(wxy), float(spec.xtarget), float(spec.ytarget), int(bool(spec.confined)),
        float(spec.duration), float(spec.tol), float(spec.maxstep), int(spec.max_attempts), R, P,
        _p(par.q), _p(par.cond), _p(par.poro), _p(par.thick), _p(par.coef), _p(start),
        g.xmin if g else 0.0, g.ymin if g else 0.0, g.deltax if g else 1.0, g.deltay if g else 1.0,
        g.nrows if g else 0, g.ncols if g else 0, float(spec.umbra), _p(counts),
        int(max_verts), _p(verts), _p(end), _p(nverts), _p(status), _p(attempts),
        int(ff["order"]), float(ff["eta"]), float(ff["x0"]), float(ff["y0"]), float(ff["tile"]), int(ff["ntx"]), int(ff["nty"]),
        _p(stats), _p(bbox), _p(pbb), _p(clip))
    if rc != 0:
        raise RuntimeError("oneka_emu_capture failed: %d" % rc)
    return dict(end_xy=end, nverts=nverts, status=status, attempts=attempts, verts=verts, counts=counts, path_bbox=pbb,
                stats=dict(attempts=int(stats[0]), steps=int(stats[1]), paths=int(stats[2]), n_not_ok=int(stats[3]),
                           n_clipped=int(stats[4]), exact_tests=int(stats[5]), bbox=tuple(bbox)))


def raster_traces(geom, umbra, traces, real_of, nreal):
    traces = [np.ascontiguousarray(t, dtype=np.float64).reshape(-1, 2) for t in traces]
    off = np.zeros(len(traces) + 1, dtype=np.int64)
    for k, t in enumerate(traces):
        off[k + 1] = off[k] + len(t)
    verts = np.ascontiguousarray(np.concatenate(traces, axis=0)) if traces else np.zeros((0, 2))
    real_of = np.ascontiguousarray(real_of, dtype=np.int32)
    counts = np.zeros((geom.nrows, geom.ncols), dtype=np.uint32)
    nexact = C.c_ulonglong(0)
    rc = lib().oneka_emu_raster_traces(geom.xmin, geom.ymin, geom.deltax, geom.deltay, geom.nrows, geom.ncols, float(umbra),
                                       len(traces), _p(off), _p(verts), _p(real_of), int(nreal), _p(counts), C.byref(nexact))
    if rc != 0:
        raise RuntimeError("oneka_emu_raster_traces failed: %d" % rc)
    return counts, int(nexact.value)

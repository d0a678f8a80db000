"""ctypes binding of liboneka_b200.so -- the C ABI declared in include/oneka_b200.h.

This is the binding a maintainer of the reference would add (INTEGRATION.md shows it in
isolation).  The library is built in-tree by `__graft_entry__.build()`; there is no CPU
fallback: if the library is missing, or no sm_100 device is usable, calls raise.
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
# ONEKA_B200_LIB lets a developer A/B an experimental build of the same ABI (never a different implementation)
LIB_PATH = os.environ.get("ONEKA_B200_LIB") or os.path.join(_HERE, "liboneka_b200.so")

OK = 0
PATH_OK, PATH_AQUIFER_DRY, PATH_MAX_ATTEMPT, PATH_NONFINITE, PATH_TRACE_FULL = 0, 1, 2, 3, 4

# every symbol include/oneka_b200.h declares (tests/test_cabi_symbols.py checks header <-> library <-> this list)
SYMBOLS = [
    "oneka_last_error", "oneka_abi_version", "oneka_create", "oneka_destroy", "oneka_set_stream",
    "oneka_set_workspace_limit", "oneka_synchronize", "oneka_launch_count", "oneka_set_profiling",
    "oneka_kernel_ms", "oneka_eval_points_host", "oneka_trace", "oneka_raster_traces", "oneka_capture",
    "oneka_read_stats", "oneka_reset_stats", "oneka_capture_host", "oneka_fp64_probe", "oneka_path_bboxes",
    "oneka_capture_clipped", "oneka_count_histogram", "oneka_gaussian_smooth", "oneka_capture_guarded",
    "oneka_set_farfield", "oneka_farfield_eval_host", "oneka_set_farfield_unconfined",
    "oneka_capture_tracked", "oneka_comm_unique_id", "oneka_comm_init_rank", "oneka_comm_attach", "oneka_comm_destroy",
    "oneka_allreduce_counts", "oneka_allreduce_f64", "oneka_distancesquared_host", "oneka_red_probe", "oneka_farfield_info", "oneka_set_raster_mode", "oneka_raster_flavour",
]


class OnekaError(RuntimeError):
    """A negative return code from the C ABI."""


class ModelDesc(C.Structure):
    """oneka_model_desc"""
    _fields_ = [("nw", C.c_int32), ("confined", C.c_int32), ("base", C.c_double), ("xo", C.c_double),
                ("yo", C.c_double), ("duration", C.c_double), ("tol", C.c_double), ("maxstep", C.c_double),
                ("max_attempts", C.c_int64)]


class Lattice(C.Structure):
    """oneka_lattice"""
    _fields_ = [("xmin", C.c_double), ("ymin", C.c_double), ("deltax", C.c_double), ("deltay", C.c_double),
                ("nrows", C.c_int32), ("ncols", C.c_int32), ("umbra", C.c_double)]


class Stats(C.Structure):
    """oneka_stats"""
    _fields_ = [("attempts", C.c_uint64), ("steps", C.c_uint64), ("paths", C.c_uint64), ("n_not_ok", C.c_uint64),
                ("n_clipped", C.c_uint64), ("exact_tests", C.c_uint64), ("bbox", C.c_double * 4)]

    def as_dict(self):
        return dict(attempts=int(self.attempts), steps=int(self.steps), paths=int(self.paths),
                    n_not_ok=int(self.n_not_ok), n_clipped=int(self.n_clipped), exact_tests=int(self.exact_tests),
                    bbox=tuple(self.bbox))


_lib = None
_vp = C.c_void_p


def load():
    """Load the shared library (once).  Raises if it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise OnekaError("%s not found: run `python -c 'import __graft_entry__ as g; g.build()'` first "
                         "(there is no CPU fallback)" % LIB_PATH)
    L = C.CDLL(LIB_PATH)
    L.oneka_last_error.restype = C.c_char_p
    L.oneka_abi_version.restype = C.c_int
    L.oneka_create.restype = _vp
    L.oneka_create.argtypes = [C.c_int]
    L.oneka_destroy.argtypes = [_vp]
    L.oneka_destroy.restype = None
    L.oneka_set_stream.argtypes = [_vp, _vp]
    L.oneka_set_workspace_limit.argtypes = [_vp, C.c_uint64]
    L.oneka_synchronize.argtypes = [_vp]
    L.oneka_launch_count.argtypes = [_vp]
    L.oneka_launch_count.restype = C.c_uint64
    L.oneka_set_profiling.argtypes = [_vp, C.c_int]
    L.oneka_kernel_ms.argtypes = [_vp, C.POINTER(C.c_double), C.POINTER(C.c_double), C.POINTER(C.c_uint64), C.c_int]
    L.oneka_eval_points_host.argtypes = [_vp, C.POINTER(ModelDesc), _vp, _vp, C.c_double, C.c_double, C.c_double,
                                         _vp, C.c_int64, _vp, _vp]
    L.oneka_trace.argtypes = [_vp, C.POINTER(ModelDesc), _vp, C.c_int64, C.c_int32, _vp, _vp, _vp, _vp, _vp, _vp,
                              C.c_int32, _vp, _vp, _vp, _vp]
    L.oneka_raster_traces.argtypes = [_vp, C.POINTER(Lattice), C.c_int64, _vp, _vp, _vp, C.c_int64, _vp]
    L.oneka_capture.argtypes = [_vp, C.POINTER(ModelDesc), C.POINTER(Lattice), _vp, C.c_int64, C.c_int32,
                                _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp]
    L.oneka_path_bboxes.argtypes = [_vp, C.POINTER(ModelDesc), _vp, C.c_int64, C.c_int32, _vp, _vp, _vp, _vp, _vp, _vp,
                                    _vp, _vp]
    L.oneka_capture_clipped.argtypes = [_vp, C.POINTER(ModelDesc), C.POINTER(Lattice), _vp, C.c_int64, C.c_int32,
                                        _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp]
    L.oneka_capture_guarded.argtypes = [_vp, C.POINTER(ModelDesc), C.POINTER(Lattice), _vp, C.c_int64, C.c_int32,
                                        _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp]
    L.oneka_count_histogram.argtypes = [_vp, _vp, C.c_int64, C.c_int32, _vp]
    L.oneka_gaussian_smooth.argtypes = [_vp, _vp, C.c_int32, C.c_int32, C.c_double, _vp, C.c_int32, _vp, _vp]
    L.oneka_read_stats.argtypes = [_vp, C.POINTER(Stats)]
    L.oneka_reset_stats.argtypes = [_vp]
    L.oneka_capture_host.argtypes = [_vp, C.POINTER(ModelDesc), C.POINTER(Lattice), _vp, C.c_int64, C.c_int32,
                                     _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, C.POINTER(Stats)]
    L.oneka_fp64_probe.argtypes = [_vp, C.c_int, C.POINTER(C.c_double), C.POINTER(C.c_double)]
    L.oneka_set_farfield.argtypes = [_vp, C.c_int32, _vp, C.c_double, C.c_double, C.c_double, C.c_double, C.c_double,
                                     C.c_int32, C.c_int32, C.c_int32, C.c_double, C.c_int32, C.POINTER(C.c_int32),
                                     C.POINTER(C.c_double)]
    L.oneka_set_farfield_unconfined.argtypes = [_vp, C.c_int]
    L.oneka_farfield_eval_host.argtypes = [C.c_int32, _vp, _vp, C.c_double, C.c_double, C.c_double, C.c_double, C.c_double,
                                           C.c_int32, C.c_int32, C.c_int32, C.c_double, C.c_int32, C.c_int64, _vp, _vp, _vp]
    L.oneka_capture_tracked.argtypes = [_vp, C.POINTER(ModelDesc), C.POINTER(Lattice), _vp, C.c_int64, C.c_int32,
                                        _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp]
    L.oneka_comm_unique_id.argtypes = [_vp]
    L.oneka_comm_init_rank.argtypes = [_vp, C.c_int32, C.c_int32, _vp]
    L.oneka_comm_attach.argtypes = [_vp, _vp, C.c_int32, C.c_int32]
    L.oneka_comm_destroy.argtypes = [_vp]
    L.oneka_allreduce_counts.argtypes = [_vp, _vp, C.c_uint64]
    L.oneka_allreduce_f64.argtypes = [_vp, _vp, C.c_uint64, C.c_int32]
    L.oneka_distancesquared_host.argtypes = [_vp, C.c_int64, _vp, _vp]
    L.oneka_farfield_info.argtypes = [_vp, C.POINTER(C.c_int32), C.POINTER(C.c_int32), C.POINTER(C.c_int32), C.POINTER(C.c_uint64),
                                      C.POINTER(C.c_uint64), C.POINTER(C.c_uint64)]
    L.oneka_set_raster_mode.argtypes = [_vp, C.c_int32]
    L.oneka_raster_flavour.argtypes = [_vp, C.c_double, C.c_double, C.c_int32]
    L.oneka_red_probe.argtypes = [_vp, C.c_int32, C.c_uint64, C.c_int32, C.POINTER(C.c_double), C.POINTER(C.c_double)]
    for name in SYMBOLS:
        getattr(L, name)                  # AttributeError here = header / library mismatch
    _lib = L
    return L


def check(rc):
    if rc != OK:
        raise OnekaError("oneka C ABI error %d: %s" % (rc, load().oneka_last_error().decode()))


def create(device):
    L = load()
    h = L.oneka_create(int(device))
    if not h:
        raise OnekaError("oneka_create(%d) failed: %s" % (device, L.oneka_last_error().decode()))
    return h

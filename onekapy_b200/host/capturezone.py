"""compute_capturezone / compute_backtrace: drop-ins for oneka/capturezone.py.

The reference hands the solver an opaque Python closure `feval(xy)`.  A kernel cannot call
it, so the velocity field is LIFTED from the closure instead: either `feval` is a
BacktraceVelocity (what this package's stochastic/deterministic drivers build), or it is a
closure of the reference's shape (oneka/stochastic.py:253-260) whose free variable `mo` is a
Model-like object; which method it calls tells confined from unconfined.  Anything else
raises TypeError -- there is no CPU integrator to fall back to.
"""
import logging

import numpy as np

from ..engine import FlowSpec, RealizationParams, default_engine, start_ring
from ..lattice import LatticeGeom
from .. import _cabi

log = logging.getLogger(__name__)


class BacktraceVelocity:
    """Callable carrying the model: feval(xy) -> -V(xy) (oneka/stochastic.py:253-260)."""

    def __init__(self, model, confined):
        self.model = model
        self.confined = bool(confined)

    def __call__(self, xy):
        if self.confined:
            Vx, Vy = self.model.compute_velocity_confined(xy[0], xy[1])
        else:
            Vx, Vy = self.model.compute_velocity(xy[0], xy[1])
        return np.array([-Vx, -Vy])


def lift_model(feval):
    """-> (model, confined) from a BacktraceVelocity or a reference-style closure."""
    if isinstance(feval, BacktraceVelocity):
        return feval.model, feval.confined
    code = getattr(feval, "__code__", None)
    cells = getattr(feval, "__closure__", None)
    if code is not None and cells:
        env = dict(zip(code.co_freevars, (c.cell_contents for c in cells)))
        for mo in env.values():
            if all(hasattr(mo, a) for a in ("conductivity", "porosity", "thickness", "wells", "xo", "yo", "coef")):
                names = code.co_names
                if "compute_velocity_confined" in names:
                    return mo, True
                if "compute_velocity" in names:
                    return mo, False
    raise TypeError("feval must be a BacktraceVelocity or a closure over a Model calling "
                    "compute_velocity[_confined] (oneka/stochastic.py:253-260); arbitrary Python "
                    "velocity functions cannot run on the GPU and there is no CPU fallback")


def _spec_and_params(mo, confined, xtarget, ytarget, rtarget, npaths, duration, spacing, umbra, tol, maxstep):
    wells = np.array([[float(w[0]), float(w[1]), float(w[3])] for w in mo.wells], dtype=np.float64).reshape(-1, 3)
    if float(mo.xo) != float(xtarget) or float(mo.yo) != float(ytarget):
        # the kernels take the regional origin separately from the start ring, so this is allowed
        pass
    spec = FlowSpec(well_xy=np.ascontiguousarray(wells[:, :2]), xtarget=float(mo.xo), ytarget=float(mo.yo),
                    rtarget=float(rtarget), npaths=int(npaths), duration=float(duration), base=float(mo.base),
                    spacing=float(spacing), umbra=float(umbra), confined=confined, tol=float(tol), maxstep=float(maxstep))
    par = RealizationParams(q=wells[None, :, 2], cond=[mo.conductivity], poro=[mo.porosity], thick=[mo.thickness],
                            coef=np.reshape(mo.coef, (1, 6)))
    return spec, par


def compute_capturezone(xtarget, ytarget, rtarget, npaths, duration, pfield, umbra, weight, tol, maxstep, feval):
    """One realization: npaths backtraces from the ring around the target, chronicled into the
    caller's `pfield`, then pfield.register(weight)  (oneka/capturezone.py:51-123).

    The reference expands the field trace by trace and clips each trace to the grid as it was at
    that moment; that order dependence is reproduced exactly (Engine.clip_windows): a tracking pass
    gives every path's bounding box, the field is expanded once to their union, and the fused pass
    rasterises each path inside the window the reference's grid had when that path was inserted.
    Truncated traces (AquiferError) are chronicled as far as they got and a warning is logged, as
    the reference's bare except does (:249-250)."""
    mo, confined = lift_model(feval)
    eng = default_engine()
    spec, par = _spec_and_params(mo, confined, xtarget, ytarget, rtarget, npaths, duration,
                                 pfield.deltax, umbra, tol, maxstep)
    start = start_ring(xtarget, ytarget, rtarget, npaths)
    dp = eng.upload(spec, par, start)
    eng.reset_stats()
    bb = eng.path_bboxes(spec, dp)                          # tracking only
    bbox = eng.read_stats()["bbox"]
    if pfield.nrows == 0 or pfield.ncols == 0:
        # un-anchored field: the reference's rasterize() of the FIRST trace takes expand()'s empty-field branch
        # (probabilityfield.py:205-220), which puts that trace's (min x, min y) at node [1, 1]; every later trace grows it
        b0 = bb[0, 0].cpu().numpy()
        pfield.expand(float(b0[0]), float(b0[1]), float(b0[2]), float(b0[3]))
    base = LatticeGeom.of_field(pfield)
    pfield.expand(bbox[0], bbox[1], bbox[2], bbox[3])       # union of the per-trace expands (:335)
    geom = LatticeGeom.of_field(pfield)
    clip = eng.clip_windows(base, geom, bb)
    counts = eng.new_counts(geom)
    eng.reset_stats()
    eng.capture(spec, dp, geom, counts, clip=clip)
    st = eng.read_stats()
    if st["n_not_ok"]:
        log.warning(' %d trace(s) terminated prematurely before duration.', st["n_not_ok"])
    pfield.rgrid |= (counts.cpu().numpy() != 0)
    pfield.register(weight)                                 # :123


def compute_backtrace(xs, ys, duration, tol, maxstep, feval):
    """One backtrace; returns the list of (x, y) vertices (oneka/capturezone.py:127-253)."""
    mo, confined = lift_model(feval)
    eng = default_engine()
    spec, par = _spec_and_params(mo, confined, mo.xo, mo.yo, 0.0, 1, duration, 1.0, 0.0, tol, maxstep)
    start = np.array([[xs, ys]], dtype=np.float64)
    dp = eng.upload(spec, par, start)
    max_verts = 4096
    while True:
        out = eng.trace(spec, dp, max_verts=max_verts)
        if out["status"][0, 0] != _cabi.PATH_TRACE_FULL:
            break
        max_verts *= 4
    if out["status"][0, 0] != _cabi.PATH_OK:
        log.warning(' trace terminated prematurely (status %d) before duration.', int(out["status"][0, 0]))
    n = int(out["nverts"][0, 0])
    v = out["verts"][0, 0, :n]
    return [(float(a), float(b)) for a, b in v]

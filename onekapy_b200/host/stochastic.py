"""create_stochastic_capturezone: drop-in for oneka/stochastic.py.

The realization loop of the reference (oneka/stochastic.py:220-265) does, per realization,
(1)-(4) sample discharges / k / n / H, fit A..F and draw them from the fitted multivariate
normal, then (5)-(6) track npaths particles and chronicle them.  Here (1)-(4) run on the
host for ALL realizations first -- with the reference's own sequence of RNG calls, so a
seeded np.random gives the same rows -- and (5)-(6) are ONE batched GPU call.
"""
import logging

import numpy as np

from .model import Model, fit_batch
from .probabilityfield import ProbabilityField
from ..engine import FlowSpec, RealizationParams, default_engine

log = logging.getLogger('Oneka')


class Error(Exception):
    """Base class for module errors (oneka/stochastic.py:60-62)."""


class DistributionError(Error):
    """Invalid distribution specification (oneka/stochastic.py:65-72)."""


def _legacy_rows(dists, nrows):
    """[nrows, len(dists)] variates, consuming np.random's global stream exactly as the reference's
    scalar calls do (oneka/stochastic.py:224-233 -> :297-304): realization by realization, one
    `generate_random_variate` per entry of `dists` in order.

    The legacy RandomState draws ONE double U per uniform / triangular variate and none for a Dirac
    value, so the whole table comes from one `random_sample((nrows, nvariates))` block pushed through
    NumPy's own formulas (legacy-distributions.c: uniform `lo + (hi - lo) U`; triangular by inversion,
    `left + sqrt(U leftprod)` if `U <= ratio` else `right - sqrt((1 - U) rightprod)`).  Bit-identical
    to the scalar loop (tests/test_host_logic.py::test_vectorised_sampling_is_the_scalar_stream)."""
    out = np.empty((nrows, len(dists)))
    rnd = []                                                  # columns that consume a double
    for j, d in enumerate(dists):
        if type(d) is not tuple:                              # Dirac (:297-298): the value itself
            out[:, j] = d
        elif len(d) == 2:
            rnd.append(j)
        elif len(d) == 3:
            left, mode, right = d
            if left > mode:                                   # the ValueErrors np.random.triangular raises
                raise ValueError("left > mode")
            if mode > right:
                raise ValueError("mode > right")
            if left == right:
                raise ValueError("left == right")
            rnd.append(j)
        else:
            raise DistributionError('<arg> must be a scalar, pair, or triple.')
    if not rnd or nrows == 0:
        return out
    # Uniform columns together, triangular columns together (per-column parameters broadcast along the rows), in row blocks
    # that stay in cache with the temporaries reused: the same elementwise operations in the same order as the scalar
    # formulas and the same row-major consumption of the stream, so the values are bit-identical.  (A Python loop over 200
    # strided columns of 16 MB arrays cost 130 ms per 10 000 realizations of the 200-well field, most of the sampling.)
    m = len(rnd)
    cu = np.array([c for c, j in enumerate(rnd) if len(dists[j]) == 2], dtype=np.intp)
    ct = np.array([c for c, j in enumerate(rnd) if len(dists[j]) == 3], dtype=np.intp)
    lo_u = np.array([float(dists[rnd[c]][0]) for c in cu])
    hi_u = np.array([float(dists[rnd[c]][1]) for c in cu])
    left = np.array([float(dists[rnd[c]][0]) for c in ct])
    mode = np.array([float(dists[rnd[c]][1]) for c in ct])
    right = np.array([float(dists[rnd[c]][2]) for c in ct])
    base = right - left
    leftbase = mode - left
    ratio = leftbase / base
    leftprod = leftbase * base
    rightprod = (right - mode) * base
    contiguous = rnd == list(range(rnd[0], rnd[0] + m))
    all_tri = len(ct) == m
    BLK = max(1, 262144 // m)                                 # ~2 MB of doubles per temporary
    for r0 in range(0, nrows, BLK):
        r1 = min(nrows, r0 + BLK)
        U = np.random.random_sample((r1 - r0, m))
        V = U if all_tri else np.empty_like(U)
        if len(cu):
            V[:, cu] = lo_u + (hi_u - lo_u) * U[:, cu]       # np.random.uniform: lo + (hi - lo) U
        if len(ct):
            u = U if all_tri else U[:, ct]
            lo = np.sqrt(u * leftprod)
            lo += left                                        # left + sqrt(U leftprod)
            hi = np.subtract(1.0, u)
            hi *= rightprod
            np.sqrt(hi, out=hi)
            np.subtract(right, hi, out=hi)                    # right - sqrt((1 - U) rightprod)
            res = np.where(u <= ratio, lo, hi)
            if all_tri:
                V = res
            else:
                V[:, ct] = res
        if contiguous:
            out[r0:r1, rnd[0]:rnd[0] + m] = V
        else:
            out[r0:r1, rnd] = V
    return out


def _mvn_rows(rng, ev, factor):
    """One `rng.multivariate_normal(ev[i], cov[i])` per row (oneka/stochastic.py:241), batched.

    Generator.multivariate_normal (method='svd') is `mean + z @ (u sqrt(s)).T` with z = 6 standard
    normals and (u, s, _) = svd(cov) -- `host.model.mvn_factor`; a [R, 6] block of normals is the same
    stream as R calls, and the stacked LAPACK svd / matmul reproduce the per-row calls to summation-order
    rounding, < 1e-12 standard deviations (tests/test_host_logic.py::test_vectorised_sampling_is_the_scalar_stream)."""
    z = rng.standard_normal(ev.shape)
    return ev + np.matmul(z[:, None, :], factor)[:, 0, :]


SAMPLE_CHUNK = 32768          # realizations per host batch: bounds WA [chunk, nobs, 6] (160 MB at 100 observations)


def iter_realizations(nrealizations, base, c_dist, p_dist, t_dist, stochastic_wells, observations,
                      xtarget, ytarget, rng=None, fit_method="auto", log_rows=True, chunk=SAMPLE_CHUNK):
    """Steps (1)-(4) of oneka/stochastic.py:186-199, `chunk` realizations at a time: yields (RealizationParams, ev, cov).

    Vectorised over the realizations of a chunk, yet the rows are the ones the reference's loop would produce: the
    discharges / conductivity / porosity / thickness consume np.random's global state in the reference's order
    (`_legacy_rows`: realization by realization, so chunks concatenate to the same stream), and A..F come from
    `rng.multivariate_normal` row by row (`_mvn_rows`; likewise sequential) -- the reference builds a fresh unseeded
    default_rng() per realization (:241), so without `rng` one unseeded Generator serves all rows; pass a seeded one for
    reproducible rows.  The two streams are independent, so interleaving them chunk by chunk changes nothing.

    fit_method: "lstsq" = the reference's LAPACK calls per realization; "qr" = one stacked factorisation (same estimator,
    rounding differs at ~1e-9 relative); "auto" = lstsq up to 4096 realizations IN TOTAL, qr above."""
    nw = len(stochastic_wells)
    R = int(nrealizations)
    dists = [w[3] for w in stochastic_wells] + [c_dist, p_dist, t_dist]
    wxy = np.array([[w[0], w[1]] for w in stochastic_wells], dtype=float).reshape(-1, 2)
    obs = np.array(observations, dtype=float).reshape(-1, 4)
    if fit_method == "auto":
        fit_method = "lstsq" if R <= 4096 else "qr"
    g = rng if rng is not None else np.random.default_rng()
    chunk = max(1, int(chunk))
    for r0 in range(0, R, chunk):
        r1 = min(R, r0 + chunk)
        rows = _legacy_rows(dists, r1 - r0)
        q, k, n, H = np.ascontiguousarray(rows[:, :nw]), rows[:, nw].copy(), rows[:, nw + 1].copy(), rows[:, nw + 2].copy()
        ev, cov, fac = fit_batch(obs, xtarget, ytarget, base, wxy, q, k, H, method=fit_method, with_factor=True)
        coef = _mvn_rows(g, ev, fac)
        if log_rows and log.isEnabledFor(logging.INFO):
            for i in range(r1 - r0):
                recharge = 2 * (coef[i, 0] + coef[i, 1])
                log.info('Realization #{0:d}: {1:.2f}, {2:.2f}, {3:.2f}, {4:.2f}, {5:.4e}'
                         .format(r0 + i, base, k[i], n[i], H[i], recharge))
        yield RealizationParams(q=q, cond=k, poro=n, thick=H, coef=coef), ev, cov


def sample_realizations(nrealizations, base, c_dist, p_dist, t_dist, stochastic_wells, observations,
                        xtarget, ytarget, rng=None, fit_method="auto", log_rows=True):
    """All realizations at once -> (RealizationParams, ev [R, 6], cov [R, 6, 6]); see iter_realizations."""
    parts = list(iter_realizations(nrealizations, base, c_dist, p_dist, t_dist, stochastic_wells, observations, xtarget, ytarget,
                                   rng=rng, fit_method=fit_method, log_rows=log_rows))
    if not parts:
        nw = len(stochastic_wells)
        return RealizationParams(q=np.zeros((0, nw)), cond=np.zeros(0), poro=np.zeros(0), thick=np.zeros(0), coef=np.zeros((0, 6))), \
            np.zeros((0, 6)), np.zeros((0, 6, 6))
    return RealizationParams.concat([p for p, _, _ in parts]), np.concatenate([e for _, e, _ in parts]), np.concatenate([c for _, _, c in parts])


import os as _os
STREAM_CHUNK = int(_os.environ.get("ONEKA_STREAM_CHUNK", "2048"))   # realizations per chunk when the drop-in call overlaps its host part with the GPU (~30 ms of GPU work at C3)


def create_stochastic_capturezone(
        target, npaths, duration, nrealizations,
        base, c_dist, p_dist, t_dist,
        stochastic_wells, observations,
        spacing, umbra, confined, tol, maxstep, rng=None, engine=None, exact_clip=True):
    """Same signature and return value as oneka/stochastic.py:76-81 (+ optional rng, engine).

    Returns a ProbabilityField whose pgrid holds, per node, the number of realizations whose
    capture zone covers it, and total_weight = nrealizations."""
    xtarget, ytarget, rtarget = stochastic_wells[target][0:3]
    spec = FlowSpec(well_xy=np.array([[w[0], w[1]] for w in stochastic_wells], dtype=float).reshape(-1, 2),
                    xtarget=float(xtarget), ytarget=float(ytarget), rtarget=float(rtarget), npaths=int(npaths),
                    duration=float(duration), base=float(base), spacing=float(spacing), umbra=float(umbra),
                    confined=bool(confined), tol=float(tol), maxstep=float(maxstep))
    eng = engine if engine is not None else default_engine()
    R = int(nrealizations)
    # exact_clip=True reproduces the reference's auto-expanding grid cell for cell (ONE fused pass + a fix-up of the few
    # realizations the growing grid clipped); False rasterises on the final lattice directly (a superset differing in
    # ~1e-3 of the cells)
    if exact_clip and R >= 2 * STREAM_CHUNK:
        # sampling and fit of chunk k + 1 run on the host while the GPU tracks chunk k (Engine.run_exact consumes the
        # generator one fused launch at a time); the rows are the ones one monolithic sampling would produce
        chunks = (p for p, _, _ in iter_realizations(R, base, c_dist, p_dist, t_dist, stochastic_wells, observations, xtarget, ytarget,
                                                     rng=rng, chunk=STREAM_CHUNK))
        res = eng.run_exact(spec, chunks, total=R)
    else:
        params, _, _ = sample_realizations(R, base, c_dist, p_dist, t_dist, stochastic_wells, observations, xtarget, ytarget, rng=rng)
        res = eng.run_exact(spec, params) if exact_clip else eng.run(spec, params)
    if res["stats"]["n_not_ok"]:
        log.warning(' %d trace(s) terminated prematurely before duration.', res["stats"]["n_not_ok"])
    return ProbabilityField.from_counts(res["geom"], res["counts"], res["total_weight"])


def generate_random_variate(arg):
    """Dirac / uniform / triangular variate (oneka/stochastic.py:274-306)."""
    if type(arg) is not tuple:
        value = arg
    elif len(arg) == 2:
        value = np.random.uniform(arg[0], arg[1])
    elif len(arg) == 3:
        value = np.random.triangular(arg[0], arg[1], arg[2])
    else:
        raise DistributionError('<arg> must be a scalar, pair, or triple.')
    return value


def compute_variate_mean(arg):
    """Mean of the Dirac / uniform / triangular distribution (oneka/stochastic.py:310-342)."""
    if type(arg) is not tuple:
        value = arg
    elif len(arg) == 2:
        value = (arg[0] + arg[1]) / 2.0
    elif len(arg) == 3:
        value = (arg[0] + arg[1] + arg[2]) / 3.0
    else:
        raise DistributionError('<arg> must be a scalar, pair, or triple.')
    return value


def isdistribution(arg, lb, ub):
    """Do the arguments define a valid distribution within [lb, ub]? (oneka/stochastic.py:345-376)."""
    if isinstance(arg, int) or isinstance(arg, float):
        return lb <= arg <= ub
    if isinstance(arg, tuple) or isinstance(arg, list):
        if len(arg) == 2:
            return lb <= arg[0] <= arg[1] <= ub
        elif len(arg) == 3:
            return lb <= arg[0] <= arg[1] <= arg[2] <= ub
    return False

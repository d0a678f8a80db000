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


def sample_realizations(nrealizations, base, c_dist, p_dist, t_dist, stochastic_wells, observations,
                        xtarget, ytarget, rng=None, fit_method="lstsq", log_rows=True):
    """Steps (1)-(4) of oneka/stochastic.py:186-199 for all realizations -> RealizationParams.

    RNG call order per realization is the reference's (:224-233): one variate per well, then
    conductivity, porosity, thickness, all from np.random's global state.  A..F come from
    `rng.multivariate_normal` -- the reference builds a fresh unseeded default_rng() per
    realization (:241); pass a seeded Generator for reproducible rows."""
    nw = len(stochastic_wells)
    q = np.zeros((nrealizations, nw))
    k = np.zeros(nrealizations)
    n = np.zeros(nrealizations)
    H = np.zeros(nrealizations)
    for i in range(nrealizations):
        for j, w in enumerate(stochastic_wells):
            q[i, j] = generate_random_variate(w[3])
        k[i] = generate_random_variate(c_dist)
        n[i] = generate_random_variate(p_dist)
        H[i] = generate_random_variate(t_dist)
    wxy = np.array([[w[0], w[1]] for w in stochastic_wells], dtype=float).reshape(-1, 2)
    obs = np.array(observations, dtype=float).reshape(-1, 4)
    ev, cov = fit_batch(obs, xtarget, ytarget, base, wxy, q, k, H, method=fit_method)
    coef = np.zeros((nrealizations, 6))
    for i in range(nrealizations):
        g = rng if rng is not None else np.random.default_rng()
        coef[i] = g.multivariate_normal(ev[i], cov[i])
        if log_rows:
            recharge = 2 * (coef[i, 0] + coef[i, 1])
            log.info('Realization #{0:d}: {1:.2f}, {2:.2f}, {3:.2f}, {4:.2f}, {5:.4e}'
                     .format(i, base, k[i], n[i], H[i], recharge))
    return RealizationParams(q=q, cond=k, poro=n, thick=H, coef=coef), ev, cov


def create_stochastic_capturezone(
        target, npaths, duration, nrealizations,
        base, c_dist, p_dist, t_dist,
        stochastic_wells, observations,
        spacing, umbra, confined, tol, maxstep, rng=None, engine=None, exact_clip=True):
    """Same signature and return value as oneka/stochastic.py:76-81 (+ optional rng, engine).

    Returns a ProbabilityField whose pgrid holds, per node, the number of realizations whose
    capture zone covers it, and total_weight = nrealizations."""
    xtarget, ytarget, rtarget = stochastic_wells[target][0:3]
    params, _, _ = sample_realizations(nrealizations, base, c_dist, p_dist, t_dist, stochastic_wells,
                                       observations, xtarget, ytarget, rng=rng)
    spec = FlowSpec(well_xy=np.array([[w[0], w[1]] for w in stochastic_wells], dtype=float).reshape(-1, 2),
                    xtarget=float(xtarget), ytarget=float(ytarget), rtarget=float(rtarget), npaths=int(npaths),
                    duration=float(duration), base=float(base), spacing=float(spacing), umbra=float(umbra),
                    confined=bool(confined), tol=float(tol), maxstep=float(maxstep))
    eng = engine if engine is not None else default_engine()
    # exact_clip=True reproduces the reference's auto-expanding grid cell for cell (one extra tracking pass);
    # False rasterises on the final lattice directly (a superset differing in ~1e-3 of the cells)
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

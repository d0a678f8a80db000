"""create_deterministic_capturezone: drop-in for oneka/deterministic.py:65-235.

One realization at the distribution means with coef = the fitted expected values (no
multivariate-normal draw, :199)."""
import logging

import numpy as np

from .model import Model
from .probabilityfield import ProbabilityField
from .stochastic import compute_variate_mean
from ..engine import FlowSpec, RealizationParams, default_engine

log = logging.getLogger('Oneka')


def mean_realization(base, c_dist, p_dist, t_dist, stochastic_wells, observations, xtarget, ytarget):
    """oneka/deterministic.py:180-199 -> (RealizationParams with one row, fitted Model)."""
    wells = []
    for w in stochastic_wells:
        xw, yw, rw = w[0:3]
        wells.append([xw, yw, rw, compute_variate_mean(w[3])])
    conductivity = compute_variate_mean(c_dist)
    porosity = compute_variate_mean(p_dist)
    thickness = compute_variate_mean(t_dist)
    mo = Model(base, conductivity, porosity, thickness, wells)
    coef_ev, coef_cov = mo.fit_regional_flow(observations, xtarget, ytarget)
    mo.coef = np.reshape(coef_ev, [6, ])
    par = RealizationParams(q=np.array([[w[3] for w in wells]], dtype=float), cond=[conductivity], poro=[porosity],
                            thick=[thickness], coef=mo.coef[None, :])
    return par, mo


def create_deterministic_capturezone(
        target, npaths, duration,
        base, c_dist, p_dist, t_dist,
        stochastic_wells, observations,
        spacing, umbra, confined, tol, maxstep, engine=None, exact_clip=True):
    """Same signature and return value as oneka/deterministic.py:65-70 (+ optional engine)."""
    xtarget, ytarget, rtarget = stochastic_wells[target][0:3]
    par, mo = mean_realization(base, c_dist, p_dist, t_dist, stochastic_wells, observations, xtarget, ytarget)
    log.info("Deterministic Capture Zone Parameters: base = %.2f conductivity = %.2f porosity = %.2f recharge = %+.4e",
             base, mo.conductivity, mo.porosity, -2 * (mo.coef[0] + mo.coef[1]))
    spec = FlowSpec(well_xy=np.array([[w[0], w[1]] for w in stochastic_wells], dtype=float).reshape(-1, 2),
                    xtarget=float(xtarget), ytarget=float(ytarget), rtarget=float(rtarget), npaths=int(npaths),
                    duration=float(duration), base=float(base), spacing=float(spacing), umbra=float(umbra),
                    confined=bool(confined), tol=float(tol), maxstep=float(maxstep))
    eng = engine if engine is not None else default_engine()
    res = eng.run_exact(spec, par) if exact_clip else eng.run(spec, par)
    return ProbabilityField.from_counts(res["geom"], res["counts"], res["total_weight"])

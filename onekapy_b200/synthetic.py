"""Synthetic well fields for the scaling configurations of BASELINE.json (C4, C5).

C4: 200 pumping wells on a jittered grid over a 10 km x 10 km square, the target well at the
centre, discharges triangular +-20 % around a per-well mode, aquifer properties distributed as
in the perham case (data/perham.py:9-11 of the reference), ~100 synthetic head observations
from a planar regional head plus noise; spacing / umbra / tol / maxstep as perham.
Everything is generated from one recorded seed; pumping wells only (a backward trace that
reaches an injection well never terminates cleanly in the reference).
"""
import numpy as np


def well_field(nwells=200, seed=2020, side=10000.0, nobs=100):
    rng = np.random.default_rng(seed)
    g = int(np.ceil(np.sqrt(nwells)))
    cell = side / g
    ij = np.stack(np.meshgrid(np.arange(g), np.arange(g), indexing="ij"), axis=-1).reshape(-1, 2)
    centre = ij[np.argmin(np.abs(ij - (g - 1) / 2.0).sum(axis=1))]
    order = [tuple(centre)] + [tuple(t) for t in ij if tuple(t) != tuple(centre)]
    order = order[:nwells]
    x0, y0 = 300000.0, 5160000.0                                # UTM-like magnitudes, as the field cases
    wells = []
    for n, (i, j) in enumerate(order):
        jx, jy = (0.0, 0.0) if n == 0 else rng.uniform(-0.3, 0.3, size=2) * cell
        x = np.round(x0 + (i + 0.5) * cell + jx)
        y = np.round(y0 + (j + 0.5) * cell + jy)
        mode = 1200.0 if n == 0 else float(np.round(rng.uniform(200.0, 900.0), 2))
        wells.append((float(x), float(y), 0.2 if n == 0 else 1.0, (0.8 * mode, mode, 1.2 * mode)))
    # planar regional head (gradient ~ 1.5 m/km towards -x) + noise; observations kept > 150 m from wells
    obs = []
    wxy = np.array([[w[0], w[1]] for w in wells])
    while len(obs) < nobs:
        x, y = x0 + rng.uniform(0, side), y0 + rng.uniform(0, side)
        if np.min(np.hypot(wxy[:, 0] - x, wxy[:, 1] - y)) < 150.0:
            continue
        z = 420.0 + 1.5e-3 * (x - x0) + 0.4e-3 * (y - y0) + rng.normal(0.0, 0.3)
        obs.append((float(np.round(x)), float(np.round(y)), float(np.round(z, 2)), 1.5))
    return dict(projectname="synthetic %d-well field (seed %d)" % (nwells, seed), target=0, npaths=1000,
                duration=10 * 365.25, nrealizations=1000000, base=380.0, c_dist=(12.0, 65.0, 120.0),
                p_dist=(0.20, 0.25), t_dist=(10.0, 20.0, 30.0), buffer=100.0, spacing=4.0, umbra=8.0, smooth=2.0,
                confined=True, tol=1.0, maxstep=10.0, wells=wells, observations=obs, seed=seed)


def sample_rows_fast(pb, nreal, seed, fit_method="qr"):
    """Vectorised sampling of parameter rows for large R (same distributions as
    host.stochastic.sample_realizations, NumPy Generator streams instead of the legacy global
    state, one stacked fit, MVN draws through one Cholesky per realization batch)."""
    from .host.model import fit_batch
    from .host.utilities import filter_obs
    from .engine import RealizationParams
    rng = np.random.default_rng(seed)

    def draw(d, size):
        if not isinstance(d, tuple):
            return np.full(size, float(d))
        if len(d) == 2:
            return rng.uniform(d[0], d[1], size=size)
        if d[0] == d[2]:
            return np.full(size, float(d[0]))
        return rng.triangular(d[0], d[1], d[2], size=size)

    wells = pb["wells"]
    q = np.stack([draw(w[3], nreal) for w in wells], axis=1)
    k = draw(pb["c_dist"], nreal)
    n = draw(pb["p_dist"], nreal)
    H = draw(pb["t_dist"], nreal)
    obs = np.array(filter_obs(pb["observations"], wells, pb["buffer"]), dtype=float)
    xt, yt = wells[pb["target"]][0:2]
    wxy = np.array([[w[0], w[1]] for w in wells], dtype=float)
    coef = np.zeros((nreal, 6))
    for r0 in range(0, nreal, 65536):
        r1 = min(nreal, r0 + 65536)
        ev, cov = fit_batch(obs, xt, yt, pb["base"], wxy, q[r0:r1], k[r0:r1], H[r0:r1], method=fit_method)
        # x = ev + L z with cov = L L^T, scaled for conditioning (columns of cov span ~1e-7 .. 1e5)
        sd = np.sqrt(np.einsum("rii->ri", cov))
        corr = cov / (sd[:, :, None] * sd[:, None, :])
        L = np.linalg.cholesky(corr + 1e-12 * np.eye(6)[None])
        z = rng.standard_normal((r1 - r0, 6))
        coef[r0:r1] = ev + sd * np.einsum("rij,rj->ri", L, z)
    return RealizationParams(q=q, cond=k, poro=n, thick=H, coef=coef)

"""Model: the Oneka-type single-layer analytic element model (drop-in for oneka/model.py).

Evaluation methods (compute_potential / discharge / head / velocity[_confined],
reference oneka/model.py:207-427) are evaluated by the CUDA field functions through
oneka_eval_points_host -- the same device code the tracker calls six times per attempt.
The fitting methods (oneka/model.py:430-606) stay on the host, as BASELINE.json's north_star
specifies, but are vectorised over observations and (fit_batch) over realizations.
"""
import logging

import numpy as np

log = logging.getLogger("Oneka")

ONE_OVER_4PI = 0.07957747154594767      # oneka/model.py:264


class Error(Exception):
    """Base class for module errors (oneka/model.py:46-48)."""


class RangeError(Error):
    """An argument is out of range (oneka/model.py:51-53)."""


class AquiferError(Error):
    """The aquifer is dry at the point (oneka/model.py:56-58)."""


class Model:
    """Same constructor and attributes as oneka/model.py:131-196."""

    def __init__(self, base, conductivity, porosity, thickness, wells, xo=0, yo=0, coef=np.zeros((6, ))):
        self.base = base
        self.conductivity = conductivity
        self.porosity = porosity
        self.thickness = thickness
        self.wells = wells
        self.xo = xo
        self.yo = yo
        self.coef = coef

    @property
    def coef(self):
        return self._coef

    @coef.setter
    def coef(self, coef):
        self._coef = np.reshape(coef, [6, ])

    def __repr__(self):
        return 'Model({0.base}, {0.conductivity}, {0.porosity}, {0.thickness})'.format(self)

    __str__ = __repr__

    # -- evaluation: CUDA ---------------------------------------------------------------------------
    def _well_arrays(self):
        w = np.array([[float(t[0]), float(t[1]), float(t[3])] for t in self.wells], dtype=np.float64).reshape(-1, 3)
        return np.ascontiguousarray(w[:, :2]), np.ascontiguousarray(w[:, 2])

    def evaluate(self, pts):
        """[npts, 8] = potential, Qx, Qy, Vx_confined, Vy_confined, head, Vx, Vy for many points
        in one launch (nan where the scalar methods would raise AquiferError)."""
        from ..engine import default_engine
        wxy, q = self._well_arrays()
        return default_engine().eval_points(wxy, q, self.base, self.conductivity, self.porosity, self.thickness,
                                            self.xo, self.yo, self.coef, pts)

    def compute_potential(self, x, y):
        """Discharge potential [m^3/d] at (x, y)  (oneka/model.py:207-237)."""
        return float(self.evaluate([(x, y)])[0, 0])

    def compute_potential_wells_only(self, x, y):
        """Wells-only part of the potential (oneka/model.py:240-266).  Host arithmetic: this is
        the fitting helper (construct_fit, :565), not part of the tracking path."""
        return float(wells_potential(np.array([[x, y]], dtype=float), *self._well_arrays())[0])

    def compute_discharge(self, x, y):
        """Vertically integrated discharge [Qx, Qy]  (oneka/model.py:269-315)."""
        o = self.evaluate([(x, y)])[0]
        return [float(o[1]), float(o[2])]

    def compute_head(self, x, y):
        """Piezometric head above the base (oneka/model.py:318-350); AquiferError if potential <= 0."""
        o = self.evaluate([(x, y)])[0]
        if np.isnan(o[5]):
            raise AquiferError("potential_to_head: potential <= 0")
        return float(o[5])

    def compute_velocity(self, x, y):
        """Seepage velocity, unconfined/confined mixed (oneka/model.py:353-389)."""
        o = self.evaluate([(x, y)])[0]
        if np.isnan(o[5]):
            raise AquiferError("potential_to_head: potential <= 0")
        if np.isnan(o[6]):
            raise AquiferError("discharge_to_velocity: head <= 0")
        return (float(o[6]), float(o[7]))

    def compute_velocity_confined(self, x, y):
        """Seepage velocity assuming confined flow (oneka/model.py:392-427)."""
        o = self.evaluate([(x, y)])[0]
        return (float(o[3]), float(o[4]))

    # -- fitting: host ----------------------------------------------------------------------------------
    def fit_regional_flow(self, obs, xo, yo):
        """Weighted least squares fit of A..F (oneka/model.py:430-494).  Sets xo, yo, coef."""
        WA, Wb = self.construct_fit(obs, xo, yo)
        coef_ev, coef_cov = self.compute_fit(WA, Wb)
        self.xo = xo
        self.yo = yo
        self.coef = coef_ev
        return (coef_ev, coef_cov)

    def construct_fit(self, obs, xo, yo):
        """WA [nobs, 6], Wb [nobs, 1]  (oneka/model.py:497-567), vectorised over observations."""
        wxy, q = self._well_arrays()
        ob = np.array(obs, dtype=np.float64).reshape(-1, 4)
        WA, Wb = construct_fit_batch(ob, xo, yo, self.base, wxy, q[None, :], np.array([self.conductivity], dtype=float),
                                     np.array([self.thickness], dtype=float))
        return (WA[0], Wb[0][:, None])

    @staticmethod
    def compute_fit(WA, Wb):
        """lstsq + inv(WA^T WA)  (oneka/model.py:570-606)."""
        try:
            coef_ev = np.linalg.lstsq(WA, Wb, rcond=-1)[0]
        except np.linalg.LinAlgError:
            log.error(' numpy.linalg.lstsq: failed')
            raise
        coef_cov = np.linalg.inv(np.matmul(WA.T, WA))
        return (coef_ev, coef_cov)


# ----------------------------------------------------------------------------------------------------------
def wells_potential(pts, well_xy, q):
    """sum_w q_w ln(r^2)/(4 pi) at pts[n, 2]; q may be [nw] or [R, nw] -> [n] or [R, n].

    Accumulated well by well in the reference's order (oneka/model.py:259-266) so that every
    element sees the same sequence of IEEE operations as the scalar loop."""
    pts = np.asarray(pts, dtype=np.float64)
    q = np.asarray(q, dtype=np.float64)
    single = q.ndim == 1
    q2 = q[None, :] if single else q
    pot = np.zeros((q2.shape[0], len(pts)))
    for j in range(len(well_xy)):
        dx = pts[:, 0] - well_xy[j, 0]
        dy = pts[:, 1] - well_xy[j, 1]
        r2 = dx * dx + dy * dy
        pot += q2[:, j, None] * np.log(r2)[None, :] * ONE_OVER_4PI
    return pot[0] if single else pot


def construct_fit_batch(obs, xo, yo, base, well_xy, q, cond, thick):
    """oneka/model.py:545-565 for R realizations at once.

    obs[nobs, 4] = (x, y, z_ev, z_std);  q[R, nw], cond[R], thick[R]
    -> WA[R, nobs, 6], Wb[R, nobs].  First-order-second-moment head -> potential, two regimes."""
    obs = np.asarray(obs, dtype=np.float64).reshape(-1, 4)
    cond = np.asarray(cond, dtype=np.float64).reshape(-1, 1)
    thick = np.asarray(thick, dtype=np.float64).reshape(-1, 1)
    x, y, z_ev, z_std = obs[:, 0], obs[:, 1], obs[:, 2], obs[:, 3]
    head = (z_ev - base)[None, :]                                         # [1, nobs]
    if np.any(head <= 0):
        raise RangeError("model.fit_coeficient: elevation < base")        # :558-559
    zs = z_std[None, :]
    conf = head >= thick                                                  # :551  [R, nobs]
    pot_ev = np.where(conf, cond * thick * (head - 0.5 * thick), 0.5 * cond * (head ** 2 + zs ** 2))
    pot_std = np.where(conf, cond * thick * zs, cond * head * zs)
    dx = x - xo
    dy = y - yo
    A = np.stack([dx ** 2, dy ** 2, dx * dy, dx, dy, np.ones_like(dx)], axis=1)      # [nobs, 6]  :564
    WA = A[None, :, :] / pot_std[:, :, None]
    Wb = (pot_ev - wells_potential(obs[:, :2], well_xy, q)) / pot_std     # :565
    return WA, Wb


def fit_batch(obs, xo, yo, base, well_xy, q, cond, thick, method="lstsq"):
    """Regional-flow fit for R realizations -> (coef_ev[R, 6], coef_cov[R, 6, 6]).

    method="lstsq": per realization np.linalg.lstsq(rcond=-1) and inv(WA^T WA), i.e. the very
                    LAPACK calls of oneka/model.py:599,604 on identically built matrices;
    method="qr":    one stacked Householder QR for all realizations (for R >~ 1e5):
                    ev = R^-1 Q^T Wb, cov = R^-1 R^-T.  Same estimator, rounding differs."""
    WA, Wb = construct_fit_batch(obs, xo, yo, base, well_xy, q, cond, thick)
    R = WA.shape[0]
    ev = np.zeros((R, 6))
    cov = np.zeros((R, 6, 6))
    if method == "lstsq":
        for i in range(R):
            e, c = Model.compute_fit(WA[i], Wb[i][:, None])
            ev[i] = e[:, 0]
            cov[i] = c
    elif method == "qr":
        Q, Rm = np.linalg.qr(WA)                                          # stacked
        rhs = np.einsum("rij,ri->rj", Q, Wb)
        ev = np.linalg.solve(Rm, rhs[:, :, None])[:, :, 0]
        Rinv = np.linalg.inv(Rm)
        cov = Rinv @ np.transpose(Rinv, (0, 2, 1))
    else:
        raise ValueError("method must be 'lstsq' or 'qr'")
    return ev, cov

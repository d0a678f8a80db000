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


def well_log_table(pts, well_xy):
    """G[nw, n] = ln(r^2)/(4 pi) between wells and points: the realization-independent part of
    compute_potential_wells_only (oneka/model.py:259-266); the wells' potential is q @ G."""
    pts = np.asarray(pts, dtype=np.float64)
    dx = pts[None, :, 0] - well_xy[:, None, 0]
    dy = pts[None, :, 1] - well_xy[:, None, 1]
    return np.log(dx * dx + dy * dy) * ONE_OVER_4PI


def _fit_shared(obs, xo, yo, base, well_xy, q, cond, thick, regime):
    """Fit for realizations whose observations are ALL in one regime of oneka/model.py:551-556.

    There the weights factor: pot_std[r, o] = c_r s_o with (c_r, s_o) = (cond_r thick_r, z_std_o) when every
    head >= thickness ("confined"), (cond_r, head_o z_std_o) when every head < thickness ("unconfined").  A common
    scalar on the weights leaves the estimate alone, so ONE Householder QR of A / s serves every realization:
        ev_r = pinv(A/s) ((pot_ev_r - wells_r)/s),   cov_r = c_r^2 inv((A/s)^T (A/s)),
    and the wells' potential is the product q @ G.  The per-realization work is two small GEMMs."""
    x, y, z_ev, z_std = obs[:, 0], obs[:, 1], obs[:, 2], obs[:, 3]
    head = z_ev - base
    cond = cond.reshape(-1, 1)
    thick = thick.reshape(-1, 1)
    if regime == "confined":
        c = cond * thick
        s = z_std
        pot_ev = c * (head[None, :] - 0.5 * thick)
    else:
        c = cond
        s = head * z_std
        pot_ev = 0.5 * cond * (head ** 2 + z_std ** 2)[None, :]
    dx = x - xo
    dy = y - yo
    A0 = np.stack([dx ** 2, dy ** 2, dx * dy, dx, dy, np.ones_like(dx)], axis=1) / s[:, None]
    Q0, R0 = np.linalg.qr(A0)
    R0inv = np.linalg.inv(R0)
    pinv = R0inv @ Q0.T                                                  # [6, nobs]
    C0 = R0inv @ R0inv.T
    rhs = (pot_ev - q @ well_log_table(obs[:, :2], well_xy)) / s[None, :]
    ev = rhs @ pinv.T
    cov = (c * c)[:, :, None] * C0[None, :, :]
    return ev, cov, c[:, 0], C0


def mvn_factor(cov):
    """F with  x = mean + z @ F  distributed N(mean, cov): numpy's Generator.multivariate_normal (method
    'svd', what oneka/stochastic.py:241 calls) uses F = (u sqrt(s))^T with (u, s, _) = svd(cov).  Stacked."""
    u, s, _ = np.linalg.svd(cov)
    return np.swapaxes(u * np.sqrt(s)[..., None, :], -1, -2)


def fit_batch(obs, xo, yo, base, well_xy, q, cond, thick, method="lstsq", with_factor=False):
    """Regional-flow fit for R realizations -> (coef_ev[R, 6], coef_cov[R, 6, 6]) and, with_factor, the
    multivariate-normal factors `mvn_factor(cov)` [R, 6, 6] (shared-QR rows: one svd, scaled by c_r).

    method="lstsq": per realization np.linalg.lstsq(rcond=-1) and inv(WA^T WA), i.e. the very
                    LAPACK calls of oneka/model.py:599,604 on identically built matrices;
    method="qr":    for R >~ 1e4.  Realizations whose observations all sit in one head regime (every
                    shipped data set: heads of 80..420 m over 5..150 m of aquifer) share one QR
                    (`_fit_shared`); the rest get one stacked Householder QR, ev = R^-1 Q^T Wb,
                    cov = R^-1 R^-T.  Same estimator, rounding differs (~1e-9 relative on ev)."""
    obs = np.asarray(obs, dtype=np.float64).reshape(-1, 4)
    q = np.asarray(q, dtype=np.float64)
    cond = np.asarray(cond, dtype=np.float64).reshape(-1)
    thick = np.asarray(thick, dtype=np.float64).reshape(-1)
    R = len(cond)
    ev = np.zeros((R, 6))
    cov = np.zeros((R, 6, 6))
    fac = np.zeros((R, 6, 6)) if with_factor else None
    if method == "lstsq":
        WA, Wb = construct_fit_batch(obs, xo, yo, base, well_xy, q, cond, thick)
        for i in range(R):
            e, c = Model.compute_fit(WA[i], Wb[i][:, None])
            ev[i] = e[:, 0]
            cov[i] = c
        if with_factor:
            fac[:] = mvn_factor(cov)
    elif method == "qr":
        head = obs[:, 2] - base
        if np.any(head <= 0):
            raise RangeError("model.fit_coeficient: elevation < base")    # :558-559
        all_conf = thick <= head.min()                                    # head >= thickness everywhere (:551)
        all_unc = thick > head.max()
        for rows, regime in ((all_conf, "confined"), (all_unc, "unconfined")):
            if rows.any():
                ev[rows], cov[rows], c, C0 = _fit_shared(obs, xo, yo, base, well_xy, q[rows], cond[rows], thick[rows],
                                                       regime)
                if with_factor:                                           # svd(c^2 C0) = (u, c^2 s, .)
                    fac[rows] = c[:, None, None] * mvn_factor(C0)[None, :, :]
        mixed = ~(all_conf | all_unc)
        if mixed.any():
            WA, Wb = construct_fit_batch(obs, xo, yo, base, well_xy, q[mixed], cond[mixed], thick[mixed])
            Q, Rm = np.linalg.qr(WA)                                      # stacked
            rhs = np.einsum("rij,ri->rj", Q, Wb)
            ev[mixed] = np.linalg.solve(Rm, rhs[:, :, None])[:, :, 0]
            Rinv = np.linalg.inv(Rm)
            cov[mixed] = Rinv @ np.transpose(Rinv, (0, 2, 1))
            if with_factor:
                fac[mixed] = mvn_factor(cov[mixed])
    else:
        raise ValueError("method must be 'lstsq' or 'qr'")
    return (ev, cov, fac) if with_factor else (ev, cov)

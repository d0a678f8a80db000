/*
 * oneka_oracle.c -- CPU restatement of OnekaPy's capture-zone hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  This file is the parity checker and the reported CPU
 * baseline.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs may load it.  The product path (onekapy_b200/, include/,
 * oneka/) never links, imports or calls anything in oracle/.
 *
 * The reference is pure Python + NumPy (no native code), so there is no compiled
 * `oracle/_ref`; parity is PINNED instead by (i) the reference's own known-answer
 * tests (tests/test_model.py:45-76, tests/test_probabilityfield.py:33-53) and
 * (ii) fixtures produced by executing the unmodified reference in the build
 * container (tests/golden/make_golden.py), which tests/test_oracle_golden.py
 * checks this file against: traces vertex by vertex, grids cell by cell.
 *
 * Every function cites the reference lines it restates.  Expressions keep the
 * reference's left-to-right evaluation order; compile with -ffp-contract=off so
 * that no multiply-add is fused (NumPy scalar arithmetic is unfused IEEE double).
 *
 * Build:  make -C oracle        (gcc -O2 -ffp-contract=off -fopenmp -shared)
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <float.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define ORACLE_OK          0
#define ORACLE_AQUIFER_DRY 1   /* AquiferError raised inside feval (model.py:343-344, 380-381) */
#define ORACLE_MAX_ATTEMPT 2   /* guard the reference lacks */
#define ORACLE_NONFINITE   3   /* guard the reference lacks (it would spin forever) */
#define ORACLE_TRACE_FULL  4   /* caller's vertex buffer too small */

typedef struct {
    int nw;
    const double *wxy;      /* [nw][2] well coordinates            (model.py:158-171) */
    const double *q;        /* [nw]    well discharges                                */
    double base, k, n, H;   /* base, conductivity, porosity, thickness (model.py:179-182) */
    double xo, yo;          /* local origin                          (model.py:185-186) */
    double c[6];            /* A..F                                  (model.py:187)     */
} omodel;

/* ---- oneka/model.py:259-266  compute_potential_wells_only --------------------------- */
static double o_potential_wells(const omodel *m, double x, double y)
{
    double potential = 0.0;
    for (int i = 0; i < m->nw; ++i) {
        double dx = x - m->wxy[2 * i];
        double dy = y - m->wxy[2 * i + 1];
        double r2 = dx * dx + dy * dy;
        potential += m->q[i] * log(r2) * 0.07957747154594767;
    }
    return potential;
}

/* ---- oneka/model.py:226-237  compute_potential -------------------------------------- */
static double o_potential(const omodel *m, double x, double y)
{
    double dx = x - m->xo;
    double dy = y - m->yo;
    double potential = (m->c[0] * (dx * dx) + m->c[1] * (dy * dy)
                        + m->c[2] * dx * dy
                        + m->c[3] * dx + m->c[4] * dy
                        + m->c[5]);
    potential += o_potential_wells(m, x, y);
    return potential;
}

/* ---- oneka/model.py:300-315  compute_discharge -------------------------------------- */
static void o_discharge(const omodel *m, double x, double y, double *Qx, double *Qy)
{
    double dx = x - m->xo;
    double dy = y - m->yo;
    double qx = -(2.0 * m->c[0] * dx + m->c[2] * dy + m->c[3]);
    double qy = -(2.0 * m->c[1] * dy + m->c[2] * dx + m->c[4]);
    for (int i = 0; i < m->nw; ++i) {
        dx = x - m->wxy[2 * i];
        dy = y - m->wxy[2 * i + 1];
        double r2 = dx * dx + dy * dy;
        qx += -m->q[i] * dx / r2 * 0.15915494309189535;
        qy += -m->q[i] * dy / r2 * 0.15915494309189535;
    }
    *Qx = qx;
    *Qy = qy;
}

/* ---- oneka/model.py:341-350  compute_head ; returns nonzero for AquiferError -------- */
static int o_head(const omodel *m, double x, double y, double *head)
{
    double potential = o_potential(m, x, y);
    if (potential <= 0)
        return ORACLE_AQUIFER_DRY;
    else if (potential < 0.5 * m->k * (m->H * m->H))
        *head = sqrt(2.0 * potential / m->k);
    else
        *head = ((potential + 0.5 * m->k * (m->H * m->H)) / (m->k * m->H));
    return ORACLE_OK;
}

/* ---- oneka/model.py:377-389  compute_velocity --------------------------------------- */
static int o_velocity(const omodel *m, double x, double y, double *Vx, double *Vy)
{
    double Qx, Qy, head;
    o_discharge(m, x, y, &Qx, &Qy);
    int rc = o_head(m, x, y, &head);
    if (rc) return rc;
    if (head <= 0) {
        return ORACLE_AQUIFER_DRY;
    } else if (head > m->H) {
        *Vx = Qx / (m->H * m->n);
        *Vy = Qy / (m->H * m->n);
    } else {
        *Vx = Qx / (head * m->n);
        *Vy = Qy / (head * m->n);
    }
    return ORACLE_OK;
}

/* ---- oneka/model.py:423-427  compute_velocity_confined ------------------------------ */
static void o_velocity_confined(const omodel *m, double x, double y, double *Vx, double *Vy)
{
    double Qx, Qy;
    o_discharge(m, x, y, &Qx, &Qy);
    *Vx = Qx / (m->H * m->n);
    *Vy = Qy / (m->H * m->n);
}

/* ---- oneka/stochastic.py:253-260, deterministic.py:221-228  feval closures ---------- */
static inline int o_feval(const omodel *m, int confined, const double xy[2], double out[2])
{
    double Vx, Vy;
    if (confined) {
        o_velocity_confined(m, xy[0], xy[1], &Vx, &Vy);
    } else {
        int rc = o_velocity(m, xy[0], xy[1], &Vx, &Vy);
        if (rc) return rc;
    }
    out[0] = -Vx;
    out[1] = -Vy;
    return ORACLE_OK;
}

/* np.linalg.norm(v) of a 2-vector is sqrt(v.dot(v)).  The dot is BLAS ddot; with the
 * OpenBLAS kernels NumPy ships the two-term sum is formed as fma(v1,v1,v0*v0) on FMA
 * hardware.  ORACLE_NORM_FMA=1 (default) selects that: with it this file reproduces the
 * executed reference's traces BIT FOR BIT (tests/test_oracle_golden.py, 152 traces);
 * with the unfused sum they agree to ~1e-11 m only. */
#ifndef ORACLE_NORM_FMA
#define ORACLE_NORM_FMA 1
#endif
static inline double o_norm2(double v0, double v1)
{
#if ORACLE_NORM_FMA
    return sqrt(fma(v1, v1, v0 * v0));
#else
    return sqrt(v0 * v0 + v1 * v1);
#endif
}

/* ---- oneka/capturezone.py:199-253  compute_backtrace -------------------------------- *
 * on_vertex(ctx, x, y) is called for every vertex appended to `vertices`
 * (the start point first, capturezone.py:215, then each accepted step, :245).     */
typedef void (*vertex_fn)(void *ctx, double x, double y);

static int o_backtrace(const omodel *m, int confined, double xs, double ys, double duration,
                       double tol, double maxstep, int64_t max_attempts,
                       vertex_fn on_vertex, void *ctx,
                       int64_t *nattempts_out, double end_xy[2])
{
    const double EPS = DBL_EPSILON;                                   /* :200 */
    /* Dormand-Prince constants, :202-209 */
    const double a20 = 1.0 / 5.0;
    const double a30 = 3.0 / 40.0, a31 = 9.0 / 40.0;
    const double a40 = 44.0 / 45.0, a41 = -56.0 / 15.0, a42 = 32.0 / 9.0;
    const double a50 = 19372.0 / 6561.0, a51 = -25360.0 / 2187.0, a52 = 64448.0 / 6561.0, a53 = -212.0 / 729.0;
    const double a60 = 9017.0 / 3168.0, a61 = -355.0 / 33.0, a62 = 46732.0 / 5247.0, a63 = 49.0 / 176.0, a64 = -5103.0 / 18656.0;
    const double a70 = 35.0 / 384.0, a72 = 500.0 / 1113.0, a73 = 125.0 / 192.0, a74 = -2187.0 / 6784.0, a75 = 11.0 / 84.0;
    const double e0 = 71.0 / 57600.0, e1 = -1.0 / 40.0, e2 = -71.0 / 16695.0, e3 = 71.0 / 1920.0, e4 = -17253.0 / 339200.0, e5 = 22.0 / 525.0;

    double t = 0;                                                      /* :212 */
    double sgn = (duration > 0) ? 1.0 : ((duration < 0) ? -1.0 : 0.0);
    double dt = 0.1 * sgn;                                             /* :213 */
    double xy[2] = {xs, ys};
    double k1[2], k2[2], k3[2], k4[2], k5[2], k6[2], xyt[2], arg[2];
    int64_t nattempts = 0;
    int rc = ORACLE_OK;

    on_vertex(ctx, xs, ys);                                            /* :215 */

    rc = o_feval(m, confined, xy, k1);                                 /* :219 */
    if (rc) goto done;

    while (fabs(t) < fabs(duration)) {                                 /* :221 */
        if (nattempts >= max_attempts) { rc = ORACLE_MAX_ATTEMPT; break; }
        if (!(isfinite(dt) && isfinite(xy[0]) && isfinite(xy[1]))) { rc = ORACLE_NONFINITE; break; }
        ++nattempts;
        if (fabs(t + dt) > fabs(duration))                             /* :223-224 */
            dt = duration - t;

        for (int c = 0; c < 2; ++c) arg[c] = xy[c] + dt * (a20 * k1[c]);                       /* :227 */
        if ((rc = o_feval(m, confined, arg, k2))) break;
        for (int c = 0; c < 2; ++c) arg[c] = xy[c] + dt * (a30 * k1[c] + a31 * k2[c]);         /* :228 */
        if ((rc = o_feval(m, confined, arg, k3))) break;
        for (int c = 0; c < 2; ++c) arg[c] = xy[c] + dt * (a40 * k1[c] + a41 * k2[c] + a42 * k3[c]);   /* :229 */
        if ((rc = o_feval(m, confined, arg, k4))) break;
        for (int c = 0; c < 2; ++c) arg[c] = xy[c] + dt * (a50 * k1[c] + a51 * k2[c] + a52 * k3[c] + a53 * k4[c]);  /* :230 */
        if ((rc = o_feval(m, confined, arg, k5))) break;
        for (int c = 0; c < 2; ++c) arg[c] = xy[c] + dt * (a60 * k1[c] + a61 * k2[c] + a62 * k3[c] + a63 * k4[c] + a64 * k5[c]);  /* :231 */
        if ((rc = o_feval(m, confined, arg, k6))) break;

        for (int c = 0; c < 2; ++c)                                                             /* :233 */
            xyt[c] = xy[c] + dt * (a70 * k1[c] + a72 * k3[c] + a73 * k4[c] + a74 * k5[c] + a75 * k6[c]);

        if ((rc = o_feval(m, confined, xyt, k2))) break;               /* :236  (k7, named k2) */
        double est = 0.0;                                              /* :237-238, inf-norm */
        for (int c = 0; c < 2; ++c) {
            double v = fabs(dt * (e0 * k1[c] + e1 * k2[c] + e2 * k3[c] + e3 * k4[c] + e4 * k5[c] + e5 * k6[c]));
            /* np.linalg.norm(.., inf) = abs(x).max(); max() propagates nan */
            if (v > est || isnan(v)) est = v;
        }
        double ds = o_norm2(xyt[0] - xy[0], xyt[1] - xy[1]);           /* :239 */

        if ((est < tol) && (ds < maxstep)) {                           /* :241-245 */
            t = t + dt;
            k1[0] = k2[0]; k1[1] = k2[1];
            xy[0] = xyt[0]; xy[1] = xyt[1];
            on_vertex(ctx, xy[0], xy[1]);
        }

        /* :247   dt = 0.9 * min((tol/(est+EPS))**(1/5), maxstep/(ds+EPS), 10) * dt
         * Python's min(a, b, 10) keeps the first argument unless a later one is smaller. */
        double mn = pow(tol / (est + EPS), 1.0 / 5.0);
        double b = maxstep / (ds + EPS);
        if (b < mn) mn = b;
        if (10.0 < mn) mn = 10.0;
        dt = 0.9 * mn * dt;
    }
done:
    if (nattempts_out) *nattempts_out = nattempts;
    if (end_xy) { end_xy[0] = xy[0]; end_xy[1] = xy[1]; }
    return rc;
}

/* ======================= oneka/probabilityfield.py ==================================== */
typedef struct {
    double deltax, deltay;
    double xmin, xmax, ymin, ymax;
    int64_t nrows, ncols;
    double total_weight;
    double *pgrid;       /* [nrows][ncols], row = y, col = x (probabilityfield.py:148) */
    uint8_t *rgrid;      /* bool per-realization registration mask (:149)              */
    int fixed;           /* 1: lattice is frozen, expand() is a no-op (fixed-lattice parity mode) */
} ofield;

/* ---- probabilityfield.py:125-151  __init__ ------------------------------------------ */
ofield *oneka_oracle_field_new(double deltax, double deltay, double xo, double yo)
{
    if (!(deltax > 0) || !(deltay > 0)) return NULL;                  /* RangeError, :127-131 */
    ofield *f = (ofield *)calloc(1, sizeof(ofield));
    f->deltax = deltax;
    f->deltay = deltay;
    if (isnan(xo) || isnan(yo)) {
        f->nrows = 0;
        f->ncols = 0;
    } else {
        f->xmin = xo - deltax;
        f->xmax = xo + deltax;
        f->ymin = yo - deltay;
        f->ymax = yo + deltay;
        f->nrows = 3;
        f->ncols = 3;
        f->pgrid = (double *)calloc(9, sizeof(double));
        f->rgrid = (uint8_t *)calloc(9, 1);
        f->total_weight = 0.0;
    }
    return f;
}

void oneka_oracle_field_free(ofield *f)
{
    if (!f) return;
    free(f->pgrid);
    free(f->rgrid);
    free(f);
}

/* ---- probabilityfield.py:175-261  expand -------------------------------------------- */
int oneka_oracle_field_expand(ofield *f, double xmin, double xmax, double ymin, double ymax)
{
    if (xmin > xmax) return -1;                                       /* RangeError :196-200 */
    if (ymin > ymax) return -1;
    if (f->fixed) return 0;

    if (f->ncols == 0 || f->nrows == 0) {                             /* :205-220 */
        f->xmin = xmin - f->deltax;
        f->ymin = ymin - f->deltay;
        int64_t nc = (int64_t)ceil((xmax - f->xmin) / f->deltax) + 2;
        int64_t nr = (int64_t)ceil((ymax - f->ymin) / f->deltay) + 2;
        f->ncols = nc > 3 ? nc : 3;
        f->nrows = nr > 3 ? nr : 3;
        f->xmax = f->xmin + (double)(f->ncols - 1) * f->deltax;
        f->ymax = f->ymin + (double)(f->nrows - 1) * f->deltay;
        f->pgrid = (double *)calloc((size_t)(f->nrows * f->ncols), sizeof(double));
        f->rgrid = (uint8_t *)calloc((size_t)(f->nrows * f->ncols), 1);
        f->total_weight = 0.0;
    } else {                                                          /* :221-261 */
        int64_t nrows = f->nrows, ncols = f->ncols;
        int64_t rshift = 0, cshift = 0;
        while (xmin <= f->xmin) { f->xmin -= f->deltax; ncols += 1; cshift += 1; }
        while (xmax >= f->xmax) { f->xmax += f->deltax; ncols += 1; }
        while (ymin <= f->ymin) { f->ymin -= f->deltay; nrows += 1; rshift += 1; }
        while (ymax >= f->ymax) { f->ymax += f->deltay; nrows += 1; }
        if (nrows != f->nrows || ncols != f->ncols) {
            double *pg = (double *)calloc((size_t)(nrows * ncols), sizeof(double));
            uint8_t *rg = (uint8_t *)calloc((size_t)(nrows * ncols), 1);
            for (int64_t i = 0; i < f->nrows; ++i) {
                memcpy(pg + (i + rshift) * ncols + cshift, f->pgrid + i * f->ncols, (size_t)f->ncols * sizeof(double));
                memcpy(rg + (i + rshift) * ncols + cshift, f->rgrid + i * f->ncols, (size_t)f->ncols);
            }
            free(f->pgrid);
            free(f->rgrid);
            f->pgrid = pg;
            f->rgrid = rg;
            f->nrows = nrows;
            f->ncols = ncols;
        }
    }
    return 0;
}

/* ---- probabilityfield.py:407-427  distancesquared ----------------------------------- */
double oneka_oracle_distancesquared(double ax, double ay, double bx, double by, double cx, double cy)
{
    double bax = bx - ax;
    double bay = by - ay;
    double cax = cx - ax;
    double cay = cy - ay;
    double perpdot = bax * cay - bay * cax;
    double dot = bax * cax + bay * cay;
    double length2 = bax * bax + bay * bay;
    double alpha2 = perpdot * perpdot / length2;
    double beta2 = dot * dot / length2;
    double d2;
    if (dot < 0)
        d2 = alpha2 + beta2;
    else if (beta2 > length2)
        d2 = alpha2 + beta2 - 2 * dot + length2;
    else
        d2 = alpha2;
    return d2;
}

static inline int64_t imax64(int64_t a, int64_t b) { return a > b ? a : b; }
static inline int64_t imin64(int64_t a, int64_t b) { return a < b ? a : b; }

/* ---- probabilityfield.py:296-310  insert -------------------------------------------- *
 * rg / ncols / nrows may be a thread-private mask on the same lattice.                 */
static void o_insert(const ofield *f, uint8_t *rg, double ax, double ay, double bx, double by, double umbra)
{
    double umbra_squared = umbra * umbra;
    double mnx = ax < bx ? ax : bx, mxx = ax > bx ? ax : bx;          /* min(ax,bx), max(ax,bx) */
    double mny = ay < by ? ay : by, mxy = ay > by ? ay : by;
    /* Python's min(a,b) returns a unless b < a; identical values either way. */
    double fl = floor((mnx - umbra - f->xmin) / f->deltax);
    double fr = floor((mxx + umbra - f->xmin) / f->deltax);
    double fb = floor((mny - umbra - f->ymin) / f->deltay);
    double ft = floor((mxy + umbra - f->ymin) / f->deltay);
    if (!(isfinite(fl) && isfinite(fr) && isfinite(fb) && isfinite(ft))) return;   /* math.floor(nan) raises -> trace lost */
    int64_t left = imax64(0, (int64_t)fl);
    int64_t right = imin64(f->ncols, (int64_t)fr + 1);
    int64_t bottom = imax64(0, (int64_t)fb);
    int64_t top = imin64(f->nrows, (int64_t)ft + 1);
    for (int64_t j = left; j < right; ++j) {
        for (int64_t i = bottom; i < top; ++i) {
            if (!rg[i * f->ncols + j]) {
                double cx = f->xmin + (double)j * f->deltax;
                double cy = f->ymin + (double)i * f->deltay;
                if (oneka_oracle_distancesquared(ax, ay, bx, by, cx, cy) < umbra_squared)
                    rg[i * f->ncols + j] = 1;
            }
        }
    }
}

void oneka_oracle_field_insert(ofield *f, double ax, double ay, double bx, double by, double umbra)
{
    o_insert(f, f->rgrid, ax, ay, bx, by, umbra);
}

/* ---- probabilityfield.py:335-339  rasterize ----------------------------------------- */
int oneka_oracle_field_rasterize(ofield *f, int64_t n, const double *x, const double *y, double umbra)
{
    if (n <= 0) return -1;
    double x0 = x[0], x1 = x[0], y0 = y[0], y1 = y[0];
    for (int64_t i = 1; i < n; ++i) {
        if (x[i] < x0) x0 = x[i];
        if (x[i] > x1) x1 = x[i];
        if (y[i] < y0) y0 = y[i];
        if (y[i] > y1) y1 = y[i];
    }
    oneka_oracle_field_expand(f, x0, x1, y0, y1);
    for (int64_t i = 0; i < n - 1; ++i)
        o_insert(f, f->rgrid, x[i], y[i], x[i + 1], y[i + 1], umbra);
    return 0;
}

/* ---- probabilityfield.py:357-359  register ;  :376 reset ----------------------------- */
void oneka_oracle_field_register(ofield *f, double weight)
{
    f->total_weight += weight;
    int64_t n = f->nrows * f->ncols;
    for (int64_t i = 0; i < n; ++i) {
        if (f->rgrid[i]) { f->pgrid[i] += weight; f->rgrid[i] = 0; }
    }
}

void oneka_oracle_field_reset(ofield *f)
{
    memset(f->rgrid, 0, (size_t)(f->nrows * f->ncols));
}

/* geometry out: xmin xmax ymin ymax deltax deltay nrows ncols total_weight */
void oneka_oracle_field_geom(const ofield *f, double *out)
{
    out[0] = f->xmin; out[1] = f->xmax; out[2] = f->ymin; out[3] = f->ymax;
    out[4] = f->deltax; out[5] = f->deltay;
    out[6] = (double)f->nrows; out[7] = (double)f->ncols; out[8] = f->total_weight;
}
void oneka_oracle_field_freeze(ofield *f, int fixed) { f->fixed = fixed; }
void oneka_oracle_field_copy_pgrid(const ofield *f, double *out) { memcpy(out, f->pgrid, (size_t)(f->nrows * f->ncols) * sizeof(double)); }
void oneka_oracle_field_copy_rgrid(const ofield *f, uint8_t *out) { memcpy(out, f->rgrid, (size_t)(f->nrows * f->ncols)); }

/* ======================= point evaluation (tests/test_model.py pins) ================== */
/* out[npts][8] = potential, Qx, Qy, Vx_conf, Vy_conf, head, Vx, Vy  (nan where AquiferError) */
int oneka_oracle_eval(int nw, const double *wxy, const double *q, const double *par /*base,k,n,H,xo,yo*/,
                      const double *coef, int64_t npts, const double *pts, double *out)
{
    omodel m = {nw, wxy, q, par[0], par[1], par[2], par[3], par[4], par[5], {coef[0], coef[1], coef[2], coef[3], coef[4], coef[5]}};
    for (int64_t i = 0; i < npts; ++i) {
        double x = pts[2 * i], y = pts[2 * i + 1];
        double *o = out + 8 * i;
        o[0] = o_potential(&m, x, y);
        o_discharge(&m, x, y, &o[1], &o[2]);
        o_velocity_confined(&m, x, y, &o[3], &o[4]);
        o[5] = o[6] = o[7] = NAN;
        double h;
        if (o_head(&m, x, y, &h) == ORACLE_OK) {
            o[5] = h;
            double vx, vy;
            if (o_velocity(&m, x, y, &vx, &vy) == ORACLE_OK) { o[6] = vx; o[7] = vy; }
        }
    }
    return 0;
}

/* ======================= single trace (capturezone.py:127) ============================ */
typedef struct { double *v; int64_t cap, n; int overflow; } vbuf;
static void vbuf_push(void *ctx, double x, double y)
{
    vbuf *b = (vbuf *)ctx;
    if (b->n < b->cap) { b->v[2 * b->n] = x; b->v[2 * b->n + 1] = y; }
    else b->overflow = 1;
    b->n++;
}

int oneka_oracle_backtrace(int nw, const double *wxy, const double *q, const double *par, const double *coef,
                           int confined, double xs, double ys, double duration, double tol, double maxstep,
                           int64_t max_attempts, int64_t max_verts, double *verts, int64_t *nverts, int64_t *nattempts)
{
    omodel m = {nw, wxy, q, par[0], par[1], par[2], par[3], par[4], par[5], {coef[0], coef[1], coef[2], coef[3], coef[4], coef[5]}};
    vbuf b = {verts, max_verts, 0, 0};
    int rc = o_backtrace(&m, confined, xs, ys, duration, tol, maxstep, max_attempts, vbuf_push, &b, nattempts, NULL);
    *nverts = b.n;
    if (b.overflow && rc == ORACLE_OK) rc = ORACLE_TRACE_FULL;
    return rc;
}

/* ======================= capture zone over R realizations ============================= *
 * capturezone.py:110-123 inside the realization loop of stochastic.py:220-265.
 * start_xy[P][2] is the start ring (capturezone.py:113-115), computed by the caller with
 * NumPy exactly as the reference does, so that cos/sin are bit-identical on both sides.
 *
 * mode 0 (auto): one ofield, traces rasterised with rasterize() (expand + insert) in
 *                (realization, path) order: the reference's order-dependent semantics.
 * mode 1 (fixed): the field is frozen (expand is a no-op, insert clips to the lattice);
 *                realizations are independent and fan out over OpenMP threads with
 *                thread-private registration masks; counts are summed.                   */
typedef struct {
    double *x, *y; int64_t cap, n;
} tbuf;
static void tbuf_push(void *ctx, double x, double y)
{
    tbuf *b = (tbuf *)ctx;
    if (b->n == b->cap) {
        b->cap = b->cap ? 2 * b->cap : 1024;
        b->x = (double *)realloc(b->x, (size_t)b->cap * sizeof(double));
        b->y = (double *)realloc(b->y, (size_t)b->cap * sizeof(double));
    }
    b->x[b->n] = x; b->y[b->n] = y; b->n++;
}

int oneka_oracle_capture(ofield *f, int mode, int nthreads,
                         int nw, const double *wxy, double base, double xo, double yo, int confined,
                         int64_t R, const double *q /*[R][nw]*/, const double *cond, const double *poro, const double *thick,
                         const double *coef /*[R][6]*/,
                         int64_t P, const double *start_xy, double duration, double umbra, double weight,
                         double tol, double maxstep, int64_t max_attempts,
                         double *end_xy /*[R][P][2] or NULL*/, int64_t *nverts /*[R][P] or NULL*/, uint8_t *status /*[R][P] or NULL*/,
                         int64_t *total_attempts, int64_t *total_steps)
{
    int64_t att_sum = 0, step_sum = 0;
    if (mode == 0) {
        tbuf tb = {0, 0, 0, 0};
        for (int64_t r = 0; r < R; ++r) {
            omodel m = {nw, wxy, q + r * nw, base, cond[r], poro[r], thick[r], xo, yo,
                        {coef[6 * r], coef[6 * r + 1], coef[6 * r + 2], coef[6 * r + 3], coef[6 * r + 4], coef[6 * r + 5]}};
            for (int64_t p = 0; p < P; ++p) {
                tb.n = 0;
                int64_t na = 0; double e[2];
                int rc = o_backtrace(&m, confined, start_xy[2 * p], start_xy[2 * p + 1], duration, tol, maxstep,
                                     max_attempts, tbuf_push, &tb, &na, e);
                oneka_oracle_field_rasterize(f, tb.n, tb.x, tb.y, umbra);          /* capturezone.py:118-120 */
                att_sum += na; step_sum += tb.n - 1;
                if (end_xy) { end_xy[2 * (r * P + p)] = e[0]; end_xy[2 * (r * P + p) + 1] = e[1]; }
                if (nverts) nverts[r * P + p] = tb.n;
                if (status) status[r * P + p] = (uint8_t)rc;
            }
            oneka_oracle_field_register(f, weight);                                /* capturezone.py:123 */
        }
        free(tb.x); free(tb.y);
    } else {
        int saved = f->fixed;
        f->fixed = 1;
        int64_t ncell = f->nrows * f->ncols;
#ifdef _OPENMP
        if (nthreads > 0) omp_set_num_threads(nthreads);
#endif
        #pragma omp parallel reduction(+:att_sum, step_sum)
        {
            uint8_t *rg = (uint8_t *)calloc((size_t)ncell, 1);
            tbuf tb = {0, 0, 0, 0};
            #pragma omp for schedule(dynamic, 1)
            for (int64_t r = 0; r < R; ++r) {
                omodel m = {nw, wxy, q + r * nw, base, cond[r], poro[r], thick[r], xo, yo,
                            {coef[6 * r], coef[6 * r + 1], coef[6 * r + 2], coef[6 * r + 3], coef[6 * r + 4], coef[6 * r + 5]}};
                int64_t i0 = f->nrows, i1 = -1;
                for (int64_t p = 0; p < P; ++p) {
                    tb.n = 0;
                    int64_t na = 0; double e[2];
                    int rc = o_backtrace(&m, confined, start_xy[2 * p], start_xy[2 * p + 1], duration, tol, maxstep,
                                         max_attempts, tbuf_push, &tb, &na, e);
                    for (int64_t i = 0; i < tb.n - 1; ++i)
                        o_insert(f, rg, tb.x[i], tb.y[i], tb.x[i + 1], tb.y[i + 1], umbra);
                    /* rows possibly touched (for a cheap register) */
                    for (int64_t i = 0; i < tb.n; ++i) {
                        double fr = floor((tb.y[i] - umbra - f->ymin) / f->deltay), ft = floor((tb.y[i] + umbra - f->ymin) / f->deltay) + 1;
                        if (isfinite(fr) && isfinite(ft)) {
                            int64_t b = imax64(0, (int64_t)fr), t = imin64(f->nrows, (int64_t)ft);
                            if (b < i0) i0 = b;
                            if (t > i1) i1 = t;
                        }
                    }
                    att_sum += na; step_sum += tb.n - 1;
                    if (end_xy) { end_xy[2 * (r * P + p)] = e[0]; end_xy[2 * (r * P + p) + 1] = e[1]; }
                    if (nverts) nverts[r * P + p] = tb.n;
                    if (status) status[r * P + p] = (uint8_t)rc;
                }
                /* register(weight) restricted to the touched rows; pgrid += weight where set */
                for (int64_t i = i0 * f->ncols; i < i1 * f->ncols; ++i) {
                    if (rg[i]) {
                        rg[i] = 0;
                        #pragma omp atomic
                        f->pgrid[i] += weight;
                    }
                }
            }
            free(rg); free(tb.x); free(tb.y);
        }
        f->total_weight += weight * (double)R;
        f->fixed = saved;
    }
    if (total_attempts) *total_attempts = att_sum;
    if (total_steps) *total_steps = step_sum;
    return 0;
}

int oneka_oracle_num_threads(void)
{
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

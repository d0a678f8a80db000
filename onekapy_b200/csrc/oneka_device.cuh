// oneka_device.cuh -- device functions of the capture-zone hot path (sm_100a).
//
// Three pieces, each written B200-first (FP64 CUDA-core pipe for the integrator, FP32 pipe +
// L2 atomics for the rasteriser; nothing here is a contraction, so no tensor cores):
//
//   field_*      analytic velocity field          reference: oneka/model.py:269-315, 318-427
//   dopri_track  Dormand-Prince 5(4) backtrace    reference: oneka/capturezone.py:199-247
//   raster_seg   segment -> registration bitmap   reference: oneka/probabilityfield.py:296-310, 407-427
//
// Results policy: the tracker reproduces the reference's step sequence (same accept/reject
// tests, same controller, same operation order where it decides anything) with rounding-level
// differences only (FMA contraction, Newton reciprocal) -- endpoints agree to ~1e-12 relative,
// far inside the 1e-6 bar.  The rasteriser is BIT-EXACT: cells are classified in FP32 with a
// rigorous error band and every cell inside the band is re-tested with the reference's own
// unfused IEEE-double formula (__dmul_rn/__dadd_rn/__ddiv_rn are never contracted).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <math.h>
#include <float.h>
#include <string.h>

namespace oneka {

// ------------------------------------------------------------------------------------------
// The approximate MUFU instructions the kernels use.  ONEKA_EMU (tests/emu: the device code compiled for the HOST to
// unit-test kernel logic without a GPU -- test infrastructure, never part of the library) swaps in libm stand-ins;
// every consumer already tolerates the difference (Newton step after the seed, error bands around the FP32 values).
#ifdef ONEKA_EMU
__device__ __forceinline__ double ptx_rcp_approx_f64(double a) { return 1.0 / a; }
__device__ __forceinline__ float ptx_lg2_approx_f32(float a) { return log2f(a); }
__device__ __forceinline__ float ptx_sqrt_approx_f32(float a) { return sqrtf(a); }
#else
__device__ __forceinline__ double ptx_rcp_approx_f64(double a)
{
    double y;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(a));
    return y;
}
__device__ __forceinline__ float ptx_lg2_approx_f32(float a)
{
    float r;
    asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(a));
    return r;
}
__device__ __forceinline__ float ptx_sqrt_approx_f32(float a)
{
    float r;
    asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(a));
    return r;
}
#endif

// ------------------------------------------------------------------------------------------
// Status words (mirror include/oneka_b200.h)
enum : int { PATH_OK = 0, PATH_AQUIFER_DRY = 1, PATH_MAX_ATTEMPT = 2, PATH_NONFINITE = 3, PATH_TRACE_FULL = 4 };

// Per-realization constants staged in shared memory by each CTA.
struct RealConsts {
    // regional gradient terms; confined: pre-divided by thickness*porosity
    double a2, b2, c, d, e;          // 2A, 2B, C, D, E  (x 1/(H n) when confined)
    // unconfined only
    double A, B, F;                  // potential terms (model.py:226-231)
    double k, H, n;                  // conductivity, thickness, porosity
    double half_kH2;                 // 0.5*k*H^2 (model.py:345)
    double inv_Hn;                   // 1/(H n): the velocity scale wherever head >= H
    double pot_err;                  // bound on |Phi_fp32 - Phi| of the screening sum (see field_feval<false>)
    double xo, yo;
};

// UNCONFINED wells in shared memory, blocks of 4 (WELL_BLK doubles = 112 bytes each, 16-byte aligned; the confined
// store is SWELL_BLK, see field_feval):
//   [0..7]  x0 y0 x1 y1 x2 y2 x3 y3      [8..11]  w0 w1 w2 w3 (scaled discharges)      [12..13]  the same w as 4 floats
// One uniform base register addresses a whole block with immediate offsets (6 LDS.128 per 4 wells, 3 uniform
// instructions of loop control); the last block may be partial (its unused slots are zero and never read by the
// hot loop, which finishes with single wells).
constexpr int WELL_BLK = 14;
__host__ __device__ __forceinline__ constexpr int well_store_doubles(int nw) { return ((nw + 3) >> 2) * WELL_BLK; }
__device__ __forceinline__ double &well_x(double *s, int i) { return s[(i >> 2) * WELL_BLK + 2 * (i & 3)]; }
__device__ __forceinline__ double &well_y(double *s, int i) { return s[(i >> 2) * WELL_BLK + 2 * (i & 3) + 1]; }
__device__ __forceinline__ double &well_w(double *s, int i) { return s[(i >> 2) * WELL_BLK + 8 + (i & 3)]; }
__device__ __forceinline__ float &well_w32(double *s, int i) { return reinterpret_cast<float *>(s + (i >> 2) * WELL_BLK + 12)[i & 3]; }
__device__ __forceinline__ double well_x(const double *s, int i) { return s[(i >> 2) * WELL_BLK + 2 * (i & 3)]; }
__device__ __forceinline__ double well_y(const double *s, int i) { return s[(i >> 2) * WELL_BLK + 2 * (i & 3) + 1]; }
__device__ __forceinline__ double well_w(const double *s, int i) { return s[(i >> 2) * WELL_BLK + 8 + (i & 3)]; }

struct TrackParams {
    int nw;
    int P;
    long long R;                     // realizations in this launch
    double duration, tol, maxstep;
    int max_attempts;
    double xo, yo;
    const double *well_xy;           // [nw][2]
    const double *q, *cond, *poro, *thick, *coef;   // per-realization rows (already offset to this launch)
    const double *start_xy;          // [P][2]
    // optional outputs (already offset)
    double *end_xy; int *nverts; unsigned char *status; int *attempts;
    double *path_bbox;               // [R][P][4] min x, max x, min y, max y of each path's vertices, or null
    const int *clip;                 // [R][P][4] per-path raster clip window (left, right, bottom, top), or null
    unsigned int *slot_flags;        // [R] set to 1 when any segment of the realization was clipped by the lattice edge
                                     //     (such realizations are not registered by the flush; oneka_capture_guarded), or null
    // oneka_trace only
    double *verts; int max_verts;
    // statistics
    unsigned long long *stats;       // see STAT_* below
};

enum : int { STAT_ATTEMPTS = 0, STAT_STEPS = 1, STAT_PATHS = 2, STAT_NOT_OK = 3, STAT_CLIPPED = 4, STAT_EXACT = 5,
             STAT_XMIN = 6, STAT_XMAX = 7, STAT_YMIN = 8, STAT_YMAX = 9, STAT_FF_MISMATCH = 10, STAT_WORDS = 11 };

struct LatticeDev {
    double xmin, ymin, dx, dy;
    int nrows, ncols, wpr;           // wpr = 32-bit words per bitmap row
    double umbra, umbra2;            // umbra2 = umbra*umbra rounded once (probabilityfield.py:296)
    float dx32, dy32, umbra2_32, umbra32, inv_dx32;
    double maxd;                     // max(dx, dy)
    double inv_dx, inv_dy;           // 1/dx, 1/dy (window pre-quotient, see floor_div)
    // fixed-point window (raster_seg): cell index * 2^16 = (coordinate - c??) * s16?;  fixed_ok when the lattice is
    // small enough for that to fit an int32 with room to saturate
    double cxl, cxr, cyb, cyt;       // xmin + umbra, xmin - umbra, ymin + umbra, ymin - umbra
    double s16x, s16y;               // 65536/dx, 65536/dy
    int fixed_ok;
    unsigned long long words;        // words per bitmap = nrows*wpr
};

// ------------------------------------------------------------------------------------------
// 1/a for a in the normal range: MUFU.RCP64H seed + one cubic Newton step (3 DFMA).
// Relative error ~1 ulp; replaces the ~20-instruction IEEE divide in the well loop.
__device__ __forceinline__ double rcp_fast(double a)
{
    double y = ptx_rcp_approx_f64(a);
    double e = fma(-a, y, 1.0);
    double t = fma(e, e, e);
    return fma(y, t, y);
}

// seed-and-correction split: returns y0 and t such that 1/a = y0 + y0*t
//   ONEKA_RCP_ORDER 2 (default): t = e        (one quadratic Newton step on the MUFU.RCP64H seed:
//                                 relative error = seed error squared, <= ~1e-12; 9 FP64 instr / well)
//   ONEKA_RCP_ORDER 3          : t = e + e^2  (cubic step, ~1 ulp; 10 FP64 instr / well, 6 % slower)
// Measured on the golden fixtures (tests/test_gpu_parity.py): max relative vertex error 4e-13 with
// order 2, 1e-14 with order 3; both leave every step count and every grid cell unchanged.
#ifndef ONEKA_RCP_ORDER
#define ONEKA_RCP_ORDER 2
#endif
__device__ __forceinline__ void rcp_parts(double a, double &y0, double &t)
{
    y0 = ptx_rcp_approx_f64(a);
    double e = fma(-a, y0, 1.0);
#if ONEKA_RCP_ORDER == 3
    t = fma(e, e, e);
#else
    t = e;
#endif
}

// ------------------------------------------------------------------------------------------
// Backtracking velocity  f(x,y) = -V(x,y).
//
// confined   (stochastic.py:254-256 -> model.py:423-427 -> 300-315):
//     -V = [ (2A dx + C dy + D) + sum_w q_w/(2 pi) (x-x_w)/r_w^2 ] / (H n)
//   SCALED WELLS: per well the store holds b = 1/w (w = q/(2 pi H n)) and c = -(well - origin) b.  With
//   X = fma(x - xo, b, cx) = (x - x_w)/w and Y likewise, the well's term  w (x - x_w)/r^2  is  X / (X^2 + Y^2):
//   the multiplication by w is gone -- 8 FP64-pipe instructions per well (4 DFMA for X, Y and the two sums, DMUL + DFMA
//   for X^2 + Y^2, 2 DFMA of Newton) + 1 MUFU.RCP64H, for the reference's 15 flops.  Coordinates relative to the
//   origin of the regional quadratic (= the target well, stochastic.py:239) keep the cancellation inside the FMA at
//   2^-52 |x_w - xo| / |x - x_w|, <= ~1e-13 relative, the level of the Newton reciprocal; the target well's own
//   c is exactly 0.  A well with q = 0 gets b = 1e100: its term is ~1e-100, i.e. nothing.
//   (Tried and dropped, profiles/r01_notes.md: serving wells from constant memory through the uniform datapath.  With the
//   unscaled 9-instruction form it removed the coordinate loads -- C4 +3.4 % -- but the scaled form needs TWO table
//   operands in one FMA and an FP64 instruction takes only one uniform register: the second costs two MOVs per
//   well, more than the LDS it replaces.  Scaled wells from shared memory are faster than either.)
// unconfined (stochastic.py:258-260 -> model.py:377-389, 341-350, 226-237, 259-266):
//   same discharge (9 FP64 per well: the squared distance itself is needed), plus
//   Phi = A dx^2 + B dy^2 + C dx dy + D dx + E dy + F + sum q ln(r^2)/(4 pi), head from Phi (two regimes), saturated
//   thickness min(head, H); Phi <= 0 or head <= 0 is the reference's AquiferError -> PATH_AQUIFER_DRY.
constexpr int SWELL_BLK = 12;        // confined store: blocks of 4 wells x {b, cx, cy} = 6 LDS.128
// confined, scaled: one well = 8 FP64-pipe instructions + MUFU.RCP64H
__device__ __forceinline__ void scaled_term(double dx0, double dy0, double b, double cx, double cy, double &gx, double &gy)
{
    const double X = fma(dx0, b, cx);
    const double Y = fma(dy0, b, cy);
    const double rho = fma(Y, Y, X * X);
    double y0, t;
    rcp_parts(rho, y0, t);
    const double yv = fma(y0, t, y0);
    gx = fma(yv, X, gx);
    gy = fma(yv, Y, gy);
}

// unconfined: one well = 9 FP64-pipe instructions + MUFU.RCP64H (+ F2F, MUFU.LG2, FFMA of the screening sum)
__device__ __forceinline__ void well_term(double x, double y, double xw, double yw, double w, float w32,
                                          double &gx, double &gy, float &lsum32)
{
    const double dx = x - xw;
    const double dy = y - yw;
    const double r2 = fma(dy, dy, dx * dx);
    double y0, t;
    rcp_parts(r2, y0, t);
    const double s0 = w * y0;
    const double s = fma(s0, t, s0);
    gx = fma(s, dx, gx);
    gy = fma(s, dy, gy);
    // bare MUFU.LG2 (no denormal scaling: (float) r2 is a normal number for 1e-19 m < r < 1e19 m; outside, the
    // screening value is inf or nan and the comparison in field_feval sends a pumping well's neighbourhood to the FP64 path)
    const float l2 = ptx_lg2_approx_f32((float)r2);
    lsum32 = fmaf(w32, l2, lsum32);
}

template <bool CONFINED>
__device__ __forceinline__ int field_feval(const RealConsts &rc, const double *__restrict__ s_wells, int nw,
                                           double x, double y, double &fx, double &fy)
{
    const double dx0 = x - rc.xo;
    const double dy0 = y - rc.yo;
    double gx = fma(rc.a2, dx0, fma(rc.c, dy0, rc.d));
    double gy = fma(rc.b2, dy0, fma(rc.c, dx0, rc.e));
    if (CONFINED) {
        const double *p = s_wells;
        const double *const pend = s_wells + (nw >> 2) * SWELL_BLK;
        // the first two LDS.128 of the NEXT block (its first well) are issued while the current block is computed (the
        // store has a block of slack): the first chain of an iteration no longer waits for shared memory.  Measured on B200
        // (profiles/r01_notes.md): +2 % over no prefetch on C3 and C4; prefetching 3 or all 6 loads is no better.
        double2 n0 = reinterpret_cast<const double2 *>(p)[0], n1 = reinterpret_cast<const double2 *>(p)[1];
#pragma unroll 1
        for (; p != pend; p += SWELL_BLK) {
            const double2 v0 = n0, v1 = n1;
            const double2 v2 = reinterpret_cast<const double2 *>(p)[2], v3 = reinterpret_cast<const double2 *>(p)[3];
            const double2 v4 = reinterpret_cast<const double2 *>(p)[4], v5 = reinterpret_cast<const double2 *>(p)[5];
            n0 = reinterpret_cast<const double2 *>(p)[6];
            n1 = reinterpret_cast<const double2 *>(p)[7];
            scaled_term(dx0, dy0, v0.x, v0.y, v1.x, gx, gy);
            scaled_term(dx0, dy0, v1.y, v2.x, v2.y, gx, gy);
            scaled_term(dx0, dy0, v3.x, v3.y, v4.x, gx, gy);
            scaled_term(dx0, dy0, v4.y, v5.x, v5.y, gx, gy);
        }
        const int rem = nw & 3;
        if (rem > 0) scaled_term(dx0, dy0, p[0], p[1], p[2], gx, gy);
        if (rem > 1) scaled_term(dx0, dy0, p[3], p[4], p[5], gx, gy);
        if (rem > 2) scaled_term(dx0, dy0, p[6], p[7], p[8], gx, gy);
        fx = gx;
        fy = gy;
        return PATH_OK;
    }
    // unconfined: FP32 screening sum of  w_i log2(r_i^2)  (MUFU.LG2 on the XU pipe; the exact FP64 logs below are
    // needed only where the aquifer is not fully saturated)
    float lsum32 = 0.0f;
    const double *p = s_wells;
    {
        const double *const pend = s_wells + (nw >> 2) * WELL_BLK;
        // the first well of the NEXT block (coordinates and discharges) is loaded ahead, as in the confined loop
        double2 nc0 = reinterpret_cast<const double2 *>(p)[0], nw01 = reinterpret_cast<const double2 *>(p)[4];
#pragma unroll 1
        for (; p != pend; p += WELL_BLK) {
            const double2 c0 = nc0, w01 = nw01;
            nc0 = reinterpret_cast<const double2 *>(p)[7];
            nw01 = reinterpret_cast<const double2 *>(p)[11];
            const double2 c1 = reinterpret_cast<const double2 *>(p)[1];
            const double2 c2 = reinterpret_cast<const double2 *>(p)[2];
            const double2 c3 = reinterpret_cast<const double2 *>(p)[3];
            const double2 w23 = reinterpret_cast<const double2 *>(p)[5];
            const float4 wf = reinterpret_cast<const float4 *>(p)[6];
            well_term(x, y, c0.x, c0.y, w01.x, wf.x, gx, gy, lsum32);
            well_term(x, y, c1.x, c1.y, w01.y, wf.y, gx, gy, lsum32);
            well_term(x, y, c2.x, c2.y, w23.x, wf.z, gx, gy, lsum32);
            well_term(x, y, c3.x, c3.y, w23.y, wf.w, gx, gy, lsum32);
        }
    }
    {   // the last, partial block: single wells
        const int rem = nw & 3;
        const float *pf = reinterpret_cast<const float *>(p + 12);
        if (rem > 0) {
            const double2 c = reinterpret_cast<const double2 *>(p)[0];
            well_term(x, y, c.x, c.y, p[8], pf[0], gx, gy, lsum32);
        }
        if (rem > 1) {
            const double2 c = reinterpret_cast<const double2 *>(p)[1];
            well_term(x, y, c.x, c.y, p[9], pf[1], gx, gy, lsum32);
        }
        if (rem > 2) {
            const double2 c = reinterpret_cast<const double2 *>(p)[2];
            well_term(x, y, c.x, c.y, p[10], pf[2], gx, gy, lsum32);
        }
    }
    {
        // regional part of the potential (model.py:226-231)
        const double pot_reg = rc.A * dx0 * dx0 + rc.B * dy0 * dy0 + rc.c * dx0 * dy0 + rc.d * dx0 + rc.e * dy0 + rc.F;
        // Phi >= k H^2/2  <=>  head >= H  (model.py:345-349), and then V = Q/(H n) whatever Phi is (model.py:382-384;
        // at head == H both branches of :382-387 give the same value).  The screening sum decides that case:
        // sum q ln(r^2)/(4 pi) = 0.5 ln2 sum w log2(r^2), |error| <= pot_err (set when the realization is staged).
        const double pot_apx = fma(0.34657359027997264, (double)lsum32, pot_reg);
        if (pot_apx - rc.pot_err > rc.half_kH2) {
            fx = gx * rc.inv_Hn;
            fy = gy * rc.inv_Hn;
            return PATH_OK;
        }
        // not (certainly) saturated: the reference's potential with FP64 logs (model.py:259-266)
        double lsum = 0.0;
        for (int i = 0; i < nw; ++i) {
            const double dx = x - well_x(s_wells, i);
            const double dy = y - well_y(s_wells, i);
            lsum = fma(well_w(s_wells, i), log(fma(dy, dy, dx * dx)), lsum);
        }
        const double pot = fma(0.5, lsum, pot_reg);                     // 0.5*lsum = sum q ln(r2)/(4 pi) since w = q/(2 pi)
        if (!(pot > 0.0)) return PATH_AQUIFER_DRY;                      // model.py:343-344 (nan also ends the trace)
        double head;
        if (pot < rc.half_kH2) head = sqrt(2.0 * pot / rc.k);           // model.py:345-346
        else head = (pot + rc.half_kH2) / (rc.k * rc.H);                // model.py:347-349
        if (!(head > 0.0)) return PATH_AQUIFER_DRY;                     // model.py:380-381
        const double sat = (head > rc.H) ? rc.H : head;                 // model.py:382-387
        const double inv = 1.0 / (sat * rc.n);
        fx = gx * inv;
        fy = gy * inv;
        return PATH_OK;
    }
}

// ------------------------------------------------------------------------------------------
// Far-field compression of the well sum (confined path): tiled local expansions.
//
// The reference adds one term per well at every velocity evaluation (model.py:307-313), 15 flops x Nw.  In complex
// form the wells' part of the backtracking velocity is  G(z) = gx - i gy = sum_w w_w / (z - z_w),  z = (x - xo) + i (y - yo).
// For a point z inside a square tile (centre z_c, half-diagonal h) and a well FAR from the tile, |z_w - z_c| >= h/eta,
//     w/(z - z_w) = - sum_k  w (z - z_c)^k / (z_w - z_c)^(k+1),      |term k| <= |w|/|z_w - z_c| eta^k,
// so ALL far wells of a tile collapse into ONE polynomial in zeta = (z - z_c)/h with coefficients
//     c_k = sum_w w_w P[tile][w][k],    P[tile][w][k] = -(h/(z_w - z_c))^k / (z_w - z_c)   (geometry only, built once on the host),
// and only the few NEAR wells of the tile are summed directly.  Truncated after `order` terms the relative error of a far
// term is <= eta^order/(1 - eta) (3e-15 for eta = 0.3, order = 28), below the Newton reciprocal of the direct sum.
// Per evaluation: ~4 order FP64 instructions of complex Horner + 8 per near well, instead of 8 Nw
// (200-well field: ~6 near wells + 28 terms ~ 20 well-equivalents instead of 200).
// The c_k depend on the realization (discharges, H, n): farfield_coef_kernel forms them per launch, each CTA copies its
// realization's table (ntiles x order complex) into shared memory next to the well store.  A particle outside the
// tile grid (or a field with too few wells to gain) takes the direct sum, so the result never depends on the grid
// beyond rounding (~1e-15 relative to sum |terms|).
struct FarFieldDev {
    int ntx, nty;
    int order;                       // terms of the polynomial (even, >= 2)
    int max_near;                    // even (lists are padded with the dummy well)
    double gx0, gy0;                 // lower-left corner of the tile grid, relative to (xo, yo)
    double inv_tile;                 // 1 / tile side
    const double2 *coef;             // [R of this launch][ntx*nty][order]  (x = re, y = im)
    const unsigned int *near_off;    // [ntx*nty][max_near]  BYTE offsets of the near wells in the confined well store
    const unsigned short *near_cnt;  // [ntx*nty]            padded (even) list lengths
    // unconfined flow (appended; null otherwise):
    const double *b0;                // [R of this launch][ntx*nty]  sum over the far wells of w ln |z_w - z_c|
    const unsigned short *near_idx;  // [ntx*nty][max_near]  near WELL INDICES (unpadded)
    const unsigned short *near_raw;  // [ntx*nty]            their counts
};
struct FarFieldShared {
    const double2 *c64; const unsigned int *off; const unsigned short *cnt;
    // unconfined flow: the potential's polynomial p_k = h c_(k-1)/k (k = 1..order) as float2, b0 per tile, near well indices
    const float2 *p32; const double *b0; const unsigned short *idx; const unsigned short *raw;
};
// shared-memory layout of a tracking CTA:
//   [well store + 256 B of slack][c64 ntiles x order double2][near_off][near_cnt];
// the dummy well {b = 1e100, c = (1, 1)} that pads odd near lists sits in the slack right behind the confined store
__host__ __device__ __forceinline__ constexpr int ff_dummy_offset(int nw) { return ((nw + 3) >> 2) * 12; }                    // doubles
__host__ __device__ __forceinline__ constexpr int ff_store_double2(int nw) { return (((nw + 3) >> 2) * 14 * 8 + 256) / 16; }  // double2s

constexpr double FF_SQRT2 = 1.4142135623730951;

// What was built, timed on B200 and dropped (C3 perham / C4 200 wells, ms per step; profiles/r01_farfield_ab*.txt, r02_knob_scan*.txt):
// high-order terms in FP32 (87.1 / 36.0 against 79.7 / 33.0: conversions, a second loop, more live values), the tile index by the
// 1.5 * 2^52 trick instead of F2I / I2F (84.8 / 34.9), the next trip's coefficients loaded ahead (86.9 / 35.5, and again slower
// with 128 registers), the evaluation as a __noinline__ call (105.6 / 42.9), the coefficient tables read in place through L1
// (93.7 / 37.9), the six Runge-Kutta stages as one loop with the stage derivatives in local memory (85.0 / 36.0).
// What won (round 2): MORE TILES and a LOWER ORDER at the same truncation -- 256-thread CTAs, two per SM, leave 104 KB of shared
// memory for a realization's table (380 tiles of order 16, eta 0.15) -- and the Horner loop unrolled at that order:
// 79.9 / 33.1 -> 64.6 / 25.6.

// sum_{k < n} c_k zeta^k: two interleaved Horner chains in w = zeta^2 (even / odd powers, half the dependency depth).
// ORD > 0: the order is the compile-time constant ORD and the loop is unrolled (no loop control, no address arithmetic: 5 of
// the 25 instructions of a trip); ORD == 0: n terms at run time.  n even, >= 2.
template <int ORD>
__host__ __device__ __forceinline__ void ff_poly_eval(const double2 *__restrict__ c, int n, double zr, double zi, double &re, double &im)
{
    if (ORD > 0) n = ORD;
    const double wr = fma(zr, zr, -(zi * zi));
    const double wi = 2.0 * (zr * zi);
    const double2 ct = c[n - 2], cu = c[n - 1];
    double er = ct.x, ei = ct.y;
    double orr = cu.x, oi = cu.y;
    auto trip = [&](int k) {
        const double2 ce = c[k], co = c[k + 1];
        const double ner = fma(er, wr, fma(-ei, wi, ce.x));
        const double nei = fma(er, wi, fma(ei, wr, ce.y));
        const double nor = fma(orr, wr, fma(-oi, wi, co.x));
        const double noi = fma(orr, wi, fma(oi, wr, co.y));
        er = ner; ei = nei; orr = nor; oi = noi;
    };
    if (ORD > 0) {
#pragma unroll
        for (int k = ORD - 4; k >= 0; k -= 2) trip(k);
    } else {
#pragma unroll 2
        for (int k = n - 4; k >= 0; k -= 2) trip(k);
    }
    re = fma(orr, zr, fma(-oi, zi, er));
    im = fma(orr, zi, fma(oi, zr, ei));
}

// tile of the point (dx0, dy0) [relative to (xo, yo)] and its scaled offset from the tile centre; false = outside the grid
// (or nan).  On an exact tile boundary either neighbour would do: both expansions hold on the closed tile.
// The tile index is rounded through the 1.5 * 2^52 constant: the low word of (u - 0.5) + MAGIC IS round-to-nearest(u - 0.5) =
// floor(u) (either neighbour on a boundary), its high word equals MAGIC's exactly when that integer lies in [0, 2^32), and
// t - MAGIC is the same integer as a double -- two additions and two integer compares per axis instead of F2I.F64, I2F.F64
// and four DSETP (7 % of the far-field kernel's instructions went into this lookup; C3 54.6 -> 53.9 ms, profiles/r02_ab_locate.txt).
__host__ __device__ __forceinline__ void ff_words(double v, unsigned int &lo, unsigned int &hi)
{
#ifdef __CUDA_ARCH__
    lo = (unsigned int)__double2loint(v); hi = (unsigned int)__double2hiint(v);
#else
    unsigned long long b; memcpy(&b, &v, 8); lo = (unsigned int)b; hi = (unsigned int)(b >> 32);
#endif
}
__host__ __device__ __forceinline__ bool ff_locate(int ntx, int nty, double gx0, double gy0, double inv_tile,
                                                   double dx0, double dy0, int &tile, double &zr, double &zi)
{
    constexpr double MAGIC = 6755399441055744.0;                                               // 1.5 * 2^52, high word 0x43380000
    const double vx = fma(dx0 - gx0, inv_tile, -0.5), vy = fma(dy0 - gy0, inv_tile, -0.5);     // u - 0.5
    const double tx = vx + MAGIC, ty = vy + MAGIC;
    unsigned int ci, cj, hx, hy;
    ff_words(tx, ci, hx);
    ff_words(ty, cj, hy);
    if (!((hx == 0x43380000u) & (hy == 0x43380000u) & (ci < (unsigned int)ntx) & (cj < (unsigned int)nty))) return false;   // also nan, inf
    tile = (int)(cj * (unsigned int)ntx + ci);
    zr = (vx - (tx - MAGIC)) * FF_SQRT2;                                                       // (x - x_c)/h,  h = tile/sqrt 2
    zi = (vy - (ty - MAGIC)) * FF_SQRT2;
    return true;
}

#if defined(__CUDACC__) || defined(ONEKA_EMU)
// the direct sum as an out-of-line call: the rare particle outside the tile grid.  Results come back BY VALUE: reference
// parameters of a real call are addresses, which put the six stage derivatives of the caller on the stack (2 STL.64 after and
// 2 LDL.64 before every evaluation of the hot path, 34 local-memory instructions in the kernel).
struct FieldValue { double fx, fy; int status; };
__device__ __noinline__ FieldValue field_direct_cold(const RealConsts &rc, const double *s_wells, int nw, double x, double y)
{
    FieldValue v;
    v.status = field_feval<true>(rc, s_wells, nw, x, y, v.fx, v.fy);
    return v;
}

template <int ORD>
__device__ __forceinline__ int field_feval_ff(const RealConsts &rc, const double *__restrict__ s_wells, int nw,
                                              const FarFieldDev &ff, const FarFieldShared &fs,
                                              double x, double y, double &fx, double &fy)
{
    const double dx0 = x - rc.xo;
    const double dy0 = y - rc.yo;
    int tile;
    double zr, zi;
    if (!ff_locate(ff.ntx, ff.nty, ff.gx0, ff.gy0, ff.inv_tile, dx0, dy0, tile, zr, zi)) {
        const FieldValue v = field_direct_cold(rc, s_wells, nw, x, y);
        fx = v.fx; fy = v.fy;
        return PATH_OK;
    }
    double gx = fma(rc.a2, dx0, fma(rc.c, dy0, rc.d));
    double gy = fma(rc.b2, dy0, fma(rc.c, dx0, rc.e));
    // near wells, two per trip (lists are padded to even length with the dummy well, whose term is ~1e-100)
    const unsigned int *po = fs.off + tile * ff.max_near;
    const char *sw = reinterpret_cast<const char *>(s_wells);
    double hx = 0.0, hy = 0.0;                                   // second accumulator pair: two independent chains
    const int n = fs.cnt[tile];
#pragma unroll 1
    for (int i = 0; i < n; i += 2) {
        const uint2 o2 = *reinterpret_cast<const uint2 *>(po + i);
        const double *p0 = reinterpret_cast<const double *>(sw + o2.x), *p1 = reinterpret_cast<const double *>(sw + o2.y);
        scaled_term(dx0, dy0, p0[0], p0[1], p0[2], gx, gy);
        scaled_term(dx0, dy0, p1[0], p1[1], p1[2], hx, hy);
    }
    // (the odd well of a list on its own instead of the dummy padding: measured 14 % SLOWER, profiles/r02_knob_scan6.txt --
    //  lanes of one warp in tiles of different parity then run the pair loop and the single-well tail one after the other)
    // far wells: the tile's polynomial
    double re, im;
    ff_poly_eval<ORD>(fs.c64 + tile * (ORD > 0 ? ORD : ff.order), ff.order, zr, zi, re, im);
    fx = (gx + hx) + re;
    fy = (gy + hy) - im;
    return PATH_OK;
}

// ---- unconfined flow through the far field (oneka_set_farfield_unconfined) -------------------------------------------------
// field_feval<false> needs, besides the discharge, the POTENTIAL -- but only to decide whether the aquifer is fully saturated at
// the point (then V = Q/(H n) whatever Phi is), which an FP32-accurate value settles almost everywhere.  The far wells' part of it,
//     sum_far w ln|z - z_w| = b0 + Re sum_{k>=1} p_k zeta^k,     b0 = sum_far w ln|z_w - z_c|,   p_k = h c_(k-1) / k,
// comes from the SAME coefficients as the discharge (d/dz of the complex potential), so it costs one more Horner, in FP32.  Where
// the screening value is within pot_err of k H^2/2 (or the point lies outside the tile grid) the whole evaluation is redone by the
// direct-sum function, FP64 logs and all -- exactly what field_feval<false> does there.
__device__ __noinline__ FieldValue field_direct_unc_cold(const RealConsts &rc, const double *s_wells, int nw, double x, double y)
{
    FieldValue v;
    v.status = field_feval<false>(rc, s_wells, nw, x, y, v.fx, v.fy);
    return v;
}

template <int ORD>
__device__ __forceinline__ int field_feval_ff_unc(const RealConsts &rc, const double *__restrict__ s_wells, int nw,
                                                  const FarFieldDev &ff, const FarFieldShared &fs,
                                                  double x, double y, double &fx, double &fy)
{
    const double dx0 = x - rc.xo;
    const double dy0 = y - rc.yo;
    int tile;
    double zr, zi;
    if (!ff_locate(ff.ntx, ff.nty, ff.gx0, ff.gy0, ff.inv_tile, dx0, dy0, tile, zr, zi)) {
        const FieldValue v = field_direct_unc_cold(rc, s_wells, nw, x, y);
        fx = v.fx; fy = v.fy;
        return v.status;
    }
    double gx = fma(rc.a2, dx0, fma(rc.c, dy0, rc.d));
    double gy = fma(rc.b2, dy0, fma(rc.c, dx0, rc.e));
    // near wells: discharge term + FP32 log term each (well_term, the unconfined store)
    float lsum32 = 0.0f;
    const unsigned short *pi = fs.idx + tile * ff.max_near;
    const int n = fs.raw[tile];
#pragma unroll 1
    for (int i = 0; i < n; ++i) {
        const int w = pi[i];
        well_term(x, y, well_x(s_wells, w), well_y(s_wells, w), well_w(s_wells, w),
                  reinterpret_cast<const float *>(s_wells + (w >> 2) * WELL_BLK + 12)[w & 3], gx, gy, lsum32);
    }
    // far wells: discharge polynomial in FP64 ...
    const int order = ORD > 0 ? ORD : ff.order;
    double re, im;
    ff_poly_eval<ORD>(fs.c64 + tile * order, order, zr, zi, re, im);
    gx += re;
    gy -= im;
    // ... and their potential in FP32:  Re(zeta T),  T = sum_j p_(j+1) zeta^j
    const float2 *pp = fs.p32 + tile * order;
    const float fr = (float)zr, fi = (float)zi;
    float ar = 0.0f, ai = 0.0f;
    auto step32 = [&](int j) {
        const float2 c = pp[j];
        const float nr = fmaf(ar, fr, fmaf(-ai, fi, c.x));
        const float ni = fmaf(ar, fi, fmaf(ai, fr, c.y));
        ar = nr; ai = ni;
    };
    if constexpr (ORD > 0) {
#pragma unroll
        for (int j = ORD - 1; j >= 0; --j) step32(j);
    } else {
#pragma unroll 2
        for (int j = order - 1; j >= 0; --j) step32(j);
    }
    const double pot_far = fs.b0[tile] + (double)fmaf(fr, ar, -(fi * ai));
    const double pot_reg = rc.A * dx0 * dx0 + rc.B * dy0 * dy0 + rc.c * dx0 * dy0 + rc.d * dx0 + rc.e * dy0 + rc.F;
    const double pot_apx = fma(0.34657359027997264, (double)lsum32, pot_reg) + pot_far;
    if (pot_apx - rc.pot_err > rc.half_kH2) {                    // certainly saturated: model.py:382-384
        fx = gx * rc.inv_Hn;
        fy = gy * rc.inv_Hn;
        return PATH_OK;
    }
    const FieldValue v = field_direct_unc_cold(rc, s_wells, nw, x, y);
    fx = v.fx; fy = v.fy;
    return v.status;
}
#endif

// ------------------------------------------------------------------------------------------
// Rasteriser.
//
// insert() (probabilityfield.py:296-310) marks every lattice node of the clipped window whose
// distancesquared() (:407-427) to the segment is < umbra^2.  Here, per accepted step:
//   1. the window is computed with the reference's own IEEE operations (sub, sub, div, floor; the quotient
//      goes through a reciprocal and is re-divided exactly only next to an integer, floor_div);
//   2. SCAN-LINE ROWS: on each row the capsule is an interval whose ends are closed-form (cap a, straight
//      edge, or cap b); every node strictly between the ends is marked with two shifts.  FP32, relative to
//      endpoint a.  Nodes within the FP32 error bound of an end (~1e-4 of the rows) go to 3;
//   3. NODE TEST (also the whole row for near-tangent rows, near-horizontal or sub-1e-10 m segments):
//      t = sat(dot/len^2), p = c - t b, d2 = |p|^2 in FP32.  |d2_fp32 - d2_exact| <= 44 eps32 L^2 with
//      L = max(|bax|,|bay|) + umbra + max(dx,dy) >= every |cax|,|cay| of the window (derivation in DESIGN.md);
//      the reference's own FP64 value is within 32 eps64 (..)^2 of exact, 2^-29 times smaller.  Band
//      E = 128 eps32 L^2:   d2 < umbra^2 - E -> inside,   d2 > umbra^2 + E -> outside   (the reference's answers),
//      otherwise (incl. nan) -> the reference's formula in unfused FP64 (exact_distancesquared);
//   4. the row's bits are OR-ed into the realization's bitmap with one or two RED.OR (L2, no return value).
// Zero-length segments set nothing (0/0 = nan in :421).  Bit-exact against the executed reference.

__device__ __forceinline__ double exact_distancesquared(double ax, double ay, double bx, double by, double cx, double cy)
{
    // probabilityfield.py:407-427, operation for operation, never contracted
    const double bax = __dsub_rn(bx, ax);
    const double bay = __dsub_rn(by, ay);
    const double cax = __dsub_rn(cx, ax);
    const double cay = __dsub_rn(cy, ay);
    const double perpdot = __dsub_rn(__dmul_rn(bax, cay), __dmul_rn(bay, cax));
    const double dot = __dadd_rn(__dmul_rn(bax, cax), __dmul_rn(bay, cay));
    const double length2 = __dadd_rn(__dmul_rn(bax, bax), __dmul_rn(bay, bay));
    const double alpha2 = __ddiv_rn(__dmul_rn(perpdot, perpdot), length2);
    const double beta2 = __ddiv_rn(__dmul_rn(dot, dot), length2);
    double d2;
    if (dot < 0)
        d2 = __dadd_rn(alpha2, beta2);
    else if (beta2 > length2)
        d2 = __dadd_rn(__dsub_rn(__dadd_rn(alpha2, beta2), __dmul_rn(2.0, dot)), length2);
    else
        d2 = alpha2;
    return d2;
}

struct RasterCounters { unsigned int clipped, exact; };

// insert() clips its window to the grid "as it is now" (probabilityfield.py:298-301).  On the fixed lattice that is
// [0, ncols) x [0, nrows); the exact emulation of the auto-expanding field passes the window the reference's grid
// had when this path was inserted (oneka_capture_clipped).
struct ClipWin { int l, r, b, t; };

// MUFU.SQRT (nan for negative arguments, as the callers expect)
__device__ __forceinline__ float sqrt_fast(float a)
{
    return ptx_sqrt_approx_f32(a);
}

// Cold path.  `lat` = {xmin, ymin, dx, dy, umbra^2} in SHARED memory: were these taken from the kernel parameters,
// ptxas hoists the ten constant-bank loads of the argument set-up out of the rare branch and into every row
// (measured: 12 of ~60 instructions per row).
__device__ __noinline__ bool exact_cell_test(const double *lat, double ax, double ay, double bx, double by, int i, int j)
{
    const double cx = __dadd_rn(lat[0], __dmul_rn((double)j, lat[2]));   // probabilityfield.py:306
    const double cy = __dadd_rn(lat[1], __dmul_rn((double)i, lat[3]));   // probabilityfield.py:307
    return exact_distancesquared(ax, ay, bx, by, cx, cy) < lat[4];       // :309 (nan -> false)
}

__device__ __forceinline__ void stage_lattice(const LatticeDev &L, double *s_lat)
{
    if (threadIdx.x == 0) { s_lat[0] = L.xmin; s_lat[1] = L.ymin; s_lat[2] = L.dx; s_lat[3] = L.dy; s_lat[4] = L.umbra2; }
}

// floor((v - org) / delta) with the reference's two IEEE operations (sub, div).  The quotient is
// first formed with the precomputed reciprocal; only when it lands within 1e-9 (relative) of an
// integer -- where a one-ulp difference could change the floor -- is the true division issued.
__device__ __forceinline__ double floor_div(double v, double org, double delta, double inv_delta)
{
    const double num = __dsub_rn(v, org);
    const double q = num * inv_delta;
    const double f = floor(q);
    const double frac = q - f;
    const double tol = 1e-9 * fmax(1.0, fabs(q));
    if (frac < tol || frac > 1.0 - tol) return floor(__ddiv_rn(num, delta));
    return f;
}

// HEAVY: the flavour for lattices where a segment meets many rows (umbra >> spacing; C5: 11-15 rows), where the kernel is
// co-limited by the bit-set traffic to L2 (one 32-byte sector per lane per row: ncu at C5, L2 tag throughput 64 %).  It issues
// fewer bit-sets, for a few more instructions per row (which is why windows of 5-7 rows, C3 / C4, keep the plain flavour:
// measured, profiles/r02_notes.md section 10):
//   * 64-bit RED.OR on the aligned word pair: a span straddles a pair half as often as a word (1.12 instead of 1.25 per row);
//   * `chained`: the path's previous chronicled segment ended at (ax, ay) and went through this function with the same clip
//     window.  Its cap b then set every node of the disc around a already, and the rows of THIS segment that lie wholly behind
//     a (both interval ends on cap a: the row's nodes are exactly that disc's chord) have nothing to add: their bit-set is
//     skipped.  Only rows on the fast path qualify -- both chord ends farther from a node than the FP32 bound, so the two
//     segments' decisions for every node of the chord are the same exact decision.
// Returns false when the segment had zero length (nothing marked, the chain is not established by it); true otherwise.
//
// (A third flavour kept each thread's rows in a shared-memory tile that followed the particle front -- OR-ed there, written to
// the bitmap once when the window had moved on: 3-4 bit-sets per segment instead of ~14.  Bit-exact, and 9-13 % SLOWER than
// RF_HEAVY on every lattice measured, profiles/r02_flavour_scan2.txt: once the bit-sets are down by a quarter the loop is bound
// by its instructions, and the tile adds ~8 per row.  Commit e42da4e has it.)
constexpr int RF_PLAIN = 0, RF_HEAVY = 1;

// bit-set of one 32-bit bitmap word on the rare rows (general_row).  The heavy flavour's common rows use 64-bit atomics on the
// aligned word pairs, so its rare rows do too: every atomic a kernel issues on the bitmap then has one size (mixed-size
// atomics on overlapping bytes are outside the PTX memory model's guarantees, whatever the L2 does with them today).
template <bool WIDE>
__device__ __forceinline__ void or_word(unsigned int *p, unsigned int mask)
{
    if constexpr (WIDE) {
        const uintptr_t a = reinterpret_cast<uintptr_t>(p);
        atomicOr(reinterpret_cast<unsigned long long *>(a & ~(uintptr_t)7), (unsigned long long)mask << ((a & 4) ? 32 : 0));
    } else {
        atomicOr(p, mask);
    }
}

template <int RF = RF_PLAIN>
__device__ __forceinline__ bool raster_seg(const LatticeDev &L, const double *s_lat, unsigned int *__restrict__ bm,
                                           const ClipWin cw, double ax, double ay, double bx, double by, RasterCounters &ctr,
                                           bool chained = false)
{
    constexpr bool HEAVY = RF != RF_PLAIN;
    // ---- window, probabilityfield.py:298-301 ----
    // left = floor((min(ax,bx) - umbra - xmin)/dx) etc.  Fast path: the four quotients in 16.16 fixed point from one
    // subtraction and one multiplication each (a few ulps away from the reference's sub, sub, div); a quotient within
    // 3 * 2^-16 of an integer -- where those ulps could change the floor -- sends the segment to the exact path
    // (floor_div: the reference's own operations).
    const double mnx = (bx < ax) ? bx : ax, mxx = (bx > ax) ? bx : ax;
    const double mny = (by < ay) ? by : ay, mxy = (by > ay) ? by : ay;
    int fl, fr, fb, ft;
    bool fast = L.fixed_ok != 0;
    if (fast) {
        const int ql = __double2int_rd((mnx - L.cxl) * L.s16x), qr = __double2int_rd((mxx - L.cxr) * L.s16x);
        const int qb = __double2int_rd((mny - L.cyb) * L.s16y), qt = __double2int_rd((mxy - L.cyt) * L.s16y);
        constexpr unsigned int T = 3u;                                   // saturated conversions stay far outside the lattice
        const bool amb = (((unsigned int)ql + T) & 0xffffu) <= 2u * T || (((unsigned int)qr + T) & 0xffffu) <= 2u * T ||
                         (((unsigned int)qb + T) & 0xffffu) <= 2u * T || (((unsigned int)qt + T) & 0xffffu) <= 2u * T;
        fl = ql >> 16; fr = qr >> 16; fb = qb >> 16; ft = qt >> 16;
        fast = !amb;
    }
    if (!fast) {
        constexpr double BIG = 1073741824.0;                             // 2^30: keeps fr + 1 inside an int
        fl = (int)fmax(-BIG, fmin(BIG, floor_div(__dsub_rn(mnx, L.umbra), L.xmin, L.dx, L.inv_dx)));
        fr = (int)fmax(-BIG, fmin(BIG, floor_div(__dadd_rn(mxx, L.umbra), L.xmin, L.dx, L.inv_dx)));
        fb = (int)fmax(-BIG, fmin(BIG, floor_div(__dsub_rn(mny, L.umbra), L.ymin, L.dy, L.inv_dy)));
        ft = (int)fmax(-BIG, fmin(BIG, floor_div(__dadd_rn(mxy, L.umbra), L.ymin, L.dy, L.inv_dy)));
    }
    const int left = max(fl, cw.l), right = min(fr + 1, cw.r), bottom = max(fb, cw.b), top = min(ft + 1, cw.t);
    if (fl < cw.l || fb < cw.b || fr + 1 > cw.r || ft + 1 > cw.t) ctr.clipped++;
    if (left >= right || bottom >= top) return true;                     // empty window (then the disc around b is outside the clip window too)

    const double bax = __dsub_rn(bx, ax), bay = __dsub_rn(by, ay);
    const double len2 = __dadd_rn(__dmul_rn(bax, bax), __dmul_rn(bay, bay));
    if (!(len2 > 0.0)) return false;                                     // 0/0 -> nan -> no node is marked
    const bool all_exact = len2 < 1e-20;

    // ---- FP32 set-up, relative to endpoint a ----
    constexpr float EPS32 = 1.1920928955078125e-07f;
    const float fbax = (float)bax, fbay = (float)bay;
    const float flen2 = fmaf(fbay, fbay, fbax * fbax);
    const float finv = __fdividef(1.0f, flen2);
    const float base_x = (float)(fma((double)left, L.dx, L.xmin) - ax);
    const float base_y = (float)(fma((double)bottom, L.dy, L.ymin) - ay);
    const float Lm = (float)(fmax(fabs(bax), fabs(bay)) + L.umbra + L.maxd);
    const float E = all_exact ? INFINITY : 128.0f * EPS32 * Lm * Lm;
    const float u2 = L.umbra2_32, u = L.umbra32;

    // FP32 distance^2 of window node (fk, row with offset cay) and its classification
    auto node_inside = [&](float fk, float cay, int i, int j) -> bool {
        const float cax = fmaf(fk, L.dx32, base_x);
        const float t = __saturatef(fmaf(fbax, cax, fbay * cay) * finv);
        const float px = fmaf(-t, fbax, cax);
        const float py = fmaf(-t, fbay, cay);
        const float d2 = fmaf(px, px, py * py);
        if (fabsf(d2 - u2) > E) return d2 < u2;
        ctr.exact++;
        return exact_cell_test(s_lat, ax, ay, bx, by, i, j);
    };

    // ---- scan-line constants (see "Scan-line rows" in DESIGN.md) ----
    // On the row at height v (relative to a) the capsule is the open interval (xl, xr):
    //   right end: t_r = (v + k)/sy;  t_r < 0 -> sqrt(u^2 - v^2) (cap a);  t_r > 1 -> sx + sqrt(u^2 - (v-sy)^2) (cap b);
    //              else m v + c (the straight edge),   k = sign(sy) u sx/len,  m = sx/sy,  c = k m + u |sy|/len
    //   left end : the mirror image with -k, -c.
    // A node is decided by the interval alone unless it lies within the FP32 error bound of an end:
    //   |x_end error| <= eps L (17 |m| + 11)   on an edge,      <= 10 eps L^2 / h   on a cap of half-chord h
    // (derivation in DESIGN.md; both are used with a 4x margin).  The at most two nodes inside those bounds are
    // classified like any node (FP32 distance + band, exact FP64 inside the band).  Rows within tau of tangency
    // and segments flatter than 1:64 use the node loop.
    const float rlen = rsqrtf(flen2);
    const float uy = u * fabsf(fbay) * rlen;
    const float kk = (fbay < 0.0f) ? -(u * fbax * rlen) : (u * fbax * rlen);
    const float isy = __fdividef(1.0f, fbay);
    const float m = fbax * isy;
    const float cr = fmaf(kk, m, uy);
    const float kisy = kk * isy;
    const float ylo = fminf(0.0f, fbay), yhi = fmaxf(0.0f, fbay);
    const float tau = fmaxf(1e-3f * u, 256.0f * EPS32 * Lm);
    const float k_edge = 4.0f * EPS32 * Lm * fmaf(17.0f, fabsf(m), 11.0f) * L.inv_dx32;     // [cells]
    const float k_cap = 40.0f * EPS32 * Lm * Lm * L.inv_dx32;                                // [cells * m]
    const int ncol = right - left;
    // A straight-edge end is trusted only while its error bound stays well below one cell (k_edge grows with the slope
    // |m| = |sx/sy|); on near-horizontal segments almost every row ends on the two caps, whose bounds do not involve m,
    // so only the rows that do cross a straight edge fall back to the node loop (edge_ok false; also for sy = 0, where
    // m and k_edge are inf or nan).
    const bool edge_ok = k_edge < 0.125f;
    // (scan-line rows for every window width: measured 8 % faster than the node loop at 7-column windows (C3), 33 % at 15 (C5))
    const bool scan_ok = !all_exact && u > 0.0f;
    // scan_ok folded into the two row thresholds: tested inside the loop, ptxas re-derived it PER ROW from the FP64 segment
    // vector (5 FP64 instructions + DSETP of the len2 < 1e-20 test, i.e. 12 issue slots of a ~100-slot row)
    const float skip_above = scan_ok ? u + tau : INFINITY;       // a row farther than this from the segment's y-extent cannot touch the capsule
    const float scan_below = scan_ok ? u - tau : -INFINITY;      // a row nearer than this takes the scan-line path

    // THE GENERAL ROW (every case): skip test, scan-line interval with ambiguous ends settled node by node, spans wider than a
    // word, and the node loop for near-tangent rows and rows the interval formulas do not cover.
    auto general_row = [&](int i, float cay, unsigned int *row) {
        const float dmin = fmaxf(fmaxf(ylo - cay, cay - yhi), 0.0f);     // distance from the row to the segment's y-extent
        if (dmin > skip_above) return;                                 // the row cannot touch the capsule
        bool done = false;
        if (dmin < scan_below) {
            const float tr = fmaf(cay, isy, kisy), tl = fmaf(cay, isy, -kisy);
            const float ha = sqrt_fast(fmaf(-cay, cay, u2));
            const float wb = cay - fbay;
            const float hb = sqrt_fast(fmaf(-wb, wb, u2));
            const bool ra = tr < 0.0f, rb = tr > 1.0f, la = tl < 0.0f, lb = tl > 1.0f;
            // (selects spelled out one by one: the nested conditional expressions became branches with BSSY / BSYNC pairs)
            float xr = fmaf(m, cay, cr), xl = fmaf(m, cay, -cr);         // the straight edges ...
            xr = rb ? fbax + hb : xr;                                    // ... cap b ...
            xl = lb ? fbax - hb : xl;
            xr = ra ? ha : xr;                                           // ... cap a
            xl = la ? -ha : xl;
            const bool cap_l = la | lb, cap_r = ra | rb;
            if (xr > xl && (edge_ok | (cap_l & cap_r))) {                // also false for nan
                const float fl = (xl - base_x) * L.inv_dx32, fr = (xr - base_x) * L.inv_dx32;
                // round-to-nearest and float->int through the 1.5 * 2^23 trick (FMA-pipe adds instead of the
                // quarter-rate FRND / F2I of the XU pipe); exact for |f| < 2^22, the window is far smaller
                constexpr float MAGIC = 12582912.0f;
                const float tl_m = fl + MAGIC, tr_m = fr + MAGIC;
                const float rl = tl_m - MAGIC, rr = tr_m - MAGIC;
                const int il = __float_as_int(tl_m) - 0x4B400000, ir = __float_as_int(tr_m) - 0x4B400000;
                const float dl = fl - rl, dr = fr - rr;
                int kl = il + (dl > 0.0f ? 1 : 0);                       // first node right of xl
                int kr = ir - (dr > 0.0f ? 0 : 1);                       // last node left of xr
                // is the node nearest to an end inside that end's error bound?
                const float hl = la ? ha : hb, hr = ra ? ha : hb;        // half-chord of the cap an end lies on (if it does)
                const bool amb_l = cap_l ? (fabsf(dl) * hl < k_cap) : (fabsf(dl) < k_edge);
                const bool amb_r = cap_r ? (fabsf(dr) * hr < k_cap) : (fabsf(dr) < k_edge);
                if (amb_l | amb_r) {                                     // rare (~1e-4 of rows)
                    if (amb_l) { const int k = il; if (k >= 0 && k < ncol) kl = node_inside(rl, cay, i, left + k) ? k : k + 1; }
                    if (amb_r) { const int k = ir; if (k >= 0 && k < ncol) kr = node_inside(rr, cay, i, left + k) ? k : k - 1; }
                }
                kl = max(kl, 0);
                kr = min(kr, ncol - 1);
                if (kl <= kr) {
                    const int ja = left + kl, span = kr - kl;            // span + 1 nodes starting at column ja
                    unsigned int *wp = row + (ja >> 5);
                    const int sh = ja & 31;
                    if (span < 32) {
                        const unsigned int bits = 0xffffffffu >> (31 - span);
                        or_word<HEAVY>(wp, bits << sh);
                        if (sh + span > 31) or_word<HEAVY>(wp + 1, bits >> (32 - sh));
                    } else {
                        const int jb = left + kr;
#pragma unroll 1
                        for (int w = ja >> 5; w <= (jb >> 5); ++w) {
                            const int lo = max(ja, w << 5), hi = min(jb, (w << 5) + 31);
                            or_word<HEAVY>(row + w, (0xffffffffu >> (31 - (hi - lo))) << (lo & 31));
                        }
                    }
                }
                done = true;
            }
        }
        if (done) return;
        // ---- node loop: columns in chunks of 32 starting at `left`; bit k of `mask` is column jc + k ----
        const float c1 = fbay * cay;
        for (int jc = left; jc < right; jc += 32) {
            const int n = min(32, right - jc);
            unsigned int mask = 0u, amb = 0u, bit = 1u;
            float fk = (float)(jc - left);
            for (int k = 0; k < n; ++k, fk += 1.0f, bit <<= 1) {
                const float cax = fmaf(fk, L.dx32, base_x);
                const float t = __saturatef(fmaf(fbax, cax, c1) * finv);
                const float px = fmaf(-t, fbax, cax);
                const float py = fmaf(-t, fbay, cay);
                const float d2 = fmaf(px, px, py * py);
                if (d2 < u2) mask |= bit;
                if (!(fabsf(d2 - u2) > E)) amb |= bit;                   // inside the error band (or nan)
            }
            if (amb) {                                                   // rare: settle with the reference's FP64 formula
                ctr.exact += __popc(amb);
                do {
                    const int k = __ffs(amb) - 1;
                    amb &= amb - 1;
                    if (exact_cell_test(s_lat, ax, ay, bx, by, i, jc + k)) mask |= 1u << k;
                    else mask &= ~(1u << k);
                } while (amb);
            }
            if (mask) {
                const int sh = jc & 31;
                unsigned int *wp = row + (jc >> 5);
                or_word<HEAVY>(wp, mask << sh);
                if (sh && (mask >> (32 - sh))) or_word<HEAVY>(wp + 1, mask >> (32 - sh));
            }
        }
    };

    // Row loop.  Almost every row is the plain case -- inside the capsule's y-range, both interval ends unambiguous, at most 32
    // nodes -- and is handled by a branch-light fast path: the same expressions as general_row evaluated unconditionally (all
    // FP32; a nan only makes `fast` false), ONE test, then the one or two bit-set atomics.  Everything else re-runs the row
    // through general_row.  (With the six tests of the general row in sequence the loop spent its time on branch resolution
    // and reconvergence: ncu, profiles/r02_track_kernel_c3_stalls_by_region.txt.)
    unsigned int *row = bm + (size_t)bottom * L.wpr;
    float fi = 0.0f;
    for (int i = bottom; i < top; ++i, row += L.wpr, fi += 1.0f) {
        const float cay = fmaf(fi, L.dy32, base_y);
        const float dmin = fmaxf(fmaxf(ylo - cay, cay - yhi), 0.0f);
        const float tr = fmaf(cay, isy, kisy), tl = fmaf(cay, isy, -kisy);
        const float ha = sqrt_fast(fmaf(-cay, cay, u2));
        const float wb = cay - fbay;
        const float hb = sqrt_fast(fmaf(-wb, wb, u2));
        const bool ra = tr < 0.0f, rb = tr > 1.0f, la = tl < 0.0f, lb = tl > 1.0f;
        float xr = fmaf(m, cay, cr), xl = fmaf(m, cay, -cr);
        xr = rb ? fbax + hb : xr;
        xl = lb ? fbax - hb : xl;
        xr = ra ? ha : xr;
        xl = la ? -ha : xl;
        const bool cap_l = la | lb, cap_r = ra | rb;
        const float fl = (xl - base_x) * L.inv_dx32, fr = (xr - base_x) * L.inv_dx32;
        constexpr float MAGIC = 12582912.0f;
        const float tl_m = fl + MAGIC, tr_m = fr + MAGIC;
        const float rl = tl_m - MAGIC, rr = tr_m - MAGIC;
        const int il = __float_as_int(tl_m) - 0x4B400000, ir = __float_as_int(tr_m) - 0x4B400000;
        const float dl = fl - rl, dr = fr - rr;
        const int kl = max(il + (dl > 0.0f ? 1 : 0), 0);                 // first node right of xl
        const int kr = min(ir - (dr > 0.0f ? 0 : 1), ncol - 1);          // last node left of xr
        const float hl = la ? ha : hb, hr = ra ? ha : hb;
        const bool amb_l = cap_l ? (fabsf(dl) * hl < k_cap) : (fabsf(dl) < k_edge);
        const bool amb_r = cap_r ? (fabsf(dr) * hr < k_cap) : (fabsf(dr) < k_edge);
        const int span = kr - kl;
        const bool fast = (dmin < scan_below) & (xr > xl) & (edge_ok | (cap_l & cap_r)) & !(amb_l | amb_r) & (span < 32);
        if (fast) {
            if constexpr (HEAVY) {
                if ((span >= 0) & !(chained & la & ra)) {
                    const int ja = left + kl;
                    unsigned long long *wp = reinterpret_cast<unsigned long long *>(row) + (ja >> 6);    // rows are 8-byte aligned (wpr is even)
                    const int sh = ja & 63;
                    const unsigned long long bits = (unsigned long long)(0xffffffffu >> (31 - span));
                    atomicOr(wp, bits << sh);
                    if (sh + span > 63) atomicOr(wp + 1, bits >> (64 - sh));
                }
            } else if (span >= 0) {
                const int ja = left + kl;
                unsigned int *wp = row + (ja >> 5);
                const int sh = ja & 31;
                const unsigned int bits = 0xffffffffu >> (31 - span);
                atomicOr(wp, bits << sh);
                if (sh + span > 31) atomicOr(wp + 1, bits >> (32 - sh));
            }
        } else {
            general_row(i, cay, row);
        }
    }
    return true;
}

// ------------------------------------------------------------------------------------------
// sortable encoding of doubles for atomicMin/atomicMax on unsigned 64-bit
__device__ __forceinline__ unsigned long long dkey(double v)
{
    unsigned long long b = (unsigned long long)__double_as_longlong(v);
    return (b >> 63) ? ~b : (b | 0x8000000000000000ULL);
}

// ------------------------------------------------------------------------------------------
#ifdef __CUDACC__
__constant__ double c_dp[26] = {1.0 / 5.0, 3.0 / 40.0, 9.0 / 40.0, 44.0 / 45.0, -56.0 / 15.0, 32.0 / 9.0,
                                19372.0 / 6561.0, -25360.0 / 2187.0, 64448.0 / 6561.0, -212.0 / 729.0,
                                9017.0 / 3168.0, -355.0 / 33.0, 46732.0 / 5247.0, 49.0 / 176.0, -5103.0 / 18656.0,
                                35.0 / 384.0, 500.0 / 1113.0, 125.0 / 192.0, -2187.0 / 6784.0, 11.0 / 84.0,
                                71.0 / 57600.0, -1.0 / 40.0, -71.0 / 16695.0, 71.0 / 1920.0, -17253.0 / 339200.0, 22.0 / 525.0};
#endif
// Dormand-Prince 5(4), capturezone.py:199-247.  One particle per thread.
//   MODE 0: track only          MODE 1: track + rasterise          MODE 2: track + store vertices
//   FF: the far-field evaluation (confined: field_feval_ff<ORD>, ORD = compile-time order or 0; unconfined: field_feval_ff_unc)
//   RF: raster_seg's flavour (RF_PLAIN, RF_HEAVY for windows of many rows)
template <bool CONFINED, int MODE, bool FF = false, int ORD = 0, int RF = RF_PLAIN>
__device__ __forceinline__ void dopri_track(const TrackParams &tp, const LatticeDev &L, const double *s_lat, unsigned int *bm,
                                            const RealConsts &rc, const double *s_wells,
                                            long long r, int p, bool active,
                                            const FarFieldDev &ff = FarFieldDev(), const FarFieldShared &fs = FarFieldShared())
{
    // the velocity: direct sum over the wells, or (FF, confined only) near wells + the tile's far-field polynomial
    auto feval = [&](double px, double py, double &ox, double &oy) -> int {
        if constexpr (FF && CONFINED) return field_feval_ff<ORD>(rc, s_wells, tp.nw, ff, fs, px, py, ox, oy);
        else if constexpr (FF) return field_feval_ff_unc<ORD>(rc, s_wells, tp.nw, ff, fs, px, py, ox, oy);
        else return field_feval<CONFINED>(rc, s_wells, tp.nw, px, py, ox, oy);
    };
    // Dormand-Prince tableau, capturezone.py:202-209
#ifdef __CUDA_ARCH__
    // the tableau from the constant bank (LDCU.64 / .128 into uniform registers) instead of pairs of UMOV immediates: a third of the
    // kernel's UMOVs gone, C3 -0.5 %, C5 -1.8 % (profiles/r02_ab_dpconst.txt)
    const double a20 = c_dp[0];
    const double a30 = c_dp[1], a31 = c_dp[2];
    const double a40 = c_dp[3], a41 = c_dp[4], a42 = c_dp[5];
    const double a50 = c_dp[6], a51 = c_dp[7], a52 = c_dp[8], a53 = c_dp[9];
    const double a60 = c_dp[10], a61 = c_dp[11], a62 = c_dp[12], a63 = c_dp[13], a64 = c_dp[14];
    const double a70 = c_dp[15], a72 = c_dp[16], a73 = c_dp[17], a74 = c_dp[18], a75 = c_dp[19];
    const double e0 = c_dp[20], e1 = c_dp[21], e2 = c_dp[22], e3 = c_dp[23], e4 = c_dp[24], e5 = c_dp[25];
#else
    constexpr double a20 = 1.0 / 5.0;
    constexpr double a30 = 3.0 / 40.0, a31 = 9.0 / 40.0;
    constexpr double a40 = 44.0 / 45.0, a41 = -56.0 / 15.0, a42 = 32.0 / 9.0;
    constexpr double a50 = 19372.0 / 6561.0, a51 = -25360.0 / 2187.0, a52 = 64448.0 / 6561.0, a53 = -212.0 / 729.0;
    constexpr double a60 = 9017.0 / 3168.0, a61 = -355.0 / 33.0, a62 = 46732.0 / 5247.0, a63 = 49.0 / 176.0, a64 = -5103.0 / 18656.0;
    constexpr double a70 = 35.0 / 384.0, a72 = 500.0 / 1113.0, a73 = 125.0 / 192.0, a74 = -2187.0 / 6784.0, a75 = 11.0 / 84.0;
    constexpr double e0 = 71.0 / 57600.0, e1 = -1.0 / 40.0, e2 = -71.0 / 16695.0, e3 = 71.0 / 1920.0, e4 = -17253.0 / 339200.0, e5 = 22.0 / 525.0;
#endif
    constexpr double EPS = DBL_EPSILON;                                   // :200

    const double duration = tp.duration, tol = tp.tol, maxstep = tp.maxstep;
    const double adur = fabs(duration);

    double x = 0.0, y = 0.0;
    double t = 0.0;                                                        // :212
    double dt = 0.1 * ((duration > 0.0) ? 1.0 : ((duration < 0.0) ? -1.0 : 0.0));   // :213
    int status = PATH_OK;
    int nattempt = 0, nvert = 0;
    double bx0 = INFINITY, bx1 = -INFINITY, by0 = INFINITY, by1 = -INFINITY;
    RasterCounters ctr = {0u, 0u};
    double *vout = nullptr;
    ClipWin cw = {0, L.ncols, 0, L.nrows};
    if (MODE == 1 && tp.clip != nullptr && active) {
        const int4 c = *reinterpret_cast<const int4 *>(tp.clip + 4 * ((size_t)r * tp.P + p));
        cw.l = max(c.x, 0); cw.r = min(c.y, L.ncols); cw.b = max(c.z, 0); cw.t = min(c.w, L.nrows);
    }

    double k1x = 0.0, k1y = 0.0;
    bool running = active;
    bool chained = false;                                                  // raster_seg: an earlier segment of this path ended where the next one starts
    if (active) {
        x = tp.start_xy[2 * p];
        y = tp.start_xy[2 * p + 1];
        nvert = 1;                                                         // :215 the start point is the first vertex
        bx0 = bx1 = x; by0 = by1 = y;
        if (MODE == 2) {
            vout = tp.verts + ((size_t)r * tp.P + p) * (size_t)tp.max_verts * 2;
            if (tp.max_verts > 0) { vout[0] = x; vout[1] = y; }
        }
        status = feval(x, y, k1x, k1y);   // :219
        if (status != PATH_OK) running = false;
    }

    // The loop is WARP-UNIFORM: every lane takes part in every iteration until no lane of the warp
    // has work left (finished lanes idle, as they would under SIMT anyway).  That lets the divergent
    // rasteriser be followed by __syncwarp(), so the FP64 tracking code of the next attempt always
    // runs with the whole warp converged (without it the compiler leaves the lanes split after the
    // raster loops and the well loop issues at half width; measured, profiles/r01_*).
    for (;;) {
        const bool go = running && fabs(t) < adur;                         // :221
        if (!__any_sync(0xffffffffu, go)) break;
        bool seg = false;
        double sax = 0.0, say = 0.0;
        if (go) {
            do {
                if (nattempt >= tp.max_attempts) { status = PATH_MAX_ATTEMPT; running = false; break; }
                ++nattempt;
                if (fabs(t + dt) > adur) dt = duration - t;                // :223-224

                double k2x, k2y, k3x, k3y, k4x, k4y, k5x, k5y, k6x, k6y, k7x, k7y;
                int st;
                st = feval(fma(dt, a20 * k1x, x), fma(dt, a20 * k1y, y), k2x, k2y);      // :227
                if (!CONFINED && st) { status = st; running = false; break; }
                st = feval(fma(dt, fma(a31, k2x, a30 * k1x), x),
                                           fma(dt, fma(a31, k2y, a30 * k1y), y), k3x, k3y);                                  // :228
                if (!CONFINED && st) { status = st; running = false; break; }
                st = feval(fma(dt, fma(a42, k3x, fma(a41, k2x, a40 * k1x)), x),
                                           fma(dt, fma(a42, k3y, fma(a41, k2y, a40 * k1y)), y), k4x, k4y);                   // :229
                if (!CONFINED && st) { status = st; running = false; break; }
                st = feval(
                                           fma(dt, fma(a53, k4x, fma(a52, k3x, fma(a51, k2x, a50 * k1x))), x),
                                           fma(dt, fma(a53, k4y, fma(a52, k3y, fma(a51, k2y, a50 * k1y))), y), k5x, k5y);    // :230
                if (!CONFINED && st) { status = st; running = false; break; }
                st = feval(
                                           fma(dt, fma(a64, k5x, fma(a63, k4x, fma(a62, k3x, fma(a61, k2x, a60 * k1x)))), x),
                                           fma(dt, fma(a64, k5y, fma(a63, k4y, fma(a62, k3y, fma(a61, k2y, a60 * k1y)))), y), k6x, k6y);  // :231
                if (!CONFINED && st) { status = st; running = false; break; }

                const double xt = fma(dt, fma(a75, k6x, fma(a74, k5x, fma(a73, k4x, fma(a72, k3x, a70 * k1x)))), x);         // :233
                const double yt = fma(dt, fma(a75, k6y, fma(a74, k5y, fma(a73, k4y, fma(a72, k3y, a70 * k1y)))), y);
                st = feval(xt, yt, k7x, k7y);                                            // :236
                if (!CONFINED && st) { status = st; running = false; break; }

                const double ex = dt * fma(e5, k6x, fma(e4, k5x, fma(e3, k4x, fma(e2, k3x, fma(e1, k7x, e0 * k1x)))));       // :237-238
                const double ey = dt * fma(e5, k6y, fma(e4, k5y, fma(e3, k4y, fma(e2, k3y, fma(e1, k7y, e0 * k1y)))));
                const double est = fmax(fabs(ex), fabs(ey));
                const double ddx = xt - x, ddy = yt - y;
                const double ds = sqrt(fma(ddy, ddy, ddx * ddx));                                                            // :239

                if (!(isfinite(est) && isfinite(ds))) { status = PATH_NONFINITE; running = false; break; }   // the reference would loop forever on nan

                if (est < tol && ds < maxstep) {                           // :241-245
                    t = t + dt;
                    k1x = k7x; k1y = k7y;
                    seg = true; sax = x; say = y;
                    x = xt; y = yt;
                    bx0 = fmin(bx0, x); bx1 = fmax(bx1, x); by0 = fmin(by0, y); by1 = fmax(by1, y);
                    if (MODE == 2) {
                        if (nvert < tp.max_verts) { vout[2 * nvert] = x; vout[2 * nvert + 1] = y; }
                        else status = PATH_TRACE_FULL;
                    }
                    ++nvert;
                }

                // :247  dt = 0.9 * min((tol/(est+EPS))**(1/5), maxstep/(ds+EPS), 10) * dt
                // x**0.2 < c  <=>  x < c^5 : the pow is evaluated only when the error term governs.
                const double cb = fmin(maxstep * rcp_fast(ds + EPS), 10.0);
                const double xr = tol * rcp_fast(est + EPS);
                const double c2 = cb * cb;
                double mn = cb;
                if (xr < c2 * c2 * cb) mn = fmin(pow(xr, 0.2), cb);
                dt = 0.9 * mn * dt;
            } while (false);
        }
        if (MODE == 1) {
            // chronicle the accepted step (capturezone.py:120 -> probabilityfield.py:338-339), then reconverge
            if (seg) chained |= raster_seg<RF>(L, s_lat, bm, cw, sax, say, x, y, ctr, chained);
            __syncwarp();
        }
    }

    // ---- per-path outputs ----
    if (active) {
        const size_t g = (size_t)r * tp.P + p;
        if (tp.end_xy) { tp.end_xy[2 * g] = x; tp.end_xy[2 * g + 1] = y; }
        if (tp.nverts) tp.nverts[g] = nvert;
        if (tp.status) tp.status[g] = (unsigned char)status;
        if (tp.attempts) tp.attempts[g] = nattempt;
        if (tp.path_bbox) { double *pb = tp.path_bbox + 4 * g; pb[0] = bx0; pb[1] = bx1; pb[2] = by0; pb[3] = by1; }
    }

    // ---- guarded mode: tell the flush (and the host) that this realization ran off the lattice ----
    if (MODE == 1 && tp.slot_flags != nullptr) {
        if (__any_sync(0xffffffffu, ctr.clipped != 0) && (threadIdx.x & 31) == 0) atomicOr(tp.slot_flags + r, 1u);
    }

    // ---- statistics: warp-reduce, one atomic per warp per word ----
    unsigned long long att = (unsigned long long)nattempt, stp = active ? (unsigned long long)(nvert - 1) : 0ull;
    unsigned int npath = active ? 1u : 0u, nbad = (active && status != PATH_OK) ? 1u : 0u;
    unsigned int ncl = ctr.clipped, nex = ctr.exact;
#ifndef ONEKA_EMU                                             // (the emulation runs one lane per "warp")
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        att += __shfl_down_sync(0xffffffffu, att, o);
        stp += __shfl_down_sync(0xffffffffu, stp, o);
        npath += __shfl_down_sync(0xffffffffu, npath, o);
        nbad += __shfl_down_sync(0xffffffffu, nbad, o);
        ncl += __shfl_down_sync(0xffffffffu, ncl, o);
        nex += __shfl_down_sync(0xffffffffu, nex, o);
        bx0 = fmin(bx0, __shfl_down_sync(0xffffffffu, bx0, o));
        bx1 = fmax(bx1, __shfl_down_sync(0xffffffffu, bx1, o));
        by0 = fmin(by0, __shfl_down_sync(0xffffffffu, by0, o));
        by1 = fmax(by1, __shfl_down_sync(0xffffffffu, by1, o));
    }
#endif
    if ((threadIdx.x & 31) == 0 && npath) {
        atomicAdd(tp.stats + STAT_ATTEMPTS, att);
        atomicAdd(tp.stats + STAT_STEPS, stp);
        atomicAdd(tp.stats + STAT_PATHS, (unsigned long long)npath);
        if (nbad) atomicAdd(tp.stats + STAT_NOT_OK, (unsigned long long)nbad);
        if (ncl) atomicAdd(tp.stats + STAT_CLIPPED, (unsigned long long)ncl);
        if (nex) atomicAdd(tp.stats + STAT_EXACT, (unsigned long long)nex);
        atomicMin(tp.stats + STAT_XMIN, dkey(bx0));
        atomicMax(tp.stats + STAT_XMAX, dkey(bx1));
        atomicMin(tp.stats + STAT_YMIN, dkey(by0));
        atomicMax(tp.stats + STAT_YMAX, dkey(by1));
    }
}

// ------------------------------------------------------------------------------------------
// Per-CTA staging of one realization: the well store in shared memory + the realization constants
// (called by every thread of the CTA; ends with a barrier)
__device__ __forceinline__ double cf_F(const TrackParams &tp, long long r) { return tp.coef[6 * r + 5]; }

template <bool CONFINED>
__device__ __forceinline__ void stage_realization(const TrackParams &tp, long long r, RealConsts &rc, double *s_wells)
{
    const double H = tp.thick[r], n = tp.poro[r], k = tp.cond[r];
    const double scale = CONFINED ? 1.0 / (H * n) : 1.0;
    if (CONFINED) {                                                                  // layout: oneka_device.cuh, SWELL_BLK
        for (int j = threadIdx.x; j < ((tp.nw + 3) & ~3); j += blockDim.x) {
            double b = 0.0, cx = 0.0, cy = 0.0;
            if (j < tp.nw) {
                const int i = j;
                const double w = tp.q[(size_t)r * tp.nw + i] * 0.15915494309189535 * scale;    // q/(2 pi H n)
                b = (w != 0.0) ? 1.0 / w : 1e100;                                    // q = 0: the term becomes ~1e-100, i.e. nothing
                cx = -(tp.well_xy[2 * i] - tp.xo) * b;
                cy = -(tp.well_xy[2 * i + 1] - tp.yo) * b;
            }
            double *d = s_wells + (j >> 2) * SWELL_BLK + 3 * (j & 3);
            d[0] = b; d[1] = cx; d[2] = cy;
        }
    } else {
        for (int i = threadIdx.x; i < ((tp.nw + 3) & ~3); i += blockDim.x) {        // layout: oneka_device.cuh, WELL_BLK
            const bool real = i < tp.nw;
            const double w = real ? tp.q[(size_t)r * tp.nw + i] * 0.15915494309189535 : 0.0;    // q/(2 pi)
            well_x(s_wells, i) = real ? tp.well_xy[2 * i] : 0.0;
            well_y(s_wells, i) = real ? tp.well_xy[2 * i + 1] : 0.0;
            well_w(s_wells, i) = w;
            well_w32(s_wells, i) = (float)w;
        }
    }
    if (threadIdx.x == 0) {
        const double *cf = tp.coef + 6 * r;
        rc.a2 = 2.0 * cf[0] * scale;
        rc.b2 = 2.0 * cf[1] * scale;
        rc.c = cf[2] * scale;
        rc.d = cf[3] * scale;
        rc.e = cf[4] * scale;
        rc.A = cf[0]; rc.B = cf[1]; rc.F = cf[5];
        rc.k = k; rc.H = H; rc.n = n;
        rc.half_kH2 = 0.5 * k * (H * H);
        rc.inv_Hn = 1.0 / (H * n);
        rc.xo = tp.xo; rc.yo = tp.yo;
    }
    __syncthreads();
    if (!CONFINED) {
        // error bound of the FP32 screening sum: per term <= |w| (47 * 2^-23 * 2 + 2^-22) log2 units, accumulation
        // <= nw * 2^-24 * 47 sum|w|; times 0.5 ln 2, with a 5x margin:  2e-5 (nw + 16) sum|w|
        if (threadIdx.x == 0) {
            double sw = 0.0;
            for (int i = 0; i < tp.nw; ++i) sw += fabs(well_w(s_wells, i));
            rc.pot_err = 2e-5 * (double)(tp.nw + 16) * sw + 1e-9 * fabs(cf_F(tp, r));
        }
        __syncthreads();
    }
}

}  // namespace oneka

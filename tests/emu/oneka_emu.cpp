// oneka_emu.cpp -- the DEVICE code of onekapy_b200/csrc/oneka_device.cuh compiled for the HOST.
//
// TEST INFRASTRUCTURE ONLY.  It lets `pytest -m "not gpu"` exercise the logic of the CUDA kernels -- the far-field
// evaluation, the Dormand-Prince tracker, the scan-line rasteriser -- against the oracle without a GPU, so that a change to
// the device source can be checked for logic errors before GPU time is spent on it.  It is not a fallback: nothing in
// onekapy_b200/ imports, links or knows about it, it is built only by tests/emu/emu.py into tests/emu/_build/, and it
// says nothing about performance.  What it does NOT cover: the kernels' launch code, shared-memory staging of the far field
// (restated below), warp-level reductions, atomics under contention -- those are what the `-m gpu` tests are for.
//
// How: one "thread" at a time.  threadIdx = 0, blockDim = 1, warp votes are the lane's own predicate, atomics are plain
// read-modify-writes, the MUFU approximations are libm calls (ONEKA_EMU in the header).  IEEE-exact intrinsics
// (__dmul_rn, ...) are the plain operators: build with -ffp-contract=off so the compiler does not fuse them.
#define ONEKA_EMU 1
#include <cuda_runtime.h>

#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <vector>

// ---- the CUDA vocabulary the device header uses -------------------------------------------------------------
#ifndef __noinline__
#define __noinline__ __attribute__((noinline))
#endif
static uint3 threadIdx = {0, 0, 0};
static dim3 blockDim(1, 1, 1);
static inline void __syncthreads() {}
static inline void __syncwarp() {}
static inline int __any_sync(unsigned, int p) { return p; }
static inline double __dadd_rn(double a, double b) { return a + b; }
static inline double __dsub_rn(double a, double b) { return a - b; }
static inline double __dmul_rn(double a, double b) { return a * b; }
static inline double __ddiv_rn(double a, double b) { return a / b; }
static inline int __double2int_rd(double a)
{
    if (!(a == a)) return 0;                                  // cvt.rmi.s32.f64: nan -> 0, saturating
    const double f = std::floor(a);
    if (f >= 2147483647.0) return 2147483647;
    if (f <= -2147483648.0) return (-2147483647 - 1);
    return (int)f;
}
static inline float __saturatef(float a) { return (a == a) ? std::min(1.0f, std::max(0.0f, a)) : 0.0f; }
static inline float __fdividef(float a, float b) { return a / b; }
static inline float rsqrtf(float a) { return 1.0f / std::sqrt(a); }
static inline int __float_as_int(float a) { int i; std::memcpy(&i, &a, 4); return i; }
static inline long long __double_as_longlong(double a) { long long i; std::memcpy(&i, &a, 8); return i; }
static inline int __double2loint(double a) { long long i; std::memcpy(&i, &a, 8); return (int)(unsigned int)(i & 0xffffffffLL); }
static inline int __popc(unsigned int v) { return __builtin_popcount(v); }
static inline int __ffs(unsigned int v) { return __builtin_ffs((int)v); }
template <typename T> static inline T __ldg(const T *p) { return *p; }
static inline unsigned int atomicOr(unsigned int *p, unsigned int v) { const unsigned int o = *p; *p = o | v; return o; }
static inline unsigned long long atomicOr(unsigned long long *p, unsigned long long v) { const unsigned long long o = *p; *p = o | v; return o; }
static inline unsigned long long atomicAdd(unsigned long long *p, unsigned long long v) { const unsigned long long o = *p; *p = o + v; return o; }
static inline unsigned long long atomicMin(unsigned long long *p, unsigned long long v) { const unsigned long long o = *p; *p = std::min(o, v); return o; }
static inline unsigned long long atomicMax(unsigned long long *p, unsigned long long v) { const unsigned long long o = *p; *p = std::max(o, v); return o; }
using std::isfinite;
using std::max;
using std::min;

#include "../../onekapy_b200/csrc/oneka_farfield_host.h"      // pulls in oneka_device.cuh

using namespace oneka;

// ---- host restatements of the few lines of oneka_api.cu the emulation needs (make_lattice, the far-field staging) ----
static void emu_lattice(double xmin, double ymin, double dx, double dy, int nrows, int ncols, double umbra, LatticeDev &L)
{
    std::memset(&L, 0, sizeof(L));
    L.xmin = xmin; L.ymin = ymin; L.dx = dx; L.dy = dy;
    L.nrows = nrows; L.ncols = ncols; L.wpr = ((ncols + 63) / 64) * 2;      // make_lattice: an even number of words per row
    L.umbra = umbra;
    L.umbra2 = umbra * umbra;
    L.dx32 = (float)dx; L.dy32 = (float)dy; L.umbra2_32 = (float)L.umbra2;
    L.umbra32 = (float)umbra; L.inv_dx32 = 1.0f / L.dx32;
    L.maxd = dx > dy ? dx : dy;
    L.inv_dx = 1.0 / dx; L.inv_dy = 1.0 / dy;
    L.cxl = xmin + umbra; L.cxr = xmin - umbra;
    L.cyb = ymin + umbra; L.cyt = ymin - umbra;
    L.s16x = 65536.0 / dx; L.s16y = 65536.0 / dy;
    L.fixed_ok = (nrows < 30000 && ncols < 30000) ? 1 : 0;
    L.words = (unsigned long long)L.nrows * (unsigned long long)L.wpr;
}

extern "C" {

// the rasteriser flavour of the next calls (oneka_set_raster_mode; launch_track's choice is by lattice): RF_PLAIN, RF_HEAVY
static int g_raster_flavour = 0;
void oneka_emu_set_raster_flavour(int rf) { g_raster_flavour = (rf == 1) ? 1 : 0; }

// mode 0: tracking only; 1: track + rasterise + register (counts[nrows][ncols] +=); 2: vertices kept (verts[R][P][max_verts][2]).
// far field: ff_order > 0 switches it on for the tile grid (ff_x0, ff_y0, ff_tile, ff_ntx, ff_nty), eta = ff_eta.
// stats[16] as the library's (attempts, steps, paths, not ok, clipped, exact re-tests, bbox keys decoded into bbox_out[4]).
int oneka_emu_capture(int mode, int nw, const double *well_xy, double xo, double yo, int confined,
                      double duration, double tol, double maxstep, long long max_attempts,
                      long long R, int P, const double *q, const double *cond, const double *poro, const double *thick,
                      const double *coef, const double *start_xy,
                      double xmin, double ymin, double dx, double dy, int nrows, int ncols, double umbra, unsigned int *counts,
                      int max_verts, double *verts, double *end_xy, int *nverts, unsigned char *status, int *attempts,
                      int ff_order, double ff_eta, double ff_x0, double ff_y0, double ff_tile, int ff_ntx, int ff_nty,
                      unsigned long long *stats_out, double *bbox_out, double *path_bbox /*[R][P][4] or null*/,
                      const int *clip /*[R][P][4] or null (16-byte aligned)*/)
{
    if (mode < 0 || mode > 2 || R < 0 || P <= 0) return -1;
    TrackParams tp;
    std::memset(&tp, 0, sizeof(tp));
    tp.nw = nw; tp.P = P; tp.R = R;
    tp.duration = duration; tp.tol = tol; tp.maxstep = maxstep;
    const long long ma = max_attempts > 0 ? max_attempts : ((long long)1 << 22);
    tp.max_attempts = (int)(ma > 0x7fffffffLL ? 0x7fffffffLL : ma);
    tp.xo = xo; tp.yo = yo;
    tp.well_xy = well_xy;
    tp.q = q; tp.cond = cond; tp.poro = poro; tp.thick = thick; tp.coef = coef; tp.start_xy = start_xy;
    tp.end_xy = end_xy; tp.nverts = nverts; tp.status = status; tp.attempts = attempts;
    tp.verts = verts; tp.max_verts = max_verts;
    tp.path_bbox = path_bbox; tp.clip = clip;
    unsigned long long stats[16];
    std::memset(stats, 0, sizeof(stats));
    stats[STAT_XMIN] = ~0ULL; stats[STAT_YMIN] = ~0ULL;
    tp.stats = stats;

    LatticeDev L;
    std::memset(&L, 0, sizeof(L));
    if (mode == 1) emu_lattice(xmin, ymin, dx, dy, nrows, ncols, umbra, L);
    double s_lat[5] = {L.xmin, L.ymin, L.dx, L.dy, L.umbra2};

    // far-field tables (geometry) once; coefficients per realization
    const bool use_ff = ff_order > 0;                          // unconfined: the opt-in path of oneka_set_farfield_unconfined
    FFTables T;
    FarFieldDev ff;
    std::memset(&ff, 0, sizeof(ff));
    if (use_ff) {
        if (build_ff_tables(nw, well_xy, xo, yo, ff_x0 - xo, ff_y0 - yo, ff_tile, ff_ntx, ff_nty, ff_order, ff_eta, T)) return -2;
        ff.ntx = ff_ntx; ff.nty = ff_nty; ff.order = ff_order; ff.max_near = T.max_near;
        ff.gx0 = ff_x0 - xo; ff.gy0 = ff_y0 - yo; ff.inv_tile = 1.0 / ff_tile;
    }

    std::vector<double> s_wells((size_t)well_store_doubles(nw) + 64, 0.0);
    std::vector<unsigned int> bitmap(mode == 1 ? (size_t)L.words : 1, 0u);
    std::vector<double> wsc(nw > 0 ? nw : 1);
    std::vector<double2> c64;
    std::vector<float2> p32;
    std::vector<double> b0;
    RealConsts rc;
    for (long long r = 0; r < R; ++r) {
        if (confined) stage_realization<true>(tp, r, rc, s_wells.data());
        else stage_realization<false>(tp, r, rc, s_wells.data());
        FarFieldShared fs = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
        if (use_ff && !confined) {                                                                       // track_kernel<false, ., true>'s staging
            for (int w = 0; w < nw; ++w) wsc[w] = q[(size_t)r * nw + w] * 0.15915494309189535;           // farfield_coef_unc_kernel
            ff_host_coefficients(T, nw, ff_order, wsc.data(), c64);
            ff_host_b0(T, nw, wsc.data(), b0);
            const double h = 0.7071067811865476 / ff.inv_tile;
            p32.resize(c64.size());
            for (size_t i = 0; i < c64.size(); ++i) {
                const double f = h / (double)(i % ff_order + 1);
                p32[i] = make_float2((float)(c64[i].x * f), (float)(c64[i].y * f));
            }
            fs.c64 = c64.data(); fs.p32 = p32.data(); fs.b0 = b0.data(); fs.idx = T.idx.data(); fs.raw = T.cnt_raw.data();
            rc.pot_err *= 1.25;
        }
        if (use_ff && confined) {
            const double scale = 1.0 / (thick[r] * poro[r]);
            for (int w = 0; w < nw; ++w) wsc[w] = q[(size_t)r * nw + w] * 0.15915494309189535 * scale;   // farfield_coef_kernel
            ff_host_coefficients(T, nw, ff_order, wsc.data(), c64);
            double *d = s_wells.data() + ff_dummy_offset(nw);                                            // track_kernel's staging
            d[0] = 1e100; d[1] = 1.0; d[2] = 1.0;
            fs.c64 = c64.data(); fs.off = T.off.data(); fs.cnt = T.cnt.data();
        }
        for (int p = 0; p < P; ++p) {
            if (confined) {
                if (use_ff && ff_order == 16) {                                                          // launch_track's choice: the unrolled order
                    if (mode == 0) dopri_track<true, 0, true, 16>(tp, L, s_lat, nullptr, rc, s_wells.data(), r, p, true, ff, fs);
                    else if (mode == 1) { if (g_raster_flavour) dopri_track<true, 1, true, 16, RF_HEAVY>(tp, L, s_lat, bitmap.data(), rc, s_wells.data(), r, p, true, ff, fs); else dopri_track<true, 1, true, 16, RF_PLAIN>(tp, L, s_lat, bitmap.data(), rc, s_wells.data(), r, p, true, ff, fs); }
                    else dopri_track<true, 2, true, 16>(tp, L, s_lat, nullptr, rc, s_wells.data(), r, p, true, ff, fs);
                } else if (use_ff) {
                    if (mode == 0) dopri_track<true, 0, true>(tp, L, s_lat, nullptr, rc, s_wells.data(), r, p, true, ff, fs);
                    else if (mode == 1) { if (g_raster_flavour) dopri_track<true, 1, true, 0, RF_HEAVY>(tp, L, s_lat, bitmap.data(), rc, s_wells.data(), r, p, true, ff, fs); else dopri_track<true, 1, true, 0, RF_PLAIN>(tp, L, s_lat, bitmap.data(), rc, s_wells.data(), r, p, true, ff, fs); }
                    else dopri_track<true, 2, true>(tp, L, s_lat, nullptr, rc, s_wells.data(), r, p, true, ff, fs);
                } else {
                    if (mode == 0) dopri_track<true, 0, false>(tp, L, s_lat, nullptr, rc, s_wells.data(), r, p, true);
                    else if (mode == 1) { if (g_raster_flavour) dopri_track<true, 1, false, 0, RF_HEAVY>(tp, L, s_lat, bitmap.data(), rc, s_wells.data(), r, p, true); else dopri_track<true, 1, false, 0, RF_PLAIN>(tp, L, s_lat, bitmap.data(), rc, s_wells.data(), r, p, true); }
                    else dopri_track<true, 2, false>(tp, L, s_lat, nullptr, rc, s_wells.data(), r, p, true);
                }
            } else if (use_ff && ff_order == 16) {                                                   // launch_track's choice: the unrolled order
                if (mode == 0) dopri_track<false, 0, true, 16>(tp, L, s_lat, nullptr, rc, s_wells.data(), r, p, true, ff, fs);
                else if (mode == 1) { if (g_raster_flavour) dopri_track<false, 1, true, 16, RF_HEAVY>(tp, L, s_lat, bitmap.data(), rc, s_wells.data(), r, p, true, ff, fs); else dopri_track<false, 1, true, 16, RF_PLAIN>(tp, L, s_lat, bitmap.data(), rc, s_wells.data(), r, p, true, ff, fs); }
                else dopri_track<false, 2, true, 16>(tp, L, s_lat, nullptr, rc, s_wells.data(), r, p, true, ff, fs);
            } else if (use_ff) {
                if (mode == 0) dopri_track<false, 0, true>(tp, L, s_lat, nullptr, rc, s_wells.data(), r, p, true, ff, fs);
                else if (mode == 1) { if (g_raster_flavour) dopri_track<false, 1, true, 0, RF_HEAVY>(tp, L, s_lat, bitmap.data(), rc, s_wells.data(), r, p, true, ff, fs); else dopri_track<false, 1, true, 0, RF_PLAIN>(tp, L, s_lat, bitmap.data(), rc, s_wells.data(), r, p, true, ff, fs); }
                else dopri_track<false, 2, true>(tp, L, s_lat, nullptr, rc, s_wells.data(), r, p, true, ff, fs);
            } else {
                if (mode == 0) dopri_track<false, 0, false>(tp, L, s_lat, nullptr, rc, s_wells.data(), r, p, true);
                else if (mode == 1) { if (g_raster_flavour) dopri_track<false, 1, false, 0, RF_HEAVY>(tp, L, s_lat, bitmap.data(), rc, s_wells.data(), r, p, true); else dopri_track<false, 1, false, 0, RF_PLAIN>(tp, L, s_lat, bitmap.data(), rc, s_wells.data(), r, p, true); }
                else dopri_track<false, 2, false>(tp, L, s_lat, nullptr, rc, s_wells.data(), r, p, true);
            }
        }
        if (mode == 1) {                                        // flush_kernel: register(1.0), bitmap back to zero
            for (int i = 0; i < L.nrows; ++i)
                for (int j = 0; j < L.ncols; ++j) {
                    unsigned int &w = bitmap[(size_t)i * L.wpr + (j >> 5)];
                    if ((w >> (j & 31)) & 1u) counts[(size_t)i * L.ncols + j] += 1u;
                }
            std::fill(bitmap.begin(), bitmap.end(), 0u);
        }
    }
    if (stats_out) std::memcpy(stats_out, stats, sizeof(stats));
    if (bbox_out) {
        auto undkey = [](unsigned long long k) {
            unsigned long long b = (k >> 63) ? (k & 0x7fffffffffffffffULL) : ~k;
            double v;
            std::memcpy(&v, &b, 8);
            return v;
        };
        bbox_out[0] = undkey(stats[STAT_XMIN]); bbox_out[1] = undkey(stats[STAT_XMAX]);
        bbox_out[2] = undkey(stats[STAT_YMIN]); bbox_out[3] = undkey(stats[STAT_YMAX]);
    }
    return 0;
}

// raster_seg on given polylines, one realization per trace group (the logic of raster_traces_kernel + flush_kernel)
int oneka_emu_raster_traces(double xmin, double ymin, double dx, double dy, int nrows, int ncols, double umbra,
                            long long ntraces, const long long *offsets, const double *verts, const int *real_of, long long nreal,
                            unsigned int *counts, unsigned long long *exact_out)
{
    LatticeDev L;
    emu_lattice(xmin, ymin, dx, dy, nrows, ncols, umbra, L);
    double s_lat[5] = {L.xmin, L.ymin, L.dx, L.dy, L.umbra2};
    std::vector<unsigned int> bitmap((size_t)L.words, 0u);
    unsigned long long nexact = 0;
    for (long long r = 0; r < nreal; ++r) {
        for (long long t = 0; t < ntraces; ++t) {
            if (real_of[t] != r) continue;
            RasterCounters ctr = {0u, 0u};
            bool chained = false;                                  // raster_traces_kernel: consecutive segments of a trace chain
            const ClipWin all = {0, L.ncols, 0, L.nrows};
            for (long long v = offsets[t]; v + 1 < offsets[t + 1]; ++v) {
                const double *a = verts + 2 * v;
                if (g_raster_flavour == RF_HEAVY) chained |= raster_seg<RF_HEAVY>(L, s_lat, bitmap.data(), all, a[0], a[1], a[2], a[3], ctr, chained);
                else raster_seg<RF_PLAIN>(L, s_lat, bitmap.data(), all, a[0], a[1], a[2], a[3], ctr);
            }

            nexact += ctr.exact;
        }
        for (int i = 0; i < L.nrows; ++i)
            for (int j = 0; j < L.ncols; ++j)
                if ((bitmap[(size_t)i * L.wpr + (j >> 5)] >> (j & 31)) & 1u) counts[(size_t)i * L.ncols + j] += 1u;
        std::fill(bitmap.begin(), bitmap.end(), 0u);
    }
    if (exact_out) *exact_out = nexact;
    return 0;
}

}  // extern "C"

// oneka_api.cu -- kernels + C ABI of liboneka_b200.so (see include/oneka_b200.h).
//
// Build (done by __graft_entry__.build()):
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -shared -Xcompiler -fPIC \
//        -Iinclude onekapy_b200/csrc/oneka_api.cu -o onekapy_b200/liboneka_b200.so
#include "oneka_device.cuh"
#include "oneka_farfield_host.h"
#include "../../include/oneka_b200.h"

#include <nccl.h>      // types and enums only: the functions are resolved at run time (nccl_api below)
#include <dlfcn.h>
#include <cstdio>
#include <cstdlib>
#include <cstdarg>
#include <cstring>
#include <vector>
#include <new>

using namespace oneka;

// ------------------------------------------------------------------------------------------
static thread_local char g_err[512] = "";

static int fail(int code, const char *fmt, ...)
{
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
    return code;
}

#define CUDA_TRY(expr)                                                                      \
    do {                                                                                    \
        cudaError_t _e = (expr);                                                            \
        if (_e != cudaSuccess)                                                              \
            return fail(ONEKA_ERR_CUDA, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, __LINE__); \
    } while (0)

// CTA shapes of the tracking kernel (every warp stays inside one realization; P = 1000 paths = 8 x 125 = 4 x 250 lanes).
//   direct sums (and everything without the far field): 128 threads, 6 CTAs per SM -- 80 registers, 24 warps: these kernels are
//     bound by the issue port and want the warps (measured 5 / 6 / 8 CTAs within 2 %; 256 x 2 loses 6 %, profiles/r02_knob_scan5.txt);
//   far field: 256 threads, 2 CTAs per SM -- 128 registers (no spills; 80 cost 200 B of them) and up to ~112 KB of shared memory
//     for the realization's coefficient table, i.e. ~6x the tiles of the 128 x 6 shape at a lower order: C3 79.9 -> 64.6 ms,
//     C4 33.1 -> 25.6 ms per step (profiles/r02_knob_scan2-4.txt).
constexpr int TRACK_THREADS = 128, TRACK_MIN_CTAS = 6;      // (5 and 4 CTAs per SM, no spills: within 1.5 % on C5 / C1 / direct C3 / C4, re-measured in round 2)
constexpr int FF_THREADS = 256, FF_MIN_CTAS = 2;
constexpr int FF_UNC_THREADS = 256, FF_UNC_MIN_CTAS = 2;     // unconfined far field: the same shape (128 x 6 measured 2x slower, scan 6)
constexpr int FF_ORDER_UNROLLED = 16;        // the order whose Horner loop is unrolled at compile time (Engine's default)
constexpr int N_STATS = 16;

struct oneka_ctx {
    int device = 0;
    int sm_count = 0;
    size_t smem_per_sm = 0;                 // cudaDevAttrMaxSharedMemoryPerMultiprocessor
    int raster_mode = 0;                    // 0: rasteriser flavour by lattice (raster_flavour), 1: plain, 2: heavy
    cudaStream_t stream = nullptr;
    unsigned int *bitmaps = nullptr;        // registration bitmaps, all-zero between calls
    size_t bitmap_bytes = 0;
    size_t workspace_limit = (size_t)16 << 30;
    unsigned long long *stats_dev = nullptr;
    uint64_t launches = 0;
    // host-buffer staging (oneka_capture_host), grow-only
    void *stage = nullptr;
    size_t stage_bytes = 0;
    // small scratch of the auxiliary entry points (Gaussian taps, point evaluation, distancesquared), grow-only
    void *aux = nullptr;
    size_t aux_bytes = 0;
    // far-field compression (oneka_set_farfield): tile geometry, static tables, per-launch coefficient workspace
    struct FarField {
        bool on = false;
        int nw = 0, ntx = 0, nty = 0, order = 0, max_near = 0;
        double xo = 0, yo = 0, gx0 = 0, gy0 = 0, tile = 0, eta = 0, mean_near = 0;
        size_t smem_confined = 0, smem_unconfined = 0;   // dynamic shared memory per CTA with these tables
        double2 *P = nullptr;                 // [ntiles][nw][order]
        unsigned int *near_off = nullptr;     // [ntiles][max_near] byte offsets into the well store
        unsigned short *near_cnt = nullptr;   // [ntiles]
        double2 *coef = nullptr;              // [realizations of a launch][ntiles][order], grow-only
        size_t coef_bytes = 0;
        double *wells = nullptr;              // [nw][2] device copy of the coordinates the tables were built from (ff_check_wells_kernel)
        // unconfined flow (oneka_set_farfield_unconfined)
        bool unconfined = false;
        double *Lg = nullptr;                 // [ntiles][nw] ln |z_w - z_c| of the far wells
        unsigned short *near_idx = nullptr, *near_raw = nullptr;
        double *b0 = nullptr;                 // [realizations of a launch][ntiles], grow-only
        size_t b0_bytes = 0;
    } ff;
    // the one collective of the path (oneka_allreduce_counts): communicator owned (comm_init_rank) or borrowed (comm_attach)
    ncclComm_t comm = nullptr;
    bool comm_owned = false;
    int comm_nranks = 1, comm_rank = 0;
    // profiling
    bool profiling = false;
    struct EvPair { cudaEvent_t a, b; int kind; };
    std::vector<EvPair> events;
    double track_ms = 0.0, flush_ms = 0.0;
    uint64_t track_launches = 0;
};

// ------------------------------------------------------------------------------------------
// Kernels
// ------------------------------------------------------------------------------------------
// stage_realization<CONFINED> (the per-CTA well store and realization constants) lives in oneka_device.cuh

// One CTA = THREADS consecutive paths of ONE realization; grid = R * ceil(P/THREADS).
// FF: the realization's far-field coefficient table and the tiles' near lists are staged behind the well store (see
// "Far-field compression" in oneka_device.cuh); ORD = its order when that is a compile-time constant, else 0.
// RF: the rasteriser's flavour (raster_seg<RF>), chosen per lattice by raster_flavour().
template <bool CONFINED, int MODE, bool FF, int ORD, int THREADS, int MIN_CTAS, int RF = RF_PLAIN>
__global__ void __launch_bounds__(THREADS, MIN_CTAS)
track_kernel(TrackParams tp, LatticeDev L, unsigned int *bitmaps, FarFieldDev ff)
{
    extern __shared__ double2 s_dyn[];
    __shared__ RealConsts rc;
    __shared__ double s_lat[5];
    double *s_wells = reinterpret_cast<double *>(s_dyn);
    if (MODE == 1) stage_lattice(L, s_lat);

    const int chunks = (tp.P + THREADS - 1) / THREADS;
    const long long r = blockIdx.x / chunks;
    const int p = (int)(blockIdx.x % chunks) * THREADS + threadIdx.x;
    FarFieldShared fs = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
    if (FF && CONFINED) {
        const int ntiles = ff.ntx * ff.nty, order = ff.order;
        double2 *s_c64 = s_dyn + ff_store_double2(tp.nw);
        unsigned int *s_off = reinterpret_cast<unsigned int *>(s_c64 + ntiles * order);
        unsigned short *s_cnt = reinterpret_cast<unsigned short *>(s_off + ntiles * ff.max_near);
        const double2 *g = ff.coef + (size_t)r * ntiles * order;
        for (int i = threadIdx.x; i < ntiles * order; i += blockDim.x) s_c64[i] = g[i];     // coalesced copy of the realization's table
        for (int i = threadIdx.x; i < ntiles * ff.max_near; i += blockDim.x) s_off[i] = ff.near_off[i];
        for (int i = threadIdx.x; i < ntiles; i += blockDim.x) s_cnt[i] = ff.near_cnt[i];
        if (threadIdx.x == 0) {                                  // the dummy well that pads odd near lists: term ~1e-100
            double *d = s_wells + ff_dummy_offset(tp.nw);
            d[0] = 1e100; d[1] = 1.0; d[2] = 1.0;
        }
        fs.c64 = s_c64; fs.off = s_off; fs.cnt = s_cnt;
    }
    if (FF && !CONFINED) {
        // unconfined: c64 (discharge) + p32 (potential, p_k = h c_(k-1)/k as float2) + b0 + near well indices
        const int ntiles = ff.ntx * ff.nty, order = ff.order;
        const double h = 0.7071067811865476 / ff.inv_tile;                        // tile / sqrt 2
        double2 *s_c64 = s_dyn + ff_store_double2(tp.nw);
        float2 *s_p32 = reinterpret_cast<float2 *>(s_c64 + ntiles * order);
        double *s_b0 = reinterpret_cast<double *>(s_p32 + ntiles * order);
        unsigned short *s_idx = reinterpret_cast<unsigned short *>(s_b0 + ntiles);
        unsigned short *s_raw = s_idx + ntiles * ff.max_near;
        const double2 *g = ff.coef + (size_t)r * ntiles * order;
        for (int i = threadIdx.x; i < ntiles * order; i += blockDim.x) {
            const int k = i % order;
            const double2 c = g[i];
            s_c64[i] = c;
            const double f = h / (double)(k + 1);
            s_p32[i] = make_float2((float)(c.x * f), (float)(c.y * f));
        }
        for (int i = threadIdx.x; i < ntiles; i += blockDim.x) { s_b0[i] = ff.b0[(size_t)r * ntiles + i]; s_raw[i] = ff.near_raw[i]; }
        for (int i = threadIdx.x; i < ntiles * ff.max_near; i += blockDim.x) s_idx[i] = ff.near_idx[i];
        fs.c64 = s_c64; fs.p32 = s_p32; fs.b0 = s_b0; fs.idx = s_idx; fs.raw = s_raw;
    }
    stage_realization<CONFINED>(tp, r, rc, s_wells);             // ends with __syncthreads()
    if (FF && !CONFINED) {                                       // the far potential is one more FP32 sum: <= 3e-6 sum|w| of error,
        if (threadIdx.x == 0) rc.pot_err *= 1.25;                // far inside a quarter of the screening bound 2e-5 (nw + 16) sum|w|
        __syncthreads();
    }
    unsigned int *bm = (MODE == 1) ? bitmaps + (size_t)r * L.words : nullptr;
    dopri_track<CONFINED, MODE, FF, ORD, RF>(tp, L, s_lat, bm, rc, s_wells, r, p, p < tp.P, ff, fs);
}

// c[r][tile][k] = sum_w w_rw P[tile][w][k]:  a (realizations x wells) . (wells x tiles*order) product with complex P, i.e. a
// small FP64 GEMM formed per launch.  w_rw = q_rw / (2 pi H_r n_r) for confined flow (the scaling of stage_realization<true>),
// q_rw / (2 pi) for unconfined flow.  One thread = one (tile, k) entry for COEF_RB realizations at once: the P rows are read
// coalesced (k is the fastest index) and every 16-byte P load feeds COEF_RB complex FMAs from registers, the scaled
// discharges of the realization block are broadcast from shared memory.  (Round 1 had one warp per (realization, tile):
// every P element re-read from L2 for every realization -- 28 GB per 10 000 realizations at 378 tiles, 3.3 ms of a 158 ms step.)
constexpr int COEF_RB = 8, COEF_THREADS = 128, COEF_WCHUNK = 512;    // wells staged per pass: 8 x 512 doubles = 32 KB

template <bool CONFINED>
__global__ void __launch_bounds__(COEF_THREADS)
farfield_coef_kernel(int nw, int ntiles, int order, long long nr, const double2 *__restrict__ P, const double *__restrict__ q,
                     const double *__restrict__ poro, const double *__restrict__ thick, double2 *__restrict__ out)
{
    __shared__ double s_w[COEF_RB * COEF_WCHUNK];                // [COEF_RB][wells of this pass]
    const long long r0 = (long long)blockIdx.y * COEF_RB;
    const int nb = (int)min((long long)COEF_RB, nr - r0);
    const int e = blockIdx.x * COEF_THREADS + threadIdx.x;       // entry (tile, k)
    const bool live = e < ntiles * order;
    const int t = live ? e / order : 0, k = live ? e - t * order : 0;
    const double2 *Pt = P + (size_t)t * nw * order + k;
    double ar[COEF_RB], ai[COEF_RB];
#pragma unroll
    for (int j = 0; j < COEF_RB; ++j) { ar[j] = 0.0; ai[j] = 0.0; }
    for (int w0 = 0; w0 < nw; w0 += COEF_WCHUNK) {               // (one pass for fields of up to 512 wells)
        const int wn = min(COEF_WCHUNK, nw - w0);
        __syncthreads();
        for (int i = threadIdx.x; i < COEF_RB * wn; i += COEF_THREADS) {
            const int j = i / wn, w = i - j * wn;
            double v = 0.0;
            if (j < nb) {
                const long long r = r0 + j;
                const double scale = CONFINED ? 1.0 / (thick[r] * poro[r]) : 1.0;
                v = q[(size_t)r * nw + w0 + w] * 0.15915494309189535 * scale;
            }
            s_w[j * COEF_WCHUNK + w] = v;
        }
        __syncthreads();
        if (live) {
#pragma unroll 2
            for (int w = 0; w < wn; ++w) {
                const double2 pk = __ldg(Pt + (size_t)(w0 + w) * order);
#pragma unroll
                for (int j = 0; j < COEF_RB; ++j) {
                    const double ww = s_w[j * COEF_WCHUNK + w];
                    ar[j] = fma(ww, pk.x, ar[j]);
                    ai[j] = fma(ww, pk.y, ai[j]);
                }
            }
        }
    }
    if (!live) return;
#pragma unroll
    for (int j = 0; j < COEF_RB; ++j)
        if (j < nb) out[((size_t)(r0 + j) * ntiles + t) * order + k] = make_double2(ar[j], ai[j]);
}

// unconfined flow: b0[r][tile] = sum_w w_rw ln|z_w - z_c| over the far wells (the constant of the potential's expansion)
__global__ void __launch_bounds__(128)
farfield_b0_kernel(int nw, int ntiles, long long nr, const double *__restrict__ Lg, const double *__restrict__ q, double *__restrict__ b0)
{
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nr * ntiles) return;
    const long long r = i / ntiles;
    const int t = (int)(i - r * ntiles);
    const double *qr = q + (size_t)r * nw;
    double a = 0.0;
    for (int w = 0; w < nw; ++w) a = fma(qr[w] * 0.15915494309189535, Lg[(size_t)t * nw + w], a);
    b0[i] = a;
}

// The far-field tables are geometry: they belong to the well coordinates oneka_set_farfield was given.  Every launch that
// uses them compares the wells it was handed with that copy (bitwise; one CTA, in stream order, nothing on the host) and
// counts a mismatch in the statistics block; oneka_read_stats then FAILS instead of returning numbers computed from near
// terms of one well field and polynomials of another.
__global__ void __launch_bounds__(128)
ff_check_wells_kernel(int nw, const double *__restrict__ built_from, const double *__restrict__ well_xy, unsigned long long *stats)
{
    __shared__ unsigned int bad;
    if (threadIdx.x == 0) bad = 0u;
    __syncthreads();
    for (int i = threadIdx.x; i < 2 * nw; i += blockDim.x)
        if (__double_as_longlong(built_from[i]) != __double_as_longlong(well_xy[i])) bad = 1u;
    __syncthreads();
    if (threadIdx.x == 0 && bad) atomicAdd(stats + STAT_FF_MISMATCH, 1ULL);
}

// register(1.0) for a batch of realizations (probabilityfield.py:357-359): one thread per bitmap word position,
// looping over its slice of realization slots.  HBM-bound (every word of every slot is read once), so the loop keeps
// FLUSH_UNROLL independent streaming loads in flight per thread (32 B; one load at a time reached ~1 TB/s).  Reads are
// coalesced (consecutive threads = consecutive words of one slot); per-bit counters live in registers, bitmap words
// are zeroed as they are consumed, counts get one RED per nonzero counter.
constexpr int FLUSH_UNROLL = 8;

__global__ void __launch_bounds__(256)
flush_kernel(unsigned int *bitmaps, long long nslots, long long slots_per_y, LatticeDev L, unsigned int *counts,
             const unsigned int *slot_flags)
{
    const unsigned long long w = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (w >= L.words) return;
    const long long s0 = (long long)blockIdx.y * slots_per_y;
    const long long s1 = min(nslots, s0 + slots_per_y);
    unsigned int cnt[32];
#pragma unroll
    for (int b = 0; b < 32; ++b) cnt[b] = 0;
    bool any = false;
    auto consume = [&](unsigned int *pw, unsigned int v, long long s) {
        *pw = 0u;
        if (slot_flags != nullptr && slot_flags[s]) return;          // guarded mode: a clipped realization is not registered
        any = true;
#pragma unroll
        for (int b = 0; b < 32; ++b) cnt[b] += (v >> b) & 1u;
    };
    long long s = s0;
    unsigned int *pw = bitmaps + (size_t)s0 * L.words + w;
    for (; s + FLUSH_UNROLL <= s1; s += FLUSH_UNROLL, pw += (size_t)FLUSH_UNROLL * L.words) {
        unsigned int v[FLUSH_UNROLL];
#pragma unroll
        for (int k = 0; k < FLUSH_UNROLL; ++k) v[k] = __ldcs(pw + (size_t)k * L.words);
        unsigned int acc = 0u;
#pragma unroll
        for (int k = 0; k < FLUSH_UNROLL; ++k) acc |= v[k];
        if (acc == 0u) continue;                                     // most of a bitmap is empty
#pragma unroll
        for (int k = 0; k < FLUSH_UNROLL; ++k)
            if (v[k]) consume(pw + (size_t)k * L.words, v[k], s + k);
    }
    for (; s < s1; ++s, pw += L.words) {
        const unsigned int v = __ldcs(pw);
        if (v) consume(pw, v, s);
    }
    if (!any) return;
    const int i = (int)(w / L.wpr);
    const int j0 = (int)(w % L.wpr) * 32;
    unsigned int *row = counts + (size_t)i * L.ncols + j0;
#pragma unroll
    for (int b = 0; b < 32; ++b)
        if (cnt[b]) atomicAdd(row + b, cnt[b]);
}

// test hook: one thread per given trace, same raster_seg as the fused kernel (consecutive segments of a trace chain)
template <int RF>
__global__ void __launch_bounds__(128)
raster_traces_kernel(LatticeDev L, long long ntraces, const long long *offsets, const double *verts,
                     const int *real_of, long long s0, long long s1, unsigned int *bitmaps, unsigned long long *stats)
{
    __shared__ double s_lat[5];
    stage_lattice(L, s_lat);
    __syncthreads();
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= ntraces) return;
    const long long r = real_of[t];
    if (r < s0 || r >= s1) return;
    unsigned int *bm = bitmaps + (size_t)(r - s0) * L.words;
    RasterCounters ctr = {0u, 0u};
    unsigned long long nseg = 0;
    bool chained = false;
    for (long long v = offsets[t]; v + 1 < offsets[t + 1]; ++v) {
        chained |= raster_seg<RF>(L, s_lat, bm, ClipWin{0, L.ncols, 0, L.nrows}, verts[2 * v], verts[2 * v + 1], verts[2 * v + 2], verts[2 * v + 3],
                                  ctr, chained);
        ++nseg;
    }

    atomicAdd(stats + STAT_STEPS, nseg);
    if (ctr.clipped) atomicAdd(stats + STAT_CLIPPED, (unsigned long long)ctr.clipped);
    if (ctr.exact) atomicAdd(stats + STAT_EXACT, (unsigned long long)ctr.exact);
}

// Model.compute_* at points (model.py:207-427); one CTA, wells staged twice (both scalings)
__global__ void __launch_bounds__(128)
eval_points_kernel(TrackParams tp, long long npts, const double *pts, double *out)
{
    extern __shared__ double2 s_dyn[];
    __shared__ RealConsts rc_c, rc_u;
    double *s_wc = reinterpret_cast<double *>(s_dyn);                    // two well stores: confined and unconfined scaling
    double *s_wu = s_wc + well_store_doubles(tp.nw);
    stage_realization<true>(tp, 0, rc_c, s_wc);
    stage_realization<false>(tp, 0, rc_u, s_wu);
    for (long long i = threadIdx.x; i < npts; i += blockDim.x) {
        const double x = pts[2 * i], y = pts[2 * i + 1];
        double *o = out + 8 * i;
        // potential, model.py:226-237 + 259-266
        const double dx0 = x - rc_u.xo, dy0 = y - rc_u.yo;
        double pot = rc_u.A * dx0 * dx0 + rc_u.B * dy0 * dy0 + rc_u.c * dx0 * dy0 + rc_u.d * dx0 + rc_u.e * dy0 + rc_u.F;
        double qx = -(rc_u.a2 * dx0 + rc_u.c * dy0 + rc_u.d);           // model.py:303-304
        double qy = -(rc_u.b2 * dy0 + rc_u.c * dx0 + rc_u.e);
        for (int w = 0; w < tp.nw; ++w) {
            const double dx = x - well_x(s_wu, w), dy = y - well_y(s_wu, w), ww = well_w(s_wu, w);
            const double r2 = dx * dx + dy * dy;
            pot += 0.5 * ww * log(r2);
            qx -= ww * dx / r2;                                          // model.py:312-313
            qy -= ww * dy / r2;
        }
        o[0] = pot; o[1] = qx; o[2] = qy;
        double fx, fy;
        field_feval<true>(rc_c, s_wc, tp.nw, x, y, fx, fy);
        o[3] = -fx; o[4] = -fy;
        o[5] = o[6] = o[7] = nan("");
        if (pot > 0.0) {
            o[5] = (pot < rc_u.half_kH2) ? sqrt(2.0 * pot / rc_u.k) : (pot + rc_u.half_kH2) / (rc_u.k * rc_u.H);   // model.py:345-349
            if (field_feval<false>(rc_u, s_wu, tp.nw, x, y, fx, fy) == PATH_OK) { o[6] = -fx; o[7] = -fy; }
        }
    }
}


// ProbabilityField.distancesquared (probabilityfield.py:379-427) through the rasteriser's own exact device function
__global__ void __launch_bounds__(128)
distsq_kernel(long long n, const double *__restrict__ abc, double *__restrict__ out)
{
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const double *a = abc + 6 * i;
    out[i] = exact_distancesquared(a[0], a[1], a[2], a[3], a[4], a[5]);
}

// Atomic bit-set probes: the memory operation of the rasteriser (one RED.OR per lattice row per segment).
//   MODE 0: RED.OR.b32 to global memory (resolved in L2), consecutive lanes on consecutive words (4 sectors per warp request)
//   MODE 1: the same, all 32 lanes of a warp on one word
//   MODE 2 / 3: ATOMS.OR on a 32 KB shared-memory tile, lane-private / one word per warp
//   MODE 4: RED.OR.b32 to global memory, EVERY LANE ITS OWN 32-BYTE SECTOR, moving on by one bitmap row (160 words) per
//           operation -- the rasteriser's own pattern (each lane is another particle; a segment's rows are 620 B apart at C5)
template <int MODE>
__global__ void __launch_bounds__(256)
red_probe_kernel(unsigned int *buf, unsigned long long words, int iters, unsigned int *sink)
{
    __shared__ unsigned int tile[8192];
    constexpr bool SHARED = (MODE == 2 || MODE == 3), CONTENDED = (MODE == 1 || MODE == 3);
    const unsigned int lane = threadIdx.x & 31u;
    const unsigned long long gthread = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (SHARED) {
        for (int i = threadIdx.x; i < 8192; i += blockDim.x) tile[i] = 0u;
        __syncthreads();
        unsigned int idx = CONTENDED ? (threadIdx.x >> 5) * 37u : threadIdx.x;
#pragma unroll 4
        for (int k = 0; k < iters; ++k) {
            atomicOr(&tile[idx & 8191u], 1u << ((k + lane) & 31));
            idx += CONTENDED ? 8u : 256u;                  // a new row of the tile, conflict-free for the lane-private case
        }
        __syncthreads();
        if (tile[threadIdx.x] == 0xdeadbeefu) sink[0] = 1u;  // never true; keeps the tile alive
        return;
    }
    if (MODE == 4) {
        unsigned long long idx = (gthread * 40503ull) % words;     // lanes far apart: one sector each
#pragma unroll 4
        for (int k = 0; k < iters; ++k) {
            atomicOr(buf + idx, 1u << ((k + lane) & 31));
            idx += 160ull;                                          // the next row of the window
            if (idx >= words) idx -= words;
        }
        return;
    }
    // global: consecutive lanes on consecutive words (the coalesced pattern of neighbouring bitmap words), each warp
    // walking its own stride through the buffer
    unsigned long long idx = CONTENDED ? (gthread >> 5) * 32ull : gthread;
    const unsigned long long step = (unsigned long long)gridDim.x * blockDim.x + 4099ull * 32ull;
#pragma unroll 4
    for (int k = 0; k < iters; ++k) {
        atomicOr(buf + (idx % words), 1u << ((k + lane) & 31));   // result unused -> RED.OR
        idx += step;
    }
}

// FP64 pipe probe: 8 independent DFMA chains per thread, registers only.
__global__ void __launch_bounds__(256)
fp64_probe_kernel(int iters, double seed, double *sink)
{
    double a0 = seed + threadIdx.x, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3, a4 = a0 + 4, a5 = a0 + 5, a6 = a0 + 6, a7 = a0 + 7;
    const double m = 0.999999, c = 1e-9;
    for (int i = 0; i < iters; ++i) {
        a0 = fma(a0, m, c); a1 = fma(a1, m, c); a2 = fma(a2, m, c); a3 = fma(a3, m, c);
        a4 = fma(a4, m, c); a5 = fma(a5, m, c); a6 = fma(a6, m, c); a7 = fma(a7, m, c);
    }
    const double s = a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7;
    if (s == 12345.678) sink[0] = s;     // never true; keeps the chains alive
}

// exceedance histogram of the count grid (visualize.py:382-386 is a sort of integer-valued data).  Bins live in shared memory
// while they fit (SMEM = true: nbins <= HIST_SMEM_BINS; shared-memory atomics run at ~5e12 word operations/s, oneka_red_probe, where
// all CTAs hammering the same few global bins ran at 48 GB/s of input on the 23 M-cell C5 grid); zeros -- most nodes -- are
// counted in a register and added once per warp.
constexpr int HIST_SMEM_BINS = 12288;
template <bool SMEM>
__global__ void __launch_bounds__(256)
count_histogram_kernel(const unsigned int *counts, long long ncell, int nbins, unsigned long long *hist)
{
    extern __shared__ unsigned int s_hist[];
    if (SMEM) {
        for (int b = threadIdx.x; b < nbins; b += blockDim.x) s_hist[b] = 0u;
        __syncthreads();
    }
    unsigned long long zeros = 0;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < ncell; i += (long long)gridDim.x * blockDim.x) {
        const unsigned int c = counts[i];
        if (c == 0) { ++zeros; continue; }                          // most nodes: aggregate per warp below
        const unsigned int b = min(c, (unsigned int)(nbins - 1));
        if (SMEM) atomicAdd(s_hist + b, 1u);
        else atomicAdd(hist + b, 1ULL);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) zeros += __shfl_down_sync(0xffffffffu, zeros, o);
    if ((threadIdx.x & 31) == 0 && zeros) atomicAdd(hist, zeros);
    if (SMEM) {
        __syncthreads();
        for (int b = threadIdx.x; b < nbins; b += blockDim.x)
            if (s_hist[b]) atomicAdd(hist + b, (unsigned long long)s_hist[b]);
    }
}

// one axis of scipy.ndimage.gaussian_filter(mode='constant', cval=0): out = sum_k w[k] in[.. + k - lw ..]
//   AXIS 0: along rows (y), AXIS 1: along columns (x).  FROM_COUNTS: input is counts * scale.
template <int AXIS, bool FROM_COUNTS>
__global__ void __launch_bounds__(256)
gaussian_axis_kernel(const unsigned int *counts, const double *in, double scale, int nrows, int ncols,
                     const double *__restrict__ w, int lw, double *out)
{
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    const int i = blockIdx.y;
    if (j >= ncols) return;
    double acc = 0.0;
    for (int k = -lw; k <= lw; ++k) {
        const int ii = AXIS == 0 ? i + k : i;
        const int jj = AXIS == 0 ? j : j + k;
        if (ii < 0 || ii >= nrows || jj < 0 || jj >= ncols) continue;   // cval = 0
        const size_t idx = (size_t)ii * ncols + jj;
        const double v = FROM_COUNTS ? (double)counts[idx] * scale : in[idx];
        acc = fma(w[k + lw], v, acc);
    }
    out[(size_t)i * ncols + j] = acc;
}

// ------------------------------------------------------------------------------------------
// Host helpers
// ------------------------------------------------------------------------------------------
static int check_model(const oneka_model_desc *m)
{
    if (!m) return fail(ONEKA_ERR_ARG, "model descriptor is NULL");
    if (m->nw < 0) return fail(ONEKA_ERR_ARG, "nw must be >= 0");
    if (!(m->tol > 0.0)) return fail(ONEKA_ERR_ARG, "tol must be > 0 (capturezone.py:139-142)");
    if (!(m->maxstep > 0.0)) return fail(ONEKA_ERR_ARG, "maxstep must be > 0 (capturezone.py:144-146)");
    if (!(m->duration == m->duration)) return fail(ONEKA_ERR_ARG, "duration is nan");
    return ONEKA_OK;
}

static int make_lattice(const oneka_lattice *lat, LatticeDev &L)
{
    if (!(lat->deltax > 0.0) || !(lat->deltay > 0.0))
        return fail(ONEKA_ERR_ARG, "<deltax>, <deltay> must be > 0 (probabilityfield.py:127-131)");
    if (lat->nrows <= 0 || lat->ncols <= 0) return fail(ONEKA_ERR_ARG, "lattice must have nrows, ncols > 0");
    if (!(lat->umbra >= 0.0)) return fail(ONEKA_ERR_ARG, "umbra must be >= 0");
    L.xmin = lat->xmin; L.ymin = lat->ymin; L.dx = lat->deltax; L.dy = lat->deltay;
    L.nrows = lat->nrows; L.ncols = lat->ncols;
    L.wpr = ((lat->ncols + 63) / 64) * 2;                        // an even number of words: every bitmap row starts on 8 bytes (raster_seg<true>'s 64-bit bit-sets)
    L.umbra = lat->umbra;
    L.umbra2 = lat->umbra * lat->umbra;
    L.dx32 = (float)lat->deltax; L.dy32 = (float)lat->deltay; L.umbra2_32 = (float)L.umbra2;
    L.umbra32 = (float)lat->umbra; L.inv_dx32 = 1.0f / L.dx32;
    L.maxd = lat->deltax > lat->deltay ? lat->deltax : lat->deltay;
    L.inv_dx = 1.0 / lat->deltax; L.inv_dy = 1.0 / lat->deltay;
    L.cxl = lat->xmin + lat->umbra; L.cxr = lat->xmin - lat->umbra;
    L.cyb = lat->ymin + lat->umbra; L.cyt = lat->ymin - lat->umbra;
    L.s16x = 65536.0 / lat->deltax; L.s16y = 65536.0 / lat->deltay;
    L.fixed_ok = (lat->nrows < 30000 && lat->ncols < 30000) ? 1 : 0;
    L.words = (unsigned long long)L.nrows * (unsigned long long)L.wpr;
    return ONEKA_OK;
}

static TrackParams make_track(const oneka_model_desc *m, const double *well_xy_dev, long long R, int P,
                              const double *q, const double *cond, const double *poro, const double *thick,
                              const double *coef, const double *start, unsigned long long *stats)
{
    TrackParams tp;
    memset(&tp, 0, sizeof(tp));
    tp.nw = m->nw; tp.P = P; tp.R = R;
    tp.duration = m->duration; tp.tol = m->tol; tp.maxstep = m->maxstep;
    long long ma = m->max_attempts > 0 ? m->max_attempts : ((long long)1 << 22);
    tp.max_attempts = (int)(ma > 0x7fffffffLL ? 0x7fffffffLL : ma);
    tp.xo = m->xo; tp.yo = m->yo;
    tp.well_xy = well_xy_dev;
    tp.q = q; tp.cond = cond; tp.poro = poro; tp.thick = thick; tp.coef = coef; tp.start_xy = start;
    tp.stats = stats;
    return tp;
}

// the well store of oneka_device.cuh (blocks of 4 wells, WELL_BLK doubles each) + 16 bytes of slack
static size_t track_smem(int nw) { return (size_t)well_store_doubles(nw) * sizeof(double) + 256; }   // slack: the loops load the first well of the block after the last

static int ensure_aux(oneka_ctx *ctx, size_t bytes)
{
    if (bytes <= ctx->aux_bytes) return ONEKA_OK;
    if (ctx->aux) { CUDA_TRY(cudaStreamSynchronize(ctx->stream)); CUDA_TRY(cudaFree(ctx->aux)); ctx->aux = nullptr; ctx->aux_bytes = 0; }
    const size_t want = bytes < 65536 ? 65536 : bytes;
    cudaError_t e = cudaMalloc(&ctx->aux, want);
    if (e != cudaSuccess) { cudaGetLastError(); return fail(ONEKA_ERR_NOMEM, "cudaMalloc(%zu) for scratch failed: %s", want, cudaGetErrorString(e)); }
    ctx->aux_bytes = want;
    return ONEKA_OK;
}

static int ensure_bitmaps(oneka_ctx *ctx, size_t bytes)
{
    if (bytes <= ctx->bitmap_bytes) return ONEKA_OK;
    if (ctx->bitmaps) { CUDA_TRY(cudaStreamSynchronize(ctx->stream)); CUDA_TRY(cudaFree(ctx->bitmaps)); ctx->bitmaps = nullptr; ctx->bitmap_bytes = 0; }
    cudaError_t e = cudaMalloc(&ctx->bitmaps, bytes);
    if (e != cudaSuccess) { cudaGetLastError(); return fail(ONEKA_ERR_NOMEM, "cudaMalloc(%zu) for registration bitmaps failed: %s", bytes, cudaGetErrorString(e)); }
    ctx->bitmap_bytes = bytes;
    CUDA_TRY(cudaMemsetAsync(ctx->bitmaps, 0, bytes, ctx->stream));
    return ONEKA_OK;
}

static void prof_begin(oneka_ctx *ctx, int kind)
{
    if (!ctx->profiling) return;
    oneka_ctx::EvPair ev;
    cudaEventCreate(&ev.a); cudaEventCreate(&ev.b); ev.kind = kind;
    cudaEventRecord(ev.a, ctx->stream);
    ctx->events.push_back(ev);
}
static void prof_end(oneka_ctx *ctx)
{
    if (!ctx->profiling) return;
    cudaEventRecord(ctx->events.back().b, ctx->stream);
}

static size_t ff_smem(const FarFieldDev &ff)
{
    const size_t nt = (size_t)ff.ntx * ff.nty;
    return (nt * ff.order * sizeof(double2) + nt * ff.max_near * 4 + nt * 2 + 15) & ~(size_t)15;
}

static size_t ff_smem_unc(const FarFieldDev &ff)
{
    const size_t nt = (size_t)ff.ntx * ff.nty;
    return (nt * ff.order * (sizeof(double2) + sizeof(float2)) + nt * 8 + nt * ff.max_near * 2 + nt * 2 + 15) & ~(size_t)15;
}

// dynamic shared memory a far-field CTA may use so that FF_MIN_CTAS of them fit on an SM (1 KB per CTA is reserved by the
// system; the kernel's static shared memory is ~200 B)
static size_t ff_smem_budget(const oneka_ctx *ctx, int min_ctas) { return ctx->smem_per_sm / (size_t)min_ctas - 1024 - 512; }

// The rasteriser's flavour for this lattice: a segment's window spans about 2 umbra / deltay + 1 rows (+ its own rise).  From
// a certain number of rows on, the bit-set traffic to L2 costs more than the heavy flavour's extra instructions -- earlier in the
// direct-sum kernels (24 resident warps per SM keep more bit-sets in flight) than in the far-field kernels (16 warps, bound by
// instruction latency).  Measured on B200, profiles/r02_flavour_scan.txt: direct kernel 7 rows 0 %, 9 rows -6 %, 11 rows -13 %,
// 15 rows -20 %; far-field kernel 9 rows +1 %, 11 rows -1 %.  ctx->raster_mode 1 / 2 force plain / heavy (oneka_set_raster_mode).
constexpr double RASTER_HEAVY_ROWS = 8.0, RASTER_HEAVY_ROWS_FF = 11.0;
static int raster_flavour(const oneka_ctx *ctx, const LatticeDev &L, bool farfield)
{
    if (ctx->raster_mode == 1) return RF_PLAIN;
    if (ctx->raster_mode == 2) return RF_HEAVY;
    return (2.0 * L.umbra / L.dy + 1.0 >= (farfield ? RASTER_HEAVY_ROWS_FF : RASTER_HEAVY_ROWS)) ? RF_HEAVY : RF_PLAIN;
}

template <bool CONFINED, int MODE, bool FF, int ORD, int THREADS, int MIN_CTAS, int RF>
static int launch_one(oneka_ctx *ctx, const TrackParams &tp, const LatticeDev &L, unsigned int *bitmaps, const FarFieldDev &ff, size_t smem)
{
    const long long nblk = tp.R * ((tp.P + THREADS - 1) / THREADS);
    if (nblk > 0x7fffffffLL) return fail(ONEKA_ERR_ARG, "too many CTAs in one launch (%lld)", nblk);
    auto kern = track_kernel<CONFINED, MODE, FF, ORD, THREADS, MIN_CTAS, RF>;
    if (smem > 48 * 1024) CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    kern<<<(unsigned)nblk, THREADS, smem, ctx->stream>>>(tp, L, bitmaps, ff);
    return ONEKA_OK;
}

template <int MODE>
static int launch_track(oneka_ctx *ctx, const oneka_model_desc *m, const TrackParams &tp, const LatticeDev &L, unsigned int *bitmaps,
                        const FarFieldDev *ff = nullptr)
{
    if (tp.R <= 0) return ONEKA_OK;
    size_t smem = track_smem(tp.nw);
    if (ff) smem += m->confined ? ff_smem(*ff) : ff_smem_unc(*ff);
    if (smem > 200 * 1024) return fail(ONEKA_ERR_ARG, "nw = %d wells do not fit in shared memory", tp.nw);
    prof_begin(ctx, 0);
    FarFieldDev none;
    memset(&none, 0, sizeof(none));
    int rc;
    // only the fused kernels rasterise: MODE 0 / 2 have one flavour (HV collapses to RF_PLAIN for them)
    constexpr int HV = (MODE == 1) ? RF_HEAVY : RF_PLAIN;
    const int rf = (MODE == 1) ? raster_flavour(ctx, L, ff != nullptr) : RF_PLAIN;
#define ONEKA_LAUNCH(C, F, O, T, M, FFV) (rf == RF_HEAVY ? launch_one<C, MODE, F, O, T, M, HV>(ctx, tp, L, bitmaps, FFV, smem) \
                                                         : launch_one<C, MODE, F, O, T, M, RF_PLAIN>(ctx, tp, L, bitmaps, FFV, smem))
    if (m->confined && ff && ff->order == FF_ORDER_UNROLLED)
        rc = ONEKA_LAUNCH(true, true, FF_ORDER_UNROLLED, FF_THREADS, FF_MIN_CTAS, *ff);
    else if (m->confined && ff)
        rc = ONEKA_LAUNCH(true, true, 0, FF_THREADS, FF_MIN_CTAS, *ff);
    else if (m->confined)
        rc = ONEKA_LAUNCH(true, false, 0, TRACK_THREADS, TRACK_MIN_CTAS, none);
    else if (ff && ff->order == FF_ORDER_UNROLLED)                 // (unrolled here too: C3 unconfined 94.8 -> 84.5 ms, profiles/r02_ab_unc16.txt)
        rc = ONEKA_LAUNCH(false, true, FF_ORDER_UNROLLED, FF_UNC_THREADS, FF_UNC_MIN_CTAS, *ff);
    else if (ff)
        rc = ONEKA_LAUNCH(false, true, 0, FF_UNC_THREADS, FF_UNC_MIN_CTAS, *ff);
    else
        rc = ONEKA_LAUNCH(false, false, 0, TRACK_THREADS, TRACK_MIN_CTAS, none);
#undef ONEKA_LAUNCH
    if (rc) return rc;
    prof_end(ctx);
    ctx->launches++;
    CUDA_TRY(cudaGetLastError());
    return ONEKA_OK;
}

// Far-field coefficients for the realizations of one launch (rows already offset) -> ctx->ff.coef; fills `out`.
// Returns ONEKA_OK with use = false when the far field is off or does not apply to this model.
static int prepare_farfield(oneka_ctx *ctx, const oneka_model_desc *m, long long nr, const double *well_xy_dev, const double *q,
                            const double *poro, const double *thick, FarFieldDev &out, bool &use)
{
    const oneka_ctx::FarField &f = ctx->ff;
    use = f.on && (m->confined || f.unconfined) && m->nw == f.nw && m->xo == f.xo && m->yo == f.yo && nr > 0;
    if (!use) return ONEKA_OK;
    if (!m->confined && f.smem_unconfined > 200 * 1024) {
        use = false;                                                  // the unconfined tables (c64 + p32) do not fit: direct sums
        return ONEKA_OK;
    }
    const int ntiles = f.ntx * f.nty;
    const size_t need = (size_t)nr * ntiles * f.order * sizeof(double2);
    if (need > ctx->ff.coef_bytes) {
        if (ctx->ff.coef) { CUDA_TRY(cudaStreamSynchronize(ctx->stream)); CUDA_TRY(cudaFree(ctx->ff.coef)); ctx->ff.coef = nullptr; ctx->ff.coef_bytes = 0; }
        cudaError_t e = cudaMalloc(&ctx->ff.coef, need);
        if (e != cudaSuccess) { cudaGetLastError(); return fail(ONEKA_ERR_NOMEM, "cudaMalloc(%zu) for far-field coefficients failed: %s", need, cudaGetErrorString(e)); }
        ctx->ff.coef_bytes = need;
    }
    const long long nby = (nr + COEF_RB - 1) / COEF_RB;
    if (nby > 65535) return fail(ONEKA_ERR_ARG, "too many realizations in one far-field launch (%lld)", nr);
    const dim3 cgrid((unsigned)((ntiles * f.order + COEF_THREADS - 1) / COEF_THREADS), (unsigned)nby);
    const size_t csmem = 0;                                          // (static shared memory: COEF_RB x COEF_WCHUNK doubles)
    ff_check_wells_kernel<<<1, 128, 0, ctx->stream>>>(f.nw, f.wells, well_xy_dev, ctx->stats_dev);
    if (m->confined) {
        farfield_coef_kernel<true><<<cgrid, COEF_THREADS, csmem, ctx->stream>>>(f.nw, ntiles, f.order, nr, f.P, q, poro, thick, ctx->ff.coef);
    } else {
        const size_t nb0 = (size_t)nr * ntiles * sizeof(double);
        if (nb0 > ctx->ff.b0_bytes) {
            if (ctx->ff.b0) { CUDA_TRY(cudaStreamSynchronize(ctx->stream)); CUDA_TRY(cudaFree(ctx->ff.b0)); ctx->ff.b0 = nullptr; ctx->ff.b0_bytes = 0; }
            cudaError_t e = cudaMalloc(&ctx->ff.b0, nb0);
            if (e != cudaSuccess) { cudaGetLastError(); return fail(ONEKA_ERR_NOMEM, "cudaMalloc(%zu) for far-field b0 failed: %s", nb0, cudaGetErrorString(e)); }
            ctx->ff.b0_bytes = nb0;
        }
        farfield_coef_kernel<false><<<cgrid, COEF_THREADS, csmem, ctx->stream>>>(f.nw, ntiles, f.order, nr, f.P, q, poro, thick, ctx->ff.coef);
        farfield_b0_kernel<<<(unsigned)((nr * ntiles + 127) / 128), 128, 0, ctx->stream>>>(f.nw, ntiles, nr, f.Lg, q, ctx->ff.b0);
        ctx->launches++;
    }
    ctx->launches += 2;
    CUDA_TRY(cudaGetLastError());
    out.ntx = f.ntx; out.nty = f.nty; out.order = f.order; out.max_near = f.max_near;
    out.gx0 = f.gx0; out.gy0 = f.gy0; out.inv_tile = 1.0 / f.tile;
    out.coef = ctx->ff.coef; out.near_off = f.near_off; out.near_cnt = f.near_cnt;
    out.b0 = m->confined ? nullptr : ctx->ff.b0; out.near_idx = f.near_idx; out.near_raw = f.near_raw;
    return ONEKA_OK;
}

// realizations per launch so that the coefficient workspace stays below 2 GiB
static long long farfield_batch(const oneka_ctx *ctx, long long want)
{
    const oneka_ctx::FarField &f = ctx->ff;
    if (!f.on) return want;
    const size_t per = (size_t)f.ntx * f.nty * f.order * sizeof(double2);
    const long long cap = (long long)(((size_t)2 << 30) / per);
    return want < cap ? want : (cap < 1 ? 1 : cap);
}

static int launch_flush(oneka_ctx *ctx, const LatticeDev &L, long long nslots, unsigned int *counts,
                        const unsigned int *slot_flags = nullptr)
{
    const unsigned gx = (unsigned)((L.words + 255) / 256);
    // aim for >= 2 waves of 148 SMs x 8 CTAs
    long long want_y = (2LL * ctx->sm_count * 8 + gx - 1) / gx;
    if (want_y < 1) want_y = 1;
    if (want_y > nslots) want_y = nslots;
    if (want_y > 65535) want_y = 65535;
    const long long per_y = (nslots + want_y - 1) / want_y;
    const unsigned gy = (unsigned)((nslots + per_y - 1) / per_y);
    prof_begin(ctx, 1);
    flush_kernel<<<dim3(gx, gy), 256, 0, ctx->stream>>>(ctx->bitmaps, nslots, per_y, L, counts, slot_flags);
    prof_end(ctx);
    ctx->launches++;
    CUDA_TRY(cudaGetLastError());
    return ONEKA_OK;
}

static long long slots_for(const oneka_ctx *ctx, const LatticeDev &L, long long want)
{
    const size_t per = (size_t)L.words * sizeof(unsigned int);
    long long cap = (long long)(ctx->workspace_limit / per);
    return want < cap ? want : cap;
}

static int reset_stats_async(oneka_ctx *ctx)
{
    unsigned long long init[N_STATS];
    memset(init, 0, sizeof(init));
    init[STAT_XMIN] = ~0ULL; init[STAT_YMIN] = ~0ULL;    // atomicMin targets
    init[STAT_XMAX] = 0ULL; init[STAT_YMAX] = 0ULL;      // atomicMax targets
    // tiny pageable copy; cudaMemcpyAsync from a stack buffer is staged by the driver before return
    CUDA_TRY(cudaMemcpyAsync(ctx->stats_dev, init, sizeof(init), cudaMemcpyHostToDevice, ctx->stream));
    return ONEKA_OK;
}

static double undkey(unsigned long long k)
{
    unsigned long long b = (k >> 63) ? (k & 0x7fffffffffffffffULL) : ~k;
    double v;
    memcpy(&v, &b, 8);
    return v;
}


// ------------------------------------------------------------------------------------------
// NCCL, resolved at run time: the library must load (and single-GPU use must work) without it, and inside a PyTorch
// process the libnccl.so.2 torch already loaded has to be the one used (two NCCL copies in a process do not share
// their topology / proxy state).
struct nccl_api {
    ncclResult_t (*GetUniqueId)(ncclUniqueId *);
    ncclResult_t (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int);
    ncclResult_t (*CommDestroy)(ncclComm_t);
    ncclResult_t (*AllReduce)(const void *, void *, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t);
    const char *(*GetErrorString)(ncclResult_t);
};

static const nccl_api *nccl()
{
    static nccl_api api;
    static int state = 0;                                      // 0 = not tried, 1 = ok, -1 = unavailable
    if (state == 0) {
        void *h = nullptr;
        if (const char *env = getenv("ONEKA_NCCL_LIB")) h = dlopen(env, RTLD_NOW | RTLD_GLOBAL);
        if (!h) h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_NOLOAD);          // the copy already in the process (torch's)
        if (!h) h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
        if (!h) h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
        state = -1;
        if (h) {
            api.GetUniqueId = (decltype(api.GetUniqueId))dlsym(h, "ncclGetUniqueId");
            api.CommInitRank = (decltype(api.CommInitRank))dlsym(h, "ncclCommInitRank");
            api.CommDestroy = (decltype(api.CommDestroy))dlsym(h, "ncclCommDestroy");
            api.AllReduce = (decltype(api.AllReduce))dlsym(h, "ncclAllReduce");
            api.GetErrorString = (decltype(api.GetErrorString))dlsym(h, "ncclGetErrorString");
            if (api.GetUniqueId && api.CommInitRank && api.CommDestroy && api.AllReduce && api.GetErrorString) state = 1;
        }
    }
    return state == 1 ? &api : nullptr;
}

#define NCCL_TRY(N, expr)                                                                   \
    do {                                                                                    \
        ncclResult_t _r = (expr);                                                           \
        if (_r != ncclSuccess)                                                              \
            return fail(ONEKA_ERR_NCCL, "%s failed: %s (%s:%d)", #expr, (N)->GetErrorString(_r), __FILE__, __LINE__); \
    } while (0)

// ------------------------------------------------------------------------------------------
// C ABI
// ------------------------------------------------------------------------------------------
extern "C" {

const char *oneka_last_error(void) { return g_err; }
int oneka_abi_version(void) { return ONEKA_ABI_VERSION; }

oneka_ctx *oneka_create(int device)
{
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n == 0) {
        cudaGetLastError();
        fail(ONEKA_ERR_NODEVICE, "no CUDA device: %s (this library has no CPU fallback)", e != cudaSuccess ? cudaGetErrorString(e) : "device count is 0");
        return nullptr;
    }
    if (device < 0 || device >= n) { fail(ONEKA_ERR_ARG, "device %d out of range [0,%d)", device, n); return nullptr; }
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) { fail(ONEKA_ERR_CUDA, "cudaGetDeviceProperties failed"); return nullptr; }
    if (prop.major != 10) {
        fail(ONEKA_ERR_NODEVICE, "device %d is sm_%d%d; this library is built for sm_100a (B200) only", device, prop.major, prop.minor);
        return nullptr;
    }
    if (cudaSetDevice(device) != cudaSuccess) { fail(ONEKA_ERR_CUDA, "cudaSetDevice(%d) failed", device); return nullptr; }
    oneka_ctx *ctx = new (std::nothrow) oneka_ctx();
    if (!ctx) { fail(ONEKA_ERR_NOMEM, "out of host memory"); return nullptr; }
    ctx->device = device;
    ctx->sm_count = prop.multiProcessorCount;
    ctx->smem_per_sm = prop.sharedMemPerMultiprocessor;
    if (cudaMalloc(&ctx->stats_dev, N_STATS * sizeof(unsigned long long)) != cudaSuccess) {
        fail(ONEKA_ERR_NOMEM, "cudaMalloc(stats) failed"); delete ctx; return nullptr;
    }
    if (reset_stats_async(ctx) != ONEKA_OK) { cudaFree(ctx->stats_dev); delete ctx; return nullptr; }
    cudaStreamSynchronize(ctx->stream);
    g_err[0] = 0;
    return ctx;
}

void oneka_destroy(oneka_ctx *ctx)
{
    if (!ctx) return;
    cudaSetDevice(ctx->device);
    cudaStreamSynchronize(ctx->stream);
    for (auto &ev : ctx->events) { cudaEventDestroy(ev.a); cudaEventDestroy(ev.b); }
    if (ctx->bitmaps) cudaFree(ctx->bitmaps);
    if (ctx->stage) cudaFree(ctx->stage);
    if (ctx->aux) cudaFree(ctx->aux);
    if (ctx->stats_dev) cudaFree(ctx->stats_dev);
    if (ctx->ff.P) cudaFree(ctx->ff.P);
    if (ctx->ff.near_off) cudaFree(ctx->ff.near_off);
    if (ctx->ff.near_cnt) cudaFree(ctx->ff.near_cnt);
    if (ctx->ff.coef) cudaFree(ctx->ff.coef);
    if (ctx->ff.Lg) cudaFree(ctx->ff.Lg);
    if (ctx->ff.near_idx) cudaFree(ctx->ff.near_idx);
    if (ctx->ff.near_raw) cudaFree(ctx->ff.near_raw);
    if (ctx->ff.b0) cudaFree(ctx->ff.b0);
    if (ctx->ff.wells) cudaFree(ctx->ff.wells);
    if (ctx->comm && ctx->comm_owned) { if (const nccl_api *N = nccl()) N->CommDestroy(ctx->comm); }
    delete ctx;
}

int oneka_set_stream(oneka_ctx *ctx, void *cuda_stream)
{
    if (!ctx) return fail(ONEKA_ERR_ARG, "ctx is NULL");
    CUDA_TRY(cudaStreamSynchronize(ctx->stream));
    ctx->stream = (cudaStream_t)cuda_stream;
    return ONEKA_OK;
}

int oneka_set_raster_mode(oneka_ctx *ctx, int32_t mode)
{
    if (!ctx || mode < 0 || mode > 2) return fail(ONEKA_ERR_ARG, "oneka_set_raster_mode: mode must be 0 (by lattice), 1 (plain) or 2 (heavy)");
    ctx->raster_mode = mode;
    return ONEKA_OK;
}

int oneka_raster_flavour(const oneka_ctx *ctx, double umbra, double deltay, int32_t farfield)
{
    if (!ctx || !(deltay > 0.0) || !(umbra >= 0.0)) return fail(ONEKA_ERR_ARG, "bad argument to oneka_raster_flavour");
    LatticeDev L;
    memset(&L, 0, sizeof(L));
    L.umbra = umbra; L.dy = deltay;
    return raster_flavour(ctx, L, farfield != 0);
}

int oneka_set_workspace_limit(oneka_ctx *ctx, uint64_t bytes)
{
    if (!ctx) return fail(ONEKA_ERR_ARG, "ctx is NULL");
    ctx->workspace_limit = (size_t)bytes;
    return ONEKA_OK;
}

int oneka_synchronize(oneka_ctx *ctx)
{
    if (!ctx) return fail(ONEKA_ERR_ARG, "ctx is NULL");
    CUDA_TRY(cudaSetDevice(ctx->device));
    CUDA_TRY(cudaStreamSynchronize(ctx->stream));
    return ONEKA_OK;
}

uint64_t oneka_launch_count(const oneka_ctx *ctx) { return ctx ? ctx->launches : 0; }

int oneka_set_profiling(oneka_ctx *ctx, int enabled)
{
    if (!ctx) return fail(ONEKA_ERR_ARG, "ctx is NULL");
    ctx->profiling = enabled != 0;
    return ONEKA_OK;
}

int oneka_kernel_ms(oneka_ctx *ctx, double *track_ms, double *flush_ms, uint64_t *track_launches, int reset)
{
    if (!ctx) return fail(ONEKA_ERR_ARG, "ctx is NULL");
    CUDA_TRY(cudaSetDevice(ctx->device));
    CUDA_TRY(cudaStreamSynchronize(ctx->stream));
    for (auto &ev : ctx->events) {
        float ms = 0.f;
        CUDA_TRY(cudaEventElapsedTime(&ms, ev.a, ev.b));
        if (ev.kind == 0) { ctx->track_ms += ms; ctx->track_launches++; }
        else ctx->flush_ms += ms;
        cudaEventDestroy(ev.a); cudaEventDestroy(ev.b);
    }
    ctx->events.clear();
    if (track_ms) *track_ms = ctx->track_ms;
    if (flush_ms) *flush_ms = ctx->flush_ms;
    if (track_launches) *track_launches = ctx->track_launches;
    if (reset) { ctx->track_ms = ctx->flush_ms = 0.0; ctx->track_launches = 0; }
    return ONEKA_OK;
}

int oneka_set_farfield(oneka_ctx *ctx, int32_t nw, const double *well_xy_host, double xo, double yo,
                       double x0, double y0, double tile, int32_t ntx, int32_t nty, int32_t order, double eta,
                       int32_t order_fp64, int32_t *max_near_out, double *mean_near_out)
{
    if (!ctx) return fail(ONEKA_ERR_ARG, "ctx is NULL");
    CUDA_TRY(cudaSetDevice(ctx->device));
    oneka_ctx::FarField &f = ctx->ff;
    CUDA_TRY(cudaStreamSynchronize(ctx->stream));                      // launches in flight may still read the old tables
    f.on = false;
    if (f.P) { cudaFree(f.P); f.P = nullptr; }
    if (f.near_off) { cudaFree(f.near_off); f.near_off = nullptr; }
    if (f.near_cnt) { cudaFree(f.near_cnt); f.near_cnt = nullptr; }
    if (f.Lg) { cudaFree(f.Lg); f.Lg = nullptr; }
    if (f.near_idx) { cudaFree(f.near_idx); f.near_idx = nullptr; }
    if (f.near_raw) { cudaFree(f.near_raw); f.near_raw = nullptr; }
    if (f.wells) { cudaFree(f.wells); f.wells = nullptr; }
    if (nw <= 0 || order <= 0) return ONEKA_OK;                        // switched off
    if (order < 2 || (order & 1)) return fail(ONEKA_ERR_ARG, "far field: order must be even and >= 2 (two interleaved Horner chains)");
    (void)order_fp64;                                                  // (ABI slot of the dropped FP32 tail)
    FFTables T;
    if (const char *why = build_ff_tables(nw, well_xy_host, xo, yo, x0 - xo, y0 - yo, tile, ntx, nty, order, eta, T))
        return fail(ONEKA_ERR_ARG, "%s", why);
    FarFieldDev probe;
    memset(&probe, 0, sizeof(probe));
    probe.ntx = ntx; probe.nty = nty; probe.order = order; probe.max_near = T.max_near;
    if (track_smem(nw) + ff_smem(probe) > 200 * 1024)
        return fail(ONEKA_ERR_ARG, "far field: %d tiles x order %d do not fit in shared memory", T.ntiles, order);
    CUDA_TRY(cudaMalloc(&f.P, T.P.size() * sizeof(double2)));
    CUDA_TRY(cudaMalloc(&f.near_off, T.off.size() * sizeof(unsigned int)));
    CUDA_TRY(cudaMalloc(&f.near_cnt, T.cnt.size() * sizeof(unsigned short)));
    CUDA_TRY(cudaMemcpy(f.P, T.P.data(), T.P.size() * sizeof(double2), cudaMemcpyHostToDevice));
    CUDA_TRY(cudaMemcpy(f.near_off, T.off.data(), T.off.size() * sizeof(unsigned int), cudaMemcpyHostToDevice));
    CUDA_TRY(cudaMemcpy(f.near_cnt, T.cnt.data(), T.cnt.size() * sizeof(unsigned short), cudaMemcpyHostToDevice));
    CUDA_TRY(cudaMalloc(&f.Lg, T.Lg.size() * sizeof(double)));
    CUDA_TRY(cudaMalloc(&f.near_idx, T.idx.size() * sizeof(unsigned short)));
    CUDA_TRY(cudaMalloc(&f.near_raw, T.cnt_raw.size() * sizeof(unsigned short)));
    CUDA_TRY(cudaMemcpy(f.Lg, T.Lg.data(), T.Lg.size() * sizeof(double), cudaMemcpyHostToDevice));
    CUDA_TRY(cudaMemcpy(f.near_idx, T.idx.data(), T.idx.size() * sizeof(unsigned short), cudaMemcpyHostToDevice));
    CUDA_TRY(cudaMemcpy(f.near_raw, T.cnt_raw.data(), T.cnt_raw.size() * sizeof(unsigned short), cudaMemcpyHostToDevice));
    CUDA_TRY(cudaMalloc(&f.wells, (size_t)nw * 2 * sizeof(double)));
    CUDA_TRY(cudaMemcpy(f.wells, well_xy_host, (size_t)nw * 2 * sizeof(double), cudaMemcpyHostToDevice));
    f.nw = nw; f.ntx = ntx; f.nty = nty; f.order = order; f.max_near = T.max_near;
    f.smem_confined = track_smem(nw) + ff_smem(probe);
    f.smem_unconfined = track_smem(nw) + ff_smem_unc(probe);
    f.xo = xo; f.yo = yo; f.gx0 = x0 - xo; f.gy0 = y0 - yo; f.tile = tile; f.eta = eta; f.mean_near = T.mean_near;
    f.on = true;
    if (max_near_out) *max_near_out = T.max_near;
    if (mean_near_out) *mean_near_out = T.mean_near;
    return ONEKA_OK;
}

int oneka_farfield_info(oneka_ctx *ctx, int32_t *ntiles, int32_t *order, int32_t *max_near, uint64_t *smem_confined,
                        uint64_t *smem_unconfined, uint64_t *smem_budget)
{
    if (!ctx) return fail(ONEKA_ERR_ARG, "ctx is NULL");
    const oneka_ctx::FarField &f = ctx->ff;
    if (ntiles) *ntiles = f.on ? f.ntx * f.nty : 0;
    if (order) *order = f.on ? f.order : 0;
    if (max_near) *max_near = f.on ? f.max_near : 0;
    if (smem_confined) *smem_confined = f.on ? f.smem_confined : 0;
    if (smem_unconfined) *smem_unconfined = f.on ? f.smem_unconfined : 0;
    if (smem_budget) *smem_budget = ff_smem_budget(ctx, FF_MIN_CTAS);
    return ONEKA_OK;
}

int oneka_set_farfield_unconfined(oneka_ctx *ctx, int enabled)
{
    if (!ctx) return fail(ONEKA_ERR_ARG, "ctx is NULL");
    ctx->ff.unconfined = enabled != 0;
    return ONEKA_OK;
}

int oneka_farfield_eval_host(int32_t nw, const double *well_xy_host, const double *w_host, double xo, double yo,
                             double x0, double y0, double tile, int32_t ntx, int32_t nty, int32_t order, double eta,
                             int32_t order_fp64, int64_t npts, const double *pts_host, double *out_host, int32_t *near_count_out)
{
    if (npts < 0 || !w_host || (npts && (!pts_host || !out_host))) return fail(ONEKA_ERR_ARG, "bad argument to oneka_farfield_eval_host");
    FFTables T;
    if (const char *why = build_ff_tables(nw, well_xy_host, xo, yo, x0 - xo, y0 - yo, tile, ntx, nty, order, eta, T))
        return fail(ONEKA_ERR_ARG, "%s", why);
    std::vector<double2> coef;
    ff_host_coefficients(T, nw, order, w_host, coef);
    if (order < 2 || (order & 1)) return fail(ONEKA_ERR_ARG, "far field: order must be even and >= 2");
    (void)order_fp64;
    const double gx0 = x0 - xo, gy0 = y0 - yo, inv_tile = 1.0 / tile;
    for (int64_t i = 0; i < npts; ++i) {
        const double dx0 = pts_host[2 * i] - xo, dy0 = pts_host[2 * i + 1] - yo;
        int tile_i;
        double zr, zi, gx = 0.0, gy = 0.0;
        auto direct = [&](int w) {
            const double dx = pts_host[2 * i] - well_xy_host[2 * w], dy = pts_host[2 * i + 1] - well_xy_host[2 * w + 1];
            const double r2 = dx * dx + dy * dy;
            gx += w_host[w] * dx / r2;
            gy += w_host[w] * dy / r2;
        };
        if (!ff_locate(ntx, nty, gx0, gy0, inv_tile, dx0, dy0, tile_i, zr, zi)) {
            for (int w = 0; w < nw; ++w) direct(w);
            if (near_count_out) near_count_out[i] = -1;
        } else {
            for (int j = T.near_begin[tile_i]; j < T.near_begin[tile_i + 1]; ++j) direct(T.near_flat[j]);
            double re, im;
            if (order == FF_ORDER_UNROLLED) ff_poly_eval<FF_ORDER_UNROLLED>(coef.data() + (size_t)tile_i * order, order, zr, zi, re, im);
            else ff_poly_eval<0>(coef.data() + (size_t)tile_i * order, order, zr, zi, re, im);
            gx += re;
            gy -= im;
            if (near_count_out) near_count_out[i] = T.near_begin[tile_i + 1] - T.near_begin[tile_i];
        }
        out_host[2 * i] = gx;
        out_host[2 * i + 1] = gy;
    }
    return ONEKA_OK;
}

int oneka_reset_stats(oneka_ctx *ctx)
{
    if (!ctx) return fail(ONEKA_ERR_ARG, "ctx is NULL");
    CUDA_TRY(cudaSetDevice(ctx->device));
    return reset_stats_async(ctx);
}

int oneka_read_stats(oneka_ctx *ctx, oneka_stats *out)
{
    if (!ctx || !out) return fail(ONEKA_ERR_ARG, "NULL argument");
    CUDA_TRY(cudaSetDevice(ctx->device));
    unsigned long long h[N_STATS];
    CUDA_TRY(cudaMemcpyAsync(h, ctx->stats_dev, sizeof(h), cudaMemcpyDeviceToHost, ctx->stream));
    CUDA_TRY(cudaStreamSynchronize(ctx->stream));
    if (h[STAT_FF_MISMATCH])
        return fail(ONEKA_ERR_ARG, "%llu launch(es) since the last oneka_reset_stats used far-field tables built for OTHER well coordinates than "
                                   "the well_xy_dev they were given: call oneka_set_farfield again (or switch it off) when the wells change; "
                                   "the results of those launches are invalid", (unsigned long long)h[STAT_FF_MISMATCH]);
    out->attempts = h[STAT_ATTEMPTS]; out->steps = h[STAT_STEPS]; out->paths = h[STAT_PATHS];
    out->n_not_ok = h[STAT_NOT_OK]; out->n_clipped = h[STAT_CLIPPED]; out->exact_tests = h[STAT_EXACT];
    if (h[STAT_PATHS] == 0) {
        out->bbox[0] = out->bbox[2] = INFINITY; out->bbox[1] = out->bbox[3] = -INFINITY;
    } else {
        out->bbox[0] = undkey(h[STAT_XMIN]); out->bbox[1] = undkey(h[STAT_XMAX]);
        out->bbox[2] = undkey(h[STAT_YMIN]); out->bbox[3] = undkey(h[STAT_YMAX]);
    }
    return ONEKA_OK;
}

int oneka_eval_points_host(oneka_ctx *ctx, const oneka_model_desc *m, const double *well_xy_host,
                           const double *q_host, double cond, double poro, double thick,
                           const double *coef_host, int64_t npts, const double *pts_host, double *out_host)
{
    if (!ctx) return fail(ONEKA_ERR_ARG, "ctx is NULL");
    if (!m || m->nw < 0 || npts < 0 || !coef_host || (npts && (!pts_host || !out_host)) || (m->nw && (!well_xy_host || !q_host)))
        return fail(ONEKA_ERR_ARG, "bad argument to oneka_eval_points_host");
    if (npts == 0) return ONEKA_OK;
    CUDA_TRY(cudaSetDevice(ctx->device));
    const int nw = m->nw;
    const size_t nd = (size_t)3 * nw + 3 + 6 + 2 * (size_t)npts + 8 * (size_t)npts;
    int rc = ensure_aux(ctx, nd * sizeof(double));
    if (rc) return rc;
    double *buf = (double *)ctx->aux;
    double *d_wxy = buf, *d_q = d_wxy + 2 * nw, *d_k = d_q + nw, *d_n = d_k + 1, *d_H = d_n + 1, *d_cf = d_H + 1;
    double *d_pts = d_cf + 6, *d_out = d_pts + 2 * npts;
    cudaStream_t s = ctx->stream;
    do {
#define TRY2(expr) if ((expr) != cudaSuccess) { rc = fail(ONEKA_ERR_CUDA, "%s failed: %s", #expr, cudaGetErrorString(cudaGetLastError())); break; }
        if (nw) { TRY2(cudaMemcpyAsync(d_wxy, well_xy_host, 2 * nw * sizeof(double), cudaMemcpyHostToDevice, s));
                  TRY2(cudaMemcpyAsync(d_q, q_host, nw * sizeof(double), cudaMemcpyHostToDevice, s)); }
        const double sc[3] = {cond, poro, thick};
        TRY2(cudaMemcpyAsync(d_k, sc, 3 * sizeof(double), cudaMemcpyHostToDevice, s));
        TRY2(cudaMemcpyAsync(d_cf, coef_host, 6 * sizeof(double), cudaMemcpyHostToDevice, s));
        TRY2(cudaMemcpyAsync(d_pts, pts_host, 2 * npts * sizeof(double), cudaMemcpyHostToDevice, s));
        oneka_model_desc mm = *m;
        if (!(mm.tol > 0)) mm.tol = 1.0;
        if (!(mm.maxstep > 0)) mm.maxstep = 1.0;
        TrackParams tp = make_track(&mm, d_wxy, 1, 1, d_q, d_k, d_n, d_H, d_cf, nullptr, ctx->stats_dev);
        const size_t smem = 2 * track_smem(nw);
        if (smem > 48 * 1024) TRY2(cudaFuncSetAttribute(eval_points_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        eval_points_kernel<<<1, 128, smem, s>>>(tp, npts, d_pts, d_out);
        ctx->launches++;
        TRY2(cudaGetLastError());
        TRY2(cudaMemcpyAsync(out_host, d_out, 8 * npts * sizeof(double), cudaMemcpyDeviceToHost, s));
        TRY2(cudaStreamSynchronize(s));
#undef TRY2
    } while (0);
    return rc;
}

int oneka_trace(oneka_ctx *ctx, const oneka_model_desc *m, const double *well_xy_dev,
                int64_t R, int32_t P,
                const double *q_dev, const double *cond_dev, const double *poro_dev, const double *thick_dev,
                const double *coef_dev, const double *start_xy_dev,
                int32_t max_verts, double *verts_dev, int32_t *nverts_dev, uint8_t *status_dev, int32_t *attempts_dev)
{
    if (!ctx) return fail(ONEKA_ERR_ARG, "ctx is NULL");
    int rc = check_model(m);
    if (rc) return rc;
    if (R < 0 || P <= 0 || max_verts < 1 || !verts_dev) return fail(ONEKA_ERR_ARG, "bad R/P/max_verts/verts");
    if (R == 0) return ONEKA_OK;
    if (!cond_dev || !poro_dev || !thick_dev || !coef_dev || !start_xy_dev || (m->nw && (!q_dev || !well_xy_dev)))
        return fail(ONEKA_ERR_ARG, "NULL parameter array");
    CUDA_TRY(cudaSetDevice(ctx->device));
    TrackParams tp = make_track(m, well_xy_dev, R, P, q_dev, cond_dev, poro_dev, thick_dev, coef_dev, start_xy_dev, ctx->stats_dev);
    tp.verts = verts_dev; tp.max_verts = max_verts;
    tp.nverts = nverts_dev; tp.status = status_dev; tp.attempts = attempts_dev;
    LatticeDev L;
    memset(&L, 0, sizeof(L));
    if (farfield_batch(ctx, R) < R) return launch_track<2>(ctx, m, tp, L, nullptr);     // test hook: too many rows for one table
    FarFieldDev ff;
    bool use_ff = false;
    rc = prepare_farfield(ctx, m, R, well_xy_dev, q_dev, poro_dev, thick_dev, ff, use_ff);
    if (rc) return rc;
    return launch_track<2>(ctx, m, tp, L, nullptr, use_ff ? &ff : nullptr);
}

int oneka_raster_traces(oneka_ctx *ctx, const oneka_lattice *lat, int64_t ntraces,
                        const int64_t *offsets_dev, const double *verts_dev, const int32_t *real_of_dev,
                        int64_t nreal, uint32_t *counts_dev)
{
    if (!ctx || !lat) return fail(ONEKA_ERR_ARG, "NULL argument");
    if (ntraces < 0 || nreal < 0) return fail(ONEKA_ERR_ARG, "negative count");
    LatticeDev L;
    int rc = make_lattice(lat, L);
    if (rc) return rc;
    if (ntraces == 0 || nreal == 0) return ONEKA_OK;
    if (!offsets_dev || !verts_dev || !real_of_dev || !counts_dev) return fail(ONEKA_ERR_ARG, "NULL array");
    CUDA_TRY(cudaSetDevice(ctx->device));
    const long long slots = slots_for(ctx, L, nreal);
    if (slots < 1) return fail(ONEKA_ERR_NOMEM, "workspace limit %zu B is smaller than one registration bitmap (%llu B)",
                               ctx->workspace_limit, (unsigned long long)(L.words * 4));
    rc = ensure_bitmaps(ctx, (size_t)slots * L.words * sizeof(unsigned int));
    if (rc) return rc;
    for (long long s0 = 0; s0 < nreal; s0 += slots) {
        const long long s1 = (s0 + slots < nreal) ? s0 + slots : nreal;
        const unsigned nb = (unsigned)((ntraces + 127) / 128);
        if (raster_flavour(ctx, L, false) == RF_HEAVY)
            raster_traces_kernel<RF_HEAVY><<<nb, 128, 0, ctx->stream>>>(L, ntraces, (const long long *)offsets_dev, verts_dev, real_of_dev, s0, s1, ctx->bitmaps, ctx->stats_dev);
        else
            raster_traces_kernel<RF_PLAIN><<<nb, 128, 0, ctx->stream>>>(L, ntraces, (const long long *)offsets_dev, verts_dev, real_of_dev, s0, s1, ctx->bitmaps, ctx->stats_dev);
        ctx->launches++;
        CUDA_TRY(cudaGetLastError());
        rc = launch_flush(ctx, L, s1 - s0, counts_dev);
        if (rc) return rc;
    }
    return ONEKA_OK;
}

static int capture_impl(oneka_ctx *ctx, const oneka_model_desc *m, const oneka_lattice *lat,
                  const double *well_xy_dev, int64_t R, int32_t P,
                  const double *q_dev, const double *cond_dev, const double *poro_dev, const double *thick_dev,
                  const double *coef_dev, const double *start_xy_dev,
                  uint32_t *counts_dev, double *end_xy_dev, int32_t *nverts_dev, uint8_t *status_dev,
                  const int32_t *clip_dev, double *path_bbox_dev, uint32_t *flags_dev = nullptr)
{
    if (!ctx) return fail(ONEKA_ERR_ARG, "ctx is NULL");
    int rc = check_model(m);
    if (rc) return rc;
    if (R < 0 || P <= 0) return fail(ONEKA_ERR_ARG, "R must be >= 0 and P > 0 (capturezone.py:68-70)");
    if (R == 0) return ONEKA_OK;
    if (!cond_dev || !poro_dev || !thick_dev || !coef_dev || !start_xy_dev || (m->nw && (!q_dev || !well_xy_dev)))
        return fail(ONEKA_ERR_ARG, "NULL parameter array");
    CUDA_TRY(cudaSetDevice(ctx->device));
    const bool raster = lat != nullptr && counts_dev != nullptr;
    LatticeDev L;
    memset(&L, 0, sizeof(L));
    long long slots = R;
    if (raster) {
        rc = make_lattice(lat, L);
        if (rc) return rc;
        slots = slots_for(ctx, L, R);
        if (slots < 1) return fail(ONEKA_ERR_NOMEM, "workspace limit %zu B is smaller than one registration bitmap (%llu B)",
                                   ctx->workspace_limit, (unsigned long long)(L.words * 4));
        rc = ensure_bitmaps(ctx, (size_t)slots * L.words * sizeof(unsigned int));
        if (rc) return rc;
    }
    // keep each launch below 2^31 CTAs
    const int chunks = (P + TRACK_THREADS - 1) / TRACK_THREADS;      // (the smallest CTA shape: the most CTAs)
    const long long max_r = 0x7fffffffLL / chunks;
    if (slots > max_r) slots = max_r;
    slots = farfield_batch(ctx, slots);
    for (long long r0 = 0; r0 < R; r0 += slots) {
        const long long nr = (r0 + slots < R) ? slots : R - r0;
        TrackParams tp = make_track(m, well_xy_dev, nr, P, q_dev + (size_t)r0 * m->nw, cond_dev + r0, poro_dev + r0,
                                    thick_dev + r0, coef_dev + 6 * r0, start_xy_dev, ctx->stats_dev);
        tp.end_xy = end_xy_dev ? end_xy_dev + 2 * (size_t)r0 * P : nullptr;
        tp.nverts = nverts_dev ? nverts_dev + (size_t)r0 * P : nullptr;
        tp.status = status_dev ? status_dev + (size_t)r0 * P : nullptr;
        tp.clip = clip_dev ? clip_dev + 4 * (size_t)r0 * P : nullptr;
        tp.path_bbox = path_bbox_dev ? path_bbox_dev + 4 * (size_t)r0 * P : nullptr;
        tp.slot_flags = (raster && flags_dev) ? flags_dev + r0 : nullptr;
        FarFieldDev ff;
        bool use_ff = false;
        rc = prepare_farfield(ctx, m, nr, well_xy_dev, tp.q, tp.poro, tp.thick, ff, use_ff);
        if (rc) return rc;
        if (raster) {
            rc = launch_track<1>(ctx, m, tp, L, ctx->bitmaps, use_ff ? &ff : nullptr);
            if (rc) return rc;
            rc = launch_flush(ctx, L, nr, counts_dev, tp.slot_flags);
            if (rc) return rc;
        } else {
            rc = launch_track<0>(ctx, m, tp, L, nullptr, use_ff ? &ff : nullptr);
            if (rc) return rc;
        }
    }
    return ONEKA_OK;
}

int oneka_capture(oneka_ctx *ctx, const oneka_model_desc *m, const oneka_lattice *lat,
                  const double *well_xy_dev, int64_t R, int32_t P,
                  const double *q_dev, const double *cond_dev, const double *poro_dev, const double *thick_dev,
                  const double *coef_dev, const double *start_xy_dev,
                  uint32_t *counts_dev, double *end_xy_dev, int32_t *nverts_dev, uint8_t *status_dev)
{
    return capture_impl(ctx, m, lat, well_xy_dev, R, P, q_dev, cond_dev, poro_dev, thick_dev, coef_dev, start_xy_dev,
                        counts_dev, end_xy_dev, nverts_dev, status_dev, nullptr, nullptr);
}

int oneka_capture_guarded(oneka_ctx *ctx, const oneka_model_desc *m, const oneka_lattice *lat,
                          const double *well_xy_dev, int64_t R, int32_t P,
                          const double *q_dev, const double *cond_dev, const double *poro_dev, const double *thick_dev,
                          const double *coef_dev, const double *start_xy_dev,
                          uint32_t *counts_dev, double *end_xy_dev, int32_t *nverts_dev, uint8_t *status_dev,
                          uint32_t *clipped_dev)
{
    if (!ctx) return fail(ONEKA_ERR_ARG, "ctx is NULL");
    if (R == 0) return ONEKA_OK;
    if (!clipped_dev || !lat || !counts_dev) return fail(ONEKA_ERR_ARG, "oneka_capture_guarded needs clipped_dev, lat and counts_dev");
    CUDA_TRY(cudaSetDevice(ctx->device));
    CUDA_TRY(cudaMemsetAsync(clipped_dev, 0, (size_t)R * sizeof(uint32_t), ctx->stream));
    return capture_impl(ctx, m, lat, well_xy_dev, R, P, q_dev, cond_dev, poro_dev, thick_dev, coef_dev, start_xy_dev,
                        counts_dev, end_xy_dev, nverts_dev, status_dev, nullptr, nullptr, clipped_dev);
}

int oneka_path_bboxes(oneka_ctx *ctx, const oneka_model_desc *m, const double *well_xy_dev, int64_t R, int32_t P,
                      const double *q_dev, const double *cond_dev, const double *poro_dev, const double *thick_dev,
                      const double *coef_dev, const double *start_xy_dev, double *bbox_dev, uint8_t *status_dev)
{
    if (R == 0) return ONEKA_OK;
    if (!bbox_dev) return fail(ONEKA_ERR_ARG, "bbox_dev is NULL");
    return capture_impl(ctx, m, nullptr, well_xy_dev, R, P, q_dev, cond_dev, poro_dev, thick_dev, coef_dev, start_xy_dev,
                        nullptr, nullptr, nullptr, status_dev, nullptr, bbox_dev);
}

int oneka_capture_clipped(oneka_ctx *ctx, const oneka_model_desc *m, const oneka_lattice *lat,
                          const double *well_xy_dev, int64_t R, int32_t P,
                          const double *q_dev, const double *cond_dev, const double *poro_dev, const double *thick_dev,
                          const double *coef_dev, const double *start_xy_dev, const int32_t *clip_dev,
                          uint32_t *counts_dev, double *end_xy_dev, int32_t *nverts_dev, uint8_t *status_dev)
{
    if (R == 0) return ONEKA_OK;
    if (!clip_dev || !lat || !counts_dev) return fail(ONEKA_ERR_ARG, "oneka_capture_clipped needs clip_dev, lat and counts_dev");
    if (((uintptr_t)clip_dev & 15) != 0) return fail(ONEKA_ERR_ARG, "clip_dev must be 16-byte aligned");
    return capture_impl(ctx, m, lat, well_xy_dev, R, P, q_dev, cond_dev, poro_dev, thick_dev, coef_dev, start_xy_dev,
                        counts_dev, end_xy_dev, nverts_dev, status_dev, clip_dev, nullptr);
}

int oneka_capture_host(oneka_ctx *ctx, const oneka_model_desc *m, const oneka_lattice *lat,
                       const double *well_xy_host, int64_t R, int32_t P,
                       const double *q_host, const double *cond_host, const double *poro_host, const double *thick_host,
                       const double *coef_host, const double *start_xy_host,
                       uint32_t *counts_host, double *end_xy_host, int32_t *nverts_host, uint8_t *status_host,
                       oneka_stats *stats_out)
{
    if (!ctx) return fail(ONEKA_ERR_ARG, "ctx is NULL");
    int rc = check_model(m);
    if (rc) return rc;
    if (R < 0 || P <= 0) return fail(ONEKA_ERR_ARG, "R must be >= 0 and P > 0");
    if (!cond_host || !poro_host || !thick_host || !coef_host || !start_xy_host || (m->nw && (!q_host || !well_xy_host)))
        return fail(ONEKA_ERR_ARG, "NULL parameter array");
    CUDA_TRY(cudaSetDevice(ctx->device));
    const bool raster = lat != nullptr && counts_host != nullptr;
    const size_t nw = (size_t)m->nw, RP = (size_t)R * P;
    const size_t ncell = raster ? (size_t)lat->nrows * lat->ncols : 0;
    // carve one staging allocation (8-byte aligned pieces first)
    size_t off = 0;
    auto carve = [&](size_t bytes) { size_t o = off; off += (bytes + 255) & ~(size_t)255; return o; };
    const size_t o_wxy = carve(2 * nw * 8), o_q = carve((size_t)R * nw * 8), o_k = carve(R * 8), o_n = carve(R * 8), o_H = carve(R * 8);
    const size_t o_cf = carve((size_t)R * 6 * 8), o_st = carve((size_t)P * 2 * 8);
    const size_t o_end = carve(end_xy_host ? RP * 16 : 0), o_nv = carve(nverts_host ? RP * 4 : 0), o_stt = carve(status_host ? RP : 0);
    const size_t o_cnt = carve(ncell * 4);
    if (off > ctx->stage_bytes) {
        if (ctx->stage) { CUDA_TRY(cudaStreamSynchronize(ctx->stream)); CUDA_TRY(cudaFree(ctx->stage)); ctx->stage = nullptr; ctx->stage_bytes = 0; }
        cudaError_t e = cudaMalloc(&ctx->stage, off);
        if (e != cudaSuccess) { cudaGetLastError(); return fail(ONEKA_ERR_NOMEM, "cudaMalloc(%zu) for staging failed", off); }
        ctx->stage_bytes = off;
    }
    char *base = (char *)ctx->stage;
    cudaStream_t s = ctx->stream;
    if (nw) {
        CUDA_TRY(cudaMemcpyAsync(base + o_wxy, well_xy_host, 2 * nw * 8, cudaMemcpyHostToDevice, s));
        if (R) CUDA_TRY(cudaMemcpyAsync(base + o_q, q_host, (size_t)R * nw * 8, cudaMemcpyHostToDevice, s));
    }
    if (R) {
        CUDA_TRY(cudaMemcpyAsync(base + o_k, cond_host, R * 8, cudaMemcpyHostToDevice, s));
        CUDA_TRY(cudaMemcpyAsync(base + o_n, poro_host, R * 8, cudaMemcpyHostToDevice, s));
        CUDA_TRY(cudaMemcpyAsync(base + o_H, thick_host, R * 8, cudaMemcpyHostToDevice, s));
        CUDA_TRY(cudaMemcpyAsync(base + o_cf, coef_host, (size_t)R * 48, cudaMemcpyHostToDevice, s));
    }
    CUDA_TRY(cudaMemcpyAsync(base + o_st, start_xy_host, (size_t)P * 16, cudaMemcpyHostToDevice, s));
    if (raster) CUDA_TRY(cudaMemsetAsync(base + o_cnt, 0, ncell * 4, s));
    if (stats_out) { rc = reset_stats_async(ctx); if (rc) return rc; }
    rc = oneka_capture(ctx, m, raster ? lat : nullptr, (const double *)(base + o_wxy), R, P,
                       (const double *)(base + o_q), (const double *)(base + o_k), (const double *)(base + o_n),
                       (const double *)(base + o_H), (const double *)(base + o_cf), (const double *)(base + o_st),
                       raster ? (uint32_t *)(base + o_cnt) : nullptr,
                       end_xy_host ? (double *)(base + o_end) : nullptr,
                       nverts_host ? (int32_t *)(base + o_nv) : nullptr,
                       status_host ? (uint8_t *)(base + o_stt) : nullptr);
    if (rc) return rc;
    if (raster) CUDA_TRY(cudaMemcpyAsync(counts_host, base + o_cnt, ncell * 4, cudaMemcpyDeviceToHost, s));
    if (end_xy_host && RP) CUDA_TRY(cudaMemcpyAsync(end_xy_host, base + o_end, RP * 16, cudaMemcpyDeviceToHost, s));
    if (nverts_host && RP) CUDA_TRY(cudaMemcpyAsync(nverts_host, base + o_nv, RP * 4, cudaMemcpyDeviceToHost, s));
    if (status_host && RP) CUDA_TRY(cudaMemcpyAsync(status_host, base + o_stt, RP, cudaMemcpyDeviceToHost, s));
    if (stats_out) return oneka_read_stats(ctx, stats_out);
    CUDA_TRY(cudaStreamSynchronize(s));
    return ONEKA_OK;
}

int oneka_count_histogram(oneka_ctx *ctx, const uint32_t *counts_dev, int64_t ncell, int32_t nbins, uint64_t *hist_dev)
{
    if (!ctx || !counts_dev || !hist_dev || ncell < 0 || nbins < 2) return fail(ONEKA_ERR_ARG, "bad argument to oneka_count_histogram");
    CUDA_TRY(cudaSetDevice(ctx->device));
    CUDA_TRY(cudaMemsetAsync(hist_dev, 0, (size_t)nbins * sizeof(unsigned long long), ctx->stream));
    if (ncell == 0) return ONEKA_OK;
    long long blocks = (ncell + 255) / 256;
    const long long cap = (long long)ctx->sm_count * 16;
    if (blocks > cap) blocks = cap;
    if (nbins <= HIST_SMEM_BINS)
        count_histogram_kernel<true><<<(unsigned)blocks, 256, (size_t)nbins * sizeof(unsigned int), ctx->stream>>>(counts_dev, ncell, nbins, (unsigned long long *)hist_dev);
    else
        count_histogram_kernel<false><<<(unsigned)blocks, 256, 0, ctx->stream>>>(counts_dev, ncell, nbins, (unsigned long long *)hist_dev);
    ctx->launches++;
    CUDA_TRY(cudaGetLastError());
    return ONEKA_OK;
}

int oneka_gaussian_smooth(oneka_ctx *ctx, const uint32_t *counts_dev, int32_t nrows, int32_t ncols, double total_weight,
                          const double *w_host, int32_t lw, double *tmp_dev, double *out_dev)
{
    if (!ctx || !counts_dev || !w_host || !tmp_dev || !out_dev || nrows <= 0 || ncols <= 0 || lw < 0 || !(total_weight > 0.0))
        return fail(ONEKA_ERR_ARG, "bad argument to oneka_gaussian_smooth");
    CUDA_TRY(cudaSetDevice(ctx->device));
    const size_t wb = (size_t)(2 * lw + 1) * sizeof(double);
    int rc = ensure_aux(ctx, wb);                                   // context-owned scratch: no allocation, no synchronisation per call
    if (rc) return rc;
    double *w_dev = (double *)ctx->aux;
    // (a pageable host source is staged by the driver before cudaMemcpyAsync returns: w_host may be released by the caller at once)
    CUDA_TRY(cudaMemcpyAsync(w_dev, w_host, wb, cudaMemcpyHostToDevice, ctx->stream));
    const dim3 grid((ncols + 255) / 256, nrows);
    gaussian_axis_kernel<0, true><<<grid, 256, 0, ctx->stream>>>(counts_dev, nullptr, 1.0 / total_weight, nrows, ncols, w_dev, lw, tmp_dev);
    gaussian_axis_kernel<1, false><<<grid, 256, 0, ctx->stream>>>(nullptr, tmp_dev, 1.0, nrows, ncols, w_dev, lw, out_dev);
    ctx->launches += 2;
    CUDA_TRY(cudaGetLastError());
    return ONEKA_OK;
}

int oneka_capture_tracked(oneka_ctx *ctx, const oneka_model_desc *m, const oneka_lattice *lat,
                          const double *well_xy_dev, int64_t R, int32_t P,
                          const double *q_dev, const double *cond_dev, const double *poro_dev, const double *thick_dev,
                          const double *coef_dev, const double *start_xy_dev,
                          uint32_t *counts_dev, double *end_xy_dev, int32_t *nverts_dev, uint8_t *status_dev,
                          uint32_t *clipped_dev, double *bbox_dev)
{
    if (!ctx) return fail(ONEKA_ERR_ARG, "ctx is NULL");
    if (R == 0) return ONEKA_OK;
    if (!lat || !counts_dev) return fail(ONEKA_ERR_ARG, "oneka_capture_tracked needs lat and counts_dev");
    CUDA_TRY(cudaSetDevice(ctx->device));
    if (clipped_dev) CUDA_TRY(cudaMemsetAsync(clipped_dev, 0, (size_t)R * sizeof(uint32_t), ctx->stream));
    return capture_impl(ctx, m, lat, well_xy_dev, R, P, q_dev, cond_dev, poro_dev, thick_dev, coef_dev, start_xy_dev,
                        counts_dev, end_xy_dev, nverts_dev, status_dev, nullptr, bbox_dev, clipped_dev);
}

// ---- the collective ---------------------------------------------------------------------------
int oneka_comm_unique_id(void *id128_out)
{
    if (!id128_out) return fail(ONEKA_ERR_ARG, "id128_out is NULL");
    const nccl_api *N = nccl();
    if (!N) return fail(ONEKA_ERR_NCCL, "NCCL is not available (libnccl.so.2 could not be loaded: %s)", dlerror() ? dlerror() : "no error text");
    ncclUniqueId id;
    NCCL_TRY(N, N->GetUniqueId(&id));
    memcpy(id128_out, &id, sizeof(id));
    return ONEKA_OK;
}

int oneka_comm_destroy(oneka_ctx *ctx)
{
    if (!ctx) return fail(ONEKA_ERR_ARG, "ctx is NULL");
    if (ctx->comm && ctx->comm_owned) {
        const nccl_api *N = nccl();
        if (N) { CUDA_TRY(cudaSetDevice(ctx->device)); CUDA_TRY(cudaStreamSynchronize(ctx->stream)); NCCL_TRY(N, N->CommDestroy(ctx->comm)); }
    }
    ctx->comm = nullptr; ctx->comm_owned = false; ctx->comm_nranks = 1; ctx->comm_rank = 0;
    return ONEKA_OK;
}

int oneka_comm_init_rank(oneka_ctx *ctx, int32_t nranks, int32_t rank, const void *id128)
{
    if (!ctx || !id128) return fail(ONEKA_ERR_ARG, "NULL argument");
    if (nranks < 1 || rank < 0 || rank >= nranks) return fail(ONEKA_ERR_ARG, "rank %d of %d", rank, nranks);
    const nccl_api *N = nccl();
    if (!N) return fail(ONEKA_ERR_NCCL, "NCCL is not available (libnccl.so.2 could not be loaded)");
    int rc = oneka_comm_destroy(ctx);
    if (rc) return rc;
    CUDA_TRY(cudaSetDevice(ctx->device));
    ncclUniqueId id;
    memcpy(&id, id128, sizeof(id));
    ncclComm_t c = nullptr;
    NCCL_TRY(N, N->CommInitRank(&c, nranks, id, rank));
    ctx->comm = c; ctx->comm_owned = true; ctx->comm_nranks = nranks; ctx->comm_rank = rank;
    return ONEKA_OK;
}

int oneka_comm_attach(oneka_ctx *ctx, void *nccl_comm, int32_t nranks, int32_t rank)
{
    if (!ctx) return fail(ONEKA_ERR_ARG, "ctx is NULL");
    if (nccl_comm && !nccl()) return fail(ONEKA_ERR_NCCL, "NCCL is not available (libnccl.so.2 could not be loaded)");
    int rc = oneka_comm_destroy(ctx);
    if (rc) return rc;
    ctx->comm = (ncclComm_t)nccl_comm; ctx->comm_owned = false;
    ctx->comm_nranks = nccl_comm ? nranks : 1; ctx->comm_rank = nccl_comm ? rank : 0;
    return ONEKA_OK;
}

int oneka_allreduce_counts(oneka_ctx *ctx, uint32_t *counts_dev, uint64_t n)
{
    if (!ctx) return fail(ONEKA_ERR_ARG, "ctx is NULL");
    if (n == 0) return ONEKA_OK;
    if (!counts_dev) return fail(ONEKA_ERR_ARG, "counts_dev is NULL");
    if (!ctx->comm) return fail(ONEKA_ERR_ARG, "no communicator: call oneka_comm_init_rank or oneka_comm_attach first");
    const nccl_api *N = nccl();
    if (!N) return fail(ONEKA_ERR_NCCL, "NCCL is not available");
    CUDA_TRY(cudaSetDevice(ctx->device));
    NCCL_TRY(N, N->AllReduce(counts_dev, counts_dev, (size_t)n, ncclUint32, ncclSum, ctx->comm, ctx->stream));
    return ONEKA_OK;
}

int oneka_allreduce_f64(oneka_ctx *ctx, double *values_dev, uint64_t n, int32_t op)
{
    if (!ctx) return fail(ONEKA_ERR_ARG, "ctx is NULL");
    if (n == 0) return ONEKA_OK;
    if (!values_dev || op < 0 || op > 2) return fail(ONEKA_ERR_ARG, "bad argument to oneka_allreduce_f64");
    if (!ctx->comm) return fail(ONEKA_ERR_ARG, "no communicator: call oneka_comm_init_rank or oneka_comm_attach first");
    const nccl_api *N = nccl();
    if (!N) return fail(ONEKA_ERR_NCCL, "NCCL is not available");
    CUDA_TRY(cudaSetDevice(ctx->device));
    const ncclRedOp_t o = op == 0 ? ncclSum : (op == 1 ? ncclMin : ncclMax);
    NCCL_TRY(N, N->AllReduce(values_dev, values_dev, (size_t)n, ncclFloat64, o, ctx->comm, ctx->stream));
    return ONEKA_OK;
}

int oneka_distancesquared_host(oneka_ctx *ctx, int64_t n, const double *abc_host, double *out_host)
{
    if (!ctx || n < 0 || (n && (!abc_host || !out_host))) return fail(ONEKA_ERR_ARG, "bad argument to oneka_distancesquared_host");
    if (n == 0) return ONEKA_OK;
    CUDA_TRY(cudaSetDevice(ctx->device));
    int rc = ensure_aux(ctx, (size_t)n * 7 * sizeof(double));
    if (rc) return rc;
    double *buf = (double *)ctx->aux;
    cudaStream_t s = ctx->stream;
    cudaError_t e = cudaMemcpyAsync(buf, abc_host, (size_t)n * 6 * sizeof(double), cudaMemcpyHostToDevice, s);
    if (e == cudaSuccess) {
        distsq_kernel<<<(unsigned)((n + 127) / 128), 128, 0, s>>>(n, buf, buf + 6 * n);
        ctx->launches++;
        e = cudaGetLastError();
    }
    if (e == cudaSuccess) e = cudaMemcpyAsync(out_host, buf + 6 * n, (size_t)n * sizeof(double), cudaMemcpyDeviceToHost, s);
    if (e == cudaSuccess) e = cudaStreamSynchronize(s);
    if (e != cudaSuccess) return fail(ONEKA_ERR_CUDA, "oneka_distancesquared_host failed: %s", cudaGetErrorString(e));
    return ONEKA_OK;
}

int oneka_red_probe(oneka_ctx *ctx, int32_t mode, uint64_t span_bytes, int32_t iters, double *gops_out, double *ms_out)
{
    if (!ctx || mode < 0 || mode > 4 || iters <= 0) return fail(ONEKA_ERR_ARG, "bad argument to oneka_red_probe");
    CUDA_TRY(cudaSetDevice(ctx->device));
    if (span_bytes < 4096) span_bytes = 4096;
    const unsigned long long words = span_bytes / 4;
    unsigned int *buf = nullptr, *sink = nullptr;
    CUDA_TRY(cudaMalloc(&buf, words * 4));
    CUDA_TRY(cudaMalloc(&sink, 4));
    CUDA_TRY(cudaMemsetAsync(buf, 0, words * 4, ctx->stream));
    const int blocks = ctx->sm_count * 8, threads = 256;
    auto launch = [&](int it) {
        switch (mode) {
        case 0: red_probe_kernel<0><<<blocks, threads, 0, ctx->stream>>>(buf, words, it, sink); break;
        case 1: red_probe_kernel<1><<<blocks, threads, 0, ctx->stream>>>(buf, words, it, sink); break;
        case 2: red_probe_kernel<2><<<blocks, threads, 0, ctx->stream>>>(buf, words, it, sink); break;
        case 3: red_probe_kernel<3><<<blocks, threads, 0, ctx->stream>>>(buf, words, it, sink); break;
        default: red_probe_kernel<4><<<blocks, threads, 0, ctx->stream>>>(buf, words, it, sink); break;
        }
    };
    cudaEvent_t a, b;
    CUDA_TRY(cudaEventCreate(&a));
    CUDA_TRY(cudaEventCreate(&b));
    launch(iters / 8 + 1);                                   // warm-up
    float best = 1e30f;
    for (int rep = 0; rep < 5; ++rep) {
        CUDA_TRY(cudaEventRecord(a, ctx->stream));
        launch(iters);
        CUDA_TRY(cudaEventRecord(b, ctx->stream));
        CUDA_TRY(cudaEventSynchronize(b));
        float ms = 0.f;
        CUDA_TRY(cudaEventElapsedTime(&ms, a, b));
        if (ms < best) best = ms;
    }
    ctx->launches += 6;
    cudaEventDestroy(a); cudaEventDestroy(b); cudaFree(buf); cudaFree(sink);
    CUDA_TRY(cudaGetLastError());
    const double ops = (double)blocks * threads * (double)iters;
    if (gops_out) *gops_out = ops / (best * 1e-3) / 1e9;
    if (ms_out) *ms_out = best;
    return ONEKA_OK;
}

int oneka_fp64_probe(oneka_ctx *ctx, int iters, double *tflops_out, double *ms_out)
{
    if (!ctx || iters <= 0) return fail(ONEKA_ERR_ARG, "bad argument");
    CUDA_TRY(cudaSetDevice(ctx->device));
    double *sink = nullptr;
    CUDA_TRY(cudaMalloc(&sink, 8));
    const int blocks = ctx->sm_count * 8, threads = 256;
    cudaEvent_t a, b;
    CUDA_TRY(cudaEventCreate(&a));
    CUDA_TRY(cudaEventCreate(&b));
    fp64_probe_kernel<<<blocks, threads, 0, ctx->stream>>>(iters / 8 + 1, 1.0, sink);     // warm-up
    float best = 1e30f;
    for (int rep = 0; rep < 5; ++rep) {
        CUDA_TRY(cudaEventRecord(a, ctx->stream));
        fp64_probe_kernel<<<blocks, threads, 0, ctx->stream>>>(iters, 1.0, sink);
        CUDA_TRY(cudaEventRecord(b, ctx->stream));
        CUDA_TRY(cudaEventSynchronize(b));
        float ms = 0.f;
        CUDA_TRY(cudaEventElapsedTime(&ms, a, b));
        if (ms < best) best = ms;
    }
    ctx->launches += 6;
    cudaEventDestroy(a); cudaEventDestroy(b); cudaFree(sink);
    CUDA_TRY(cudaGetLastError());
    const double flops = (double)blocks * threads * (double)iters * 8.0 * 2.0;
    if (tflops_out) *tflops_out = flops / (best * 1e-3) / 1e12;
    if (ms_out) *ms_out = best;
    return ONEKA_OK;
}

}  // extern "C"

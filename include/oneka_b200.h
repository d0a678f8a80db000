/*
 * oneka_b200.h -- C ABI of the B200-native capture-zone hot path (liboneka_b200.so).
 *
 * The reference (RandalJBarnes/OnekaPy) is pure Python and has no FFI layer; its boundary
 * for this path is the Python function API.  Each entry point below names the reference
 * interface it replaces (file:line under the reference tree).  A maintainer binds these
 * with ctypes; onekapy_b200/_cabi.py is that binding.
 *
 * Conventions
 *   - plain pointers and sizes only; no C++/torch types cross this boundary;
 *   - every function returns ONEKA_OK (0) or a negative error code and never throws;
 *     oneka_last_error() gives the message for the calling thread's last failure;
 *   - pointers named *_dev are DEVICE pointers (caller-owned, e.g. torch tensors' data_ptr());
 *     pointers named *_host are host pointers; the *_host entry points do their own copies;
 *   - one oneka_ctx per GPU; work is enqueued on the context's stream (oneka_set_stream) and
 *     is asynchronous unless stated otherwise; a context is not thread-safe;
 *   - there is no CPU fallback: oneka_create() fails when no sm_100 CUDA device is usable.
 *
 * Array layouts (row-major, C order)
 *   well_xy[nw][2]      well coordinates x, y                       (oneka/model.py:158-171)
 *   q[R][nw]            well discharges per realization             (oneka/stochastic.py:224-228)
 *   cond[R] poro[R] thick[R]  conductivity, porosity, thickness     (oneka/stochastic.py:231-233)
 *   coef[R][6]          regional coefficients A..F                  (oneka/stochastic.py:241)
 *   start_xy[P][2]      start ring around the target well           (oneka/capturezone.py:113-115)
 *   counts[nrows][ncols] uint32, row = y index, col = x index       (oneka/probabilityfield.py:148; pgrid with weight 1)
 */
#ifndef ONEKA_B200_H
#define ONEKA_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define ONEKA_ABI_VERSION 1

/* return codes */
#define ONEKA_OK            0
#define ONEKA_ERR_ARG      -1
#define ONEKA_ERR_CUDA     -2
#define ONEKA_ERR_NOMEM    -3
#define ONEKA_ERR_NODEVICE -4
#define ONEKA_ERR_NCCL     -5

/* per-path status words (status[R][P]) */
#define ONEKA_PATH_OK          0
#define ONEKA_PATH_AQUIFER_DRY 1   /* AquiferError inside feval: oneka/model.py:343-344, 380-381; trace truncated as by capturezone.py:249-253 */
#define ONEKA_PATH_MAX_ATTEMPT 2   /* guard the reference lacks (capturezone.py:221 has no iteration cap) */
#define ONEKA_PATH_NONFINITE   3   /* guard the reference lacks (it would spin forever on nan) */
#define ONEKA_PATH_TRACE_FULL  4   /* oneka_trace only: max_verts reached */

typedef struct oneka_ctx oneka_ctx;

/* The aquifer model that is constant over a run: oneka/model.py:131-196 (Model attributes)
 * plus the solver settings of oneka/capturezone.py:127 (compute_backtrace arguments). */
typedef struct {
    int32_t nw;            /* number of wells                                                  */
    int32_t confined;      /* 1: compute_velocity_confined (model.py:392-427); 0: compute_velocity (model.py:353-389) */
    double  base;          /* aquifer base elevation (model.py:179); does not enter the velocity */
    double  xo, yo;        /* origin of the regional quadratic = target well (stochastic.py:239 -> model.py:490-491) */
    double  duration;      /* capturezone.py:127; negative = forward tracking                  */
    double  tol;           /* absolute local error bound [m]                                   */
    double  maxstep;       /* space-step cap [m]                                               */
    int64_t max_attempts;  /* per-path cap on DOPRI attempts; <= 0 selects the default (1<<22)   */
} oneka_model_desc;

/* A fixed node-centred lattice: node (i, j) sits at (xmin + j*deltax, ymin + i*deltay)
 * exactly as oneka/probabilityfield.py:306-307 computes cx, cy. */
typedef struct {
    double  xmin, ymin;
    double  deltax, deltay;
    int32_t nrows, ncols;
    double  umbra;         /* vector-to-raster range, probabilityfield.py:264 (insert) */
} oneka_lattice;

/* Run statistics, filled by oneka_read_stats (which synchronises the stream). */
typedef struct {
    uint64_t attempts;         /* DOPRI5 attempts summed over all particles ("particle-steps") */
    uint64_t steps;            /* accepted steps (= segments rasterised)                       */
    uint64_t paths;            /* particles tracked                                            */
    uint64_t n_not_ok;         /* paths whose status != ONEKA_PATH_OK                          */
    uint64_t n_clipped;        /* segments whose raster window was clipped by the lattice edge */
    uint64_t exact_tests;      /* cells whose FP32 classification fell inside the error band and were re-tested in exact FP64 */
    double   bbox[4];          /* min x, max x, min y, max y over every vertex (incl. start points) */
} oneka_stats;

const char *oneka_last_error(void);
int         oneka_abi_version(void);

/* ---- context ------------------------------------------------------------------------ */
oneka_ctx *oneka_create(int device);                    /* NULL on failure; see oneka_last_error */
void       oneka_destroy(oneka_ctx *ctx);
int        oneka_set_stream(oneka_ctx *ctx, void *cuda_stream);         /* cudaStream_t; NULL = legacy default */
int        oneka_set_workspace_limit(oneka_ctx *ctx, uint64_t bytes);   /* cap for the per-realization registration bitmaps */
/* The rasteriser (ProbabilityField.insert, oneka/probabilityfield.py:296-310) has two flavours with identical results: the
 * plain one, and one for lattices where a segment's window spans many rows (umbra >> deltay), which issues fewer bit-set
 * operations (64-bit RED.OR; rows wholly behind the start of a segment, which the previous segment of the path already set,
 * are not set again) for a few more instructions per row.  mode 0 (default): chosen from the lattice handed to each capture
 * call (2 umbra / deltay + 1 >= 8 rows with direct well sums, >= 11 with the far field: heavy);  1: always plain;  2: always heavy. */
int        oneka_set_raster_mode(oneka_ctx *ctx, int32_t mode);
int        oneka_raster_flavour(const oneka_ctx *ctx, double umbra, double deltay, int32_t farfield);   /* the flavour a capture on such a lattice runs: 0 plain, 1 heavy (< 0: error) */
int        oneka_synchronize(oneka_ctx *ctx);
uint64_t   oneka_launch_count(const oneka_ctx *ctx);    /* kernels launched by this context so far */
/* When enabled, CUDA events bracket every launch of the tracking/raster kernel and of the flush
 * kernel on the context's stream; oneka_kernel_ms returns the accumulated device time. */
int        oneka_set_profiling(oneka_ctx *ctx, int enabled);
int        oneka_kernel_ms(oneka_ctx *ctx, double *track_ms, double *flush_ms, uint64_t *track_launches, int reset);

/* ---- Model evaluation at points ----------------------------------------------------- *
 * Replaces Model.compute_potential / compute_discharge / compute_velocity_confined /
 * compute_head / compute_velocity (oneka/model.py:207-427) for npts points of ONE model.
 * out_host[npts][8] = potential, Qx, Qy, Vx_confined, Vy_confined, head, Vx, Vy
 * (head, Vx, Vy are nan where the reference raises AquiferError).  Synchronous.          */
int oneka_eval_points_host(oneka_ctx *ctx, const oneka_model_desc *m, const double *well_xy_host,
                           const double *q_host, double cond, double poro, double thick,
                           const double *coef_host, int64_t npts, const double *pts_host, double *out_host);

/* ---- Backtraces with stored vertices (test hook for the tracking kernel alone) -------- *
 * Replaces compute_backtrace (oneka/capturezone.py:127-253) for R x P particles.
 * verts_dev[R][P][max_verts][2]; nverts_dev[R][P] counts vertices incl. the start point. */
int oneka_trace(oneka_ctx *ctx, const oneka_model_desc *m, const double *well_xy_dev,
                int64_t R, int32_t P,
                const double *q_dev, const double *cond_dev, const double *poro_dev, const double *thick_dev,
                const double *coef_dev, const double *start_xy_dev,
                int32_t max_verts, double *verts_dev, int32_t *nverts_dev, uint8_t *status_dev, int32_t *attempts_dev);

/* ---- Rasterise given traces (test hook for the rasteriser alone) ---------------------- *
 * Replaces ProbabilityField.insert per segment + register(1.0) per realization
 * (oneka/probabilityfield.py:264-310, 342-359) on a FIXED lattice:
 * trace t = verts_dev[offsets[t] .. offsets[t+1]) belongs to realization real_of_dev[t]
 * (0 <= real_of < nreal); counts_dev[nrows][ncols] += 1 per realization that marks the cell. */
int oneka_raster_traces(oneka_ctx *ctx, const oneka_lattice *lat, int64_t ntraces,
                        const int64_t *offsets_dev, const double *verts_dev, const int32_t *real_of_dev,
                        int64_t nreal, uint32_t *counts_dev);

/* ---- The hot path: track + rasterise + register, R realizations x P paths ------------- *
 * Replaces the body of the realization loop, i.e. R calls of compute_capturezone
 * (oneka/capturezone.py:51-123, called at oneka/stochastic.py:264-265 and
 * oneka/deterministic.py:232-233) with weight 1.0 on a fixed lattice.
 * counts_dev is accumulated (+=).  lat == NULL or counts_dev == NULL: tracking only
 * (bounding box / end points / statistics; no rasterisation).
 * Optional per-path outputs may be NULL: end_xy_dev[R][P][2], nverts_dev[R][P], status_dev[R][P]. */
int oneka_capture(oneka_ctx *ctx, const oneka_model_desc *m, const oneka_lattice *lat,
                  const double *well_xy_dev, int64_t R, int32_t P,
                  const double *q_dev, const double *cond_dev, const double *poro_dev, const double *thick_dev,
                  const double *coef_dev, const double *start_xy_dev,
                  uint32_t *counts_dev, double *end_xy_dev, int32_t *nverts_dev, uint8_t *status_dev);

/* ---- Guarded capture: the lattice was only estimated ------------------------------------- *
 * Same as oneka_capture, but a realization one of whose segments had its window clipped by the
 * LATTICE EDGE (it left the estimated extents) is not registered: clipped_dev[r] is set to 1 (0
 * otherwise) and its bitmap is discarded -- the analogue of ProbabilityField.reset()
 * (oneka/probabilityfield.py:362-376, "discarding a partially processed invalid realization").
 * The caller re-runs just those realizations on a lattice grown to the bounding box reported by
 * oneka_read_stats, instead of repeating the whole run (Engine.run).                            */
int oneka_capture_guarded(oneka_ctx *ctx, const oneka_model_desc *m, const oneka_lattice *lat,
                          const double *well_xy_dev, int64_t R, int32_t P,
                          const double *q_dev, const double *cond_dev, const double *poro_dev, const double *thick_dev,
                          const double *coef_dev, const double *start_xy_dev,
                          uint32_t *counts_dev, double *end_xy_dev, int32_t *nverts_dev, uint8_t *status_dev,
                          uint32_t *clipped_dev);

/* ---- Guarded capture that also reports every path's bounding box ----------------------------- *
 * oneka_capture_guarded plus bbox_dev[R][P][4] = min x, max x, min y, max y of each path's vertices (either of
 * clipped_dev / bbox_dev may be NULL).  One fused pass then serves BOTH the count grid and the running union of
 * bounding boxes from which the reference's order-dependent clip (next section) is decided: only the realizations
 * containing a path whose windows the reference's grid-at-that-moment would have clipped are re-rasterised with
 * oneka_capture_clipped (Engine.run_exact) -- instead of a full tracking pass before the fused one.               */
int oneka_capture_tracked(oneka_ctx *ctx, const oneka_model_desc *m, const oneka_lattice *lat,
                          const double *well_xy_dev, int64_t R, int32_t P,
                          const double *q_dev, const double *cond_dev, const double *poro_dev, const double *thick_dev,
                          const double *coef_dev, const double *start_xy_dev,
                          uint32_t *counts_dev, double *end_xy_dev, int32_t *nverts_dev, uint8_t *status_dev,
                          uint32_t *clipped_dev, double *bbox_dev);

/* ---- Exact emulation of the reference's auto-expanding field ----------------------------- *
 * The reference expands its grid to each trace's bounding box just before inserting it
 * (ProbabilityField.rasterize, oneka/probabilityfield.py:335) and insert() clips every segment's
 * window to the grid as it is at that moment (:298-301), so what a path marks depends on all paths
 * before it, in (realization, path) order.  Two calls reproduce that exactly:
 *   oneka_path_bboxes      tracking only; bbox_dev[R][P][4] = min x, max x, min y, max y of each path
 *                          (the host turns their running union into per-path lattice windows);
 *   oneka_capture_clipped  oneka_capture with clip_dev[R][P][4] = left, right, bottom, top (half-open
 *                          lattice index ranges, 16-byte aligned) applied to every segment of the path. */
int oneka_path_bboxes(oneka_ctx *ctx, const oneka_model_desc *m, const double *well_xy_dev, int64_t R, int32_t P,
                      const double *q_dev, const double *cond_dev, const double *poro_dev, const double *thick_dev,
                      const double *coef_dev, const double *start_xy_dev, double *bbox_dev, uint8_t *status_dev);
int oneka_capture_clipped(oneka_ctx *ctx, const oneka_model_desc *m, const oneka_lattice *lat,
                          const double *well_xy_dev, int64_t R, int32_t P,
                          const double *q_dev, const double *cond_dev, const double *poro_dev, const double *thick_dev,
                          const double *coef_dev, const double *start_xy_dev, const int32_t *clip_dev,
                          uint32_t *counts_dev, double *end_xy_dev, int32_t *nverts_dev, uint8_t *status_dev);

/* ---- Far-field compression of the well sum (confined flow) ------------------------------------- *
 * The reference adds one term per well at EVERY velocity evaluation (Model.compute_discharge,
 * oneka/model.py:307-313: 15 flops x nw, six times per DOPRI5 attempt).  With this switched on, the capture
 * entry points above evaluate the wells' part as  [direct sum over the few wells near the particle's tile]
 * + [one complex polynomial of `order` terms for all the others]  on a grid of ntx x nty square tiles of side
 * `tile` whose lower-left corner is (x0, y0): tiled local (Taylor) expansions of sum_w w/(z - z_w), wells farther
 * than tile/(sqrt(2) eta) from a tile's centre being "far".  Truncation <= eta^order/(1 - eta) relative to a far
 * term (3e-15 for eta = 0.3, order = 28); particles outside the grid, unconfined flow and models whose nw / xo / yo
 * differ from the ones given here use the direct sum.  The tables depend on the well COORDINATES only (host pointer;
 * must be the wells later passed as well_xy_dev: every launch checks that on the device, and oneka_read_stats fails
 * with ONEKA_ERR_ARG when a launch since the last oneka_reset_stats was handed other coordinates); the realization-dependent coefficients are formed on the device
 * per launch.  order must be even; order 16 runs an unrolled evaluation (Engine's default: eta = 0.15, order 16, truncation
 * 8e-14 of a far term -- below the 1e-12 of the Newton reciprocal in the direct sum).  order_fp64 is ignored (the slot of
 * an FP32 tail that measured slower and was removed).  nw = 0 or order = 0 switches the far field off (the default).  Synchronous.
 * max_near_out / mean_near_out (may be NULL): padded length of the longest near list, mean near wells per tile.   */
int oneka_set_farfield(oneka_ctx *ctx, int32_t nw, const double *well_xy_host, double xo, double yo,
                       double x0, double y0, double tile, int32_t ntx, int32_t nty, int32_t order, double eta,
                       int32_t order_fp64, int32_t *max_near_out, double *mean_near_out);
/* What the tables of the last oneka_set_farfield need: tiles, order, longest near list, dynamic shared memory per CTA of the
 * confined / unconfined tracking kernels, and the budget per CTA that keeps two of those CTAs on an SM (beyond it the kernel
 * still runs, at half the occupancy: Engine sizes the tile grid to stay inside).  Any pointer may be NULL.               */
int oneka_farfield_info(oneka_ctx *ctx, int32_t *ntiles, int32_t *order, int32_t *max_near, uint64_t *smem_confined,
                        uint64_t *smem_unconfined, uint64_t *smem_budget);
/* Also use the far field for
 * UNCONFINED flow (Model.compute_velocity, oneka/model.py:353-389).  The far wells' part of the potential -- needed only to
 * decide whether the aquifer is fully saturated at the point -- comes from the same coefficients in FP32; where that decision
 * is not certain the evaluation falls back to the direct sums with FP64 logs, exactly as without the far field.            */
int oneka_set_farfield_unconfined(oneka_ctx *ctx, int enabled);
/* Host restatement of the same tables and evaluation (NO GPU needed; test hook for the expansion's accuracy):
 * out_host[npts][2] = sum_w w_host[w] (x - x_w)/r_w^2, sum_w w_host[w] (y - y_w)/r_w^2 evaluated the far-field way;
 * near_count_out[npts] (may be NULL) = near wells summed directly, or -1 where the point lies outside the grid.   */
int oneka_farfield_eval_host(int32_t nw, const double *well_xy_host, const double *w_host, double xo, double yo,
                             double x0, double y0, double tile, int32_t ntx, int32_t nty, int32_t order, double eta,
                             int32_t order_fp64, int64_t npts, const double *pts_host, double *out_host, int32_t *near_count_out);

/* Statistics of everything enqueued since the last oneka_reset_stats.  Synchronises. */
int oneka_read_stats(oneka_ctx *ctx, oneka_stats *out);
int oneka_reset_stats(oneka_ctx *ctx);

/* Same as oneka_capture with HOST buffers: copies the parameter rows to the device, runs,
 * and copies counts (nrows*ncols uint32, overwritten) back.  Synchronous.
 * This is the call the reference-facing Python layer makes for one batch (bench `e2e`). */
int oneka_capture_host(oneka_ctx *ctx, const oneka_model_desc *m, const oneka_lattice *lat,
                       const double *well_xy_host, int64_t R, int32_t P,
                       const double *q_host, const double *cond_host, const double *poro_host, const double *thick_host,
                       const double *coef_host, const double *start_xy_host,
                       uint32_t *counts_host, double *end_xy_host, int32_t *nverts_host, uint8_t *status_host,
                       oneka_stats *stats_out);

/* ---- Post-processing of the count grid (the step right after the path) -------------------- *
 * oneka_count_histogram: hist_dev[c] = number of lattice nodes whose count is c, 0 <= c < nbins
 *   (counts >= nbins land in the last bin).  With weight-1 realizations pgrid is integer valued, so
 *   this histogram IS the sorted exceedance curve of create_impact_plot (oneka/visualize.py:382-386:
 *   pr = flip(sort(pgrid/total_weight)), area = (arange(n)+1)*spacing^2) and, through hist[0], the
 *   deterministic capture-zone area (oneka/visualize.py:316-331).
 * oneka_gaussian_smooth: out_dev[nrows][ncols] = scipy.ndimage.gaussian_filter(counts/total_weight,
 *   sigma, mode='constant', cval=0) as used by create_probability_plot (oneka/visualize.py:228-233):
 *   separable, axis 0 then axis 1, taps w_host[2*lw+1] (normalised, lw = int(4 sigma + 0.5)), FP64.
 *   tmp_dev is a caller-provided scratch grid of the same size as out_dev.                       */
int oneka_count_histogram(oneka_ctx *ctx, const uint32_t *counts_dev, int64_t ncell, int32_t nbins, uint64_t *hist_dev);
int oneka_gaussian_smooth(oneka_ctx *ctx, const uint32_t *counts_dev, int32_t nrows, int32_t ncols, double total_weight,
                          const double *w_host, int32_t lw, double *tmp_dev, double *out_dev);

/* ---- The one collective of the path: sum of the per-GPU count grids ------------------------- *
 * Realizations shard over GPUs with no exchange until the end; the only shared state of the reference's loop is the
 * additive grid (ProbabilityField.register, oneka/probabilityfield.py:357-358: pgrid[rgrid] += weight, total_weight += weight).
 * oneka_allreduce_counts is that sum: ncclAllReduce(counts, uint32, sum), in place, enqueued on the context's stream
 * (so it orders after the captures that filled `counts` without a host synchronisation).  Integer sums are
 * order-independent: the reduced grid is bit-identical for any number of GPUs.
 *   oneka_comm_unique_id   fills 128 bytes (ncclUniqueId) on ONE rank; the caller ships them to the others by any means
 *                          (torch.distributed broadcast, MPI, a file ...);
 *   oneka_comm_init_rank   joins the communicator (collective over all nranks contexts; the context owns it);
 *   oneka_comm_attach      alternatively borrow an existing ncclComm_t (e.g. the host framework's); not destroyed here;
 *   oneka_allreduce_f64    small packed reductions on the same communicator (op: 0 sum, 1 min, 2 max), e.g. the
 *                          bounding box every rank needs to agree on the lattice.
 * NCCL is resolved at run time (the libnccl.so.2 already in the process, else the system one; ONEKA_NCCL_LIB
 * overrides), so the library loads -- and single-GPU use works -- without it.                                    */
int oneka_comm_unique_id(void *id128_out);
int oneka_comm_init_rank(oneka_ctx *ctx, int32_t nranks, int32_t rank, const void *id128);
int oneka_comm_attach(oneka_ctx *ctx, void *nccl_comm, int32_t nranks, int32_t rank);
int oneka_comm_destroy(oneka_ctx *ctx);
int oneka_allreduce_counts(oneka_ctx *ctx, uint32_t *counts_dev, uint64_t n);
int oneka_allreduce_f64(oneka_ctx *ctx, double *values_dev, uint64_t n, int32_t op);

/* ---- ProbabilityField.distancesquared on the device (test hook) ------------------------------- *
 * out_host[n] = distance^2 from c to the segment [a, b] for abc_host[n][6] = ax, ay, bx, by, cx, cy, evaluated by the
 * very device function the rasteriser uses inside its error band (exact_distancesquared: the operations of
 * oneka/probabilityfield.py:407-427 in unfused IEEE double).  Bit-exact against the reference.  Synchronous.      */
int oneka_distancesquared_host(oneka_ctx *ctx, int64_t n, const double *abc_host, double *out_host);

/* ---- Atomic-throughput probes: the roofline of the rasteriser -------------------------------- *
 * The rasteriser's memory operation is a bit-set: RED.OR of one 32-bit word per lattice row per segment
 * (insert(), oneka/probabilityfield.py:296-310 sets rgrid[i, j] node by node).  mode 0: RED.OR to L2, every lane its own
 * word of a `span_bytes` buffer (uncontended);  mode 1: the same with all 32 lanes of a warp on ONE word (contended);
 * mode 2: atomicOr on shared memory, every lane its own word;  mode 3: shared memory, one word per warp;
 * mode 4: RED.OR to L2 with every lane on its OWN 32-byte sector, moving on by one bitmap row per operation -- the rasteriser's
 * own access pattern (each lane tracks another particle; mode 0 packs 8 lanes into a sector and needs 8x fewer L2 requests).
 * gops_out = 1e9 atomic word-operations per second (best of 5).  Synchronous.                                    */
int oneka_red_probe(oneka_ctx *ctx, int32_t mode, uint64_t span_bytes, int32_t iters, double *gops_out, double *ms_out);

/* ---- FP64 pipe probe ------------------------------------------------------------------ *
 * Times a register-resident DFMA kernel (no memory traffic) and reports the achieved
 * FP64 rate; bench.py uses it as the measured FP64 roofline denominator.  Synchronous.  */
int oneka_fp64_probe(oneka_ctx *ctx, int iters, double *tflops_out, double *ms_out);

#ifdef __cplusplus
}
#endif
#endif /* ONEKA_B200_H */

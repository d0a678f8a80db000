// oneka_farfield_host.h -- host-side tables of the far-field compression (geometry only; see "Far-field compression" in
// oneka_device.cuh).  Shared by the library (oneka_api.cu: oneka_set_farfield, oneka_farfield_eval_host) and by the
// host emulation of the device code that the CPU tests build (tests/emu).  No CUDA runtime calls in here.
#pragma once
#include "oneka_device.cuh"

#include <cmath>
#include <cstdio>
#include <vector>

namespace oneka {

struct FFTables {
    int ntiles = 0, max_near = 0;
    double mean_near = 0.0;
    std::vector<double2> P;                    // [ntiles][nw][order]; zero rows for near wells
    std::vector<unsigned int> off;             // [ntiles][max_near] byte offsets into the well store (padded with the dummy well)
    std::vector<unsigned short> cnt;           // [ntiles] padded (even) lengths
    std::vector<int> near_flat, near_begin;    // unpadded near lists (host evaluator)
    // unconfined flow only (the potential's far part, see field_feval_ff_unc):
    std::vector<double> Lg;                    // [ntiles][nw] ln |z_w - z_c| for far wells, 0 for near wells
    std::vector<unsigned short> idx, cnt_raw;  // [ntiles][max_near] near WELL INDICES (unpadded lists), [ntiles] their lengths
};

// returns nullptr, or the reason the arguments are unusable
static inline const char *build_ff_tables(int nw, const double *well_xy, double xo, double yo, double gx0, double gy0, double tile,
                                          int ntx, int nty, int order, double eta, FFTables &T)
{
    if (nw < 1 || !well_xy) return "far field needs nw >= 1 wells";
    if (!(tile > 0.0) || ntx < 1 || nty < 1 || (long long)ntx * nty > 4096) return "far field: bad tile grid (tile > 0, 1 <= ntx * nty <= 4096)";
    if (order < 4 || order > 64 || (order & 1)) return "far field: order must be even and in [4, 64]";
    if (!(eta > 0.0 && eta < 0.9)) return "far field: eta must be in (0, 0.9)";
    const int ntiles = ntx * nty;
    const long double h = (long double)tile / sqrtl(2.0L);
    const long double rfar = h / (long double)eta;
    T.ntiles = ntiles;
    T.P.assign((size_t)ntiles * nw * order, make_double2(0.0, 0.0));
    T.Lg.assign((size_t)ntiles * nw, 0.0);
    T.near_flat.clear();
    T.near_begin.assign(ntiles + 1, 0);
    int maxn = 0;
    for (int tj = 0; tj < nty; ++tj)
        for (int ti = 0; ti < ntx; ++ti) {
            const int t = tj * ntx + ti;
            const long double cx = (long double)gx0 + ((long double)ti + 0.5L) * tile;     // tile centre relative to (xo, yo)
            const long double cy = (long double)gy0 + ((long double)tj + 0.5L) * tile;
            for (int w = 0; w < nw; ++w) {
                const long double dx = ((long double)well_xy[2 * w] - xo) - cx, dy = ((long double)well_xy[2 * w + 1] - yo) - cy;
                const long double d2 = dx * dx + dy * dy;
                if (!(sqrtl(d2) >= rfar)) { T.near_flat.push_back(w); continue; }          // near (or nan): summed directly
                T.Lg[(size_t)t * nw + w] = (double)(0.5L * logl(d2));
                const long double ir = dx / d2, ii = -dy / d2;                              // 1/(z_w - z_c)
                const long double ur = h * ir, ui = h * ii;                                 // h/(z_w - z_c)
                long double tr = -ir, tim = -ii;                                            // term_0 = -1/(z_w - z_c)
                double2 *row = &T.P[((size_t)t * nw + w) * order];
                for (int k = 0; k < order; ++k) {
                    row[k] = make_double2((double)tr, (double)tim);
                    const long double nr = tr * ur - tim * ui, ni = tr * ui + tim * ur;
                    tr = nr; tim = ni;
                }
            }
            T.near_begin[t + 1] = (int)T.near_flat.size();
            const int n = T.near_begin[t + 1] - T.near_begin[t];
            if (n > maxn) maxn = n;
        }
    T.max_near = (maxn + 1) & ~1;
    if (T.max_near < 2) T.max_near = 2;
    T.mean_near = (double)T.near_flat.size() / ntiles;
    if (T.max_near > 65534) return "far field: near list too long";
    const unsigned int dummy = (unsigned int)ff_dummy_offset(nw) * 8u;
    T.off.assign((size_t)ntiles * T.max_near, dummy);
    T.cnt.assign(ntiles, 0);
    T.cnt_raw.assign(ntiles, 0);
    T.idx.assign((size_t)ntiles * T.max_near, 0);
    if (nw > 65535) return "far field: too many wells for 16-bit near lists";
    for (int t = 0; t < ntiles; ++t) {
        const int n = T.near_begin[t + 1] - T.near_begin[t];
        for (int i = 0; i < n; ++i) {
            const int w = T.near_flat[T.near_begin[t] + i];
            T.off[(size_t)t * T.max_near + i] = (unsigned int)((w >> 2) * SWELL_BLK + 3 * (w & 3)) * 8u;
        }
        T.cnt[t] = (unsigned short)((n + 1) & ~1);
        T.cnt_raw[t] = (unsigned short)n;
        for (int i = 0; i < n; ++i) T.idx[(size_t)t * T.max_near + i] = (unsigned short)T.near_flat[T.near_begin[t] + i];
    }
    return nullptr;
}

// c[tile][k] = sum_w w_w P[tile][w][k] as farfield_coef_kernel forms them (w = the scaled discharges q/(2 pi H n))
static inline void ff_host_coefficients(const FFTables &T, int nw, int order, const double *w, std::vector<double2> &coef)
{
    coef.assign((size_t)T.ntiles * order, make_double2(0.0, 0.0));
    for (int t = 0; t < T.ntiles; ++t)
        for (int k = 0; k < order; ++k) {
            double ar = 0.0, ai = 0.0;
            for (int j = 0; j < nw; ++j) {
                const double2 pk = T.P[((size_t)t * nw + j) * order + k];
                ar = fma(w[j], pk.x, ar);
                ai = fma(w[j], pk.y, ai);
            }
            coef[(size_t)t * order + k] = make_double2(ar, ai);
        }
}

// unconfined flow: c[tile][k] with w = q/(2 pi) and b0[tile] = sum_far w ln |z_w - z_c|, as farfield_coef_unc_kernel forms them
static inline void ff_host_b0(const FFTables &T, int nw, const double *w, std::vector<double> &b0)
{
    b0.assign(T.ntiles, 0.0);
    for (int t = 0; t < T.ntiles; ++t) {
        double a = 0.0;
        for (int j = 0; j < nw; ++j) a = fma(w[j], T.Lg[(size_t)t * nw + j], a);
        b0[t] = a;
    }
}

}  // namespace oneka

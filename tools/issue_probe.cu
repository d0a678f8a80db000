// issue_probe.cu -- what does one well of the tracking loop cost on the B200 issue port / FP64 pipe?
//
// Stand-alone micro-benchmark (not part of the library):
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o build/issue_probe tools/issue_probe.cu && build/issue_probe
// Each variant runs the 4-well block of field_feval (oneka_device.cuh) with one ingredient removed or replaced and
// reports cycles per well per SM sub-partition (4 warps-per-SMSP settings).  The differences are the marginal cost of
// the LDS.128 loads, the MUFU.RCP64H seed (+ the MOV that zeroes its low word) and the loop control.
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>

constexpr int BLK = 14;   // doubles per block of 4 wells (x0 y0 .. x3 y3, w0..w3, 4 floats)

__device__ __forceinline__ double seed_mufu(double a)
{
    double y;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(a));
    return y;
}
__device__ __forceinline__ double seed_int(double a)      // one integer op instead of MUFU (a rough 1/a, good enough here)
{
    return __hiloint2double(0x7fde0000 - __double2hiint(a), 0);
}

template <int SEED>   // 0 = MUFU, 1 = integer, 2 = none (y0 = a * const: one more FP64 instead)
__device__ __forceinline__ void well(double x, double y, double xw, double yw, double w, double &gx, double &gy)
{
    const double dx = x - xw, dy = y - yw;
    const double r2 = fma(dy, dy, dx * dx);
    const double y0 = SEED == 0 ? seed_mufu(r2) : (SEED == 1 ? seed_int(r2) : r2 * 1e-7);
    const double e = fma(-r2, y0, 1.0);
    const double s0 = w * y0;
    const double s = fma(s0, e, s0);
    gx = fma(s, dx, gx);
    gy = fma(s, dy, gy);
}

// V: 0 full (LDS + MUFU)   1 registers instead of LDS   2 LDS + integer seed   3 registers + integer seed
//    4 registers + no seed instruction (10 FP64 per well)   5 LDS + no seed     6 pure DFMA chains (36 per "block")
template <int V>
__global__ void __launch_bounds__(128) probe(int iters, int nblk, const double *src, double *sink)
{
    extern __shared__ double s_w[];
    for (int i = threadIdx.x; i < nblk * BLK; i += blockDim.x) s_w[i] = src[i];
    __syncthreads();
    double x = 1000.0 + threadIdx.x, y = 2000.0 + blockIdx.x % 97, gx = 0.0, gy = 0.0;
    if (V == 6) {
        double a[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) a[k] = x + k;
        for (int it = 0; it < iters; ++it)
            for (int b = 0; b < nblk; ++b) {
#pragma unroll
                for (int k = 0; k < 36; ++k) a[k & 7] = fma(a[k & 7], 0.999999, 1e-9);
            }
#pragma unroll
        for (int k = 0; k < 8; ++k) gx += a[k];
    } else {
        constexpr bool LDS = (V == 0 || V == 2 || V == 5);
        constexpr int SEED = (V == 0 || V == 1) ? 0 : ((V == 2 || V == 3) ? 1 : 2);
        // register operands: loop-invariant but opaque values
        double rx[4], ry[4], rw[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) { rx[k] = src[2 * k]; ry[k] = src[2 * k + 1]; rw[k] = src[8 + k]; }
        for (int it = 0; it < iters; ++it) {
            const double *p = s_w;
            const double *pend = s_w + nblk * BLK;
#pragma unroll 1
            for (; p != pend; p += BLK) {
                if (LDS) {
                    const double2 c0 = reinterpret_cast<const double2 *>(p)[0], c1 = reinterpret_cast<const double2 *>(p)[1];
                    const double2 c2 = reinterpret_cast<const double2 *>(p)[2], c3 = reinterpret_cast<const double2 *>(p)[3];
                    const double2 w01 = reinterpret_cast<const double2 *>(p)[4], w23 = reinterpret_cast<const double2 *>(p)[5];
                    well<SEED>(x, y, c0.x, c0.y, w01.x, gx, gy);
                    well<SEED>(x, y, c1.x, c1.y, w01.y, gx, gy);
                    well<SEED>(x, y, c2.x, c2.y, w23.x, gx, gy);
                    well<SEED>(x, y, c3.x, c3.y, w23.y, gx, gy);
                } else {
                    well<SEED>(x, y, rx[0], ry[0], rw[0], gx, gy);
                    well<SEED>(x, y, rx[1], ry[1], rw[1], gx, gy);
                    well<SEED>(x, y, rx[2], ry[2], rw[2], gx, gy);
                    well<SEED>(x, y, rx[3], ry[3], rw[3], gx, gy);
                }
            }
            x += 1e-3 * gx; y += 1e-3 * gy;      // a new evaluation point per pass, as in the integrator
        }
    }
    if (gx + gy == 12345.678) sink[0] = gx;
}

template <int V>
static void run(const char *name, int ctas_per_sm, const double *src, double *sink, int sms, double clk_hz)
{
    const int iters = 200, nblk = 50;            // 200 wells per pass
    // occupancy is set with dynamic shared memory: 227 KB / ctas_per_sm each
    const size_t smem = (size_t)(200 * 1024 / ctas_per_sm) & ~(size_t)15;
    cudaFuncSetAttribute(probe<V>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    cudaEvent_t a, b;
    cudaEventCreate(&a); cudaEventCreate(&b);
    const int grid = sms * ctas_per_sm;
    probe<V><<<grid, 128, smem>>>(10, nblk, src, sink);
    cudaEventRecord(a);
    probe<V><<<grid, 128, smem>>>(iters, nblk, src, sink);
    cudaEventRecord(b);
    cudaEventSynchronize(b);
    float ms = 0;
    cudaEventElapsedTime(&ms, a, b);
    const double cycles = ms * 1e-3 * clk_hz;
    const double wells_per_smsp = (double)iters * nblk * 4 * ctas_per_sm;      // 4 warps per CTA, one per SMSP
    printf("%-44s %2d warps/SMSP  %7.3f ms  %6.2f cycles per well per SMSP\n", name, ctas_per_sm, ms, cycles / wells_per_smsp);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) printf("  CUDA error: %s\n", cudaGetErrorString(e));
}

int main()
{
    cudaDeviceProp prop;
    cudaGetDeviceProperties(&prop, 0);
    int clk_khz = 0;
    cudaDeviceGetAttribute(&clk_khz, cudaDevAttrClockRate, 0);
    const double clk = clk_khz * 1e3;
    printf("%s, %d SMs, %.0f MHz (attribute; the run may boost differently)\n", prop.name, prop.multiProcessorCount, clk / 1e6);
    double h[50 * BLK];
    for (int i = 0; i < 50 * BLK; ++i) h[i] = 10.0 + 37.0 * (i % 13) + i;
    double *src, *sink;
    cudaMalloc(&src, sizeof(h)); cudaMalloc(&sink, 8);
    cudaMemcpy(src, h, sizeof(h), cudaMemcpyHostToDevice);
    const int sms = prop.multiProcessorCount;
    for (int c : {3, 6, 8, 12}) {
        run<0>("0 full: LDS + MUFU seed + 9 FP64", c, src, sink, sms, clk);
        run<1>("1 registers (no LDS) + MUFU + 9 FP64", c, src, sink, sms, clk);
        run<2>("2 LDS + integer seed + 9 FP64", c, src, sink, sms, clk);
        run<3>("3 registers + integer seed + 9 FP64", c, src, sink, sms, clk);
        run<4>("4 registers + 10 FP64, nothing else", c, src, sink, sms, clk);
        run<5>("5 LDS + 10 FP64", c, src, sink, sms, clk);
        run<6>("6 DFMA chains, 9 per 'well'", c, src, sink, sms, clk);
    }
    return 0;
}

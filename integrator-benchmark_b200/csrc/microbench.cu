// FMA-pipe microbenchmarks: the measured FP32 / FP64 non-tensor peaks that bench.py uses as
// roofline denominators for the FMA-bound kernels (MEASURED_PEAKS.json has HBM and tensor peaks only).
#include <cuda_runtime.h>

#include "lsi_internal.h"
#include "rng.cuh"

namespace {

template <typename T>
__global__ void __launch_bounds__(256) fma_kernel(int64_t iters, T seed, T* __restrict__ sink)
{
    T a0 = seed + (T)threadIdx.x, a1 = a0 + (T)1, a2 = a0 + (T)2, a3 = a0 + (T)3;
    T a4 = a0 + (T)4, a5 = a0 + (T)5, a6 = a0 + (T)6, a7 = a0 + (T)7;
    const T m = (T)0.999993, c = (T)1.0e-3;
    for (int64_t i = 0; i < iters; ++i) {
        a0 = a0 * m + c; a1 = a1 * m + c; a2 = a2 * m + c; a3 = a3 * m + c;
        a4 = a4 * m + c; a5 = a5 * m + c; a6 = a6 * m + c; a7 = a7 * m + c;
    }
    const T s = ((a0 + a1) + (a2 + a3)) + ((a4 + a5) + (a6 + a7));
    if (s == (T)-1.2345) sink[0] = s;  // never true; keeps the chains alive
}

// Dispatch test behind the toy kernels' roofline argument: per iteration 8 independent DFMAs interleaved with
// `N_INT` x 8 independent 32-bit integer multiply-adds (the instruction class of Philox).  If a warp-wide fp64
// instruction only occupied the 16-lane fp64 pipe for two cycles, the integer work would hide under it (time = the
// fp64-only time until N_INT > 1); if it holds the sub-partition's dispatch port for both cycles, every integer
// instruction adds a cycle (time = fp64 time x (1 + N_INT / 2)).
template <int N_INT>
__global__ void __launch_bounds__(256) dispatch_mix_kernel(int64_t iters, double seed, double* __restrict__ sink)
{
    double a0 = seed + (double)threadIdx.x, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3, a4 = a0 + 4, a5 = a0 + 5, a6 = a0 + 6, a7 = a0 + 7;
    unsigned k[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) k[j] = threadIdx.x * 2654435761u + j;
    const double m = 0.999993, c = 1.0e-3;
    for (int64_t i = 0; i < iters; ++i) {
        a0 = a0 * m + c; a1 = a1 * m + c; a2 = a2 * m + c; a3 = a3 * m + c;
        a4 = a4 * m + c; a5 = a5 * m + c; a6 = a6 * m + c; a7 = a7 * m + c;
#pragma unroll
        for (int r = 0; r < N_INT; ++r)
#pragma unroll
            for (int j = 0; j < 8; ++j) k[j] = k[j] * 0xD2511F53u + 0x9E3779B9u;
    }
    const double s = ((a0 + a1) + (a2 + a3)) + ((a4 + a5) + (a6 + a7));
    unsigned x = 0;
#pragma unroll
    for (int j = 0; j < 8; ++j) x ^= k[j];
    if (s == -1.2345 || x == 0x12345u) sink[0] = s + (double)x;  // never true; keeps the chains alive
}

// test hook: the lean fp64 functions of fastmath64.cuh, element-wise
__global__ void fastmath_kernel(int which, int64_t n, const double* __restrict__ x, double* __restrict__ out0, double* __restrict__ out1)
{
    __shared__ fm::Tables T;
    if (which >= 7) {
        fm::load_tables(T);
        __syncthreads();
    }
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    double a = 0.0, b = 0.0;
    if (which == 0) a = fm::log_pos(x[i]);
    else if (which == 7) a = fm::neg2log_tab(T, x[i]);
    else if (which == 8) fm::sincos2pi_u32_tab(T, (unsigned)x[i], a, b);
    else if (which == 10) a = fm::sin_tab(T, x[i]);
    else if (which == 11) a = fm::sqrt_pos_scaled(x[i], 1.0);
    else if (which == 9) {  // Box-Muller, table-driven vs library
        const unsigned ra = (unsigned)x[i], rb = ra * 2654435761u;
        double c, d;
        lsi::box_muller_tab(T, ra, rb, 1.0, a, b);
        lsi::box_muller(ra, rb, c, d);
        a -= c; b -= d;
    }
    else if (which == 1) fm::sincospi_small(x[i], a, b);
    else if (which == 2) {  // sin from the single-polynomial version, cos from the pair
        double s2;
        fm::sincos_medium(x[i], s2, b);
        a = fm::sin_medium(x[i]);
        if (fabs(a - s2) > 2.3e-16) a = 2.0;  // the two versions agree to an ulp
    }
    else if (which == 3) a = fm::sqrt_pos(x[i]);
    else if (which == 4) a = fm::rcp(x[i]);
    else if (which == 5) {  // x holds integers < 2^32: the uniforms of the Box-Muller pair
        a = fm::uniform_from_u32((unsigned)x[i], false);
        b = fm::uniform_from_u32((unsigned)x[i], true);
    } else {  // 6: Box-Muller, lean vs library, from the words (x[i], x[i] * 2654435761 mod 2^32)
        const unsigned ra = (unsigned)x[i], rb = ra * 2654435761u;
        double c, d;
        lsi::box_muller_lean(ra, rb, a, b);
        lsi::box_muller(ra, rb, c, d);
        a -= c; b -= d;
    }
    out0[i] = a;
    if (out1) out1[i] = b;
}

}  // namespace

// test hook for tests/test_toy_gpu.py::test_fastmath_accuracy (device pointers)
LSI_EXPORT int lsi_test_fastmath(int which, int64_t n, const double* x, double* out0, double* out1, void* stream)
{
    if (!x || !out0 || n < 0 || which < 0 || which > 11) return lsi_set_error(LSI_ERR_ARG, "lsi_test_fastmath: bad argument");
    if (n == 0) return LSI_OK;
    fastmath_kernel<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(which, n, x, out0, out1);
    LSI_CUDA_CHECK(cudaGetLastError());
    return LSI_OK;
}

// Launches blocks x 256 threads, each doing iters x 8 FMAs; flops = blocks * 256 * iters * 16.
LSI_EXPORT int lsi_microbench_fma(int precision, int64_t iters, int blocks, void* sink, void* stream)
{
    if (!sink || iters <= 0 || blocks <= 0) return lsi_set_error(LSI_ERR_ARG, "lsi_microbench_fma: bad argument");
    if (precision >= 10 && precision <= 14) {  // dispatch test: 8 DFMA + (precision - 10) x 8 IMAD per iteration
        switch (precision - 10) {
        case 0: dispatch_mix_kernel<0><<<blocks, 256, 0, (cudaStream_t)stream>>>(iters, 1.0, (double*)sink); break;
        case 1: dispatch_mix_kernel<1><<<blocks, 256, 0, (cudaStream_t)stream>>>(iters, 1.0, (double*)sink); break;
        case 2: dispatch_mix_kernel<2><<<blocks, 256, 0, (cudaStream_t)stream>>>(iters, 1.0, (double*)sink); break;
        case 3: dispatch_mix_kernel<3><<<blocks, 256, 0, (cudaStream_t)stream>>>(iters, 1.0, (double*)sink); break;
        default: dispatch_mix_kernel<4><<<blocks, 256, 0, (cudaStream_t)stream>>>(iters, 1.0, (double*)sink); break;
        }
    } else if (precision == LSI_PRECISION_DOUBLE)
        fma_kernel<double><<<blocks, 256, 0, (cudaStream_t)stream>>>(iters, 1.0, (double*)sink);
    else
        fma_kernel<float><<<blocks, 256, 0, (cudaStream_t)stream>>>(iters, 1.0f, (float*)sink);
    LSI_CUDA_CHECK(cudaGetLastError());
    return LSI_OK;
}

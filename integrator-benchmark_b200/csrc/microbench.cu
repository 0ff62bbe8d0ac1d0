// FMA-pipe microbenchmarks: the measured FP32 / FP64 non-tensor peaks that bench.py uses as
// roofline denominators for the FMA-bound kernels (MEASURED_PEAKS.json has HBM and tensor peaks only).
#include <cuda_runtime.h>

#include "lsi_internal.h"

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

}  // namespace

// Launches blocks x 256 threads, each doing iters x 8 FMAs; flops = blocks * 256 * iters * 16.
LSI_EXPORT int lsi_microbench_fma(int precision, int64_t iters, int blocks, void* sink, void* stream)
{
    if (!sink || iters <= 0 || blocks <= 0) return lsi_set_error(LSI_ERR_ARG, "lsi_microbench_fma: bad argument");
    if (precision == LSI_PRECISION_DOUBLE)
        fma_kernel<double><<<blocks, 256, 0, (cudaStream_t)stream>>>(iters, 1.0, (double*)sink);
    else
        fma_kernel<float><<<blocks, 256, 0, (cudaStream_t)stream>>>(iters, 1.0f, (float*)sink);
    LSI_CUDA_CHECK(cudaGetLastError());
    return LSI_OK;
}

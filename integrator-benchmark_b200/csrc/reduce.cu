// K5: work reductions.
//   * six-moment reduction for estimate_nonequilibrium_free_energy
//     (reference: benchmark/evaluation/analysis.py:8-17)
//   * row-wise log-mean-exp(-w) and its stderr heuristic
//     (reference: benchmark/evaluation/compare_near_eq_and_exact.py:62-97)
//   * two-level bootstrap of the nested estimator (same file :293-320)
// All sums are fp64, reduced in a fixed order (deterministic, independent of scheduling).
#include <cmath>
#include <cuda_runtime.h>

#include "lsi_internal.h"
#include "rng.cuh"

namespace {

constexpr int kThreads = 256;

__device__ __forceinline__ double warp_sum(double v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

__device__ __forceinline__ double warp_max(double v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}

// block-wide sum; result valid in thread 0 (and broadcast through smem to all)
__device__ __forceinline__ double block_sum(double v, double* sh)
{
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    v = warp_sum(v);
    __syncthreads();
    if (lane == 0) sh[warp] = v;
    __syncthreads();
    double s = 0.0;
    for (int w = 0; w < (int)(blockDim.x >> 5); ++w) s += sh[w];
    return s;
}

__device__ __forceinline__ double block_max(double v, double* sh)
{
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    v = warp_max(v);
    __syncthreads();
    if (lane == 0) sh[warp] = v;
    __syncthreads();
    double s = sh[0];
    for (int w = 1; w < (int)(blockDim.x >> 5); ++w) s = fmax(s, sh[w]);
    return s;
}

// partials: [n][8]; one block sums them column-wise in a fixed order
__global__ void __launch_bounds__(kThreads) finalize_moments_kernel(const double* __restrict__ partials, int n,
                                                                    double* __restrict__ moments)
{
    __shared__ double sh[kThreads / 32];
    for (int j = 0; j < 8; ++j) {
        double s = 0.0;
        for (int i = threadIdx.x; i < n; i += kThreads) s += partials[(int64_t)i * 8 + j];
        s = block_sum(s, sh);
        if (threadIdx.x == 0) moments[j] = (j < 7) ? s : 0.0;
    }
}

struct FinalizeBatch {
    const double* partials[8];
    double* moments[8];
};

// block b folds partials[b] exactly as finalize_moments_kernel does (same order, same result)
__global__ void __launch_bounds__(kThreads) finalize_moments_batch_kernel(const __grid_constant__ FinalizeBatch F, int n)
{
    __shared__ double sh[kThreads / 32];
    const double* __restrict__ partials = F.partials[blockIdx.x];
    double* __restrict__ moments = F.moments[blockIdx.x];
    for (int j = 0; j < 8; ++j) {
        double s = 0.0;
        for (int i = threadIdx.x; i < n; i += kThreads) s += partials[(int64_t)i * 8 + j];
        s = block_sum(s, sh);
        if (threadIdx.x == 0) moments[j] = (j < 7) ? s : 0.0;
    }
}

__global__ void __launch_bounds__(kThreads) moments_partials_kernel(const double* __restrict__ wF,
                                                                    const double* __restrict__ wR, int64_t n,
                                                                    double* __restrict__ partials)
{
    double m[7] = {0, 0, 0, 0, 0, 0, 0};
    for (int64_t i = (int64_t)blockIdx.x * kThreads + threadIdx.x; i < n; i += (int64_t)gridDim.x * kThreads) {
        const double a = wF[i], b = wR[i];
        if (isfinite(a) && isfinite(b)) {
            m[0] += 1.0; m[1] += a; m[2] += b; m[3] += a * a; m[4] += b * b; m[5] += a * b;
        } else {
            m[6] += 1.0;
        }
    }
    __shared__ double sh[kThreads / 32];
    for (int j = 0; j < 7; ++j) {
        const double s = block_sum(m[j], sh);
        if (threadIdx.x == 0) partials[(int64_t)blockIdx.x * 8 + j] = s;
    }
    if (threadIdx.x == 0) partials[(int64_t)blockIdx.x * 8 + 7] = 0.0;
}

// one block per row
__global__ void __launch_bounds__(kThreads) logmeanexp_kernel(const double* __restrict__ w,
                                                              const int64_t* __restrict__ offs, int64_t n_rows,
                                                              double* __restrict__ estimate,
                                                              double* __restrict__ stdev)
{
    __shared__ double sh[kThreads / 32];
    for (int64_t row = blockIdx.x; row < n_rows; row += gridDim.x) {
        const int64_t lo = offs[row], hi = offs[row + 1];
        const double n = (double)(hi - lo);
        double mx = -INFINITY;
        for (int64_t i = lo + threadIdx.x; i < hi; i += kThreads) mx = fmax(mx, -w[i]);
        mx = block_max(mx, sh);
        // rows whose largest term is +-inf (all works +inf, or a work of -inf): no shift, so that the sums come out as the
        // reference's naive formula gives them (0 -> log 0 = -inf, inf -> inf) instead of inf - inf = NaN
        if (!isfinite(mx)) mx = 0.0;
        double s1 = 0.0, s2 = 0.0;
        for (int64_t i = lo + threadIdx.x; i < hi; i += kThreads) {
            const double e = exp(-w[i] - mx);
            s1 += e;
            s2 += e * e;
        }
        s1 = block_sum(s1, sh);
        s2 = block_sum(s2, sh);
        if (threadIdx.x == 0) {
            const double mean = s1 / n;  // of exp(-w - mx)
            estimate[row] = log(mean) + mx;
            if (stdev) {
                double var = s2 / n - mean * mean;  // ddof = 0, scale exp(-2 mx) cancels in the ratio
                if (var < 0.0) var = 0.0;
                stdev[row] = sqrt(var) / (mean * sqrt(n));
            }
        }
    }
}

// one block per bootstrap replicate; thread t walks resampled rows t, t + 256, ...
// draws: row pick  = Philox(seed, replica = b, draw = i,           TAG_BOOT, unit 0).x
//        col picks = Philox(seed, replica = b, draw = i,           TAG_BOOT, unit 1 + (j >> 2)) lane j & 3
__global__ void __launch_bounds__(kThreads) bootstrap_kernel(const double* __restrict__ w,
                                                             const int64_t* __restrict__ offs, int64_t n_rows,
                                                             int n_boot, uint64_t seed, double* __restrict__ out)
{
    __shared__ double sh[kThreads / 32];
    for (int b = blockIdx.x; b < n_boot; b += gridDim.x) {
        double acc = 0.0;
        for (int64_t i = threadIdx.x; i < n_rows; i += kThreads) {
            const lsi::Philox4 pr = lsi::philox_block(seed, (uint64_t)b, (uint32_t)i, LSI_TAG_BOOT, 0u);
            const int64_t row = (int64_t)(((uint64_t)pr.x * (uint64_t)n_rows) >> 32);
            const int64_t lo = offs[row];
            const int64_t nc = offs[row + 1] - lo;
            // stable log-mean-exp over nc resampled columns, single pass with running max
            double mx = -INFINITY, s = 0.0;
            for (int64_t j = 0; j < nc; j += 4) {
                const lsi::Philox4 pc = lsi::philox_block(seed, (uint64_t)b, (uint32_t)i, LSI_TAG_BOOT,
                                                          1u + (uint32_t)(j >> 2));
                const uint32_t rr[4] = {pc.x, pc.y, pc.z, pc.w};
#pragma unroll
                for (int l = 0; l < 4; ++l) {
                    if (j + l < nc) {
                        const int64_t col = (int64_t)(((uint64_t)rr[l] * (uint64_t)nc) >> 32);
                        const double v = -w[lo + col];
                        if (v > mx) { s = s * exp(mx - v) + 1.0; mx = v; }
                        else s += exp(v - mx);
                    }
                }
            }
            acc += log(s / (double)nc) + mx;
        }
        acc = block_sum(acc, sh);
        if (threadIdx.x == 0) out[b] = acc / (double)n_rows;
        __syncthreads();
    }
}

}  // namespace

int lsi_finalize_moments(const double* partials, int n_partials, double* moments, void* stream)
{
    finalize_moments_kernel<<<1, kThreads, 0, (cudaStream_t)stream>>>(partials, n_partials, moments);
    LSI_CUDA_CHECK(cudaGetLastError());
    return LSI_OK;
}

int lsi_finalize_moments_batch(const double* const* partials, int n_partials, double* const* moments, int n, void* stream)
{
    if (n <= 0) return LSI_OK;
    if (n > 8) return lsi_set_error(LSI_ERR_ARG, "lsi_finalize_moments_batch: at most 8 at a time");
    FinalizeBatch F;
    for (int i = 0; i < 8; ++i) { F.partials[i] = partials[i < n ? i : 0]; F.moments[i] = moments[i < n ? i : 0]; }
    finalize_moments_batch_kernel<<<n, kThreads, 0, (cudaStream_t)stream>>>(F, n_partials);
    LSI_CUDA_CHECK(cudaGetLastError());
    return LSI_OK;
}

LSI_EXPORT int lsi_reduce_moments(const double* W_F, const double* W_R, int64_t n, double* moments,
                                  void* workspace, int64_t workspace_bytes, void* stream)
{
    if (!W_F || !W_R || !moments || n < 0) return lsi_set_error(LSI_ERR_ARG, "lsi_reduce_moments: bad argument");
    const int grid = (int)std::min<int64_t>(1024, std::max<int64_t>(1, (n + kThreads - 1) / kThreads));
    if (!workspace || workspace_bytes < (int64_t)grid * 8 * (int64_t)sizeof(double))
        return lsi_set_error(LSI_ERR_ARG, "lsi_reduce_moments: workspace too small (need %lld bytes)",
                             (long long)(grid * 8 * sizeof(double)));
    moments_partials_kernel<<<grid, kThreads, 0, (cudaStream_t)stream>>>(W_F, W_R, n, (double*)workspace);
    LSI_CUDA_CHECK(cudaGetLastError());
    return lsi_finalize_moments((const double*)workspace, grid, moments, stream);
}

// estimate_nonequilibrium_free_energy on the moment vector: np.mean, np.var (ddof = 0),
// np.cov(...)[0, 1] (ddof = 1), exactly as analysis.py:14-16 combines them.
LSI_EXPORT int lsi_neq_free_energy(const double m[8], double N_eff, double* dF, double* sq_unc)
{
    if (!m || !dF || !sq_unc) return lsi_set_error(LSI_ERR_ARG, "lsi_neq_free_energy: NULL argument");
    const double n = m[0];
    if (!(n > 0.0)) { *dF = NAN; *sq_unc = NAN; return LSI_OK; }
    if (!(N_eff > 0.0)) N_eff = n;
    const double mF = m[1] / n, mR = m[2] / n;
    const double varF = m[3] / n - mF * mF;
    const double varR = m[4] / n - mR * mR;
    const double cov = (m[5] - n * mF * mR) / (n - 1.0);
    *dF = 0.5 * (mF - mR);
    *sq_unc = (varF + varR - 2.0 * cov) / (4.0 * N_eff);
    return LSI_OK;
}

LSI_EXPORT int lsi_reduce_logmeanexp(const double* w, const int64_t* row_offsets, int64_t n_rows,
                                     double* estimate, double* stdev, void* stream)
{
    if (!w || !row_offsets || !estimate || n_rows < 0) return lsi_set_error(LSI_ERR_ARG, "lsi_reduce_logmeanexp: bad argument");
    if (n_rows == 0) return LSI_OK;
    const int grid = (int)std::min<int64_t>(n_rows, 148 * 8);
    logmeanexp_kernel<<<grid, kThreads, 0, (cudaStream_t)stream>>>(w, row_offsets, n_rows, estimate, stdev);
    LSI_CUDA_CHECK(cudaGetLastError());
    return LSI_OK;
}

LSI_EXPORT int lsi_bootstrap_nested(const double* w, const int64_t* row_offsets, int64_t n_rows, int32_t n_boot,
                                    uint64_t seed, double* out, void* stream)
{
    if (!w || !row_offsets || !out || n_rows <= 0 || n_boot < 0) return lsi_set_error(LSI_ERR_ARG, "lsi_bootstrap_nested: bad argument");
    if (n_boot == 0) return LSI_OK;
    const int grid = std::min<int>(n_boot, 148 * 8);
    bootstrap_kernel<<<grid, kThreads, 0, (cudaStream_t)stream>>>(w, row_offsets, n_rows, n_boot, seed, out);
    LSI_CUDA_CHECK(cudaGetLastError());
    return LSI_OK;
}

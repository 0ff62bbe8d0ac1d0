// K1: fused whole-protocol kernels for 1-DOF toy systems (quartic, double well).
//
// One thread = one replica.  A thread carries x, v, the cached force, the heat accumulator,
// the energy bookkeeping and its Philox noise block in registers through the WHOLE protocol
// (forward leg, midpoint velocity redraw, reverse leg) -- the loop the reference runs in
// Python around numba closures (benchmark/testsystems/low_dimensional_systems.py:160-184,
// benchmark/integrators/numba_integrators.py:17-225).  HBM traffic is one read of the start
// state and one write of the two works per replica; block-level fp64 partial moments for
// estimate_nonequilibrium_free_energy (benchmark/evaluation/analysis.py:8-17) leave with
// the same launch.
//
// Two code paths, same arithmetic:
//   * toy_fused_kernel<POT, CODE, NOPS, NOISE32>: the splitting string is a template
//     argument (2 bits per substep), so the substep sequence is unrolled at compile time,
//     force evaluations are cached across substeps that do not move x (statically), the
//     step loop is unrolled so that every noise lane is a fixed register, and substep
//     lengths are four scalars.  Instantiated for the strings the reference uses.
//   * toy_protocol_kernel: a substep interpreter for everything else (long multi-timestep
//     strings, explicit-fraction programs, traces, direct shadow work).  The program
//     travels as a __grid_constant__ parameter and is read from the constant bank with
//     warp-uniform indices.
#include <cuda_runtime.h>

#include <cstring>

#include "lsi_internal.h"
#include "rng.cuh"

// fp64 Box-Muller and the double well's sin / cos: lean versions of fastmath64.cuh (define LSI_TOY_LIBM for the
// CUDA math library instead; same formulas, results agree to ~1 ulp)
#if defined(LSI_TOY_LIBM)
#define LSI_BM64(T, ra, rb, sc, n0, n1) do { lsi::box_muller(ra, rb, n0, n1); n0 *= sc; n1 *= sc; } while (0)
#elif defined(LSI_TOY_POLY)
#define LSI_BM64(T, ra, rb, sc, n0, n1) do { lsi::box_muller_lean(ra, rb, n0, n1); n0 *= sc; n1 *= sc; } while (0)
#else
#define LSI_BM64(T, ra, rb, sc, n0, n1) lsi::box_muller_tab(T, ra, rb, sc, n0, n1)
#define LSI_TOY_TABLES 1
#endif

namespace {

constexpr int kMaxOps = 160;  // fits the 4 KB parameter space; longer programs use the global copy
#ifndef LSI_TOY_THREADS
#define LSI_TOY_THREADS 256
#endif
constexpr int kThreads = LSI_TOY_THREADS;

struct ToyProgram {
    int n_ops;
    int n_O;
    const unsigned char* ext_kind;  // used when n_ops > kMaxOps
    const double* ext_c0;
    const double* ext_c1;
    unsigned char kind[kMaxOps];
    double c0[kMaxOps];  // R: h      V: h / m     O: a
    double c1[kMaxOps];  //                        O: b * sqrt(kT / m)
};

struct ToyParams {
    lsi::PhiloxKeys keys;  // round keys of `seed`
    uint64_t seed, replica0;
    int64_t n;
    int n_steps, n_legs, marginal;
    double mass, kT, inv_kT, sigma, p0;
    double hR, hV, oa, oc;  // fused path: R length, V length / m, O coefficients a and b * sigma
    const double* x_cache;
    int64_t n_cache;
    const double* x0;
    const double* v0;
    double *W, *heat, *W_direct, *x_final, *v_final, *trace_x, *trace_v, *trace_w;
    int trace_rm;      // 1: traces replica-major, trace_w [leg][r][step], trace_x = (x, v) pairs [leg][r][step][2]
                       // 2: step-major, trace_x = (x, v) pairs [leg][step][r][2]
    double* partials;  // [gridDim.x][8]
};

// one trace record of replica r at step index s (0 = start of the leg) of `leg`; T = n_steps + 1
__device__ __forceinline__ void store_trace(const ToyParams& P, int leg, int64_t T, int64_t s, int64_t r, double x, double v,
                                            double w)
{
    if (P.trace_rm) {
        const int64_t idx = P.trace_rm == 1 ? ((int64_t)leg * P.n + r) * T + s : ((int64_t)leg * T + s) * P.n + r;
        if (P.trace_x) reinterpret_cast<double2*>(P.trace_x)[idx] = make_double2(x, v);
        if (P.trace_w) P.trace_w[idx] = w;
    } else {
        const int64_t idx = ((int64_t)leg * T + s) * P.n + r;
        if (P.trace_x) P.trace_x[idx] = x;
        if (P.trace_v) P.trace_v[idx] = v;
        if (P.trace_w) P.trace_w[idx] = w;
    }
}

template <int POT>
__device__ __forceinline__ double potential(const fm::Tables& T, double x, double p0)
{
    if (POT == LSI_POT_QUARTIC) {
        const double x2 = x * x;
        return x2 * x2;
    } else if (POT == LSI_POT_DOUBLE_WELL) {
        const double x2 = x * x;
#ifdef LSI_TOY_LIBM
        return x2 * x2 * x2 + 2.0 * cos(5.0 * (x + 1.0));
#elif defined(LSI_TOY_POLY)
        double s, c;
        fm::sincos_medium(fma(5.0, x, 5.0), s, c);
        return fma(x2 * x2, x2, 2.0 * c);
#else
        return fma(x2 * x2, x2, 2.0 * fm::cos_tab(T, fma(5.0, x, 5.0)));
#endif
    } else {
        return 0.5 * p0 * x * x;
    }
}

template <int POT>
__device__ __forceinline__ double force(const fm::Tables& T, double x, double p0)
{
    if (POT == LSI_POT_QUARTIC) {
        return -4.0 * x * x * x;
    } else if (POT == LSI_POT_DOUBLE_WELL) {
        const double x2 = x * x;
#ifdef LSI_TOY_LIBM
        return -(6.0 * x2 * x2 * x - 10.0 * sin(5.0 * (x + 1.0)));
#elif defined(LSI_TOY_POLY)
        return fma(-6.0 * x2 * x2, x, 10.0 * fm::sin_medium(fma(5.0, x, 5.0)));
#else
        return fma(-6.0 * x2 * x2, x, 10.0 * fm::sin_tab(T, fma(5.0, x, 5.0)));
#endif
    } else {
        return -p0 * x;
    }
}

// four normals of one block of a replica's scalar stream, each multiplied by `scale`
template <bool NOISE32>
__device__ __forceinline__ void normal_block(const fm::Tables& T, const lsi::PhiloxKeys& seed, uint64_t replica, uint32_t draw, uint32_t tag,
                                             double scale, double& n0, double& n1, double& n2, double& n3)
{
    const lsi::Philox4 r = lsi::philox_block(seed, replica, draw, tag, 0u);
    if (NOISE32) {
        float a, b, c, d;
        lsi::box_muller(r.x, r.y, a, b);
        lsi::box_muller(r.z, r.w, c, d);
        n0 = scale * (double)a; n1 = scale * (double)b; n2 = scale * (double)c; n3 = scale * (double)d;
    } else {
        LSI_BM64(T, r.x, r.y, scale, n0, n1);
        LSI_BM64(T, r.z, r.w, scale, n2, n3);
    }
}

// normal number k of a (replica, tag) scalar stream: block k >> 2, lane k & 3 (always fp64:
// start and midpoint velocities are drawn once per leg)
__device__ __forceinline__ double scalar_normal(const fm::Tables& T, const lsi::PhiloxKeys& seed, uint64_t replica, uint32_t tag, uint32_t k)
{
    const lsi::Philox4 q = lsi::philox_block(seed, replica, k >> 2, tag, 0u);
    double a, b;
    if ((k & 2u) == 0u) LSI_BM64(T, q.x, q.y, 1.0, a, b);
    else LSI_BM64(T, q.z, q.w, 1.0, a, b);
    return (k & 1u) ? b : a;
}

__device__ __forceinline__ double warp_sum(double v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// block partial of the six moments (+ count of non-finite pairs) -> partials[blockIdx.x][8]
__device__ __forceinline__ void block_moments(double* __restrict__ partials, bool active, bool finite,
                                              double wF, double wR)
{
    const bool ok = active && finite;
    double m[7];
    m[0] = ok ? 1.0 : 0.0;
    m[1] = ok ? wF : 0.0;
    m[2] = ok ? wR : 0.0;
    m[3] = ok ? wF * wF : 0.0;
    m[4] = ok ? wR * wR : 0.0;
    m[5] = ok ? wF * wR : 0.0;
    m[6] = (active && !finite) ? 1.0 : 0.0;
    __shared__ double sh[kThreads / 32][7];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int j = 0; j < 7; ++j) {
        const double s = warp_sum(m[j]);
        if (lane == 0) sh[warp][j] = s;
    }
    __syncthreads();
    if (threadIdx.x < 8) {
        double s = 0.0;
        if (threadIdx.x < 7)
            for (int w = 0; w < kThreads / 32; ++w) s += sh[w][threadIdx.x];
        partials[(int64_t)blockIdx.x * 8 + threadIdx.x] = s;
    }
}

// persistent variant: each thread adds its replicas' contributions to its own column of a shared accumulator
// (no shuffles per chunk); one block reduction at the end of the kernel
struct MomentAcc { double m[7][kThreads]; };

__device__ __forceinline__ void moments_add(MomentAcc& A, bool finite, double wF, double wR)
{
    const int t = threadIdx.x;
    if (finite) {
        A.m[0][t] += 1.0; A.m[1][t] += wF; A.m[2][t] += wR;
        A.m[3][t] += wF * wF; A.m[4][t] += wR * wR; A.m[5][t] += wF * wR;
    } else {
        A.m[6][t] += 1.0;
    }
}

__device__ __forceinline__ void moments_flush(const MomentAcc& A, double* __restrict__ partials)
{
    __shared__ double sh[kThreads / 32][7];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int j = 0; j < 7; ++j) {
        const double s = warp_sum(A.m[j][threadIdx.x]);
        if (lane == 0) sh[warp][j] = s;
    }
    __syncthreads();
    if (threadIdx.x < 8) {
        double s = 0.0;
        if (threadIdx.x < 7)
            for (int w = 0; w < kThreads / 32; ++w) s += sh[w][threadIdx.x];
        partials[(int64_t)blockIdx.x * 8 + threadIdx.x] = s;
    }
}

// ------------------------------------------------------------------------------------------
// compile-time splitting: CODE holds 2 bits per substep (substep i at bits 2i, 2i+1)
constexpr int code_op(uint32_t code, int i) { return (int)((code >> (2 * i)) & 3u); }
constexpr int code_count(uint32_t code, int nops, int kind)
{
    int n = 0;
    for (int i = 0; i < nops; ++i) n += (code_op(code, i) == kind);
    return n;
}
// is the force known at step entry in steady state?  yes iff a V follows the last R
constexpr bool code_entry_valid(uint32_t code, int nops)
{
    bool valid = false;
    for (int i = 0; i < nops; ++i) {
        if (code_op(code, i) == LSI_OP_R) valid = false;
        if (code_op(code, i) == LSI_OP_V) valid = true;
    }
    return valid;
}

template <int POT, uint32_t CODE, int NOPS>
struct FusedStep {
    // one application of the splitting; xi[] are this step's normals in substep order, already multiplied by b sigma
    template <int NO>
    static __device__ __forceinline__ void run(const fm::Tables& FT, double& x, double& v, double& f, double& heat,
                                               const ToyParams& P, double half_m, const double (&xi)[NO])
    {
        bool valid = code_entry_valid(CODE, NOPS);  // folded at compile time after unrolling
        int o = 0;
#pragma unroll
        for (int i = 0; i < NOPS; ++i) {
            constexpr uint32_t code = CODE;
            const int kind = (int)((code >> (2 * i)) & 3u);
            if (kind == LSI_OP_R) {
                x = x + P.hR * v;
                valid = false;
            } else if (kind == LSI_OP_V) {
                if (!valid) f = force<POT>(FT, x, P.p0);
                valid = true;
                v = v + P.hV * f;
            } else {
                // heat += m/2 (v_new^2 - v_old^2): three fp64 instructions instead of six
                const double v_old2 = v * v;
                v = fma(P.oa, v, xi[o]);
                heat = fma(half_m, fma(v, v, -v_old2), heat);
                ++o;
            }
        }
    }
};

#ifndef LSI_TOY_MINB
#define LSI_TOY_MINB 4
#endif
#ifndef LSI_TOY_PIPE
#define LSI_TOY_PIPE 1
#endif
template <int POT, uint32_t CODE, int NOPS, bool NOISE32, bool TRACE = false>
__global__ void __launch_bounds__(kThreads, LSI_TOY_MINB) toy_fused_kernel(const __grid_constant__ ToyParams P)
{
    constexpr int NO = code_count(CODE, NOPS, LSI_OP_O);  // normals per step
    constexpr int U = (NO == 0) ? 1 : ((4 % NO == 0) ? 4 / NO : 4);  // steps per noise refill when NO | 4
    constexpr bool ALIGNED = (NO > 0) && (4 % NO == 0);
    __shared__ fm::Tables FT;
    __shared__ MomentAcc ACC;
#ifdef LSI_TOY_TABLES
    fm::load_tables(FT);
#endif
#pragma unroll
    for (int j = 0; j < 7; ++j) ACC.m[j][threadIdx.x] = 0.0;
    __syncthreads();

    // persistent CTAs: the grid is (at most) one full wave; CTA b takes the 256-replica chunks b, b + gridDim.x, ...
    // (tables staged and moments reduced once per CTA, no launch tail)
    const int64_t n_chunks = (P.n + kThreads - 1) / kThreads;
    for (int64_t chunk = blockIdx.x; chunk < n_chunks; chunk += gridDim.x) {
        const int64_t r = chunk * kThreads + threadIdx.x;
        if (r >= P.n) continue;
        double wF = 0.0, wR = 0.0;
        bool finite = true;
        const uint64_t rid = P.replica0 + (uint64_t)r;
        double x = P.x0 ? P.x0[r] : P.x_cache[lsi::cache_index(P.seed, rid, P.n_cache)];
        // start-of-leg velocities of a scalar system: normal number `leg` of the replica's TAG_V0 stream, so the start
        // and the midpoint velocity of a two-leg protocol are the two halves of ONE Box-Muller pair
        double n_start, n_mid;
        {
            const lsi::Philox4 q = lsi::philox_block(P.keys, rid, 0u, LSI_TAG_V0, 0u);
            LSI_BM64(FT, q.x, q.y, 1.0, n_start, n_mid);
        }
        double v = P.v0 ? P.v0[r] : P.sigma * n_start;
        double heat = 0.0;
        const double half_m = 0.5 * P.mass;
        double f = code_entry_valid(CODE, NOPS) ? force<POT>(FT, x, P.p0) : 0.0;
        double pe = potential<POT>(FT, x, P.p0);  // x does not move between the legs: one evaluation per leg boundary

        for (int leg = 0; leg < P.n_legs; ++leg) {
            if (leg > 0 && P.marginal == LSI_MARGINAL_CONFIGURATION)
                v = P.sigma * (leg == 1 ? n_mid : scalar_normal(FT, P.keys, rid, LSI_TAG_V0, (uint32_t)leg));
            const double E0 = pe + half_m * v * v;
            const double Q0 = heat;
            const uint32_t leg_base = (uint32_t)leg << 26;
            const int64_t T = (int64_t)P.n_steps + 1;
            // cumulative work after a step, for the traces (one extra potential evaluation per step)
            auto record = [&](int64_t s_done) {
                const double w = P.trace_w ? ((potential<POT>(FT, x, P.p0) + half_m * v * v - E0) - (heat - Q0)) * P.inv_kT : 0.0;
                store_trace(P, leg, T, s_done, r, x, v, w);
            };
            if (TRACE) store_trace(P, leg, T, 0, r, x, v, 0.0);

            if (ALIGNED) {
                // one Philox block feeds exactly U steps: every lane index is a constant.  The block of the NEXT U steps
                // is generated in the same basic block as the current U steps (LSI_TOY_PIPE): the noise (integer Philox
                // rounds, then Box-Muller with 2-4 independent fp64 chains) does not depend on x and v, so the scheduler
                // interleaves it with the strictly sequential dynamics chain and a single warp keeps both the integer
                // and the fp64 pipe fed.  Every block is still computed exactly once (the last one without a prefetch).
                const int n_blk = (P.n_steps + U - 1) / U;
                if (n_blk > 0) {
                    double cur[4];
                    normal_block<NOISE32>(FT, P.keys, rid, leg_base, LSI_TAG_O, P.oc, cur[0], cur[1], cur[2], cur[3]);
                    int s = 0;
#if LSI_TOY_PIPE
#pragma unroll 2
                    for (uint32_t blk = 1; blk < (uint32_t)n_blk; ++blk, s += U) {
                        double nxt[4];
                        normal_block<NOISE32>(FT, P.keys, rid, leg_base | blk, LSI_TAG_O, P.oc, nxt[0], nxt[1], nxt[2], nxt[3]);
#pragma unroll
                        for (int u = 0; u < U; ++u) {
                            double xi[NO > 0 ? NO : 1];
#pragma unroll
                            for (int o = 0; o < NO; ++o) xi[o] = cur[u * NO + o];
                            FusedStep<POT, CODE, NOPS>::run(FT, x, v, f, heat, P, half_m, xi);
                            if (TRACE) record(s + u + 1);
                        }
#pragma unroll
                        for (int k = 0; k < 4; ++k) cur[k] = nxt[k];
                    }
#else
                    for (uint32_t blk = 1; blk < (uint32_t)n_blk; ++blk, s += U) {
#pragma unroll
                        for (int u = 0; u < U; ++u) {
                            double xi[NO > 0 ? NO : 1];
#pragma unroll
                            for (int o = 0; o < NO; ++o) xi[o] = cur[u * NO + o];
                            FusedStep<POT, CODE, NOPS>::run(FT, x, v, f, heat, P, half_m, xi);
                            if (TRACE) record(s + u + 1);
                        }
                        normal_block<NOISE32>(FT, P.keys, rid, leg_base | blk, LSI_TAG_O, P.oc, cur[0], cur[1], cur[2], cur[3]);
                    }
#endif
                    // last block: U steps, or fewer when n_steps is not a multiple of U
#pragma unroll
                    for (int u = 0; u < U; ++u) {
                        if (s + u < P.n_steps) {
                            double xi[NO > 0 ? NO : 1];
#pragma unroll
                            for (int o = 0; o < NO; ++o) xi[o] = cur[u * NO + o];
                            FusedStep<POT, CODE, NOPS>::run(FT, x, v, f, heat, P, half_m, xi);
                            if (TRACE) record(s + u + 1);
                        }
                    }
                }
            } else {
                // NO does not divide 4 (e.g. 3 O substeps): walk the stream lane by lane
                uint32_t q = 0;
                double n0 = 0, n1 = 0, n2 = 0, n3 = 0;
                for (int s = 0; s < P.n_steps; ++s) {
                    double xi[NO > 0 ? NO : 1];
#pragma unroll
                    for (int o = 0; o < NO; ++o) {
                        const uint32_t lane = q & 3u;
                        if (lane == 0u) normal_block<NOISE32>(FT, P.keys, rid, leg_base | (q >> 2), LSI_TAG_O, P.oc, n0, n1, n2, n3);
                        xi[o] = lane == 0u ? n0 : (lane == 1u ? n1 : (lane == 2u ? n2 : n3));
                        ++q;
                    }
                    FusedStep<POT, CODE, NOPS>::run(FT, x, v, f, heat, P, half_m, xi);
                    if (TRACE) record(s + 1);
                }
            }

            pe = potential<POT>(FT, x, P.p0);
            const double E1 = pe + half_m * v * v;
            const double w = ((E1 - E0) - (heat - Q0)) * P.inv_kT;
            P.W[(int64_t)leg * P.n + r] = w;
            if (P.heat) P.heat[(int64_t)leg * P.n + r] = heat - Q0;
            if (leg == 0) wF = w;
            if (leg == 1) wR = w;
            finite = finite && isfinite(w);
        }
        if (P.x_final) P.x_final[r] = x;
        if (P.v_final) P.v_final[r] = v;
        if (P.partials) moments_add(ACC, finite, wF, wR);
    }
    if (P.partials) moments_flush(ACC, P.partials);
}

// ------------------------------------------------------------------------------------------
// interpreter
template <bool EXT>
struct ProgRead {
    static __device__ __forceinline__ int kind(const ToyProgram& p, int i) { return EXT ? p.ext_kind[i] : p.kind[i]; }
    static __device__ __forceinline__ double c0(const ToyProgram& p, int i) { return EXT ? p.ext_c0[i] : p.c0[i]; }
    static __device__ __forceinline__ double c1(const ToyProgram& p, int i) { return EXT ? p.ext_c1[i] : p.c1[i]; }
};

template <int POT, bool DIRECT, bool TRACE, bool EXT, bool NOISE32>
__global__ void __launch_bounds__(kThreads) toy_protocol_kernel(const __grid_constant__ ToyProgram prog,
                                                                const __grid_constant__ ToyParams P)
{
    __shared__ fm::Tables FT;
#ifdef LSI_TOY_TABLES
    fm::load_tables(FT);
    __syncthreads();
#endif
    const int64_t r = (int64_t)blockIdx.x * kThreads + threadIdx.x;
    const bool active = r < P.n;
    double wF = 0.0, wR = 0.0;
    bool finite = true;

    if (active) {
        const uint64_t rid = P.replica0 + (uint64_t)r;
        double x = P.x0 ? P.x0[r] : P.x_cache[lsi::cache_index(P.seed, rid, P.n_cache)];
        double v = P.v0 ? P.v0[r] : P.sigma * scalar_normal(FT, P.keys, rid, LSI_TAG_V0, 0u);
        double heat = 0.0;
        const double half_m = 0.5 * P.mass;
        const int64_t T = (int64_t)P.n_steps + 1;

        for (int leg = 0; leg < P.n_legs; ++leg) {
            if (leg > 0 && P.marginal == LSI_MARGINAL_CONFIGURATION)
                v = P.sigma * scalar_normal(FT, P.keys, rid, LSI_TAG_V0, (uint32_t)leg);
            const double E0 = potential<POT>(FT, x, P.p0) + half_m * v * v;
            const double Q0 = heat;
            const uint32_t leg_base = (uint32_t)leg << 26;
            double wd = 0.0;
            uint32_t q = 0;  // O draws so far in this leg (warp-uniform)
            double n0 = 0, n1 = 0, n2 = 0, n3 = 0;
            if (TRACE) store_trace(P, leg, T, 0, r, x, v, 0.0);
            for (int s = 0; s < P.n_steps; ++s) {
                for (int i = 0; i < prog.n_ops; ++i) {
                    const int kind = ProgRead<EXT>::kind(prog, i);
                    const double c0 = ProgRead<EXT>::c0(prog, i);
                    if (kind == LSI_OP_R) {
                        if (DIRECT) {
                            const double pe_old = potential<POT>(FT, x, P.p0);
                            x = x + c0 * v;
                            wd += potential<POT>(FT, x, P.p0) - pe_old;
                        } else {
                            x = x + c0 * v;
                        }
                    } else if (kind == LSI_OP_V) {
                        const double v_old = v;
                        v = v + c0 * force<POT>(FT, x, P.p0);
                        if (DIRECT) wd += half_m * v * v - half_m * v_old * v_old;
                    } else {
                        const double c1 = ProgRead<EXT>::c1(prog, i);
                        const uint32_t lane = q & 3u;
                        if (lane == 0u) normal_block<NOISE32>(FT, P.keys, rid, leg_base | (q >> 2), LSI_TAG_O, 1.0, n0, n1, n2, n3);
                        ++q;
                        const double xi = lane == 0u ? n0 : (lane == 1u ? n1 : (lane == 2u ? n2 : n3));
                        const double ke_old = half_m * v * v;
                        v = c0 * v + c1 * xi;
                        heat += half_m * v * v - ke_old;
                    }
                }
                if (TRACE) {
                    const double w = P.trace_w ? ((potential<POT>(FT, x, P.p0) + half_m * v * v - E0) - (heat - Q0)) * P.inv_kT : 0.0;
                    store_trace(P, leg, T, s + 1, r, x, v, w);
                }
            }
            const double E1 = potential<POT>(FT, x, P.p0) + half_m * v * v;
            const double w = ((E1 - E0) - (heat - Q0)) * P.inv_kT;
            P.W[(int64_t)leg * P.n + r] = w;
            if (P.heat) P.heat[(int64_t)leg * P.n + r] = heat - Q0;
            if (DIRECT && P.W_direct) P.W_direct[(int64_t)leg * P.n + r] = wd * P.inv_kT;
            if (leg == 0) wF = w;
            if (leg == 1) wR = w;
            finite = finite && isfinite(w);
        }
        if (P.x_final) P.x_final[r] = x;
        if (P.v_final) P.v_final[r] = v;
    }
    if (P.partials) block_moments(P.partials, active, finite, wF, wR);
}

template <int POT>
__global__ void __launch_bounds__(kThreads) toy_metropolis_kernel(uint64_t seed, uint64_t replica0, int64_t n,
                                                                  int n_burn, double inv_kT, double p0,
                                                                  double* __restrict__ x_out)
{
    __shared__ fm::Tables FT;
#ifdef LSI_TOY_TABLES
    fm::load_tables(FT);
    __syncthreads();
#endif
    const int64_t c = (int64_t)blockIdx.x * kThreads + threadIdx.x;
    if (c >= n) return;
    const uint64_t rid = replica0 + (uint64_t)c;
    double n0, n1;
    {
        const lsi::Philox4 q = lsi::philox_block(seed, rid, 0u, LSI_TAG_EQ, 0u);
        lsi::box_muller(q.x, q.y, n0, n1);
    }
    double x = n0;
    double u_x = potential<POT>(FT, x, p0);
    for (int i = 1; i <= n_burn; ++i) {
        const lsi::Philox4 q = lsi::philox_block(seed, rid, (uint32_t)i, LSI_TAG_EQ, 0u);
        lsi::box_muller(q.x, q.y, n0, n1);
        const double accept_eps = ((double)q.z + 0.5) * 2.3283064365386962890625e-10;
        const double x_prop = x + n0;
        const double u_prop = potential<POT>(FT, x_prop, p0);
        if (exp(-(u_prop - u_x) * inv_kT) > accept_eps) { x = x_prop; u_x = u_prop; }
    }
    x_out[c] = x;
}

// ---- dispatch --------------------------------------------------------------------------
template <int POT, bool EXT, bool NOISE32>
cudaError_t launch_interp3(const ToyProgram& prog, const ToyParams& P, bool direct, bool trace, cudaStream_t st)
{
    const unsigned grid = (unsigned)((P.n + kThreads - 1) / kThreads);
    if (direct && trace) toy_protocol_kernel<POT, true, true, EXT, NOISE32><<<grid, kThreads, 0, st>>>(prog, P);
    else if (direct) toy_protocol_kernel<POT, true, false, EXT, NOISE32><<<grid, kThreads, 0, st>>>(prog, P);
    else if (trace) toy_protocol_kernel<POT, false, true, EXT, NOISE32><<<grid, kThreads, 0, st>>>(prog, P);
    else toy_protocol_kernel<POT, false, false, EXT, NOISE32><<<grid, kThreads, 0, st>>>(prog, P);
    return cudaGetLastError();
}

template <int POT>
cudaError_t launch_interp(const ToyProgram& prog, const ToyParams& P, bool direct, bool trace, bool ext,
                          bool noise32, cudaStream_t st)
{
    if (ext) return noise32 ? launch_interp3<POT, true, true>(prog, P, direct, trace, st)
                            : launch_interp3<POT, true, false>(prog, P, direct, trace, st);
    return noise32 ? launch_interp3<POT, false, true>(prog, P, direct, trace, st)
                   : launch_interp3<POT, false, false>(prog, P, direct, trace, st);
}

constexpr uint32_t pack(const char* s)
{
    uint32_t code = 0;
    int i = 0;
    for (; *s; ++s) {
        const uint32_t k = (*s == 'R') ? LSI_OP_R : (*s == 'V') ? LSI_OP_V : LSI_OP_O;
        code |= k << (2 * i);
        ++i;
    }
    return code;
}
constexpr int slen(const char* s)
{
    int n = 0;
    while (s[n]) ++n;
    return n;
}

// the strings the reference itself uses: six 5-letter palindromes (tests/test_shadow_work.py:47-55,
// evaluation/compare_near_eq_and_exact.py:21-25), the substep sequences of the four numba closures,
// "R V O" / "R O V", and the g-BAOAB strings of utilities/mts_utilities.py:generate_gbaoab_string
#define LSI_FUSED_STRINGS(X)                                                                       \
    X("OVRVO") X("ORVRO") X("RVOVR") X("ROVOR") X("VRORV") X("VOROV")                              \
    X("OVRRVO") X("VROORV") X("RVOOVR") X("RVO") X("ROV") X("VRRORRV") X("VRRRORRRV")

// grid of the persistent fused kernels: one full wave (resident CTAs per SM x SMs), or fewer when there are fewer chunks
unsigned fused_grid(int64_t n)
{
    int dev = 0, sms = 0;  // two attribute queries per launch (no cache: devices of one process may differ)
    if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || sms <= 0)
        sms = 148;
    const int64_t wave = (int64_t)sms * LSI_TOY_MINB;
    const int64_t chunks = (n + kThreads - 1) / kThreads;
    return (unsigned)(chunks < wave ? chunks : wave);
}

template <int POT, bool NOISE32>
bool launch_fused2(uint32_t code, int nops, bool trace, const ToyParams& P, cudaStream_t st, cudaError_t* err)
{
    const unsigned grid = fused_grid(P.n);
    if (trace && NOISE32) return false;  // traced mixed-precision runs go through the interpreter
#define X(S)                                                                                       \
    if (nops == slen(S) && code == pack(S)) {                                                      \
        if (trace) toy_fused_kernel<POT, pack(S), slen(S), false, true><<<grid, kThreads, 0, st>>>(P);  \
        else toy_fused_kernel<POT, pack(S), slen(S), NOISE32><<<grid, kThreads, 0, st>>>(P);       \
        *err = cudaGetLastError();                                                                 \
        return true;                                                                               \
    }
    LSI_FUSED_STRINGS(X)
#undef X
    return false;
}

bool launch_fused(int pot, bool noise32, uint32_t code, int nops, bool trace, const ToyParams& P, cudaStream_t st,
                  cudaError_t* err)
{
    switch (pot) {
    case LSI_POT_QUARTIC:
        return noise32 ? launch_fused2<LSI_POT_QUARTIC, true>(code, nops, trace, P, st, err)
                       : launch_fused2<LSI_POT_QUARTIC, false>(code, nops, trace, P, st, err);
    case LSI_POT_DOUBLE_WELL:
        return noise32 ? launch_fused2<LSI_POT_DOUBLE_WELL, true>(code, nops, trace, P, st, err)
                       : launch_fused2<LSI_POT_DOUBLE_WELL, false>(code, nops, trace, P, st, err);
    default:
        return noise32 ? launch_fused2<LSI_POT_HARMONIC, true>(code, nops, trace, P, st, err)
                       : launch_fused2<LSI_POT_HARMONIC, false>(code, nops, trace, P, st, err);
    }
}

}  // namespace

LSI_EXPORT int64_t lsi_protocol_workspace_bytes(int64_t n_replicas)
{
    if (n_replicas < 0) return 0;
    const int64_t blocks = (n_replicas + 31) / 32;  // upper bound over all kernels' block sizes
    return (blocks * 8 + 8) * (int64_t)sizeof(double) + 3 * 1024 * (int64_t)sizeof(double);
}

LSI_EXPORT int lsi_sample_equilibrium_scalar(const lsi_system* sys, double kT, int64_t n, int32_t n_burn,
                                             uint64_t seed, uint64_t replica0, double* x_out, void* stream)
{
    if (!sys || sys->kind != LSI_SYS_SCALAR || !x_out || n < 0 || n_burn < 0 || !(kT > 0.0))
        return lsi_set_error(LSI_ERR_ARG, "lsi_sample_equilibrium_scalar: bad argument");
    if (n == 0) return LSI_OK;
    cudaStream_t st = (cudaStream_t)stream;
    const unsigned grid = (unsigned)((n + kThreads - 1) / kThreads);
    switch (sys->potential) {
    case LSI_POT_QUARTIC:
        toy_metropolis_kernel<LSI_POT_QUARTIC><<<grid, kThreads, 0, st>>>(seed, replica0, n, n_burn, 1.0 / kT, sys->params[0], x_out);
        break;
    case LSI_POT_DOUBLE_WELL:
        toy_metropolis_kernel<LSI_POT_DOUBLE_WELL><<<grid, kThreads, 0, st>>>(seed, replica0, n, n_burn, 1.0 / kT, sys->params[0], x_out);
        break;
    default:
        toy_metropolis_kernel<LSI_POT_HARMONIC><<<grid, kThreads, 0, st>>>(seed, replica0, n, n_burn, 1.0 / kT, sys->params[0], x_out);
        break;
    }
    LSI_CUDA_CHECK(cudaGetLastError());
    return LSI_OK;
}

int lsi_toy_run_protocol(const lsi_system* sys, const lsi_program* prog, const lsi_protocol_args* a, void* stream)
{
    cudaStream_t st = (cudaStream_t)stream;
    if (a->n_replicas == 0) return LSI_OK;
    if (!a->W) return lsi_set_error(LSI_ERR_ARG, "lsi_run_protocol: W output is required");
    if (!a->x0 && !(a->x_cache && a->n_cache > 0))
        return lsi_set_error(LSI_ERR_ARG, "lsi_run_protocol: need x0 or a non-empty x_cache");
    if (a->precision != LSI_PRECISION_DOUBLE && a->precision != LSI_PRECISION_MIXED)
        return lsi_set_error(LSI_ERR_ARG, "unknown precision %d", a->precision);
    if (a->W_direct && !prog->measure_shadow_work)
        return lsi_set_error(LSI_ERR_ARG, "W_direct needs a program compiled with measure_shadow_work");
    if (a->n_legs > 64) return lsi_set_error(LSI_ERR_ARG, "at most 64 legs");
    for (const lsi_op& op : prog->ops)
        if (op.kind == LSI_OP_V && op.group >= 0)
            return lsi_set_error(LSI_ERR_UNSUPPORTED, "scalar systems have a single force group; got V%d", op.group);

    const int n_ops = (int)prog->ops.size();
    const double sigma = sqrt(a->kT / sys->mass);
    const bool noise32 = a->precision == LSI_PRECISION_MIXED;
    if ((int64_t)a->n_steps * (int64_t)(prog->n_O > 0 ? prog->n_O : 1) >= ((int64_t)1 << 28))
        return lsi_set_error(LSI_ERR_ARG, "n_steps * n_O exceeds the 2^28 draws a leg's noise stream holds");

    ToyParams P;
    P.seed = a->seed; P.replica0 = a->replica0; P.n = a->n_replicas;
    P.keys = lsi::philox_round_keys(a->seed);
    P.n_steps = a->n_steps; P.n_legs = a->n_legs; P.marginal = a->marginal;
    P.mass = sys->mass; P.kT = a->kT; P.inv_kT = 1.0 / a->kT; P.sigma = sigma; P.p0 = sys->params[0];
    P.hR = P.hV = P.oa = P.oc = 0.0;
    P.x_cache = a->x_cache; P.n_cache = a->n_cache; P.x0 = a->x0; P.v0 = a->v0;
    P.W = a->W; P.heat = a->heat; P.W_direct = a->W_direct; P.x_final = a->x_final; P.v_final = a->v_final;
    P.trace_x = a->trace_x; P.trace_v = a->trace_v; P.trace_w = a->trace_w;
    P.trace_rm = (a->flags & LSI_FLAG_TRACE_REPLICA_MAJOR) ? 1 : ((a->flags & LSI_FLAG_TRACE_XV_PAIRS) ? 2 : 0);
    if (P.trace_rm && a->trace_v) return lsi_set_error(LSI_ERR_ARG, "these trace layouts keep (x, v) pairs in trace_x; trace_v must be NULL");
    P.partials = nullptr;
    const int64_t grid = (a->n_replicas + kThreads - 1) / kThreads;
    if (a->moments) {
        if (a->n_legs != 2) return lsi_set_error(LSI_ERR_ARG, "moments need n_legs == 2");
        if (!a->workspace || a->workspace_bytes < lsi_protocol_workspace_bytes(a->n_replicas))
            return lsi_set_error(LSI_ERR_ARG, "workspace too small: need %lld bytes",
                                 (long long)lsi_protocol_workspace_bytes(a->n_replicas));
        P.partials = (double*)a->workspace;
    }
    const bool direct = a->W_direct != nullptr;
    const bool trace = a->trace_x || a->trace_v || a->trace_w;

    // ---- fused path: every substep of a kind has the same length (true for lsi_compile output)
    bool launched = false;
    cudaError_t err = cudaSuccess;
    if (!direct && n_ops <= 16) {
        uint32_t code = 0;
        bool uniform = true;
        double fr[3] = {-1, -1, -1}, oa = 0, ob = 0;
        for (int i = 0; i < n_ops && uniform; ++i) {
            const lsi_op& op = prog->ops[i];
            code |= (uint32_t)op.kind << (2 * i);
            if (fr[op.kind] < 0) fr[op.kind] = op.fraction;
            uniform = uniform && fr[op.kind] == op.fraction;
            if (op.kind == LSI_OP_O) {
                if (oa == 0 && ob == 0) { oa = op.a; ob = op.b; }
                uniform = uniform && oa == op.a && ob == op.b;
            }
        }
        if (uniform) {
            P.hR = prog->dt * (fr[LSI_OP_R] > 0 ? fr[LSI_OP_R] : 0.0);
            P.hV = prog->dt * (fr[LSI_OP_V] > 0 ? fr[LSI_OP_V] : 0.0) / sys->mass;
            P.oa = oa; P.oc = ob * sigma;
            launched = launch_fused(sys->potential, noise32, code, n_ops, trace, P, st, &err);
        }
    }

    void* ext = nullptr;
    if (!launched) {
        ToyProgram tp;
        tp.n_ops = n_ops;
        tp.n_O = prog->n_O;
        tp.ext_kind = nullptr; tp.ext_c0 = nullptr; tp.ext_c1 = nullptr;
        std::vector<unsigned char> kinds(n_ops);
        std::vector<double> c0(n_ops), c1(n_ops);
        for (int i = 0; i < n_ops; ++i) {
            const lsi_op& op = prog->ops[i];
            kinds[i] = (unsigned char)op.kind;
            if (op.kind == LSI_OP_R) { c0[i] = prog->dt * op.fraction; c1[i] = 0.0; }
            else if (op.kind == LSI_OP_V) { c0[i] = prog->dt * op.fraction / sys->mass; c1[i] = 0.0; }
            else { c0[i] = op.a; c1[i] = op.b * sigma; }
        }
        if (n_ops <= kMaxOps) {
            for (int i = 0; i < n_ops; ++i) { tp.kind[i] = kinds[i]; tp.c0[i] = c0[i]; tp.c1[i] = c1[i]; }
        } else {
            // long multi-timestep strings: the tables live in device memory, cached on the program handle
            // (lsi_program_stage: no allocation, copy or synchronisation after the first launch)
            const size_t nb = (size_t)n_ops * (2 * sizeof(double)) + (size_t)n_ops;
            std::vector<unsigned char> bytes(nb);
            std::memcpy(bytes.data(), c0.data(), n_ops * sizeof(double));
            std::memcpy(bytes.data() + n_ops * sizeof(double), c1.data(), n_ops * sizeof(double));
            std::memcpy(bytes.data() + 2 * n_ops * sizeof(double), kinds.data(), n_ops);
            const int src = lsi_program_stage(prog, bytes.data(), nb, &ext);
            if (src != LSI_OK) return src;
            double* d0 = (double*)ext;
            double* d1 = d0 + n_ops;
            unsigned char* dk = (unsigned char*)(d1 + n_ops);
            tp.ext_c0 = d0; tp.ext_c1 = d1; tp.ext_kind = dk;
        }
        switch (sys->potential) {
        case LSI_POT_QUARTIC: err = launch_interp<LSI_POT_QUARTIC>(tp, P, direct, trace, ext != nullptr, noise32, st); break;
        case LSI_POT_DOUBLE_WELL: err = launch_interp<LSI_POT_DOUBLE_WELL>(tp, P, direct, trace, ext != nullptr, noise32, st); break;
        case LSI_POT_HARMONIC: err = launch_interp<LSI_POT_HARMONIC>(tp, P, direct, trace, ext != nullptr, noise32, st); break;
        default: return lsi_set_error(LSI_ERR_ARG, "unknown potential kind %d", sys->potential);
        }
    }
    if (err != cudaSuccess) return lsi_set_error(LSI_ERR_CUDA, "toy kernel launch failed: %s", cudaGetErrorString(err));
    if (a->moments) return lsi_finalize_moments(P.partials, launched ? (int)fused_grid(a->n_replicas) : (int)grid, a->moments, stream);
    return LSI_OK;
}

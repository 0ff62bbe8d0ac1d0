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
#include <type_traits>
#include <utility>

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

template <int POT, class TAB>
__device__ __forceinline__ double potential(const TAB& T, double x, double p0)
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

template <int POT, class TAB>
__device__ __forceinline__ double force(const TAB& T, double x, double p0)
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
template <bool NOISE32, class TAB>
__device__ __forceinline__ void normal_block(const TAB& T, const lsi::PhiloxKeys& seed, uint64_t replica, uint32_t draw, uint32_t tag,
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

// start-of-leg velocity normal of leg >= 2 (legs 0 and 1 share the pair of the replica's start block, see rng.cuh):
// normal leg - 2 of the TAG_V0 stream that begins at draw 1
template <class TAB>
__device__ __forceinline__ double later_leg_normal(const TAB& T, const lsi::PhiloxKeys& seed, uint64_t replica, uint32_t leg)
{
    const uint32_t k = leg - 2u;
    const lsi::Philox4 q = lsi::philox_block(seed, replica, 1u + (k >> 2), LSI_TAG_V0, 0u);
    double a, b;
    if ((k & 2u) == 0u) LSI_BM64(T, q.x, q.y, 1.0, a, b);
    else LSI_BM64(T, q.z, q.w, 1.0, a, b);
    return (k & 1u) ? b : a;
}

// the replica's start block: start / midpoint velocity normals and the equilibrium-cache index
template <class TAB>
__device__ __forceinline__ void start_block(const TAB& T, const lsi::PhiloxKeys& seed, uint64_t replica, int64_t n_cache,
                                            double& n_start, double& n_mid, int64_t& index)
{
    const lsi::Philox4 q = lsi::philox_block(seed, replica, 0u, LSI_TAG_V0, 0u);
    LSI_BM64(T, q.x, q.y, 1.0, n_start, n_mid);
    index = lsi::index_from_word(q.z, n_cache);
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
        double n_start, n_mid;
        int64_t idx0;
        start_block(FT, P.keys, rid, P.n_cache, n_start, n_mid, idx0);
        double x = P.x0 ? P.x0[r] : P.x_cache[idx0];
        double v = P.v0 ? P.v0[r] : P.sigma * n_start;
        double heat = 0.0;
        const double half_m = 0.5 * P.mass;
        double f = code_entry_valid(CODE, NOPS) ? force<POT>(FT, x, P.p0) : 0.0;
        double pe = potential<POT>(FT, x, P.p0);  // x does not move between the legs: one evaluation per leg boundary

        for (int leg = 0; leg < P.n_legs; ++leg) {
            if (leg > 0 && P.marginal == LSI_MARGINAL_CONFIGURATION)
                v = P.sigma * (leg == 1 ? n_mid : later_leg_normal(FT, P.keys, rid, (uint32_t)leg));
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
// K1 (production path, no traces): the same protocol with
//   * conflict-free replicated tables in dynamic shared memory (fm::TabView: 80 KB per CTA);
//   * forces kept as g with F = c g (double well: g = sin(5 x + 5) - 0.6 x^5, c = 10; quartic: g = x^3, c = -4), the
//     constant folded into the kick length;
//   * the heat of a leg accumulated as h = sum over O substeps of (v_after^2 - v_before^2), multiplied by m / 2 once;
//     two O substeps with nothing between them (the end of one step and the start of the next in "O V R V O", the
//     "O O" of the reference's baoab closure) share their middle term, which cancels and is never computed;
//   * the last noise block of a leg generating only the Box-Muller pair the remaining steps consume;
//   * ILP replicas per thread, their chains interleaved by the compiler.
// Arithmetic differs from the plain formulas only in rounding (1e-16 relative per operation).
#ifndef LSI_TOY2_THREADS
#define LSI_TOY2_THREADS 1024
#endif
#ifndef LSI_TOY2_MINB
#define LSI_TOY2_MINB 1
#endif
#ifndef LSI_TOY2_ILP
#define LSI_TOY2_ILP 1
#endif
#ifndef LSI_TOY2_UNROLL
#define LSI_TOY2_UNROLL 2
#endif
constexpr int kF2Threads = LSI_TOY2_THREADS;
constexpr int kF2Unroll = LSI_TOY2_UNROLL;
constexpr size_t kF2Smem = (size_t)fm::TabView::kBytes + 7u * kF2Threads * sizeof(double);
static_assert(7u * kF2Threads * sizeof(double) >= (128 + 512) * sizeof(double2), "the accumulator region doubles as the table staging scratch");

template <int POT>
__device__ __forceinline__ double force_scale(double p0)
{
    return POT == LSI_POT_QUARTIC ? -4.0 : (POT == LSI_POT_DOUBLE_WELL ? 10.0 : -p0);
}

template <int POT, class TAB>
__device__ __forceinline__ double force_g(const TAB& T, double x)
{
    if (POT == LSI_POT_QUARTIC) {
        return x * x * x;
    } else if (POT == LSI_POT_DOUBLE_WELL) {
        const double x2 = x * x;
        return fma(-0.6, x2 * x2 * x, fm::sin_tab1(T, fma(5.0, x, 5.0)));
    } else {
        return x;
    }
}

constexpr bool code_is_O(uint32_t code, int nops, int i) { return code_op(code, ((i % nops) + nops) % nops) == LSI_OP_O; }

// two normals (first pair) or four of one O-stream block, each multiplied by `scale`
template <bool NOISE32, class TAB>
__device__ __forceinline__ void normal_block_part(const TAB& T, const lsi::PhiloxKeys& seed, uint64_t replica, uint32_t draw,
                                                  double scale, bool second_pair, double (&n)[4])
{
    const lsi::Philox4 r = lsi::philox_block(seed, replica, draw, LSI_TAG_O, 0u);
    if (NOISE32) {
        float a, b;
        lsi::box_muller(r.x, r.y, a, b);
        n[0] = scale * (double)a; n[1] = scale * (double)b;
        if (second_pair) {
            lsi::box_muller(r.z, r.w, a, b);
            n[2] = scale * (double)a; n[3] = scale * (double)b;
        }
    } else {
        LSI_BM64(T, r.x, r.y, scale, n[0], n[1]);
        if (second_pair) LSI_BM64(T, r.z, r.w, scale, n[2], n[3]);
    }
}

template <int POT, uint32_t CODE, int NOPS, int ILP>
struct FusedStep2 {
    // one application of the splitting to ILP replicas; xi[j][] are replica j's normals of this step in substep order,
    // already multiplied by b sigma; hVs = V length / m * force_scale
    template <int NO>
    static __device__ __forceinline__ void run(const fm::TabView& FT, double (&x)[ILP], double (&v)[ILP], double (&g)[ILP],
                                               double (&h)[ILP], double hR, double hVs, double oa, const double (&xi)[ILP][NO])
    {
        bool valid = code_entry_valid(CODE, NOPS);  // folded at compile time after unrolling
        int o = 0;
#pragma unroll
        for (int i = 0; i < NOPS; ++i) {
            constexpr uint32_t code = CODE;
            const int kind = (int)((code >> (2 * i)) & 3u);
            if (kind == LSI_OP_R) {
#pragma unroll
                for (int j = 0; j < ILP; ++j) x[j] = fma(hR, v[j], x[j]);
                valid = false;
            } else if (kind == LSI_OP_V) {
                if (!valid) {
#pragma unroll
                    for (int j = 0; j < ILP; ++j) g[j] = force_g<POT>(FT, x[j]);
                }
                valid = true;
#pragma unroll
                for (int j = 0; j < ILP; ++j) v[j] = fma(hVs, g[j], v[j]);
            } else {
                const bool sub = !code_is_O(CODE, NOPS, i - 1), add = !code_is_O(CODE, NOPS, i + 1);
#pragma unroll
                for (int j = 0; j < ILP; ++j) {
                    if (sub) h[j] = fma(-v[j], v[j], h[j]);
                    v[j] = fma(oa, v[j], xi[j][o]);
                    if (add) h[j] = fma(v[j], v[j], h[j]);
                }
                ++o;
            }
        }
    }
};

// the protocol of ONE program over this CTA's chunks; `acc` is the thread's column of the [7][kF2Threads] moment accumulator
template <int POT, uint32_t CODE, int NOPS, bool NOISE32, int ILP>
__device__ __forceinline__ void fused2_run(const fm::TabView& FT, double* acc, const ToyParams& P)
{
    constexpr int NO = code_count(CODE, NOPS, LSI_OP_O);  // normals per step: 1 or 2 for every instantiated string
    static_assert(NO == 1 || NO == 2, "the fused kernel expects one or two O substeps per step");
    constexpr int U = 4 / NO;                              // steps fed by one Philox block
    constexpr bool CYC_OO = code_is_O(CODE, NOPS, 0) && code_is_O(CODE, NOPS, NOPS - 1);
#pragma unroll
    for (int j = 0; j < 7; ++j) acc[j * kF2Threads] = 0.0;

    const double half_m = 0.5 * P.mass;
    const double hVs = P.hV * force_scale<POT>(P.p0);
    const int n_blk = (P.n_steps + U - 1) / U;
    const int rem = P.n_steps - (n_blk - 1) * U;       // steps fed by the last block (1..U)
    const bool last_needs_second = rem * NO > 2;
    // persistent CTAs: CTA b takes the (ILP x kF2Threads)-replica chunks b, b + gridDim.x, ...
    const int64_t per = (int64_t)kF2Threads * ILP;
    const int64_t n_chunks = (P.n + per - 1) / per;
    for (int64_t chunk = blockIdx.x; chunk < n_chunks; chunk += gridDim.x) {
        int64_t r[ILP];
        bool act[ILP];
        uint64_t rid[ILP];
        double x[ILP], v[ILP], g[ILP], h[ILP], pe[ILP], n_mid[ILP], wF[ILP], wR[ILP];
        bool finite[ILP];
#pragma unroll
        for (int j = 0; j < ILP; ++j) {
            r[j] = chunk * per + (int64_t)j * kF2Threads + threadIdx.x;
            act[j] = r[j] < P.n;
            if (!act[j]) r[j] = P.n - 1;  // computed like a live replica, never stored
            rid[j] = P.replica0 + (uint64_t)r[j];
            double n_start;
            int64_t idx0;
            start_block(FT, P.keys, rid[j], P.n_cache, n_start, n_mid[j], idx0);
            x[j] = P.x0 ? P.x0[r[j]] : P.x_cache[idx0];
            v[j] = P.v0 ? P.v0[r[j]] : P.sigma * n_start;
            wF[j] = wR[j] = 0.0;
            finite[j] = true;
        }
#pragma unroll
        for (int j = 0; j < ILP; ++j) {
            g[j] = code_entry_valid(CODE, NOPS) ? force_g<POT>(FT, x[j]) : 0.0;
            pe[j] = potential<POT>(FT, x[j], P.p0);  // x does not move between the legs: one evaluation per leg boundary
        }

        for (int leg = 0; leg < P.n_legs; ++leg) {
            double E0[ILP];
#pragma unroll
            for (int j = 0; j < ILP; ++j) {
                if (leg > 0 && P.marginal == LSI_MARGINAL_CONFIGURATION)
                    v[j] = P.sigma * (leg == 1 ? n_mid[j] : later_leg_normal(FT, P.keys, rid[j], (uint32_t)leg));
                const double v2 = v[j] * v[j];
                E0[j] = fma(half_m, v2, pe[j]);
                h[j] = CYC_OO ? -v2 : 0.0;  // the first O of the leg has no O before it
            }
            const uint32_t leg_base = (uint32_t)leg << 26;
            if (n_blk > 0) {
                double cur[ILP][4];
#pragma unroll
                for (int j = 0; j < ILP; ++j)
                    normal_block_part<NOISE32>(FT, P.keys, rid[j], leg_base, P.oc, n_blk > 1 || last_needs_second, cur[j]);
                auto steps = [&](int count) {
#pragma unroll
                    for (int u = 0; u < U; ++u) {
                        if (u < count) {
                            double xi[ILP][NO];
#pragma unroll
                            for (int j = 0; j < ILP; ++j)
#pragma unroll
                                for (int o = 0; o < NO; ++o) xi[j][o] = cur[j][u * NO + o];
                            FusedStep2<POT, CODE, NOPS, ILP>::run(FT, x, v, g, h, P.hR, hVs, P.oa, xi);
                        }
                    }
                };
                // the block of the NEXT U steps is generated in the same basic block as the current U steps: the noise
                // (integer Philox rounds, then Box-Muller) does not depend on x and v and interleaves with the dynamics chain
#pragma unroll kF2Unroll
                for (int blk = 1; blk < n_blk - 1; ++blk) {
                    double nxt[ILP][4];
#pragma unroll
                    for (int j = 0; j < ILP; ++j) normal_block_part<NOISE32>(FT, P.keys, rid[j], leg_base | (uint32_t)blk, P.oc, true, nxt[j]);
                    steps(U);
#pragma unroll
                    for (int j = 0; j < ILP; ++j)
#pragma unroll
                        for (int k = 0; k < 4; ++k) cur[j][k] = nxt[j][k];
                }
                if (n_blk > 1) {
                    double nxt[ILP][4];
#pragma unroll
                    for (int j = 0; j < ILP; ++j)
                        normal_block_part<NOISE32>(FT, P.keys, rid[j], leg_base | (uint32_t)(n_blk - 1), P.oc, last_needs_second, nxt[j]);
                    steps(U);
#pragma unroll
                    for (int j = 0; j < ILP; ++j)
#pragma unroll
                        for (int k = 0; k < 4; ++k) cur[j][k] = nxt[j][k];
                }
                steps(rem);
            }
#pragma unroll
            for (int j = 0; j < ILP; ++j) {
                pe[j] = potential<POT>(FT, x[j], P.p0);
                const double v2 = v[j] * v[j];
                const double E1 = fma(half_m, v2, pe[j]);
                const double heat = half_m * (CYC_OO ? h[j] + v2 : h[j]);  // the last O of the leg has no O after it
                const double w = ((E1 - E0[j]) - heat) * P.inv_kT;
                if (act[j]) {
                    P.W[(int64_t)leg * P.n + r[j]] = w;
                    if (P.heat) P.heat[(int64_t)leg * P.n + r[j]] = heat;
                }
                if (leg == 0) wF[j] = w;
                if (leg == 1) wR[j] = w;
                finite[j] = finite[j] && isfinite(w);
            }
        }
#pragma unroll
        for (int j = 0; j < ILP; ++j) {
            if (!act[j]) continue;
            if (P.x_final) P.x_final[r[j]] = x[j];
            if (P.v_final) P.v_final[r[j]] = v[j];
            if (P.partials) {
                if (finite[j]) {
                    acc[0] += 1.0; acc[kF2Threads] += wF[j]; acc[2 * kF2Threads] += wR[j];
                    acc[3 * kF2Threads] += wF[j] * wF[j]; acc[4 * kF2Threads] += wR[j] * wR[j]; acc[5 * kF2Threads] += wF[j] * wR[j];
                } else {
                    acc[6 * kF2Threads] += 1.0;
                }
            }
        }
    }
    if (P.partials) {
        __shared__ double sh[kF2Threads / 32][7];
        const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
        __syncthreads();  // a previous program's flush (batch kernel) has been read
#pragma unroll
        for (int j = 0; j < 7; ++j) {
            const double s = warp_sum(acc[j * kF2Threads]);
            if (lane == 0) sh[warp][j] = s;
        }
        __syncthreads();
        if (threadIdx.x < 8) {
            double s = 0.0;
            if (threadIdx.x < 7)
                for (int w = 0; w < kF2Threads / 32; ++w) s += sh[w][threadIdx.x];
            P.partials[(int64_t)blockIdx.x * 8 + threadIdx.x] = s;
        }
    }
}

template <int POT, uint32_t CODE, int NOPS, bool NOISE32, int ILP>
__global__ void __launch_bounds__(kF2Threads, LSI_TOY2_MINB) toy_fused2_kernel(const __grid_constant__ ToyParams P)
{
    extern __shared__ __align__(16) unsigned char toy2_smem[];
    // the moment accumulators' region doubles as the staging scratch of the tables
    const fm::TabView FT = fm::load_tables_replicated(toy2_smem, toy2_smem + fm::TabView::kBytes);
    double* acc = reinterpret_cast<double*>(toy2_smem + fm::TabView::kBytes) + threadIdx.x;  // column of [7][kF2Threads]
    fused2_run<POT, CODE, NOPS, NOISE32, ILP>(FT, acc, P);
}

// Several programs of one FAMILY (a compile-time list of splitting strings) over the same ensemble in ONE launch: slot i
// of the family runs when bit i of `enabled` is set, with its own lengths, coefficients and output pointers in P[i].
// Tables are staged once, the CTAs never drain between programs, and a CTA that finishes a program early starts the next.
constexpr int kBatchMax = 6;
struct ToyBatchParams {
    unsigned enabled;
    ToyParams P[kBatchMax];
};

template <int POT, bool NOISE32, int ILP, int NOPS, uint32_t... CODES>
__global__ void __launch_bounds__(kF2Threads, LSI_TOY2_MINB) toy_fused2_batch_kernel(const __grid_constant__ ToyBatchParams B)
{
    extern __shared__ __align__(16) unsigned char toy2_smem[];
    const fm::TabView FT = fm::load_tables_replicated(toy2_smem, toy2_smem + fm::TabView::kBytes);
    double* acc = reinterpret_cast<double*>(toy2_smem + fm::TabView::kBytes) + threadIdx.x;
    int slot = 0;
    // a fold over the family: every body is inlined with its slot's parameters at constant offsets of the parameter bank
    ((((B.enabled >> slot) & 1u) ? fused2_run<POT, CODES, NOPS, NOISE32, ILP>(FT, acc, B.P[slot]) : (void)0, ++slot), ...);
}

// ------------------------------------------------------------------------------------------
// K1-family: ONE replica per thread pushed through EVERY program of a family at once.  The programs of a batch share the
// ensemble (same seed, same global replica ids, same start states), hence the replica's start block, its gathered start
// position, the potential there AND its whole O-noise stream (Philox blocks + Box-Muller pairs, more than half of all fp64
// instructions of a single-program launch): all of that is computed once per replica and consumed by all programs.  The
// programs advance in lockstep over the noise blocks: a program with two O substeps per step consumes a block in two steps,
// one with a single O substep in four; each keeps its own (x, v, g, h).  Besides removing the redundant work this gives
// every thread as many independent dependency chains as there are programs (the single-program kernel has one).
// Per-program arithmetic is the single-program kernel's, instruction for instruction: works are bit-identical.
// Requires (checked on the host): every slot of the family enabled, and all programs with the same number of O
// substeps share (a, b sigma) -- true when they were compiled with the same time step and collision rate.
// threads per CTA (one CTA per SM) by the number of programs a thread carries: six chains per thread need 164 registers
// (384 threads); two or three chains leave room for more warps
#ifndef LSI_TOY3_THREADS
#define LSI_TOY3_THREADS 384
#endif
#ifndef LSI_TOY3_THREADS_NP3
#define LSI_TOY3_THREADS_NP3 512
#endif
#ifndef LSI_TOY3_THREADS_NP2
#define LSI_TOY3_THREADS_NP2 1024
#endif
#ifndef LSI_TOY3_PIPE
#define LSI_TOY3_PIPE 0
#endif
constexpr int family_threads(int np) { return np >= 4 ? LSI_TOY3_THREADS : (np == 3 ? LSI_TOY3_THREADS_NP3 : LSI_TOY3_THREADS_NP2); }

template <int NOPS, uint32_t... CODES>
struct Family {
    static constexpr int NP = (int)sizeof...(CODES);
    static constexpr uint32_t codes[NP] = {CODES...};
    static constexpr int no(int p) { return code_count(codes[p], NOPS, LSI_OP_O); }
    static constexpr bool has_class(int n_o)
    {
        for (int p = 0; p < NP; ++p)
            if (no(p) == n_o) return true;
        return false;
    }
    static constexpr int first_of_class(int n_o)
    {
        for (int p = 0; p < NP; ++p)
            if (no(p) == n_o) return p;
        return 0;
    }
};

template <class F, int... Is>
__device__ __forceinline__ void for_each_slot(F&& f, std::integer_sequence<int, Is...>)
{
    (f(std::integral_constant<int, Is>{}), ...);
}

constexpr int kF3Acc = 6;  // count, sum wF, sum wR, sum wF^2, sum wR^2, sum wF wR per program (non-finite pairs = replicas seen - count)
constexpr size_t family_smem(int np) { return (size_t)fm::TabView::kBytes + (size_t)np * kF3Acc * family_threads(np) * sizeof(double); }

// the shared part of one Box-Muller pair, and its completion for one scale (b sigma of a class of programs): the same
// instruction sequence as box_muller_tab / the fp32 path of normal_block_part, split where the scale enters
struct PairCore { double t, q, c, s; };

template <bool NOISE32, class TAB>
__device__ __forceinline__ PairCore pair_core(const TAB& T, uint32_t ra, uint32_t rb)
{
    PairCore k;
    if (NOISE32) {
        float a, b;
        lsi::box_muller(ra, rb, a, b);
        k.t = (double)a; k.q = (double)b; k.c = 0.0; k.s = 0.0;
    } else {
        const double x = fm::neg2log_tab(T, fm::uniform_from_u32(ra, false));
        double y;
        asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
        k.t = x * y;
        const double e = fma(-k.t, y, 1.0);
        k.q = e * fma(e, 0.375, 0.5);
        fm::sincos2pi_u32_tab(T, rb, k.s, k.c);
    }
    return k;
}

template <bool NOISE32>
__device__ __forceinline__ void pair_scaled(const PairCore& k, double scale, double& n0, double& n1)
{
    if (NOISE32) {
        n0 = scale * k.t; n1 = scale * k.q;
    } else {
        const double ts = k.t * scale;
        const double rad = fma(ts, k.q, ts);
        n0 = rad * k.c; n1 = rad * k.s;
    }
}

template <int POT, bool NOISE32, int NOPS, uint32_t... CODES>
__global__ void __launch_bounds__(family_threads((int)sizeof...(CODES)), 1) toy_family_kernel(const __grid_constant__ ToyBatchParams B)
{
    using Fam = Family<NOPS, CODES...>;
    constexpr int NP = Fam::NP;
    constexpr int kF3Threads = family_threads(NP);
    constexpr bool C1 = Fam::has_class(1), C2 = Fam::has_class(2);
    constexpr int S1 = Fam::first_of_class(1), S2 = Fam::first_of_class(2);
    static_assert(C1 || C2, "every program of a family has one or two O substeps per step");
    const auto slots = std::make_integer_sequence<int, NP>{};
    extern __shared__ __align__(16) unsigned char toy3_smem[];
    const fm::TabView FT = fm::load_tables_replicated(toy3_smem, toy3_smem + fm::TabView::kBytes);
    double* acc = reinterpret_cast<double*>(toy3_smem + fm::TabView::kBytes) + threadIdx.x;  // [NP][kF3Acc][kF3Threads]
#pragma unroll
    for (int j = 0; j < NP * kF3Acc; ++j) acc[j * kF3Threads] = 0.0;
    int n_seen = 0;  // replicas this thread has finished
    __syncthreads();

    const ToyParams& P0 = B.P[0];  // the ensemble: identical in every slot
    const double half_m = 0.5 * P0.mass;
    const int n_steps = P0.n_steps;
    const int n_blk = C2 ? (n_steps + 1) / 2 : (n_steps + 3) / 4;  // noise blocks of a leg: the class with two O substeps consumes the most
    const double oc1 = B.P[S1].oc, oc2 = B.P[S2].oc;
    // normals of one block, already multiplied by the b sigma of each class
    struct Normals { double n1[4], n2[4]; };
    auto generate_from = [&](const lsi::Philox4& r, bool second_pair, Normals& N) {
        const PairCore a = pair_core<NOISE32>(FT, r.x, r.y);
        if (C1) pair_scaled<NOISE32>(a, oc1, N.n1[0], N.n1[1]);
        if (C2) pair_scaled<NOISE32>(a, oc2, N.n2[0], N.n2[1]);
        if (second_pair) {
            const PairCore b = pair_core<NOISE32>(FT, r.z, r.w);
            if (C1) pair_scaled<NOISE32>(b, oc1, N.n1[2], N.n1[3]);
            if (C2) pair_scaled<NOISE32>(b, oc2, N.n2[2], N.n2[3]);
        }
    };
    auto generate = [&](uint64_t rid, uint32_t draw, bool second_pair, Normals& N) {
        generate_from(lsi::philox_block(P0.keys, rid, draw, LSI_TAG_O, 0u), second_pair, N);
    };

    const int64_t n_chunks = (P0.n + kF3Threads - 1) / kF3Threads;
    for (int64_t chunk = blockIdx.x; chunk < n_chunks; chunk += gridDim.x) {
        const int64_t r = chunk * kF3Threads + threadIdx.x;
        if (r >= P0.n) continue;
        const uint64_t rid = P0.replica0 + (uint64_t)r;
        double n_start, n_mid;
        int64_t idx0;
        start_block(FT, P0.keys, rid, P0.n_cache, n_start, n_mid, idx0);
        const double x0 = P0.x0 ? P0.x0[r] : P0.x_cache[idx0];
        const double v0 = P0.v0 ? P0.v0[r] : P0.sigma * n_start;
        const double pe0 = potential<POT>(FT, x0, P0.p0);
        const double g0 = force_g<POT>(FT, x0);
        double x[NP][1], v[NP][1], g[NP][1], h[NP][1], pe[NP], wF[NP];
#pragma unroll
        for (int p = 0; p < NP; ++p) { x[p][0] = x0; v[p][0] = v0; g[p][0] = g0; pe[p] = pe0; wF[p] = 0.0; }

        for (int leg = 0; leg < P0.n_legs; ++leg) {
            double E0[NP];
            if (leg > 0 && P0.marginal == LSI_MARGINAL_CONFIGURATION) {
                const double vm = P0.sigma * (leg == 1 ? n_mid : later_leg_normal(FT, P0.keys, rid, (uint32_t)leg));
#pragma unroll
                for (int p = 0; p < NP; ++p) v[p][0] = vm;
            }
            for_each_slot([&](auto ic) {
                constexpr int p = decltype(ic)::value;
                constexpr uint32_t CODE = Fam::codes[p];
                const double v2 = v[p][0] * v[p][0];
                E0[p] = fma(half_m, v2, pe[p]);
                h[p][0] = (code_is_O(CODE, NOPS, 0) && code_is_O(CODE, NOPS, NOPS - 1)) ? -v2 : 0.0;
            }, slots);
            const uint32_t leg_base = (uint32_t)leg << 26;
            // steps [2 b + u] of the two-O class and [4 b + u] of the one-O class, for one u
            auto step_class = [&](auto n_o, const Normals& N, int u) {
                constexpr int NO = decltype(n_o)::value;
                for_each_slot([&](auto ic) {
                    constexpr int p = decltype(ic)::value;
                    constexpr uint32_t CODE = Fam::codes[p];
                    if constexpr (code_count(CODE, NOPS, LSI_OP_O) == NO) {
                        double xi[1][NO];
#pragma unroll
                        for (int o = 0; o < NO; ++o) xi[0][o] = NO == 1 ? N.n1[u] : N.n2[2 * u + o];
                        FusedStep2<POT, CODE, NOPS, 1>::run(FT, x[p], v[p], g[p], h[p], B.P[p].hR, B.P[p].hV * force_scale<POT>(P0.p0),
                                                            B.P[p].oa, xi);
                    }
                }, slots);
            };
            // `wn`: with LSI_TOY3_PIPE == 2 the Philox words of the NEXT block are computed inside the first basic block of this
            // one (whichever branch runs at u = 0): ten dependent integer rounds that need no fp64 slot and no result of the
            // dynamics, so they issue in the shadow of the block's fp64 instructions (four registers, not a second set of normals)
            auto block_steps = [&](int b, const Normals& N, lsi::Philox4& wn) {
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    const bool a2 = C2 && u < 2 && 2 * b + u < n_steps;
                    const bool a1 = C1 && 4 * b + u < n_steps;
                    if (a2 && a1) {  // one basic block: all programs' chains interleave
                        if (LSI_TOY3_PIPE == 2 && u == 0) wn = lsi::philox_block(P0.keys, rid, leg_base | (uint32_t)(b + 1), LSI_TAG_O, 0u);
                        step_class(std::integral_constant<int, 2>{}, N, u);
                        step_class(std::integral_constant<int, 1>{}, N, u);
                    } else if (a2) {
                        if (LSI_TOY3_PIPE == 2 && u == 0) wn = lsi::philox_block(P0.keys, rid, leg_base | (uint32_t)(b + 1), LSI_TAG_O, 0u);
                        step_class(std::integral_constant<int, 2>{}, N, u);
                    } else if (a1) {
                        if (LSI_TOY3_PIPE == 2 && u == 0) wn = lsi::philox_block(P0.keys, rid, leg_base | (uint32_t)(b + 1), LSI_TAG_O, 0u);
                        step_class(std::integral_constant<int, 1>{}, N, u);
                    }
                }
            };
            // does block b need its second Box-Muller pair?  (normals 2, 3: the one-O class' steps 4 b + 2, 4 b + 3, the
            // two-O class' step 2 b + 1)
            auto needs_second = [&](int b) { return (C2 && 2 * b + 1 < n_steps) || (C1 && 4 * b + 2 < n_steps); };
            if (n_blk > 0) {
                Normals cur;
                generate(rid, leg_base, needs_second(0), cur);
                lsi::Philox4 wn = {0u, 0u, 0u, 0u};
#if LSI_TOY3_PIPE == 1
                for (int b = 0; b + 1 < n_blk; ++b) {
                    Normals nxt;
                    generate(rid, leg_base | (uint32_t)(b + 1), needs_second(b + 1), nxt);
                    block_steps(b, cur, wn);
                    cur = nxt;
                }
                block_steps(n_blk - 1, cur, wn);
#elif LSI_TOY3_PIPE == 2
                for (int b = 0; b < n_blk; ++b) {
                    block_steps(b, cur, wn);
                    if (b + 1 < n_blk) generate_from(wn, needs_second(b + 1), cur);
                }
#else
                for (int b = 0; b < n_blk; ++b) {
                    if (b > 0) generate(rid, leg_base | (uint32_t)b, needs_second(b), cur);
                    block_steps(b, cur, wn);
                }
#endif
            }
            for_each_slot([&](auto ic) {
                constexpr int p = decltype(ic)::value;
                constexpr uint32_t CODE = Fam::codes[p];
                constexpr bool CYC_OO = code_is_O(CODE, NOPS, 0) && code_is_O(CODE, NOPS, NOPS - 1);
                const ToyParams& P = B.P[p];
                pe[p] = potential<POT>(FT, x[p][0], P0.p0);
                const double v2 = v[p][0] * v[p][0];
                const double E1 = fma(half_m, v2, pe[p]);
                const double heat = half_m * (CYC_OO ? h[p][0] + v2 : h[p][0]);
                const double w = ((E1 - E0[p]) - heat) * P0.inv_kT;
                P.W[(int64_t)leg * P0.n + r] = w;
                if (P.heat) P.heat[(int64_t)leg * P0.n + r] = heat;
                if (leg == 0) wF[p] = w;
                if (leg == 1 && P.partials) {
                    double* a = acc + p * kF3Acc * kF3Threads;
                    const double wf = wF[p];
                    if (isfinite(wf) && isfinite(w)) {
                        a[0] += 1.0; a[kF3Threads] += wf; a[2 * kF3Threads] += w;
                        a[3 * kF3Threads] += wf * wf; a[4 * kF3Threads] += w * w; a[5 * kF3Threads] += wf * w;
                    }
                }
            }, slots);
        }
        for_each_slot([&](auto ic) {
            constexpr int p = decltype(ic)::value;
            const ToyParams& P = B.P[p];
            if (P.x_final) P.x_final[r] = x[p][0];
            if (P.v_final) P.v_final[r] = v[p][0];
        }, slots);
        ++n_seen;
    }
    // block partials of every program
    __shared__ double sh[kF3Threads / 32][7];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for_each_slot([&](auto ic) {
        constexpr int p = decltype(ic)::value;
        const ToyParams& P = B.P[p];
        if (P.partials) {
            __syncthreads();
#pragma unroll
            for (int j = 0; j < 7; ++j) {
                const double s = warp_sum(j < kF3Acc ? acc[(p * kF3Acc + j) * kF3Threads] : (double)n_seen - acc[p * kF3Acc * kF3Threads]);
                if (lane == 0) sh[warp][j] = s;
            }
            __syncthreads();
            if (threadIdx.x < 8) {
                double s = 0.0;
                if (threadIdx.x < 7)
                    for (int w = 0; w < kF3Threads / 32; ++w) s += sh[w][threadIdx.x];
                P.partials[(int64_t)blockIdx.x * 8 + threadIdx.x] = s;
            }
        }
    }, slots);
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
        double n_start, n_mid;
        int64_t idx0;
        start_block(FT, P.keys, rid, P.n_cache, n_start, n_mid, idx0);
        double x = P.x0 ? P.x0[r] : P.x_cache[idx0];
        double v = P.v0 ? P.v0[r] : P.sigma * n_start;
        double heat = 0.0;
        const double half_m = 0.5 * P.mass;
        const int64_t T = (int64_t)P.n_steps + 1;

        for (int leg = 0; leg < P.n_legs; ++leg) {
            if (leg > 0 && P.marginal == LSI_MARGINAL_CONFIGURATION)
                v = P.sigma * (leg == 1 ? n_mid : later_leg_normal(FT, P.keys, rid, (uint32_t)leg));
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
unsigned persistent_grid(int64_t n, int64_t per_chunk, int ctas_per_sm)
{
    int dev = 0, sms = 0;  // two attribute queries per launch (no cache: devices of one process may differ)
    if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || sms <= 0)
        sms = 148;
    const int64_t wave = (int64_t)sms * ctas_per_sm;
    const int64_t chunks = (n + per_chunk - 1) / per_chunk;
    return (unsigned)(chunks < wave ? chunks : wave);
}
unsigned fused_grid(int64_t n) { return persistent_grid(n, kThreads, LSI_TOY_MINB); }
unsigned fused2_grid(int64_t n) { return persistent_grid(n, (int64_t)kF2Threads * LSI_TOY2_ILP, LSI_TOY2_MINB); }

template <class K>
cudaError_t launch_fused2_kernel(K kernel, unsigned grid, const ToyParams& P, cudaStream_t st)
{
    // more than 48 KB of dynamic shared memory is an opt-in per function (and per device): cheap, so set on every launch
    cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kF2Smem);
    if (e != cudaSuccess) return e;
    kernel<<<grid, kF2Threads, kF2Smem, st>>>(P);
    return cudaGetLastError();
}

template <int POT, bool NOISE32>
bool launch_fused2(uint32_t code, int nops, bool trace, const ToyParams& P, cudaStream_t st, cudaError_t* err, unsigned* grid_out)
{
    if (trace && NOISE32) return false;  // traced mixed-precision runs go through the interpreter
#define X(S)                                                                                       \
    if (nops == slen(S) && code == pack(S)) {                                                      \
        if (trace) {                                                                               \
            *grid_out = fused_grid(P.n);                                                           \
            toy_fused_kernel<POT, pack(S), slen(S), false, true><<<*grid_out, kThreads, 0, st>>>(P);  \
            *err = cudaGetLastError();                                                             \
        } else {                                                                                   \
            *grid_out = fused2_grid(P.n);                                                          \
            *err = launch_fused2_kernel(toy_fused2_kernel<POT, pack(S), slen(S), NOISE32, LSI_TOY2_ILP>, *grid_out, P, st);  \
        }                                                                                          \
        return true;                                                                               \
    }
    LSI_FUSED_STRINGS(X)
#undef X
    return false;
}

bool launch_fused(int pot, bool noise32, uint32_t code, int nops, bool trace, const ToyParams& P, cudaStream_t st,
                  cudaError_t* err, unsigned* grid_out)
{
    switch (pot) {
    case LSI_POT_QUARTIC:
        return noise32 ? launch_fused2<LSI_POT_QUARTIC, true>(code, nops, trace, P, st, err, grid_out)
                       : launch_fused2<LSI_POT_QUARTIC, false>(code, nops, trace, P, st, err, grid_out);
    case LSI_POT_DOUBLE_WELL:
        return noise32 ? launch_fused2<LSI_POT_DOUBLE_WELL, true>(code, nops, trace, P, st, err, grid_out)
                       : launch_fused2<LSI_POT_DOUBLE_WELL, false>(code, nops, trace, P, st, err, grid_out);
    default:
        return noise32 ? launch_fused2<LSI_POT_HARMONIC, true>(code, nops, trace, P, st, err, grid_out)
                       : launch_fused2<LSI_POT_HARMONIC, false>(code, nops, trace, P, st, err, grid_out);
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

// validation + kernel parameters of one (system, program, args); `fusable`: every substep of a kind has the same length
// (true for lsi_compile output) and the string is short enough for a compile-time code
static int toy_prepare(const lsi_system* sys, const lsi_program* prog, const lsi_protocol_args* a, ToyParams& P, bool& fusable,
                       uint32_t& code, int& n_ops_out)
{
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
    n_ops_out = n_ops;
    const double sigma = sqrt(a->kT / sys->mass);
    if ((int64_t)a->n_steps * (int64_t)(prog->n_O > 0 ? prog->n_O : 1) >= ((int64_t)1 << 28))
        return lsi_set_error(LSI_ERR_ARG, "n_steps * n_O exceeds the 2^28 draws a leg's noise stream holds");

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
    if (a->moments) {
        if (a->n_legs != 2) return lsi_set_error(LSI_ERR_ARG, "moments need n_legs == 2");
        if (!a->workspace || a->workspace_bytes < lsi_protocol_workspace_bytes(a->n_replicas))
            return lsi_set_error(LSI_ERR_ARG, "workspace too small: need %lld bytes",
                                 (long long)lsi_protocol_workspace_bytes(a->n_replicas));
        P.partials = (double*)a->workspace;
    }
    fusable = false;
    code = 0;
    if (!a->W_direct && n_ops <= 16) {
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
            fusable = true;
        }
    }
    return LSI_OK;
}

// families of the batch kernel (toy_fused2_batch_kernel): slot order is the order of the strings
#define LSI_FAMILY_FIVE(X) X("OVRVO") X("ORVRO") X("RVOVR") X("ROVOR") X("VRORV") X("VOROV")
#define LSI_FAMILY_SIX(X) X("OVRRVO") X("VROORV") X("RVOOVR")
// whole-family (shared-noise) kernel only: the pair the reference's quartic validation compares, vvvr and baoab
// (experiments/toy/quartic_kl_validation.py:33-36); slots are a prefix of LSI_FAMILY_SIX
#define LSI_FAMILY_PAIR(X) X("OVRRVO") X("VROORV")

static int family_slot(int family, uint32_t code, int nops)
{
    int slot = 0;
#define X(S)                                                \
    if (nops == slen(S) && code == pack(S)) return slot;   \
    ++slot;
    if (family == 0) { LSI_FAMILY_FIVE(X) } else { LSI_FAMILY_SIX(X) }
#undef X
    return -1;
}

template <int POT, bool NOISE32>
static cudaError_t launch_batch2(int family, unsigned grid, const ToyBatchParams& B, cudaStream_t st)
{
#define X(S) , pack(S)
    if (family == 0) {
        auto k = toy_fused2_batch_kernel<POT, NOISE32, LSI_TOY2_ILP, 5 LSI_FAMILY_FIVE(X)>;
        cudaError_t e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kF2Smem);
        if (e != cudaSuccess) return e;
        k<<<grid, kF2Threads, kF2Smem, st>>>(B);
    } else {
        auto k = toy_fused2_batch_kernel<POT, NOISE32, LSI_TOY2_ILP, 6 LSI_FAMILY_SIX(X)>;
        cudaError_t e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kF2Smem);
        if (e != cudaSuccess) return e;
        k<<<grid, kF2Threads, kF2Smem, st>>>(B);
    }
#undef X
    return cudaGetLastError();
}

template <int POT, bool NOISE32>
static cudaError_t launch_family2(int family, unsigned grid, const ToyBatchParams& B, cudaStream_t st)
{
#define X(S) , pack(S)
    if (family == 0) {
        auto k = toy_family_kernel<POT, NOISE32, 5 LSI_FAMILY_FIVE(X)>;
        cudaError_t e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)family_smem(6));
        if (e != cudaSuccess) return e;
        k<<<grid, family_threads(6), family_smem(6), st>>>(B);
    } else if (family == 1) {
        auto k = toy_family_kernel<POT, NOISE32, 6 LSI_FAMILY_SIX(X)>;
        cudaError_t e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)family_smem(3));
        if (e != cudaSuccess) return e;
        k<<<grid, family_threads(3), family_smem(3), st>>>(B);
    } else {
        auto k = toy_family_kernel<POT, NOISE32, 6 LSI_FAMILY_PAIR(X)>;
        cudaError_t e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)family_smem(2));
        if (e != cudaSuccess) return e;
        k<<<grid, family_threads(2), family_smem(2), st>>>(B);
    }
#undef X
    return cudaGetLastError();
}

static cudaError_t launch_family(int pot, bool noise32, int family, unsigned grid, const ToyBatchParams& B, cudaStream_t st)
{
    switch (pot) {
    case LSI_POT_QUARTIC:
        return noise32 ? launch_family2<LSI_POT_QUARTIC, true>(family, grid, B, st) : launch_family2<LSI_POT_QUARTIC, false>(family, grid, B, st);
    case LSI_POT_DOUBLE_WELL:
        return noise32 ? launch_family2<LSI_POT_DOUBLE_WELL, true>(family, grid, B, st) : launch_family2<LSI_POT_DOUBLE_WELL, false>(family, grid, B, st);
    default:
        return noise32 ? launch_family2<LSI_POT_HARMONIC, true>(family, grid, B, st) : launch_family2<LSI_POT_HARMONIC, false>(family, grid, B, st);
    }
}

static cudaError_t launch_batch(int pot, bool noise32, int family, unsigned grid, const ToyBatchParams& B, cudaStream_t st)
{
    switch (pot) {
    case LSI_POT_QUARTIC:
        return noise32 ? launch_batch2<LSI_POT_QUARTIC, true>(family, grid, B, st) : launch_batch2<LSI_POT_QUARTIC, false>(family, grid, B, st);
    case LSI_POT_DOUBLE_WELL:
        return noise32 ? launch_batch2<LSI_POT_DOUBLE_WELL, true>(family, grid, B, st) : launch_batch2<LSI_POT_DOUBLE_WELL, false>(family, grid, B, st);
    default:
        return noise32 ? launch_batch2<LSI_POT_HARMONIC, true>(family, grid, B, st) : launch_batch2<LSI_POT_HARMONIC, false>(family, grid, B, st);
    }
}

// lsi_run_protocol_batch for scalar systems: ONE launch when the programs are distinct members of one family and share
// the ensemble (replicas, seed, start states, lengths, precision); otherwise one launch per program.  Same results either way.
int lsi_toy_run_protocol_batch(const lsi_system* sys, const lsi_program* const* progs, int n_progs, const lsi_protocol_args* args,
                               void* stream)
{
    cudaStream_t st = (cudaStream_t)stream;
    bool one_launch = n_progs >= 2 && n_progs <= kBatchMax && args[0].n_replicas > 0;
    ToyBatchParams B;
    std::memset(&B, 0, sizeof(B));  // slots that stay disabled travel as zeros
    int family = -1;
    const lsi_protocol_args& a0 = args[0];
    // (a, b sigma) of the programs with one / two O substeps per step: the whole-family kernel needs one pair per class
    double class_oa[3] = {0, 0, 0}, class_oc[3] = {0, 0, 0};
    bool class_seen[3] = {false, false, false}, classes_ok = true;
    for (int p = 0; p < n_progs && one_launch; ++p) {
        const lsi_protocol_args& a = args[p];
        ToyParams P;
        bool fusable = false;
        uint32_t code = 0;
        int nops = 0;
        const int rc = toy_prepare(sys, progs[p], &a, P, fusable, code, nops);
        if (rc != LSI_OK) return rc;
        const bool same = a.n_replicas == a0.n_replicas && a.n_steps == a0.n_steps && a.n_legs == a0.n_legs && a.marginal == a0.marginal &&
                          a.precision == a0.precision && a.kT == a0.kT && a.seed == a0.seed && a.replica0 == a0.replica0 &&
                          a.x_cache == a0.x_cache && a.n_cache == a0.n_cache && a.x0 == a0.x0 && a.v0 == a0.v0;
        const bool plain = !a.trace_x && !a.trace_v && !a.trace_w && !a.W_direct;
        if (!same || !plain || !fusable) { one_launch = false; break; }
        int slot = family_slot(0, code, nops), fam = 0;
        if (slot < 0) { slot = family_slot(1, code, nops); fam = 1; }
        if (slot < 0 || (family >= 0 && fam != family) || ((B.enabled >> slot) & 1u)) { one_launch = false; break; }
        family = fam;
        B.enabled |= 1u << slot;
        B.P[slot] = P;
        const int n_o = progs[p]->n_O;
        if (n_o < 1 || n_o > 2) classes_ok = false;
        else if (!class_seen[n_o]) { class_seen[n_o] = true; class_oa[n_o] = P.oa; class_oc[n_o] = P.oc; }
        else if (class_oa[n_o] != P.oa || class_oc[n_o] != P.oc) classes_ok = false;
    }
    if (!one_launch) {
        for (int p = 0; p < n_progs; ++p) {
            const int rc = lsi_toy_run_protocol(sys, progs[p], &args[p], stream);
            if (rc != LSI_OK) return rc;
        }
        return LSI_OK;
    }
    // the whole family, and one (a, b sigma) per number of O substeps: the shared-noise kernel (one replica per thread through
    // all programs); otherwise the programs run one after the other inside one launch
    const bool whole_family = classes_ok && (B.enabled == (family == 0 ? 0x3Fu : 0x7u) || (family == 1 && B.enabled == 0x3u));
    const int family_kernel = (family == 1 && B.enabled == 0x3u) ? 2 : family;
    unsigned grid;
    cudaError_t err;
    if (whole_family) {
        grid = persistent_grid(a0.n_replicas, family_threads(family_kernel == 0 ? 6 : (family_kernel == 1 ? 3 : 2)), 1);
        err = launch_family(sys->potential, a0.precision == LSI_PRECISION_MIXED, family_kernel, grid, B, st);
    } else {
        grid = fused2_grid(a0.n_replicas);
        err = launch_batch(sys->potential, a0.precision == LSI_PRECISION_MIXED, family, grid, B, st);
    }
    if (err != cudaSuccess) return lsi_set_error(LSI_ERR_CUDA, "toy batch kernel launch failed: %s", cudaGetErrorString(err));
    const double* partials[kBatchMax];
    double* moments[kBatchMax];
    int n_fin = 0;
    for (int p = 0; p < n_progs; ++p)
        if (args[p].moments) { partials[n_fin] = (const double*)args[p].workspace; moments[n_fin] = args[p].moments; ++n_fin; }
    if (n_fin > 0) return lsi_finalize_moments_batch(partials, (int)grid, moments, n_fin, stream);
    return LSI_OK;
}

int lsi_toy_run_protocol(const lsi_system* sys, const lsi_program* prog, const lsi_protocol_args* a, void* stream)
{
    cudaStream_t st = (cudaStream_t)stream;
    if (a->n_replicas == 0) return LSI_OK;
    ToyParams P;
    bool fusable = false;
    uint32_t code = 0;
    int n_ops = 0;
    {
        const int rc = toy_prepare(sys, prog, a, P, fusable, code, n_ops);
        if (rc != LSI_OK) return rc;
    }
    const double sigma = P.sigma;
    const bool noise32 = a->precision == LSI_PRECISION_MIXED;
    const int64_t grid = (a->n_replicas + kThreads - 1) / kThreads;
    const bool direct = a->W_direct != nullptr;
    const bool trace = a->trace_x || a->trace_v || a->trace_w;

    bool launched = false;
    unsigned fused_blocks = 0;
    cudaError_t err = cudaSuccess;
    if (fusable) launched = launch_fused(sys->potential, noise32, code, n_ops, trace, P, st, &err, &fused_blocks);

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
    if (a->moments) return lsi_finalize_moments(P.partials, launched ? (int)fused_blocks : (int)grid, a->moments, stream);
    return LSI_OK;
}

// Counter-based Philox-4x32-10 noise (Salmon et al., SC'11) and Box-Muller normals.
//
// Canonical layout (DESIGN.md "RNG layout"; the CPU oracle restates the same layout
// independently so a run can be replayed exactly):
//     key     = (seed_lo, seed_hi)
//     counter = (replica_lo, replica_hi, draw, (tag << 28) | unit)
// One block gives four 32-bit words -> two Box-Muller pairs -> four normals:
//     (r0, r1) -> (n0, n1) = rad(r0) * (cos, sin)(2 pi u(r1)),   (r2, r3) -> (n2, n3)
// with u(r) = (r + 0.5) * 2^-32, rad(r) = sqrt(-2 ln u(r)).
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

#include "fastmath64.cuh"

namespace lsi {

struct Philox4 { uint32_t x, y, z, w; };

__device__ __forceinline__ Philox4 philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3,
                                                 uint32_t k0, uint32_t k1)
{
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        const uint32_t hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
        const uint32_t hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
        c0 = hi1 ^ c1 ^ k0;
        c2 = hi0 ^ c3 ^ k1;
        c1 = lo1;
        c3 = lo0;
        k0 += 0x9E3779B9u;
        k1 += 0xBB67AE85u;
    }
    return Philox4{c0, c1, c2, c3};
}

// Round keys of one seed, computed once on the host (philox_round_keys) and carried in the kernel parameters:
// the key operand of every round's XOR then comes straight from the constant bank instead of 20 additions
// per block.  Same function as philox4x32_10.
struct PhiloxKeys { uint32_t k[20]; };

inline __host__ __device__ PhiloxKeys philox_round_keys(uint64_t seed)
{
    PhiloxKeys K;
    uint32_t k0 = (uint32_t)seed, k1 = (uint32_t)(seed >> 32);
    for (int r = 0; r < 10; ++r) {
        K.k[2 * r] = k0;
        K.k[2 * r + 1] = k1;
        k0 += 0x9E3779B9u;
        k1 += 0xBB67AE85u;
    }
    return K;
}

__device__ __forceinline__ Philox4 philox_block(const PhiloxKeys& K, uint64_t replica, uint32_t draw, uint32_t tag,
                                                uint32_t unit)
{
    uint32_t c0 = (uint32_t)replica, c1 = (uint32_t)(replica >> 32), c2 = draw, c3 = (tag << 28) | unit;
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        const uint32_t hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
        const uint32_t hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
        c0 = hi1 ^ c1 ^ K.k[2 * r];
        c2 = hi0 ^ c3 ^ K.k[2 * r + 1];
        c1 = lo1;
        c3 = lo0;
    }
    return Philox4{c0, c1, c2, c3};
}

__device__ __forceinline__ Philox4 philox_block(uint64_t seed, uint64_t replica, uint32_t draw, uint32_t tag,
                                                uint32_t unit)
{
    return philox4x32_10((uint32_t)replica, (uint32_t)(replica >> 32), draw, (tag << 28) | unit,
                         (uint32_t)seed, (uint32_t)(seed >> 32));
}

// fp64 Box-Muller on one pair of words, lean arithmetic (fastmath64.cuh): same formula, ~1 ulp from the
// library version below; used by the fused toy kernels
__device__ __forceinline__ void box_muller_lean(uint32_t ra, uint32_t rb, double& n0, double& n1)
{
    const double u1 = fm::uniform_from_u32(ra, false);
    const double u2x2 = fm::uniform_from_u32(rb, true);
    const double rad = fm::sqrt_pos(-2.0 * fm::log_pos(u1));
    double s, c;
    fm::sincospi_small(u2x2, s, c);
    n0 = rad * c;
    n1 = rad * s;
}

// the same pair with the table-driven -2 log and sincos (tables staged in shared memory by the caller)
// and the five-instruction sqrt; `scale` multiplies both normals
template <class TAB>
__device__ __forceinline__ void box_muller_tab(const TAB& T, uint32_t ra, uint32_t rb, double scale, double& n0, double& n1)
{
    const double rad = fm::sqrt_pos_scaled(fm::neg2log_tab(T, fm::uniform_from_u32(ra, false)), scale);
    double s, c;
    fm::sincos2pi_u32_tab(T, rb, s, c);
    n0 = rad * c;
    n1 = rad * s;
}

// fp64 Box-Muller on one pair of words
__device__ __forceinline__ void box_muller(uint32_t ra, uint32_t rb, double& n0, double& n1)
{
    const double u1 = ((double)ra + 0.5) * 2.3283064365386962890625e-10;
    const double u2x2 = ((double)rb + 0.5) * 4.656612873077392578125e-10;  // 2 * u2
    const double rad = sqrt(-2.0 * log(u1));
    double s, c;
    sincospi(u2x2, &s, &c);
    n0 = rad * c;
    n1 = rad * s;
}

// fp32 Box-Muller (mixed precision paths: the O noise of LSI_PRECISION_MIXED, as OpenMM's mixed mode draws single-precision
// gaussians).  Same uniforms and the same formula as the oracle's fp32 restatement (24-bit words, rad = sqrt(-2 ln u1),
// (cos, sin)(2 pi u2)), evaluated with the special-function unit: lg2.approx (absolute error 2^-22 on [0.5, 2], 2 ulp elsewhere),
// sqrt.approx, and sin / cos.approx on the angle shifted into [-pi, pi] (absolute error 2^-21.4) -- a normal carries an absolute
// error of at most ~2e-6, ten fp32 ulps of a typical draw, against the libm evaluation; 14 instructions per pair instead of ~85
// (the O noise was 10 % of all instructions of the water cluster and alanine dipeptide, 20 % for two-O splittings).
__device__ __forceinline__ void box_muller_sfu(uint32_t ra, uint32_t rb, float& n0, float& n1)
{
    const float u1 = ((float)(ra >> 8) + 0.5f) * 5.9604644775390625e-8f;         // 24 bits -> (0, 1]
    const float u2x2 = ((float)(rb >> 8) + 0.5f) * 1.1920928955078125e-7f;       // 2 * u2 in (0, 2]
    float lg, rad, s, c;
    asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(lg) : "f"(u1));
    asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(rad) : "f"(lg * -1.3862943611198906f));  // -2 ln u1
    const float x = fmaf(u2x2, 3.14159265358979f, -3.14159265358979f);              // 2 pi u2 - pi in (-pi, pi]
    asm("sin.approx.ftz.f32 %0, %1;" : "=f"(s) : "f"(x));
    asm("cos.approx.ftz.f32 %0, %1;" : "=f"(c) : "f"(x));
    n0 = -rad * c;  // cos(x + pi) = -cos x
    n1 = -rad * s;
}

// the same pair with the accurate library functions (scalar systems' "mixed" mode: a 1-DOF double well amplifies a 1e-6
// difference in a normal beyond the 1e-5 parity bar within ten steps, so its fp32 noise stays within an ulp of the oracle's)
__device__ __forceinline__ void box_muller(uint32_t ra, uint32_t rb, float& n0, float& n1)
{
    const float u1 = ((float)(ra >> 8) + 0.5f) * 5.9604644775390625e-8f;         // 24 bits -> (0, 1)
    const float u2x2 = ((float)(rb >> 8) + 0.5f) * 1.1920928955078125e-7f;       // 2 * u2
    const float rad = sqrtf(-2.0f * logf(u1));
    float s, c;
    sincospif(u2x2, &s, &c);
    n0 = rad * c;
    n1 = rad * s;
}

// particle systems: uniform index into the equilibrium cache from lane 0 of the replica's own TAG_X0 block
__device__ __forceinline__ int64_t cache_index(uint64_t seed, uint64_t replica, int64_t n)
{
    const Philox4 r = philox_block(seed, replica, 0u, LSI_TAG_X0, 0u);
    return (int64_t)(((uint64_t)r.x * (uint64_t)n) >> 32);
}

// scalar (toy) systems: ONE start block per replica, (replica, draw 0, TAG_V0, unit 0), feeds everything a two-leg
// protocol draws outside its O substeps: words 0, 1 -> the Box-Muller pair (start velocity, midpoint velocity),
// word 2 -> the cache index (w2 * n) >> 32, word 3 spare.  Start velocities of legs >= 2 (rare) are normal leg - 2
// of the stream that begins at draw 1: block 1 + ((leg - 2) >> 2), lane (leg - 2) & 3.
__device__ __forceinline__ int64_t index_from_word(uint32_t w, int64_t n)
{
    return (int64_t)(((uint64_t)w * (uint64_t)n) >> 32);
}

}  // namespace lsi

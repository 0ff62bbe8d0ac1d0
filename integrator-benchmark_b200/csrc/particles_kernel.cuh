// K2/K3: whole-protocol kernel for particle systems (water cluster, alanine dipeptide, tiny water
// box).  Included by particles_f32.cu / particles_f64.cu, which instantiate it for one `real`.
//
// Reference semantics (paths relative to the reference repository):
//   substeps R / V / V{g} / O with constraints     benchmark/integrators/langevin.py:124-175
//   protocol, work = (dE - dheat)/kT               benchmark/testsystems/bookkeepers.py:184-295
//   systems                                        benchmark/testsystems/{watercluster,waterbox,alanine_dipeptide}.py
// OpenMM supplies forces, SETTLE/CCMA and the `gaussian` stream there.  Here:
//
//   * a replica belongs to a TEAM: half a warp when N <= 32 (two replicas in lockstep per warp), one warp
//     when N <= 64 (several replicas per CTA, only __syncwarp over the team's lanes between phases),
//     otherwise one CTA.  Its state (x, v fp64) stays in shared memory for the whole protocol (forward
//     leg, midpoint, reverse leg).
//   * FORCE phases are atom-mapped: a lane owns APL atoms and runs a full-shell loop over all partners,
//     whose records {x, y, z, q sqrt(K)}, {sigma/2, 2 sqrt(eps)} are shared-memory broadcasts staged in
//     SLOT order (Lennard-Jones-active atoms first; one LDS.128 feeds APL pair evaluations).  Forces
//     accumulate in registers: no atomics, no cross-thread reduction, bit-reproducible.  The first n_lj
//     partners get the Coulomb + LJ body (only in warps that own LJ-active atoms), the rest (TIP3P
//     hydrogens) a Coulomb-only body with the lane's charge factored out; exclusions are per-slot
//     bitmasks; minimum image by magic-number rounding (full-rate FP32).
//   * bonded terms and 1-4 exceptions: one lane per TERM computes all role gradients once into a
//     shared buffer; every atom then gathers its entries in a fixed order (deterministic).
//     Periodic torsions use cos/sin of the dihedral and the multiple-angle recurrence (no atan2).
//   * R / V-kick / O substeps are UNIT-mapped: a unit is a free atom, a rigid water or a star of
//     pair constraints; its owner lane holds the unit in registers, updates it, applies analytic
//     SETTLE / a Newton solve of the star (positions) and an exact 3x3 solve (velocities), and
//     stages the new positions for the next force phase.  Consecutive unit-mapped substeps need no
//     synchronisation; heat and shadow work are accumulated per lane and reduced once per leg.
//   * every phase is a non-inlined function that receives shared-memory OFFSETS (loads stay LDS) and
//     reads its constants from a per-team context in shared memory: small kernels (instruction cache),
//     separate register allocation for the fp64 constraint algebra.
//   * per-force-group force caches with validity bits: a group is re-evaluated only after R.
//   * O noise: Philox-4x32-10 keyed by (replica, leg, draw, atom).
#pragma once
#include <cuda_runtime.h>

#include <cstdlib>

#include "lsi_internal.h"
#include "particles_shared.h"
#include "rng.cuh"

// resident CTAs per SM the register allocation is capped for: warp teams (128 threads) / CTA teams (<= 384 threads)
#ifndef LSI_PK_WT_MINB
#define LSI_PK_WT_MINB 4
#endif
#ifndef LSI_PK_CTA_MINB
#define LSI_PK_CTA_MINB 2
#endif

namespace pk {

template <typename real> struct Vec;
template <> struct Vec<float> { using r4 = float4; };
template <> struct Vec<double> { using r4 = double4; };

// MUFU.RSQ without the denormal pre/post-scaling of rsqrtf (squared distances are never denormal; 2 ulp)
static __device__ __forceinline__ float rsqrt_(float x)
{
    float y;
    asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
static __device__ __forceinline__ double rsqrt_(double x) { return rsqrt(x); }
// round to nearest integer by adding and removing 1.5 * 2^(mantissa bits): two full-rate adds
static __device__ __forceinline__ float round_fma(float d, float inv) { return __fadd_rn(__fmaf_rn(d, inv, 12582912.0f), -12582912.0f); }
static __device__ __forceinline__ double round_fma(double d, double inv) { return __dadd_rn(__fma_rn(d, inv, 6755399441055744.0), -6755399441055744.0); }

// ---- shared-memory layout of one team (host and device agree through this function)
struct Layout {
    uint32_t x, v, xref, xold, vold, mass, invm, sigv, red, ctx, xq, qk, f, tf, total;
};
template <typename real>
__host__ __device__ inline Layout layout(int N, int n_slots, int n_tf, bool generic, bool warp_team, bool hmc = false)
{
    Layout L;
    uint32_t o = 0;
    L.x = o; o += (uint32_t)sizeof(double) * 3 * N;
    L.v = o; o += (uint32_t)sizeof(double) * 3 * N;
    L.xref = o; if (generic) o += (uint32_t)sizeof(double) * 3 * N;
    L.xold = o; if (hmc) o += (uint32_t)sizeof(double) * 3 * N;  // HMC sampler: state before the proposal
    L.vold = o; if (hmc) o += (uint32_t)sizeof(double) * 3 * N;
    L.mass = o; o += (uint32_t)sizeof(double) * N;
    L.invm = o; o += (uint32_t)sizeof(double) * N;
    L.sigv = o; o += (uint32_t)sizeof(double) * N;
    L.red = o; if (!warp_team) o += (uint32_t)sizeof(double) * 32;
    L.ctx = o; o += 768;  // TeamCtx<real> (static_assert in the kernel file)
    o = (o + 31u) & ~31u;
    L.xq = o; o += 8 * (uint32_t)sizeof(real) * N;  // per atom {x, y, z, q sqrt(K)}, {sigma/2, 2 sqrt(eps), -, -}
    L.qk = o; o += (uint32_t)sizeof(real) * N;
    o = (o + 15u) & ~15u;
    L.f = o; o += (uint32_t)sizeof(real) * 3 * N * n_slots;
    L.tf = o; o += (uint32_t)sizeof(real) * 3 * n_tf;
    L.total = (o + 31u) & ~31u;
    return L;
}

// the dynamic shared memory of every kernel in this file; pointers derived from it keep the shared
// address space inside the non-inlined phase functions (LDS / STS instead of generic loads)
extern __shared__ __align__(32) unsigned char pk_smem[];
template <class T>
__device__ __forceinline__ T* sm(uint32_t off) { return reinterpret_cast<T*>(pk_smem + off); }

template <typename real>
struct Smem {
    double *x, *v, *xref, *mass, *invm, *sigv, *red;
    typename Vec<real>::r4* xq;
    real *qk, *f, *tf;
};

template <typename real>
__device__ __forceinline__ Smem<real> carve(unsigned char* base, const Layout& L)
{
    Smem<real> s;
    s.x = reinterpret_cast<double*>(base + L.x);
    s.v = reinterpret_cast<double*>(base + L.v);
    s.xref = reinterpret_cast<double*>(base + L.xref);
    s.mass = reinterpret_cast<double*>(base + L.mass);
    s.invm = reinterpret_cast<double*>(base + L.invm);
    s.sigv = reinterpret_cast<double*>(base + L.sigv);
    s.red = reinterpret_cast<double*>(base + L.red);
    s.xq = reinterpret_cast<typename Vec<real>::r4*>(base + L.xq);
    s.qk = reinterpret_cast<real*>(base + L.qk);
    s.f = reinterpret_cast<real*>(base + L.f);
    s.tf = reinterpret_cast<real*>(base + L.tf);
    return s;
}

// lanes of this thread's team within its warp: all 32, or one half when two 16-lane teams share the warp
// (their control flow may differ -- e.g. accept / reject in the HMC sampler -- so warp primitives name only the team)
template <bool WT>
__device__ __forceinline__ unsigned team_mask(int T)
{
    return (WT && T == 16) ? (0xffffu << (threadIdx.x & 16)) : 0xffffffffu;
}

template <bool WT>
__device__ __forceinline__ void team_sync(int T)
{
    if (WT) __syncwarp(team_mask<WT>(T));
    else __syncthreads();
}

// sum over the team, result in every thread (fixed order: deterministic)
template <bool WT>
__device__ __forceinline__ double team_sum(double val, double* red, int tid, int T)
{
    if (WT) {  // sub-warp team: T = 16 or 32 lanes
        const unsigned mask = team_mask<WT>(T);
        for (int o = T >> 1; o > 0; o >>= 1) val += __shfl_xor_sync(mask, val, o);
        return val;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) val += __shfl_xor_sync(0xffffffffu, val, o);
    __syncthreads();
    if ((tid & 31) == 0) red[tid >> 5] = val;
    __syncthreads();
    double s = 0.0;
    for (int w = 0; w < (T >> 5); ++w) s += red[w];
    return s;
}

// Per-team context in shared memory: everything the non-inlined phases (force evaluation, unit-mapped
// substeps) need, with the force-loop constants already in force precision.  Keeping those phases out
// of line keeps the kernel small enough for the instruction cache and gives the fp64 constraint
// algebra its own register allocation.
template <typename real>
struct TeamCtx {
    Layout L;
    uint32_t team_off;
    // units
    const PKUnit* units;
    const int32_t *gc_off, *gc_pairs, *gc_atom_off, *gc_atoms;
    const double* gc_dist;
    double box[3], inv_box[3], tol;
    double settle_c[8], settle_m[9];
    uint64_t seed;
    int periodic;
    // forces
    int N, excl_words, n_terms, n_sets, g_nb, g_restr, n_lj;
    int set_group[kMaxSets], set_lo[kMaxSets], set_hi[kMaxSets];
    const uint32_t* excl;
    const int32_t *perm, *iperm;
    const PKTerm<real>* terms;
    const int32_t *gather_off, *gather_idx;
    real boxr[3], inv_boxr[3], rc2, krf, krf2, crf, restraint_k, switch_r, inv_switch_w;
};

template <typename real>
__device__ __forceinline__ void fill_ctx(TeamCtx<real>* c, const PKParams& P, const Layout& L, uint32_t team_off)
{
    c->L = L;
    c->team_off = team_off;
    c->units = P.units;
    c->gc_off = P.gc_off; c->gc_pairs = P.gc_pairs; c->gc_atom_off = P.gc_atom_off; c->gc_atoms = P.gc_atoms;
    c->gc_dist = P.gc_dist;
    for (int k = 0; k < 3; ++k) {
        c->box[k] = P.box[k]; c->inv_box[k] = P.inv_box[k];
        c->boxr[k] = (real)P.box[k]; c->inv_boxr[k] = (real)P.inv_box[k];
    }
    c->tol = P.tol;
    for (int k = 0; k < 8; ++k) c->settle_c[k] = P.settle_c[k];
    for (int k = 0; k < 9; ++k) c->settle_m[k] = P.settle_m[k];
    c->seed = P.seed;
    c->periodic = P.nb_method == LSI_NB_CUTOFF_PERIODIC;
    c->N = P.N; c->excl_words = P.excl_words; c->n_terms = P.n_terms; c->n_sets = P.n_sets;
    c->g_nb = P.g_nb; c->g_restr = P.g_restr;
    for (int k = 0; k < kMaxSets; ++k) { c->set_group[k] = P.set_group[k]; c->set_lo[k] = P.set_lo[k]; c->set_hi[k] = P.set_hi[k]; }
    c->excl = P.excl; c->perm = P.perm; c->iperm = P.iperm;
    c->n_lj = P.n_lj;
    c->terms = reinterpret_cast<const PKTerm<real>*>(P.terms);
    c->gather_off = P.gather_off; c->gather_idx = P.gather_idx;
    c->rc2 = (real)P.cutoff2; c->krf = (real)P.krf; c->krf2 = (real)(2.0 * P.krf); c->crf = (real)P.crf;
    c->restraint_k = (real)P.restraint_k;
    c->switch_r = (real)P.switch_r; c->inv_switch_w = (real)P.inv_switch_w;
}

template <typename real, bool PERIODIC>
__device__ __forceinline__ void min_image(const TeamCtx<real>& C, real& dx, real& dy, real& dz)
{
    if (PERIODIC) {
        dx = dx - C.boxr[0] * round_fma(dx, C.inv_boxr[0]);
        dy = dy - C.boxr[1] * round_fma(dy, C.inv_boxr[1]);
        dz = dz - C.boxr[2] * round_fma(dz, C.inv_boxr[2]);
    }
}

// ------------------------------------------------------------------------------------------ forces
// Nonbonded full-shell loop for the APL atoms of this lane.  Positions are staged in SLOT order:
// slot k holds atom perm[k], Lennard-Jones-active atoms first.  Partners k < n_lj get the Coulomb + LJ
// body (only in warps that own LJ-active atoms -- a warp-uniform choice), the rest (TIP3P hydrogens)
// the Coulomb-only body, whose sums leave the lane's own charge out of the loop.
template <typename real, bool PERIODIC, bool ENERGY>
struct PairBody {
    real bx, by, bz, ibx, iby, ibz, rc2, krf, krf2, crf, sw_r, sw_iw;
    // geometry shared by both bodies
    __device__ __forceinline__ bool geom(const typename Vec<real>::r4& pj, real xi, real yi, real zi, uint32_t mw, int jj,
                                         real& dx, real& dy, real& dz, real& r2v, real& inv_r, real& inv_r2) const
    {
        dx = pj.x - xi; dy = pj.y - yi; dz = pj.z - zi;
        if (PERIODIC) {
            dx = dx - bx * round_fma(dx, ibx);
            dy = dy - by * round_fma(dy, iby);
            dz = dz - bz * round_fma(dz, ibz);
        }
        r2v = dx * dx + dy * dy + dz * dz;
        bool ok = ((mw >> jj) & 1u) == 0u;  // not excluded, not an exception, not self
        if (PERIODIC) ok = ok && (r2v < rc2);
        inv_r = rsqrt_(r2v);
        inv_r2 = inv_r * inv_r;
        return ok;
    }
};

template <typename real, bool PERIODIC, int APL, bool ENERGY>
__device__ __forceinline__ void pair_forces(const TeamCtx<real>& C, const typename Vec<real>::r4* __restrict__ xq, int slot0, int T, unsigned mask,
                                            const int (&atom)[APL], real (&fx)[APL], real (&fy)[APL], real (&fz)[APL],
                                            double (&en)[APL])
{
    using r4 = typename Vec<real>::r4;
    const int N = C.N, W = C.excl_words, n_lj = C.n_lj;
    real xi[APL], yi[APL], zi[APL], qi[APL], hsi[APL], sei[APL];
    bool wlj[APL];
#pragma unroll
    for (int a = 0; a < APL; ++a) {
        const int k = atom[a] >= 0 ? slot0 + a * T : 0;
        const r4 p = xq[2 * k], l = xq[2 * k + 1];
        xi[a] = p.x; yi[a] = p.y; zi[a] = p.z;
        qi[a] = atom[a] >= 0 ? p.w : (real)0;
        hsi[a] = l.x;
        sei[a] = atom[a] >= 0 ? l.y : (real)0;
        wlj[a] = __any_sync(mask, sei[a] != (real)0);  // uniform over the team's lanes (a whole warp unless T = 16)
    }
    PairBody<real, PERIODIC, ENERGY> B;
    B.bx = C.boxr[0]; B.by = C.boxr[1]; B.bz = C.boxr[2];
    B.ibx = C.inv_boxr[0]; B.iby = C.inv_boxr[1]; B.ibz = C.inv_boxr[2];
    B.rc2 = C.rc2; B.krf = C.krf; B.krf2 = C.krf2; B.crf = C.crf; B.sw_r = C.switch_r; B.sw_iw = C.inv_switch_w;
    const bool use_switch = PERIODIC && C.inv_switch_w != (real)0;  // team-uniform
    real gx[APL], gy[APL], gz[APL];  // Coulomb sums without the lane's own charge
    double ge[APL];              // energies: fp32 pair terms summed in fp64 (as OpenMM's mixed mode)
#pragma unroll
    for (int a = 0; a < APL; ++a) { gx[a] = gy[a] = gz[a] = (real)0; ge[a] = 0.0; }

    for (int jb = 0; jb < N; jb += 32) {
        uint32_t mw[APL];
#pragma unroll
        for (int a = 0; a < APL; ++a)
            mw[a] = atom[a] >= 0 ? __ldg(C.excl + (size_t)(slot0 + a * T) * W + (jb >> 5)) : 0xffffffffu;
        const int jn = min(32, N - jb);
        const int jl = max(0, min(jn, n_lj - jb));  // partners [0, jl) of this block are LJ-active
        const r4* __restrict__ xj = xq + 2 * jb;
        int jj = 0;
#pragma unroll 2
        for (; jj < jl; ++jj) {
            const r4 pj = xj[2 * jj], lj = xj[2 * jj + 1];
#pragma unroll
            for (int a = 0; a < APL; ++a) {
                real dx, dy, dz, r2v, inv_r, inv_r2;
                const bool ok = B.geom(pj, xi[a], yi[a], zi[a], mw[a], jj, dx, dy, dz, r2v, inv_r, inv_r2);
                if (wlj[a]) {
                    const real qq = qi[a] * pj.w;
                    real fr = PERIODIC ? qq * (inv_r * inv_r2 - B.krf2) : qq * inv_r * inv_r2;
                    const real sig = hsi[a] + lj.x;
                    const real s2 = sig * sig * inv_r2, s6 = s2 * s2 * s2;
                    const real t = sei[a] * lj.y * s6;  // 4 eps s6
                    real f_lj = t * ((real)12 * s6 - (real)6) * inv_r2;
                    real e_lj = t * (s6 - (real)1);
                    if (use_switch) {
                        // OpenMM's switching function S(x) = 1 - 6 x^5 + 15 x^4 - 10 x^3 on [switch_r, cutoff]; x = 0 below
                        const real xs = fmax((r2v * inv_r - B.sw_r) * B.sw_iw, (real)0);
                        const real xs2 = xs * xs;
                        const real sw = (real)1 + xs2 * xs * ((real)-10 + xs * ((real)15 - (real)6 * xs));
                        const real dsw = xs2 * ((real)-30 + xs * ((real)60 - (real)30 * xs)) * B.sw_iw;
                        f_lj = f_lj * sw - e_lj * dsw * inv_r;
                        e_lj *= sw;
                    }
                    fr += f_lj;
                    fr = ok ? fr : (real)0;
                    fx[a] -= fr * dx; fy[a] -= fr * dy; fz[a] -= fr * dz;
                    if (ENERGY) {
                        const real e = (PERIODIC ? qq * (inv_r + B.krf * r2v - B.crf) : qq * inv_r) + e_lj;
                        en[a] += ok ? (double)e : 0.0;
                    }
                } else {
                    real fr = PERIODIC ? pj.w * (inv_r * inv_r2 - B.krf2) : pj.w * inv_r * inv_r2;
                    fr = ok ? fr : (real)0;
                    gx[a] -= fr * dx; gy[a] -= fr * dy; gz[a] -= fr * dz;
                    if (ENERGY) {
                        const real e = PERIODIC ? pj.w * (inv_r + B.krf * r2v - B.crf) : pj.w * inv_r;
                        ge[a] += ok ? (double)e : 0.0;
                    }
                }
            }
        }
#pragma unroll 4
        for (; jj < jn; ++jj) {
            const r4 pj = xj[2 * jj];
#pragma unroll
            for (int a = 0; a < APL; ++a) {
                real dx, dy, dz, r2v, inv_r, inv_r2;
                const bool ok = B.geom(pj, xi[a], yi[a], zi[a], mw[a], jj, dx, dy, dz, r2v, inv_r, inv_r2);
                real fr = PERIODIC ? pj.w * (inv_r * inv_r2 - B.krf2) : pj.w * inv_r * inv_r2;
                fr = ok ? fr : (real)0;
                gx[a] -= fr * dx; gy[a] -= fr * dy; gz[a] -= fr * dz;
                if (ENERGY) {
                    const real e = PERIODIC ? pj.w * (inv_r + B.krf * r2v - B.crf) : pj.w * inv_r;
                    ge[a] += ok ? (double)e : 0.0;
                }
            }
        }
    }
#pragma unroll
    for (int a = 0; a < APL; ++a) {
        fx[a] += qi[a] * gx[a]; fy[a] += qi[a] * gy[a]; fz[a] += qi[a] * gz[a];
        if (ENERGY) en[a] += (double)qi[a] * ge[a];
    }
}

// one lane per bonded term / exception: every role gradient once, into tf[3 * (slot + role) + c]
template <typename real, bool PERIODIC, bool ENERGY>
__device__ __forceinline__ real eval_terms(const TeamCtx<real>& C, const typename Vec<real>::r4* __restrict__ xq,
                                           real* __restrict__ tf, int lo, int hi, int tid, int T)
{
    using r4 = typename Vec<real>::r4;
    real e_acc = (real)0;
    for (int k = lo + tid; k < hi; k += T) {
        const PKTerm<real> t = C.terms[k];
        real* out = tf + 3 * (size_t)t.slot;
        const r4 p0 = xq[2 * t.a0], p1 = xq[2 * t.a1];
        if (t.kind == TERM_BOND || t.kind == TERM_EXCEPTION) {
            real dx = p1.x - p0.x, dy = p1.y - p0.y, dz = p1.z - p0.z;  // from atom 0 to atom 1
            min_image<real, PERIODIC>(C, dx, dy, dz);
            const real r2v = dx * dx + dy * dy + dz * dz;
            const real inv_r = rsqrt_(r2v);
            real fr, e;
            if (t.kind == TERM_BOND) {
                const real dr = r2v * inv_r - t.p[0];
                e = (real)0.5 * t.p[1] * dr * dr;
                fr = t.p[1] * dr * inv_r;  // force on atom 0 = +fr d
            } else {
                const real inv_r2 = inv_r * inv_r;
                const real s2 = t.p[1] * t.p[1] * inv_r2, s6 = s2 * s2 * s2;
                e = t.p[2] * (s6 * s6 - s6) + t.p[0] * inv_r;
                fr = -(t.p[2] * ((real)12 * s6 * s6 - (real)6 * s6) * inv_r2 + t.p[0] * inv_r * inv_r2);
                if (PERIODIC && r2v >= C.rc2) { fr = (real)0; e = (real)0; }
            }
            out[0] = fr * dx; out[1] = fr * dy; out[2] = fr * dz;
            out[3] = -fr * dx; out[4] = -fr * dy; out[5] = -fr * dz;
            if (ENERGY) e_acc += e;
        } else if (t.kind == TERM_ANGLE) {
            const r4 p2 = xq[2 * t.a2];
            real ax = p0.x - p1.x, ay = p0.y - p1.y, az = p0.z - p1.z;
            real cx = p2.x - p1.x, cy = p2.y - p1.y, cz = p2.z - p1.z;
            min_image<real, PERIODIC>(C, ax, ay, az);
            min_image<real, PERIODIC>(C, cx, cy, cz);
            const real ira = rsqrt_(ax * ax + ay * ay + az * az), irc = rsqrt_(cx * cx + cy * cy + cz * cz);
            real ct = (ax * cx + ay * cy + az * cz) * ira * irc;
            ct = fmin(fmax(ct, (real)-1), (real)1);
            const real dth = acos(ct) - t.p[0];
            const real st = sqrt(fmax((real)1 - ct * ct, (real)1e-24));
            const real c = t.p[1] * dth / st;  // force = +c * d cos(theta)/d r
            const real iac = ira * irc, ia2 = ira * ira, ic2 = irc * irc;
            const real gax = cx * iac - ct * ax * ia2, gay = cy * iac - ct * ay * ia2, gaz = cz * iac - ct * az * ia2;
            const real gcx = ax * iac - ct * cx * ic2, gcy = ay * iac - ct * cy * ic2, gcz = az * iac - ct * cz * ic2;
            out[0] = c * gax; out[1] = c * gay; out[2] = c * gaz;
            out[3] = -c * (gax + gcx); out[4] = -c * (gay + gcy); out[5] = -c * (gaz + gcz);
            out[6] = c * gcx; out[7] = c * gcy; out[8] = c * gcz;
            if (ENERGY) e_acc += (real)0.5 * t.p[1] * dth * dth;
        } else {
            const r4 p2 = xq[2 * t.a2], p3 = xq[2 * t.a3];
            real b1x = p1.x - p0.x, b1y = p1.y - p0.y, b1z = p1.z - p0.z;
            real b2x = p2.x - p1.x, b2y = p2.y - p1.y, b2z = p2.z - p1.z;
            real b3x = p3.x - p2.x, b3y = p3.y - p2.y, b3z = p3.z - p2.z;
            min_image<real, PERIODIC>(C, b1x, b1y, b1z);
            min_image<real, PERIODIC>(C, b2x, b2y, b2z);
            min_image<real, PERIODIC>(C, b3x, b3y, b3z);
            const real n1x = b1y * b2z - b1z * b2y, n1y = b1z * b2x - b1x * b2z, n1z = b1x * b2y - b1y * b2x;
            const real n2x = b2y * b3z - b2z * b3y, n2y = b2z * b3x - b2x * b3z, n2z = b2x * b3y - b2y * b3x;
            const real b2sq = b2x * b2x + b2y * b2y + b2z * b2z;
            const real ib2 = rsqrt_(b2sq), b2n = b2sq * ib2;
            const real yy = b2n * (b1x * n2x + b1y * n2y + b1z * n2z);
            const real xx = n1x * n2x + n1y * n2y + n1z * n2z;
            const real n1sq = n1x * n1x + n1y * n1y + n1z * n1z, n2sq = n2x * n2x + n2y * n2y + n2z * n2z;
            // cos(phi) = xx / (|n1||n2|), sin(phi) = yy / (|n1||n2|)  (phi = atan2(yy, xx), IUPAC sign)
            const real inorm = rsqrt_(fmax(n1sq * n2sq, (real)1e-36));
            const real c1 = xx * inorm, s1 = yy * inorm;
            real cn = c1, sn = s1;
            const int nper = (int)(t.p[0] + (real)0.5);
            for (int q = 1; q < nper; ++q) {  // multiple-angle recurrence
                const real c2 = cn * c1 - sn * s1;
                sn = sn * c1 + cn * s1;
                cn = c2;
            }
            if (nper == 0) { cn = (real)1; sn = (real)0; }
            // E = k (1 + cos(n phi - phase)), dE/dphi = -k n sin(n phi - phase)
            const real dE = -t.p[3] * (real)nper * (sn * t.p[1] - cn * t.p[2]);
            if (ENERGY) e_acc += t.p[3] * ((real)1 + cn * t.p[1] + sn * t.p[2]);
            const real k1 = -b2n / n1sq, k4 = b2n / n2sq;
            const real g1x = k1 * n1x, g1y = k1 * n1y, g1z = k1 * n1z;
            const real g4x = k4 * n2x, g4y = k4 * n2y, g4z = k4 * n2z;
            const real ib2sq = ib2 * ib2;
            const real u = (b1x * b2x + b1y * b2y + b1z * b2z) * ib2sq, w = (b3x * b2x + b3y * b2y + b3z * b2z) * ib2sq;
            out[0] = -dE * g1x; out[1] = -dE * g1y; out[2] = -dE * g1z;
            out[3] = -dE * (-g1x - u * g1x + w * g4x); out[4] = -dE * (-g1y - u * g1y + w * g4y);
            out[5] = -dE * (-g1z - u * g1z + w * g4z);
            out[6] = -dE * (-g4x + u * g1x - w * g4x); out[7] = -dE * (-g4y + u * g1y - w * g4y);
            out[8] = -dE * (-g4z + u * g1z - w * g4z);
            out[9] = -dE * g4x; out[10] = -dE * g4y; out[11] = -dE * g4z;
        }
    }
    return e_acc;
}

// forces of force group g (-1 = all) into fout[3N] (team shared memory); ENERGY: returns the
// potential energy (team sum, fp64).  Entry: positions staged in s.xq by the unit owners (the
// leading team_sync orders those writes).  Exit: fout complete and visible to the team.
template <typename real, bool PERIODIC, bool WT, int APL, bool ENERGY>
__device__ __noinline__ double compute_forces(uint32_t ctx_off, int g, int out_slot, int tid, int T)
{
    const TeamCtx<real>& C = *sm<TeamCtx<real>>(ctx_off);
    const Smem<real> s = carve<real>(pk_smem + C.team_off, C.L);
    real* __restrict__ fout = s.f + (size_t)out_slot * 3 * C.N;
    const bool do_nb = g < 0 || g == C.g_nb;
    const bool do_restr = (g < 0 || g == C.g_restr) && C.restraint_k != (real)0;
    int set = -1;
    if (g < 0) set = C.n_terms > 0 ? 0 : -1;
    else
        for (int q = 1; q < C.n_sets; ++q)
            if (C.set_group[q] == g) set = q;
    int atom[APL];
    real fx[APL], fy[APL], fz[APL];
    double en[APL];
#pragma unroll
    for (int a = 0; a < APL; ++a) {
        const int slot = tid + a * T;
        atom[a] = slot < C.N ? __ldg(C.perm + slot) : -1;
        fx[a] = fy[a] = fz[a] = (real)0;
        en[a] = 0.0;
    }
    team_sync<WT>(T);
    if (do_nb) pair_forces<real, PERIODIC, APL, ENERGY>(C, s.xq, tid, T, team_mask<WT>(T), atom, fx, fy, fz, en);
    double e_thread = 0.0;
    if (set >= 0) {
        const real et = eval_terms<real, PERIODIC, ENERGY>(C, s.xq, s.tf, C.set_lo[set], C.set_hi[set], tid, T);
        if (ENERGY) e_thread += (double)et;
        team_sync<WT>(T);
    }
#pragma unroll
    for (int a = 0; a < APL; ++a) {
        const int i = atom[a];
        if (i < 0) continue;
        if (set >= 0) {
            const int32_t* off = C.gather_off + (size_t)set * (C.N + 1) + i;
            const int k0 = __ldg(off), k1 = __ldg(off + 1);
            for (int k = k0; k < k1; ++k) {
                const real* tfp = s.tf + 3 * (size_t)__ldg(C.gather_idx + k);
                fx[a] += tfp[0]; fy[a] += tfp[1]; fz[a] += tfp[2];
            }
        }
        real er = (real)0;
        if (do_restr) {  // restraint acts on the true (unwrapped) coordinates
            const real k = C.restraint_k;
            const real px = (real)s.x[3 * i], py = (real)s.x[3 * i + 1], pz = (real)s.x[3 * i + 2];
            fx[a] -= k * px; fy[a] -= k * py; fz[a] -= k * pz;
            if (ENERGY) er = (real)0.5 * k * (px * px + py * py + pz * pz);
        }
        fout[3 * i] = fx[a]; fout[3 * i + 1] = fy[a]; fout[3 * i + 2] = fz[a];
        if (ENERGY) e_thread += 0.5 * en[a] + (double)er;  // every pair is visited from both ends
    }
    team_sync<WT>(T);
    if (ENERGY) return team_sum<WT>(e_thread, s.red, tid, T);
    return 0.0;
}

// ------------------------------------------------------------------------------------------ constraints (fp64)
// fp64 reciprocal square root / square root / reciprocal without the special-case slow paths of the
// CUDA math library: hardware seed (2^-22.9 relative) + two Newton steps; ~1 ulp for normal inputs
// One higher-order correction of the hardware seed y (relative error e, |e| < 2^-22) instead of two Newton steps:
// 1/sqrt(x) = y (1 + e/2 + 3 e^2/8 + O(e^3)) with e = 1 - x y^2, and 1/x = y (1 + e + e^2 + O(e^3)) with e = 1 - x y.
static __device__ __forceinline__ double fast_rsqrt(double x)
{
    double y;
    asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
    const double e = fma(-(x * y), y, 1.0);
    return fma(y, e * fma(e, 0.375, 0.5), y);
}
static __device__ __forceinline__ double fast_sqrt(double x)
{
    double y;
    asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
    const double t = x * y;
    const double e = fma(-t, y, 1.0);
    return fma(t, e * fma(e, 0.375, 0.5), t);
}
static __device__ __forceinline__ double fast_rcp(double x)
{
    double y;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
    const double e = fma(-x, y, 1.0);
    return fma(y, fma(e, e, e), y);
}

static __device__ __forceinline__ double dot3(const double* a, const double* b) { return a[0] * b[0] + a[1] * b[1] + a[2] * b[2]; }

// analytic SETTLE (Miyamoto & Kollman 1992): x0 = reference triangle (O, H1, H2), x1 = unconstrained
// new positions, overwritten with the constrained ones.  K = constants of the rigid geometry
// (PKParams::settle_c): {mO / M, mH / M, ra, rb, rc, 1 / ra, 1 / (2 rc)}
static __device__ __forceinline__ void settle_positions(const double (&x0)[3][3], double (&x1)[3][3], const double* __restrict__ K)
{
    const double wO = K[0], wH = K[1], ra = K[2], rb = K[3], rc = K[4], inv_ra = K[5], inv_2rc = K[6];
    double xb0[3], xc0[3], xcom[3], xa1[3], xb1[3], xc1[3];
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        xb0[c] = x0[1][c] - x0[0][c];
        xc0[c] = x0[2][c] - x0[0][c];
        xcom[c] = wO * x1[0][c] + wH * x1[1][c] + wH * x1[2][c];
        xa1[c] = x1[0][c] - xcom[c];
        xb1[c] = x1[1][c] - xcom[c];
        xc1[c] = x1[2][c] - xcom[c];
    }
    double Z[3] = {xb0[1] * xc0[2] - xb0[2] * xc0[1], xb0[2] * xc0[0] - xb0[0] * xc0[2], xb0[0] * xc0[1] - xb0[1] * xc0[0]};
    double X[3] = {xa1[1] * Z[2] - xa1[2] * Z[1], xa1[2] * Z[0] - xa1[0] * Z[2], xa1[0] * Z[1] - xa1[1] * Z[0]};
    double Y[3] = {Z[1] * X[2] - Z[2] * X[1], Z[2] * X[0] - Z[0] * X[2], Z[0] * X[1] - Z[1] * X[0]};
    const double nz = fast_rsqrt(dot3(Z, Z)), nx = fast_rsqrt(dot3(X, X)), ny = fast_rsqrt(dot3(Y, Y));
#pragma unroll
    for (int c = 0; c < 3; ++c) { X[c] *= nx; Y[c] *= ny; Z[c] *= nz; }
    const double xb0d = dot3(X, xb0), yb0d = dot3(Y, xb0), xc0d = dot3(X, xc0), yc0d = dot3(Y, xc0);
    const double za1d = dot3(Z, xa1);
    const double xb1d = dot3(X, xb1), yb1d = dot3(Y, xb1), zb1d = dot3(Z, xb1);
    const double xc1d = dot3(X, xc1), yc1d = dot3(Y, xc1), zc1d = dot3(Z, xc1);
    const double sinphi = za1d * inv_ra;
    const double c2 = fma(-sinphi, sinphi, 1.0);
    const double icos = fast_rsqrt(c2), cosphi = c2 * icos;
    const double sinpsi = (zb1d - zc1d) * inv_2rc * icos, cospsi = fast_sqrt(fma(-sinpsi, sinpsi, 1.0));
    const double ya2d = ra * cosphi;
    const double xb2d = -rc * cospsi;
    const double yb2d = -rb * cosphi - rc * sinpsi * sinphi;
    const double yc2d = -rb * cosphi + rc * sinpsi * sinphi;
    const double alpha = xb2d * (xb0d - xc0d) + yb0d * yb2d + yc0d * yc2d;
    const double beta = xb2d * (yc0d - yb0d) + xb0d * yb2d + xc0d * yc2d;
    const double gamma = xb0d * yb1d - xb1d * yb0d + xc0d * yc1d - xc1d * yc0d;
    const double al2be2 = alpha * alpha + beta * beta;
    const double sintheta = (alpha * gamma - beta * fast_sqrt(al2be2 - gamma * gamma)) * fast_rcp(al2be2);
    const double costheta = fast_sqrt(fma(-sintheta, sintheta, 1.0));
    const double xa3d = -ya2d * sintheta, ya3d = ya2d * costheta, za3d = za1d;
    const double xb3d = xb2d * costheta - yb2d * sintheta, yb3d = xb2d * sintheta + yb2d * costheta, zb3d = zb1d;
    const double xc3d = -xb2d * costheta - yc2d * sintheta, yc3d = -xb2d * sintheta + yc2d * costheta, zc3d = zc1d;
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        x1[0][c] = xcom[c] + X[c] * xa3d + Y[c] * ya3d + Z[c] * za3d;
        x1[1][c] = xcom[c] + X[c] * xb3d + Y[c] * yb3d + Z[c] * zb3d;
        x1[2][c] = xcom[c] + X[c] * xc3d + Y[c] * yc3d + Z[c] * zc3d;
    }
}

static __device__ __forceinline__ void solve3(double A00, double A01, double A02, double A10, double A11, double A12, double A20,
                                       double A21, double A22, double r0, double r1, double r2, double& l0, double& l1,
                                       double& l2)
{
    const double c00 = A11 * A22 - A12 * A21, c01 = A10 * A22 - A12 * A20, c02 = A10 * A21 - A11 * A20;
    const double inv_det = fast_rcp(A00 * c00 - A01 * c01 + A02 * c02);
    l0 = (r0 * c00 - A01 * (r1 * A22 - A12 * r2) + A02 * (r1 * A21 - A11 * r2)) * inv_det;
    l1 = (A00 * (r1 * A22 - A12 * r2) - r0 * c01 + A02 * (A10 * r2 - r1 * A20)) * inv_det;
    l2 = (A00 * (A11 * r2 - r1 * A21) - A01 * (A10 * r2 - r1 * A20) + r0 * c02) * inv_det;
}

// exact removal of the velocity components along the three edges of a rigid water.  Edges k -> (k+1)%3
// with e_k = x_h - x_t; impulses lambda = M b, b_k = e_k . (v_h - v_t).  For a rigid triangle the
// edge dot products are constants, so M = -C^-1 is precomputed on the host (PKParams::settle_m):
//   C_kk = d_k^2 (1/m_h + 1/m_t),  C_01 = -(e0.e1)/m_1,  C_12 = -(e1.e2)/m_2,  C_02 = -(e0.e2)/m_0
static __device__ __forceinline__ void settle_velocities(const double (&x)[3][3], double (&v)[3][3], const double* __restrict__ M,
                                                         double imO, double imH)
{
    const double im[3] = {imO, imH, imH};
    double e[3][3], b[3];
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        const int h = (k + 1) % 3;
        b[k] = 0.0;
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            e[k][c] = x[h][c] - x[k][c];
            b[k] += (v[h][c] - v[k][c]) * e[k][c];
        }
    }
    double lam[3];
#pragma unroll
    for (int k = 0; k < 3; ++k) lam[k] = M[3 * k] * b[0] + M[3 * k + 1] * b[1] + M[3 * k + 2] * b[2];
#pragma unroll
    for (int l = 0; l < 3; ++l) {
        const int h = (l + 1) % 3;
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            v[h][c] += lam[l] * e[l][c] * im[h];
            v[l][c] -= lam[l] * e[l][c] * im[l];
        }
    }
}

// star of m <= 3 pair constraints around atom 0: Newton iteration on the multipliers of SHAKE
// (displacements along the reference bond vectors s_k = x0[0] - x0[k+1]); converges quadratically
static __device__ __forceinline__ void star_positions(const double (&x0)[4][3], double (&x1)[4][3], int m, const double (&d)[3],
                                            const double (&im)[4], double tol)
{
    double sv[3][3], r[3][3], lam[3] = {0.0, 0.0, 0.0};
#pragma unroll
    for (int k = 0; k < 3; ++k)
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            sv[k][c] = k < m ? x0[0][c] - x0[k + 1][c] : 0.0;
            r[k][c] = k < m ? x1[0][c] - x1[k + 1][c] : 0.0;
        }
    for (int it = 0; it < 30; ++it) {
        // r'_k = r_k - im0 sum_l lam_l s_l - im_k lam_k s_k
        double sum[3];
#pragma unroll
        for (int c = 0; c < 3; ++c) sum[c] = im[0] * (lam[0] * sv[0][c] + lam[1] * sv[1][c] + lam[2] * sv[2][c]);
        double rp[3][3], sig[3];
        bool done = true;
#pragma unroll
        for (int k = 0; k < 3; ++k) {
#pragma unroll
            for (int c = 0; c < 3; ++c) rp[k][c] = r[k][c] - sum[c] - im[k + 1] * lam[k] * sv[k][c];
            const double d2 = d[k] * d[k];
            sig[k] = k < m ? dot3(rp[k], rp[k]) - d2 : 0.0;
            if (k < m && fabs(sig[k]) > 2.0 * tol * d2) done = false;
        }
        if (done && it > 0) break;
        // J_kl = d sig_k / d lam_l = -2 r'_k . s_l (im0 + delta_kl im_k)
        double J[3][3];
#pragma unroll
        for (int k = 0; k < 3; ++k)
#pragma unroll
            for (int l = 0; l < 3; ++l) {
                if (k < m && l < m) J[k][l] = -2.0 * dot3(rp[k], sv[l]) * (im[0] + (k == l ? im[k + 1] : 0.0));
                else J[k][l] = k == l ? 1.0 : 0.0;
            }
        double d0, d1, d2;
        solve3(J[0][0], J[0][1], J[0][2], J[1][0], J[1][1], J[1][2], J[2][0], J[2][1], J[2][2], -sig[0], -sig[1], -sig[2], d0, d1, d2);
        lam[0] += d0; lam[1] += d1; lam[2] += d2;
        if (done) break;  // one polishing step past the tolerance
    }
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        x1[0][c] -= im[0] * (lam[0] * sv[0][c] + lam[1] * sv[1][c] + lam[2] * sv[2][c]);
#pragma unroll
        for (int k = 0; k < 3; ++k)
            if (k < m) x1[k + 1][c] += im[k + 1] * lam[k] * sv[k][c];
    }
}

// star: exact removal of the relative velocity along each constrained bond (3x3 linear solve)
static __device__ __forceinline__ void star_velocities(const double (&x)[4][3], double (&v)[4][3], int m, const double (&im)[4])
{
    double r[3][3], b[3];
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        b[k] = 0.0;
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            r[k][c] = k < m ? x[0][c] - x[k + 1][c] : 0.0;
            b[k] += k < m ? r[k][c] * (v[0][c] - v[k + 1][c]) : 0.0;
        }
    }
    double A[3][3];
#pragma unroll
    for (int k = 0; k < 3; ++k)
#pragma unroll
        for (int l = 0; l < 3; ++l) {
            if (k < m && l < m) A[k][l] = dot3(r[k], r[l]) * (im[0] + (k == l ? im[k + 1] : 0.0));
            else A[k][l] = k == l ? 1.0 : 0.0;
        }
    double mu[3];
    solve3(A[0][0], A[0][1], A[0][2], A[1][0], A[1][1], A[1][2], A[2][0], A[2][1], A[2][2], b[0], b[1], b[2], mu[0], mu[1], mu[2]);
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        v[0][c] -= im[0] * (mu[0] * r[0][c] + mu[1] * r[1][c] + mu[2] * r[2][c]);
#pragma unroll
        for (int k = 0; k < 3; ++k)
            if (k < m) v[k + 1][c] += im[k + 1] * mu[k] * r[k][c];
    }
}

// generic clusters (anything that is not a water or a small star): Gauss-Seidel SHAKE / RATTLE in
// shared memory by the cluster's owner lane.  A cluster that has not converged after 500 sweeps is not left silently
// off the constraint manifold: its first coordinate becomes NaN, so the replica's works are NaN and the replica is counted
// as crashed (moments[6]), like any other diverged trajectory.
static __device__ __noinline__ void shake_cluster(const int32_t* pairs, const double* dist, double tol, const double* mass, const double* x0,
                                           double* x1, int c0, int c1)
{
    bool converged = c1 <= c0;
    for (int it = 0; it < 500; ++it) {
        bool done = true;
        for (int k = c0; k < c1; ++k) {
            const int i = pairs[2 * k], j = pairs[2 * k + 1];
            const double d2 = dist[k] * dist[k];
            double r[3], sv[3];
#pragma unroll
            for (int c = 0; c < 3; ++c) {
                r[c] = x1[3 * i + c] - x1[3 * j + c];
                sv[c] = x0[3 * i + c] - x0[3 * j + c];
            }
            const double diff = dot3(r, r) - d2;
            if (fabs(diff) > 2.0 * tol * d2) done = false;
            const double imi = 1.0 / mass[i], imj = 1.0 / mass[j];
            const double gl = diff / (2.0 * dot3(r, sv) * (imi + imj));
#pragma unroll
            for (int c = 0; c < 3; ++c) {
                x1[3 * i + c] -= gl * sv[c] * imi;
                x1[3 * j + c] += gl * sv[c] * imj;
            }
        }
        if (done) { converged = true; break; }
    }
    if (!converged) x1[3 * pairs[2 * c0]] = __longlong_as_double(0x7ff8000000000000LL);
}

static __device__ __noinline__ void rattle_cluster(const int32_t* pairs, const double* dist, double tol, const double* mass, const double* x,
                                            double* v, int c0, int c1)
{
    bool converged = c1 <= c0;
    for (int it = 0; it < 500; ++it) {
        bool done = true;
        for (int k = c0; k < c1; ++k) {
            const int i = pairs[2 * k], j = pairs[2 * k + 1];
            const double d2 = dist[k] * dist[k];
            double r[3], w[3];
#pragma unroll
            for (int c = 0; c < 3; ++c) {
                r[c] = x[3 * i + c] - x[3 * j + c];
                w[c] = v[3 * i + c] - v[3 * j + c];
            }
            const double imi = 1.0 / mass[i], imj = 1.0 / mass[j];
            const double gl = dot3(r, w) / (d2 * (imi + imj));
            if (fabs(gl) > tol) done = false;
#pragma unroll
            for (int c = 0; c < 3; ++c) {
                v[3 * i + c] -= gl * r[c] * imi;
                v[3 * j + c] += gl * r[c] * imj;
            }
        }
        if (done) { converged = true; break; }
    }
    if (!converged) v[3 * pairs[2 * c0]] = __longlong_as_double(0x7ff8000000000000LL);
}

// ------------------------------------------------------------------------------------------ units
struct N3 { double x, y, z; };

// three normals for atom `unit` (fp64 Box-Muller): start and midpoint velocities
static __device__ __noinline__ N3 normals3_f64(uint64_t seed, uint64_t rid, uint32_t draw, uint32_t tag, uint32_t unit)
{
    const lsi::Philox4 r = lsi::philox_block(seed, rid, draw, tag, unit);
    N3 n;
    double d;
    lsi::box_muller(r.x, r.y, n.x, n.y);
    lsi::box_muller(r.z, r.w, n.z, d);
    return n;
}
// fp32 Box-Muller ("mixed" O noise, as OpenMM's mixed mode draws single-precision gaussians)
static __device__ __noinline__ N3 normals3_f32(uint64_t seed, uint64_t rid, uint32_t draw, uint32_t tag, uint32_t unit)
{
    const lsi::Philox4 r = lsi::philox_block(seed, rid, draw, tag, unit);
    float a, b, c, d;
    lsi::box_muller_sfu(r.x, r.y, a, b);
    lsi::box_muller_sfu(r.z, r.w, c, d);
    N3 n;
    n.x = (double)a; n.y = (double)b; n.z = (double)c;
    return n;
}

template <typename real>
struct UnitOps {
    const TeamCtx<real>& C;
    const Smem<real>& s;

    template <int NA>
    __device__ __forceinline__ void load(const double* arr, const PKUnit& U, double (&a)[NA][3]) const
    {
#pragma unroll
        for (int k = 0; k < NA; ++k)
#pragma unroll
            for (int c = 0; c < 3; ++c) a[k][c] = k < U.n ? arr[3 * U.a[k] + c] : 0.0;
    }
    template <int NA>
    __device__ __forceinline__ void store(double* arr, const PKUnit& U, const double (&a)[NA][3]) const
    {
#pragma unroll
        for (int k = 0; k < NA; ++k)
            if (k < U.n) {
#pragma unroll
                for (int c = 0; c < 3; ++c) arr[3 * U.a[k] + c] = a[k][c];
            }
    }
    // positions for the force loop: force precision, wrapped into the box when periodic
    __device__ __forceinline__ void stage_atom(int i, double px, double py, double pz) const
    {
        if (C.periodic) {
            px -= C.box[0] * floor(px * C.inv_box[0]);
            py -= C.box[1] * floor(py * C.inv_box[1]);
            pz -= C.box[2] * floor(pz * C.inv_box[2]);
        }
        typename Vec<real>::r4 q;
        q.x = (real)px; q.y = (real)py; q.z = (real)pz; q.w = s.qk[i];
        s.xq[2 * __ldg(C.iperm + i)] = q;  // slot order: LJ-active atoms first
    }
    template <int NA>
    __device__ __forceinline__ void stage(const PKUnit& U, const double (&x)[NA][3]) const
    {
#pragma unroll
        for (int k = 0; k < NA; ++k)
            if (k < U.n) stage_atom(U.a[k], x[k][0], x[k][1], x[k][2]);
    }
    __device__ __forceinline__ void inv_masses(const PKUnit& U, double (&im)[4]) const
    {
#pragma unroll
        for (int k = 0; k < 4; ++k) im[k] = k < U.n ? s.invm[U.a[k]] : 0.0;
    }
    template <int NA>
    __device__ __forceinline__ double kinetic(const PKUnit& U, const double (&v)[NA][3]) const
    {
        double ke = 0.0;
#pragma unroll
        for (int k = 0; k < NA; ++k)
            if (k < U.n) ke += 0.5 * s.mass[U.a[k]] * (v[k][0] * v[k][0] + v[k][1] * v[k][1] + v[k][2] * v[k][2]);
        return ke;
    }
    // O noise on the unit's atoms, in shared memory (a loop, not unrolled: Philox + Box-Muller is large)
    __device__ __forceinline__ void noise(const PKUnit& U, int u, double a, double b, uint64_t rid, uint32_t draw) const
    {
        constexpr bool NOISE32 = sizeof(real) == 4;
#pragma unroll 1
        for (int k = 0; k < U.n; ++k) {
            const int i = C.units[u].a[k];
            const N3 n = NOISE32 ? normals3_f32(C.seed, rid, draw, LSI_TAG_O, (uint32_t)i)
                                 : normals3_f64(C.seed, rid, draw, LSI_TAG_O, (uint32_t)i);
            const double sg = b * s.sigv[i];
            s.v[3 * i] = a * s.v[3 * i] + sg * n.x;
            s.v[3 * i + 1] = a * s.v[3 * i + 1] + sg * n.y;
            s.v[3 * i + 2] = a * s.v[3 * i + 2] + sg * n.z;
        }
    }

    // ---- generic clusters: atoms gc_atoms[a0..a1), constraints [c0..c1)
    __device__ __forceinline__ double generic_kinetic(int a0, int a1) const
    {
        double ke = 0.0;
        for (int q = a0; q < a1; ++q) {
            const int i = C.gc_atoms[q];
            ke += 0.5 * s.mass[i] * (s.v[3 * i] * s.v[3 * i] + s.v[3 * i + 1] * s.v[3 * i + 1] + s.v[3 * i + 2] * s.v[3 * i + 2]);
        }
        return ke;
    }
    __device__ __forceinline__ void generic_stage(int a0, int a1) const
    {
        for (int q = a0; q < a1; ++q) {
            const int i = C.gc_atoms[q];
            stage_atom(i, s.x[3 * i], s.x[3 * i + 1], s.x[3 * i + 2]);
        }
    }
    __device__ __forceinline__ void generic_shake(int c) const
    {
        shake_cluster(C.gc_pairs, C.gc_dist, C.tol, s.mass, s.xref, s.x, C.gc_off[c], C.gc_off[c + 1]);
    }
    __device__ __forceinline__ void generic_rattle(int c) const
    {
        rattle_cluster(C.gc_pairs, C.gc_dist, C.tol, s.mass, s.x, s.v, C.gc_off[c], C.gc_off[c + 1]);
    }
};

// ---- rigid water (SETTLE).  R: x <- x + h v; constrain positions; v <- v + (x - x1)/h; constrain
// velocities (langevin.py:124-139).  h = 0: Context.applyConstraints / applyVelocityConstraints at the
// start of a leg (bookkeepers.py:206-208).  Stages the new positions.  Returns KE after (h = 0) or its change.
template <typename real>
__device__ __noinline__ double water_R(uint32_t ctx_off, int u, double h)
{
    const TeamCtx<real>& C = *sm<TeamCtx<real>>(ctx_off);
    const Smem<real> s = carve<real>(pk_smem + C.team_off, C.L);
    const UnitOps<real> ops{C, s};
    const PKUnit U = C.units[u];
    double x[3][3];
    double ke0 = 0.0;
    {
        double x0[3][3], v[3][3];
        ops.load(s.x, U, x0);
        ops.load(s.v, U, v);
        if (h != 0.0) ke0 = ops.kinetic(U, v);
#pragma unroll
        for (int k = 0; k < 3; ++k)
#pragma unroll
            for (int c = 0; c < 3; ++c) x[k][c] = x0[k][c] + h * v[k][c];
        settle_positions(x0, x, C.settle_c);
    }
    double v[3][3];
    ops.load(s.v, U, v);
    if (h != 0.0) {
        const double inv_h = fast_rcp(h);
        double x0[3][3];
        ops.load(s.x, U, x0);
#pragma unroll
        for (int k = 0; k < 3; ++k)
#pragma unroll
            for (int c = 0; c < 3; ++c) {
                const double x1 = x0[k][c] + h * v[k][c];  // the unconstrained update
                v[k][c] = v[k][c] + (x[k][c] - x1) * inv_h;
            }
    }
    ops.store(s.x, U, x);
    ops.stage(U, x);
    settle_velocities(x, v, C.settle_m, s.invm[U.a[0]], s.invm[U.a[1]]);
    ops.store(s.v, U, v);
    return ops.kinetic(U, v) - ke0;
}

// velocity part of V and O for a rigid water: kick (f != nullptr) or noise, then project; returns change of KE
template <typename real>
__device__ __noinline__ double water_VO(uint32_t ctx_off, int u, double h, int f_slot, double a, double b,
                                        uint64_t rid, uint32_t draw, int want_dke)
{
    const TeamCtx<real>& C = *sm<TeamCtx<real>>(ctx_off);
    const Smem<real> s = carve<real>(pk_smem + C.team_off, C.L);
    const UnitOps<real> ops{C, s};
    const PKUnit U = C.units[u];
    const real* __restrict__ f = f_slot >= 0 ? s.f + (size_t)f_slot * 3 * C.N : nullptr;
    double v[3][3];
    double ke0 = 0.0;
    if (f) {
        ops.load(s.v, U, v);
        if (want_dke) ke0 = ops.kinetic(U, v);
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            const int i = U.a[k];
            const double hm = h * s.invm[i];
#pragma unroll
            for (int c = 0; c < 3; ++c) v[k][c] = v[k][c] + hm * (double)f[3 * i + c];
        }
    } else {
        ops.load(s.v, U, v);
        ke0 = ops.kinetic(U, v);
        ops.noise(U, u, a, b, rid, draw);
        ops.load(s.v, U, v);
    }
    double x[3][3];
    ops.load(s.x, U, x);
    settle_velocities(x, v, C.settle_m, s.invm[U.a[0]], s.invm[U.a[1]]);
    ops.store(s.v, U, v);
    return want_dke ? ops.kinetic(U, v) - ke0 : 0.0;
}

// ---- star of pair constraints (H-bonds): same contracts as the water functions
template <typename real>
__device__ __noinline__ double star_R(uint32_t ctx_off, int u, double h)
{
    const TeamCtx<real>& C = *sm<TeamCtx<real>>(ctx_off);
    const Smem<real> s = carve<real>(pk_smem + C.team_off, C.L);
    const UnitOps<real> ops{C, s};
    const PKUnit U = C.units[u];
    double im[4];
    ops.inv_masses(U, im);
    const double d[3] = {U.d[0], U.d[1], U.d[2]};
    double x[4][3];
    double ke0 = 0.0;
    {
        double x0[4][3], v[4][3];
        ops.load(s.x, U, x0);
        ops.load(s.v, U, v);
        if (h != 0.0) ke0 = ops.kinetic(U, v);
#pragma unroll
        for (int k = 0; k < 4; ++k)
#pragma unroll
            for (int c = 0; c < 3; ++c) x[k][c] = x0[k][c] + h * v[k][c];
        star_positions(x0, x, U.n - 1, d, im, C.tol);
    }
    double v[4][3];
    ops.load(s.v, U, v);
    if (h != 0.0) {
        const double inv_h = fast_rcp(h);
        double x0[4][3];
        ops.load(s.x, U, x0);
#pragma unroll
        for (int k = 0; k < 4; ++k)
#pragma unroll
            for (int c = 0; c < 3; ++c) {
                const double x1 = x0[k][c] + h * v[k][c];
                v[k][c] = v[k][c] + (x[k][c] - x1) * inv_h;
            }
    }
    ops.store(s.x, U, x);
    ops.stage(U, x);
    star_velocities(x, v, U.n - 1, im);
    ops.store(s.v, U, v);
    return ops.kinetic(U, v) - ke0;
}

template <typename real>
__device__ __noinline__ double star_VO(uint32_t ctx_off, int u, double h, int f_slot, double a, double b,
                                       uint64_t rid, uint32_t draw, int want_dke)
{
    const TeamCtx<real>& C = *sm<TeamCtx<real>>(ctx_off);
    const Smem<real> s = carve<real>(pk_smem + C.team_off, C.L);
    const UnitOps<real> ops{C, s};
    const PKUnit U = C.units[u];
    double im[4];
    ops.inv_masses(U, im);
    const real* __restrict__ f = f_slot >= 0 ? s.f + (size_t)f_slot * 3 * C.N : nullptr;
    double v[4][3];
    double ke0 = 0.0;
    ops.load(s.v, U, v);
    if (f) {
        if (want_dke) ke0 = ops.kinetic(U, v);
#pragma unroll
        for (int k = 0; k < 4; ++k)
            if (k < U.n) {
                const int i = U.a[k];
                const double hm = h * im[k];
#pragma unroll
                for (int c = 0; c < 3; ++c) v[k][c] = v[k][c] + hm * (double)f[3 * i + c];
            }
    } else {
        ke0 = ops.kinetic(U, v);
        ops.noise(U, u, a, b, rid, draw);
        ops.load(s.v, U, v);
    }
    double x[4][3];
    ops.load(s.x, U, x);
    star_velocities(x, v, U.n - 1, im);
    ops.store(s.v, U, v);
    return want_dke ? ops.kinetic(U, v) - ke0 : 0.0;
}

// ---- generic clusters (Gauss-Seidel SHAKE / RATTLE in shared memory by the owner lane)
template <typename real>
__device__ __noinline__ double generic_R(uint32_t ctx_off, int u, double h)
{
    const TeamCtx<real>& C = *sm<TeamCtx<real>>(ctx_off);
    const Smem<real> s = carve<real>(pk_smem + C.team_off, C.L);
    const UnitOps<real> ops{C, s};
    const int c = C.units[u].a[0], a0 = C.gc_atom_off[c], a1 = C.gc_atom_off[c + 1];
    const double ke0 = h != 0.0 ? ops.generic_kinetic(a0, a1) : 0.0;
    for (int q = a0; q < a1; ++q) {
        const int i = C.gc_atoms[q];
        for (int k = 0; k < 3; ++k) {
            s.xref[3 * i + k] = s.x[3 * i + k];
            s.x[3 * i + k] = s.x[3 * i + k] + h * s.v[3 * i + k];
        }
    }
    ops.generic_shake(c);
    if (h != 0.0)
        for (int q = a0; q < a1; ++q) {
            const int i = C.gc_atoms[q];
            for (int k = 0; k < 3; ++k) {
                const double x1 = s.xref[3 * i + k] + h * s.v[3 * i + k];
                s.v[3 * i + k] = s.v[3 * i + k] + (s.x[3 * i + k] - x1) / h;
            }
        }
    ops.generic_rattle(c);
    ops.generic_stage(a0, a1);
    return ops.generic_kinetic(a0, a1) - ke0;
}

template <typename real>
__device__ __noinline__ double generic_VO(uint32_t ctx_off, int u, double h, int f_slot, double a, double b,
                                          uint64_t rid, uint32_t draw, int want_dke)
{
    constexpr bool NOISE32 = sizeof(real) == 4;
    const TeamCtx<real>& C = *sm<TeamCtx<real>>(ctx_off);
    const Smem<real> s = carve<real>(pk_smem + C.team_off, C.L);
    const UnitOps<real> ops{C, s};
    const int c = C.units[u].a[0], a0 = C.gc_atom_off[c], a1 = C.gc_atom_off[c + 1];
    const real* __restrict__ f = f_slot >= 0 ? s.f + (size_t)f_slot * 3 * C.N : nullptr;
    const double ke0 = want_dke ? ops.generic_kinetic(a0, a1) : 0.0;
    for (int q = a0; q < a1; ++q) {
        const int i = C.gc_atoms[q];
        if (f) {
            const double hm = h * s.invm[i];
            for (int k = 0; k < 3; ++k) s.v[3 * i + k] = s.v[3 * i + k] + hm * (double)f[3 * i + k];
        } else {
            const N3 n = NOISE32 ? normals3_f32(C.seed, rid, draw, LSI_TAG_O, (uint32_t)i)
                                 : normals3_f64(C.seed, rid, draw, LSI_TAG_O, (uint32_t)i);
            const double sg = b * s.sigv[i];
            s.v[3 * i] = a * s.v[3 * i] + sg * n.x;
            s.v[3 * i + 1] = a * s.v[3 * i + 1] + sg * n.y;
            s.v[3 * i + 2] = a * s.v[3 * i + 2] + sg * n.z;
        }
    }
    ops.generic_rattle(c);
    return want_dke ? ops.generic_kinetic(a0, a1) - ke0 : 0.0;
}

// ---- free atom
template <typename real>
__device__ __noinline__ double free_O(uint32_t ctx_off, int i, double a, double b, uint64_t rid, uint32_t draw)
{
    constexpr bool NOISE32 = sizeof(real) == 4;
    const TeamCtx<real>& C = *sm<TeamCtx<real>>(ctx_off);
    const Smem<real> s = carve<real>(pk_smem + C.team_off, C.L);
    const N3 n = NOISE32 ? normals3_f32(C.seed, rid, draw, LSI_TAG_O, (uint32_t)i) : normals3_f64(C.seed, rid, draw, LSI_TAG_O, (uint32_t)i);
    const double sg = b * s.sigv[i];
    const double v0 = s.v[3 * i], v1 = s.v[3 * i + 1], v2 = s.v[3 * i + 2];
    const double w0 = a * v0 + sg * n.x, w1 = a * v1 + sg * n.y, w2 = a * v2 + sg * n.z;
    s.v[3 * i] = w0; s.v[3 * i + 1] = w1; s.v[3 * i + 2] = w2;
    return 0.5 * s.mass[i] * (w0 * w0 + w1 * w1 + w2 * w2) - 0.5 * s.mass[i] * (v0 * v0 + v1 * v1 + v2 * v2);
}

// ---- dispatch by unit kind (units are sorted by kind, so warps rarely diverge here)
template <typename real>
__device__ __forceinline__ double unit_R(uint32_t ctx_off, int u, double h)
{
    const TeamCtx<real>& C = *sm<TeamCtx<real>>(ctx_off);
    const int kind = C.units[u].kind;
    if (kind == UNIT_WATER) return water_R<real>(ctx_off, u, h);
    if (kind == UNIT_STAR) return star_R<real>(ctx_off, u, h);
    if (kind == UNIT_GENERIC) return generic_R<real>(ctx_off, u, h);
    const Smem<real> s = carve<real>(pk_smem + C.team_off, C.L);
    const UnitOps<real> ops{C, s};
    const int i = C.units[u].a[0];
    const double px = s.x[3 * i] + h * s.v[3 * i], py = s.x[3 * i + 1] + h * s.v[3 * i + 1], pz = s.x[3 * i + 2] + h * s.v[3 * i + 2];
    s.x[3 * i] = px; s.x[3 * i + 1] = py; s.x[3 * i + 2] = pz;
    ops.stage_atom(i, px, py, pz);
    if (h != 0.0) return 0.0;
    return 0.5 * s.mass[i] * (s.v[3 * i] * s.v[3 * i] + s.v[3 * i + 1] * s.v[3 * i + 1] + s.v[3 * i + 2] * s.v[3 * i + 2]);
}

template <typename real>
__device__ __forceinline__ double unit_V(uint32_t ctx_off, int u, double h, int f_slot, int want_dke)
{
    const TeamCtx<real>& C = *sm<TeamCtx<real>>(ctx_off);
    const int kind = C.units[u].kind;
    if (kind == UNIT_WATER) return water_VO<real>(ctx_off, u, h, f_slot, 0.0, 0.0, 0ull, 0u, want_dke);
    if (kind == UNIT_STAR) return star_VO<real>(ctx_off, u, h, f_slot, 0.0, 0.0, 0ull, 0u, want_dke);
    if (kind == UNIT_GENERIC) return generic_VO<real>(ctx_off, u, h, f_slot, 0.0, 0.0, 0ull, 0u, want_dke);
    const Smem<real> s = carve<real>(pk_smem + C.team_off, C.L);
    const real* __restrict__ f = s.f + (size_t)f_slot * 3 * C.N;
    const int i = C.units[u].a[0];
    const double hm = h * s.invm[i];
    const double v0 = s.v[3 * i], v1 = s.v[3 * i + 1], v2 = s.v[3 * i + 2];
    const double w0 = v0 + hm * (double)f[3 * i], w1 = v1 + hm * (double)f[3 * i + 1], w2 = v2 + hm * (double)f[3 * i + 2];
    s.v[3 * i] = w0; s.v[3 * i + 1] = w1; s.v[3 * i + 2] = w2;
    return want_dke ? 0.5 * s.mass[i] * ((w0 * w0 + w1 * w1 + w2 * w2) - (v0 * v0 + v1 * v1 + v2 * v2)) : 0.0;
}

template <typename real>
__device__ __forceinline__ double unit_O(uint32_t ctx_off, int u, double a, double b, uint64_t rid, uint32_t draw)
{
    const TeamCtx<real>& C = *sm<TeamCtx<real>>(ctx_off);
    const int kind = C.units[u].kind;
    if (kind == UNIT_WATER) return water_VO<real>(ctx_off, u, 0.0, -1, a, b, rid, draw, 1);
    if (kind == UNIT_STAR) return star_VO<real>(ctx_off, u, 0.0, -1, a, b, rid, draw, 1);
    if (kind == UNIT_GENERIC) return generic_VO<real>(ctx_off, u, 0.0, -1, a, b, rid, draw, 1);
    return free_O<real>(ctx_off, C.units[u].a[0], a, b, rid, draw);
}

template <typename real>
__device__ __forceinline__ double unit_kinetic(uint32_t ctx_off, int u)
{
    const TeamCtx<real>& C = *sm<TeamCtx<real>>(ctx_off);
    const Smem<real> s = carve<real>(pk_smem + C.team_off, C.L);
    const PKUnit U = C.units[u];
    const UnitOps<real> ops{C, s};
    if (U.kind == UNIT_GENERIC) return ops.generic_kinetic(C.gc_atom_off[U.a[0]], C.gc_atom_off[U.a[0] + 1]);
    double ke = 0.0;
#pragma unroll
    for (int k = 0; k < 4; ++k)
        if (k < U.n) {
            const int i = U.a[k];
            ke += 0.5 * s.mass[i] * (s.v[3 * i] * s.v[3 * i] + s.v[3 * i + 1] * s.v[3 * i + 1] + s.v[3 * i + 2] * s.v[3 * i + 2]);
        }
    return ke;
}

// tables -> team shared memory
template <typename real>
__device__ __forceinline__ void load_tables(const PKParams& P, const Smem<real>& s, int tid, int T)
{
    const typename Vec<real>::r4* nbp = reinterpret_cast<const typename Vec<real>::r4*>(P.nbp);
    for (int i = tid; i < P.N; i += T) {
        const double m = P.mass[i];
        s.mass[i] = m;
        s.invm[i] = 1.0 / m;
        s.sigv[i] = sqrt(P.kT / m);
        const typename Vec<real>::r4 p = nbp[i];
        s.qk[i] = p.x;
        typename Vec<real>::r4 l;
        l.x = p.y; l.y = p.z; l.z = (real)0; l.w = (real)0;
        s.xq[2 * __ldg(P.iperm + i) + 1] = l;
    }
}

// Maxwell-Boltzmann velocities for every atom (start of the protocol, midpoint of the "configuration" marginal)
template <typename real>
__device__ __noinline__ void draw_velocities(uint32_t ctx_off, uint64_t rid, uint32_t draw, uint32_t tag, const double* mass,
                                             double kT, int tid, int T)
{
    const TeamCtx<real>& C = *sm<TeamCtx<real>>(ctx_off);
    const Smem<real> s = carve<real>(pk_smem + C.team_off, C.L);
    for (int i = tid; i < C.N; i += T) {
        const N3 n = normals3_f64(C.seed, rid, draw, tag, (uint32_t)i);
        const double sg = sqrt(kT / mass[i]);
        s.v[3 * i] = sg * n.x; s.v[3 * i + 1] = sg * n.y; s.v[3 * i + 2] = sg * n.z;
    }
}

#define PK_FOR_UNITS(u) for (int u = tid; u < P.n_units; u += T)

template <typename real, bool PERIODIC, bool WT, int APL>
__global__ void __launch_bounds__(WT ? 128 : 384, WT ? LSI_PK_WT_MINB : LSI_PK_CTA_MINB) particles_protocol_kernel(const __grid_constant__ PKParams P, int teams)
{
    const int N = P.N, D = 3 * N;
    // warp-team mode: teams of P.team_lanes = 32 or 16 lanes (two small replicas share a warp and run in lockstep)
    const int T = WT ? P.team_lanes : (int)blockDim.x;
    const int tid = WT ? (int)(threadIdx.x & (T - 1)) : (int)threadIdx.x;
    const int team = WT ? (int)(threadIdx.x / T) : 0;
    int64_t r = WT ? (int64_t)blockIdx.x * teams + team : (int64_t)blockIdx.x;
    if (r >= P.n) {
        if (!WT || (T == 32)) return;  // a whole warp (warp-team mode never uses CTA barriers)
        r = P.n - 1;                   // half-warp team without a replica: shadow the last one (same values written twice)
    }
    const bool DIAG = P.diag != 0;
    const Layout L = layout<real>(N, P.n_slots, P.n_tf, P.n_generic > 0, WT);
    const uint32_t team_off = (uint32_t)team * L.total;
    const Smem<real> s = carve<real>(pk_smem + team_off, L);
    const uint32_t ctx = team_off + L.ctx;  // offset of this team's TeamCtx in shared memory
    if (tid == 0) fill_ctx<real>(sm<TeamCtx<real>>(ctx), P, L, team_off);
    const uint64_t rid = P.replica0 + (uint64_t)r;
    const int64_t TT = (int64_t)P.n_steps + 1;

    // ---- start state
    load_tables<real>(P, s, tid, T);
    {
        const double* xs = P.x0 ? P.x0 + r * D : P.x_cache + lsi::cache_index(P.seed, rid, P.n_cache) * D;
        for (int k = tid; k < D; k += T) s.x[k] = xs[k];
        if (P.v0)
            for (int k = tid; k < D; k += T) s.v[k] = P.v0[r * D + k];
    }
    team_sync<WT>(T);
    if (!P.v0) {
        draw_velocities<real>(ctx, rid, 0u, LSI_TAG_V0, P.mass, P.kT, tid, T);
        team_sync<WT>(T);
    }

    const int scratch = P.n_slots - 1;  // DIAG: an extra slot that is never a cache
    double heat_run = 0.0;  // per-lane share of the heat accumulated so far (all legs)
    for (int leg = 0; leg < P.n_legs; ++leg) {
        if (leg > 0) {
            if (P.flags & LSI_FLAG_ROUND_MIDPOINT_X_FP32)
                for (int k = tid; k < D; k += T) s.x[k] = (double)(float)s.x[k];
            if (P.marginal == LSI_MARGINAL_CONFIGURATION)
                draw_velocities<real>(ctx, rid, (uint32_t)(leg - 1), LSI_TAG_VMID, P.mass, P.kT, tid, T);
            team_sync<WT>(T);
        }
        double ke0 = 0.0;
        PK_FOR_UNITS(u) ke0 += unit_R<real>(ctx, u, 0.0);  // h = 0: project onto the constraints, stage
        unsigned valid = 0u;  // bit per force-cache slot (team-uniform)
        double pe_cur = compute_forces<real, PERIODIC, WT, APL, true>(ctx, -1, scratch, tid, T);
        const double pe0 = pe_cur;
        const double heat0 = heat_run;
        double wd_ke = 0.0, wd_pe = 0.0;
        double E0 = 0.0;
        uint32_t q = 0;
        if (DIAG) {
            E0 = pe0 + team_sum<WT>(ke0, s.red, tid, T);
            const int64_t base = ((int64_t)leg * TT) * P.n + r;
            if (P.trace_x) for (int k = tid; k < D; k += T) P.trace_x[base * D + k] = s.x[k];
            if (P.trace_v) for (int k = tid; k < D; k += T) P.trace_v[base * D + k] = s.v[k];
            if (P.trace_w && tid == 0) P.trace_w[base] = 0.0;
            team_sync<WT>(T);
        }

        for (int st = 0; st < P.n_steps; ++st) {
            for (int op = 0; op < P.n_ops; ++op) {
                const int kind = P.ext_kind ? P.ext_kind[op] : P.op_kind[op];
                const double c0 = P.ext_c0 ? P.ext_c0[op] : P.op_c0[op];
                if (kind == LSI_OP_R) {
                    double dke = 0.0;
                    PK_FOR_UNITS(u) dke += unit_R<real>(ctx, u, c0);
                    valid = 0u;
                    if (DIAG) {
                        wd_ke += dke;
                        const double pe_new = compute_forces<real, PERIODIC, WT, APL, true>(ctx, -1, scratch, tid, T);
                        wd_pe += pe_new - pe_cur;
                        pe_cur = pe_new;
                    }
                } else if (kind == LSI_OP_V) {
                    const int slot = P.ext_slot ? P.ext_slot[op] : P.op_slot[op];
                    if (!((valid >> slot) & 1u)) {
                        compute_forces<real, PERIODIC, WT, APL, false>(ctx, P.slot_group[slot], slot, tid, T);
                        valid |= 1u << slot;
                    }
                    double dke = 0.0;
                    PK_FOR_UNITS(u) dke += unit_V<real>(ctx, u, c0, slot, DIAG ? 1 : 0);
                    if (DIAG) wd_ke += dke;
                } else {
                    const double a = c0, b = P.ext_c1 ? P.ext_c1[op] : P.op_c1[op];
                    const uint32_t draw = ((uint32_t)leg << 26) | q;
                    ++q;
                    double dke = 0.0;
                    PK_FOR_UNITS(u) dke += unit_O<real>(ctx, u, a, b, rid, draw);
                    heat_run += dke;
                }
            }
            if (DIAG) {
                team_sync<WT>(T);
                const int64_t idx = ((int64_t)leg * TT + st + 1) * P.n + r;
                if (P.trace_x) for (int k = tid; k < D; k += T) P.trace_x[idx * D + k] = s.x[k];
                if (P.trace_v) for (int k = tid; k < D; k += T) P.trace_v[idx * D + k] = s.v[k];
                if (P.trace_w) {
                    double ke = 0.0;
                    PK_FOR_UNITS(u) ke += unit_kinetic<real>(ctx, u);
                    // pe_cur is current: DIAG re-evaluates the energy after every R
                    const double tot = team_sum<WT>(ke - (heat_run - heat0), s.red, tid, T);
                    if (tid == 0) P.trace_w[idx] = ((pe_cur + tot) - E0) * P.inv_kT;
                }
                team_sync<WT>(T);  // trace reads of x, v finish before the owners move on
            }
        }
        // ---- end of leg: W = ((E1 - E0) - (heat1 - heat0)) / kT
        {
            const double pe1 = DIAG ? pe_cur : compute_forces<real, PERIODIC, WT, APL, true>(ctx, -1, scratch, tid, T);
            double ke1 = 0.0;
            PK_FOR_UNITS(u) ke1 += unit_kinetic<real>(ctx, u);
            const double dq = heat_run - heat0;
            const double part = team_sum<WT>((ke1 - ke0) - dq, s.red, tid, T);
            double heat_tot = 0.0, wd_tot = 0.0;
            if (P.heat) heat_tot = team_sum<WT>(dq, s.red, tid, T);
            if (DIAG && P.W_direct) wd_tot = team_sum<WT>(wd_ke, s.red, tid, T) + wd_pe;
            if (tid == 0) {
                P.W[(int64_t)leg * P.n + r] = ((pe1 - pe0) + part) * P.inv_kT;
                if (P.heat) P.heat[(int64_t)leg * P.n + r] = heat_tot;
                if (DIAG && P.W_direct) P.W_direct[(int64_t)leg * P.n + r] = wd_tot * P.inv_kT;
            }
        }
        team_sync<WT>(T);
    }
    if (P.x_final) for (int k = tid; k < D; k += T) P.x_final[r * D + k] = s.x[k];
    if (P.v_final) for (int k = tid; k < D; k += T) P.v_final[r * D + k] = s.v[k];
}

// Equilibrium configurations by extra-chance HMC (the sampler behind the reference's sample caches:
// XCGHMCIntegrator with collision_rate = None, benchmark/integrators/kyle/xchmc.py:84-116, driven by
// EquilibriumSimulator.collect_equilibrium_samples, benchmark/testsystems/bookkeepers.py:66-113).
// Per round: fresh Maxwell-Boltzmann velocities (projected on the constraints), E_old, one uniform u; up to
// 1 + extra chances of `hmc_steps` velocity-Verlet steps each (V(h/2) R(h) V(h/2), projections after every
// substep); after each, mu1 = max(mu1, min(1, exp(-(E_new - E_old)/kT))); the first chance with u <= mu1 is
// accepted, otherwise the configuration is restored.  Targets exp(-U/kT) on the constraint manifold exactly.
// Streams: velocities Philox(replica, draw = 2 round, TAG_EQ, unit = atom); u = Philox(replica, 2 round + 1, TAG_EQ, 0).x
template <typename real, bool PERIODIC, bool WT, int APL>
__global__ void __launch_bounds__(WT ? 128 : 384, WT ? LSI_PK_WT_MINB : LSI_PK_CTA_MINB) particles_hmc_kernel(const __grid_constant__ PKParams P, int teams)
{
    const int N = P.N, D = 3 * N;
    // warp-team mode: teams of P.team_lanes = 32 or 16 lanes (two small replicas share a warp and run in lockstep)
    const int T = WT ? P.team_lanes : (int)blockDim.x;
    const int tid = WT ? (int)(threadIdx.x & (T - 1)) : (int)threadIdx.x;
    const int team = WT ? (int)(threadIdx.x / T) : 0;
    int64_t r = WT ? (int64_t)blockIdx.x * teams + team : (int64_t)blockIdx.x;
    if (r >= P.n) {
        if (!WT || (T == 32)) return;
        r = P.n - 1;  // half-warp team without a replica: shadow the last one
    }
    const Layout L = layout<real>(N, 2, P.n_tf, P.n_generic > 0, WT, true);
    const uint32_t team_off = (uint32_t)team * L.total;
    const Smem<real> s = carve<real>(pk_smem + team_off, L);
    double* const xold = sm<double>(team_off + L.xold);
    double* const vold = sm<double>(team_off + L.vold);
    const uint32_t ctx = team_off + L.ctx;
    if (tid == 0) fill_ctx<real>(sm<TeamCtx<real>>(ctx), P, L, team_off);
    const uint64_t rid = P.replica0 + (uint64_t)r;
    load_tables<real>(P, s, tid, T);
    for (int k = tid; k < D; k += T) { s.x[k] = P.x0[r * D + k]; s.v[k] = 0.0; }
    team_sync<WT>(T);
    PK_FOR_UNITS(u) unit_R<real>(ctx, u, 0.0);  // project the start configuration, stage it
    double pe = compute_forces<real, PERIODIC, WT, APL, true>(ctx, -1, 0, tid, T);
    const double h = P.hmc_dt;
    int n_acc = 0, n_first = 0;
    for (int round = 0; round < P.hmc_rounds; ++round) {
        draw_velocities<real>(ctx, rid, 2u * (uint32_t)round, LSI_TAG_EQ, P.mass, P.kT, tid, T);
        team_sync<WT>(T);
        double ke = 0.0;
        PK_FOR_UNITS(u) ke += unit_R<real>(ctx, u, 0.0);  // velocity constraints (positions already satisfy theirs)
        const double e_old = pe + team_sum<WT>(ke, s.red, tid, T);
        team_sync<WT>(T);
        for (int k = tid; k < D; k += T) { xold[k] = s.x[k]; vold[k] = s.v[k]; s.f[D + k] = s.f[k]; }
        const lsi::Philox4 ur = lsi::philox_block(P.seed, rid, 2u * (uint32_t)round + 1u, LSI_TAG_EQ, 0u);
        const double uni = ((double)ur.x + 0.5) * 2.3283064365386962890625e-10;
        double mu1 = 0.0, pe_new = pe;
        for (int c = 0; c <= P.hmc_extra; ++c) {
            if (!(uni > mu1)) break;
            for (int st = 0; st < P.hmc_steps; ++st) {
                PK_FOR_UNITS(u) unit_V<real>(ctx, u, 0.5 * h, 0, 0);
                PK_FOR_UNITS(u) unit_R<real>(ctx, u, h);
                if (st + 1 < P.hmc_steps) compute_forces<real, PERIODIC, WT, APL, false>(ctx, -1, 0, tid, T);
                else pe_new = compute_forces<real, PERIODIC, WT, APL, true>(ctx, -1, 0, tid, T);
                PK_FOR_UNITS(u) unit_V<real>(ctx, u, 0.5 * h, 0, 0);
            }
            double ke1 = 0.0;
            PK_FOR_UNITS(u) ke1 += unit_kinetic<real>(ctx, u);
            double e_new = pe_new + team_sum<WT>(ke1, s.red, tid, T);
            if (!(e_new == e_new)) e_new = INFINITY;  // nan_to_inf (xchmc.py:92)
            mu1 = fmax(mu1, fmin(1.0, exp(-(e_new - e_old) * P.inv_kT)));
            if (c == 0 && uni <= mu1) ++n_first;
        }
        team_sync<WT>(T);
        if (uni > mu1) {  // every chance failed: back to the start of the round
            for (int k = tid; k < D; k += T) { s.x[k] = xold[k]; s.v[k] = vold[k]; s.f[k] = s.f[D + k]; }
            team_sync<WT>(T);
            PK_FOR_UNITS(u) unit_R<real>(ctx, u, 0.0);  // re-stage the restored positions
            team_sync<WT>(T);
        } else {
            pe = pe_new;
            ++n_acc;
        }
    }
    team_sync<WT>(T);
    if (P.x_final) for (int k = tid; k < D; k += T) P.x_final[r * D + k] = s.x[k];
    if (P.v_final) for (int k = tid; k < D; k += T) P.v_final[r * D + k] = s.v[k];
    if (P.hmc_accept && tid == 0) {
        P.hmc_accept[2 * r] = (double)n_acc / (double)(P.hmc_rounds > 0 ? P.hmc_rounds : 1);
        P.hmc_accept[2 * r + 1] = (double)n_first / (double)(P.hmc_rounds > 0 ? P.hmc_rounds : 1);
    }
}

// potential energy and forces of force group q_group for n configurations
template <typename real, bool PERIODIC, bool WT, int APL>
__global__ void __launch_bounds__(WT ? 128 : 384) particles_energy_kernel(const __grid_constant__ PKParams P, int teams)
{
    const int N = P.N, D = 3 * N;
    // warp-team mode: teams of P.team_lanes = 32 or 16 lanes (two small replicas share a warp and run in lockstep)
    const int T = WT ? P.team_lanes : (int)blockDim.x;
    const int tid = WT ? (int)(threadIdx.x & (T - 1)) : (int)threadIdx.x;
    const int team = WT ? (int)(threadIdx.x / T) : 0;
    int64_t r = WT ? (int64_t)blockIdx.x * teams + team : (int64_t)blockIdx.x;
    if (r >= P.n) {
        if (!WT || (T == 32)) return;
        r = P.n - 1;  // half-warp team without a replica: shadow the last one
    }
    const Layout L = layout<real>(N, 1, P.n_tf, false, WT);
    const uint32_t team_off = (uint32_t)team * L.total;
    const Smem<real> s = carve<real>(pk_smem + team_off, L);
    const uint32_t ctx = team_off + L.ctx;  // offset of this team's TeamCtx in shared memory
    if (tid == 0) fill_ctx<real>(sm<TeamCtx<real>>(ctx), P, L, team_off);
    load_tables<real>(P, s, tid, T);
    for (int k = tid; k < D; k += T) s.x[k] = P.q_x[r * D + k];
    team_sync<WT>(T);
    {
        const UnitOps<real> ops{*sm<TeamCtx<real>>(ctx), s};
        for (int i = tid; i < N; i += T) ops.stage_atom(i, s.x[3 * i], s.x[3 * i + 1], s.x[3 * i + 2]);
    }
    const double e = compute_forces<real, PERIODIC, WT, APL, true>(ctx, P.q_group, 0, tid, T);
    if (P.q_energy && tid == 0) P.q_energy[r] = e;
    if (P.q_forces) for (int k = tid; k < D; k += T) P.q_forces[r * D + k] = (double)s.f[k];
}

// ------------------------------------------------------------------------------------------ launch
template <typename K>
int set_smem(K kernel, size_t bytes)
{
    if (bytes > 48 * 1024) LSI_CUDA_CHECK(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
    // tuning knob: preferred shared-memory carve-out in percent (measured: the driver's default choice is the fastest)
    if (const char* c = getenv("LSI_PK_CARVEOUT"))
        if (*c) LSI_CUDA_CHECK(cudaFuncSetAttribute(kernel, cudaFuncAttributePreferredSharedMemoryCarveout, atoi(c)));
    return LSI_OK;
}

template <typename real, bool PERIODIC, bool WT, int APL>
int launch_protocol3(const PKParams& P, const PKLaunch& L, cudaStream_t st)
{
    static_assert(sizeof(TeamCtx<real>) <= 768, "Layout reserves 768 bytes for TeamCtx");
    const unsigned grid = (unsigned)(WT ? (L.n + L.teams - 1) / L.teams : L.n);
    auto k = particles_protocol_kernel<real, PERIODIC, WT, APL>;
    int rc;
    if ((rc = set_smem(k, L.smem)) != LSI_OK) return rc;
    k<<<grid, L.threads, L.smem, st>>>(P, L.teams);
    LSI_CUDA_CHECK(cudaGetLastError());
    return LSI_OK;
}

template <typename real, bool PERIODIC, bool WT, int APL>
int launch_hmc3(const PKParams& P, const PKLaunch& L, cudaStream_t st)
{
    const unsigned grid = (unsigned)(WT ? (L.n + L.teams - 1) / L.teams : L.n);
    auto k = particles_hmc_kernel<real, PERIODIC, WT, APL>;
    int rc;
    if ((rc = set_smem(k, L.smem)) != LSI_OK) return rc;
    k<<<grid, L.threads, L.smem, st>>>(P, L.teams);
    LSI_CUDA_CHECK(cudaGetLastError());
    return LSI_OK;
}

template <typename real, bool PERIODIC, bool WT, int APL>
int launch_energy3(const PKParams& P, const PKLaunch& L, cudaStream_t st)
{
    const unsigned grid = (unsigned)(WT ? (L.n + L.teams - 1) / L.teams : L.n);
    auto k = particles_energy_kernel<real, PERIODIC, WT, APL>;
    int rc;
    if ((rc = set_smem(k, L.smem)) != LSI_OK) return rc;
    k<<<grid, L.threads, L.smem, st>>>(P, L.teams);
    LSI_CUDA_CHECK(cudaGetLastError());
    return LSI_OK;
}

#define PK_DISPATCH(FN, real)                                                                     \
    if (L.warp_team) {                                                                            \
        if (L.apl == 1) return L.periodic ? FN<real, true, true, 1>(P, L, st) : FN<real, false, true, 1>(P, L, st); \
        return L.periodic ? FN<real, true, true, 2>(P, L, st) : FN<real, false, true, 2>(P, L, st); \
    }                                                                                             \
    if (L.apl == 1) return L.periodic ? FN<real, true, false, 1>(P, L, st) : FN<real, false, false, 1>(P, L, st); \
    return L.periodic ? FN<real, true, false, 2>(P, L, st) : FN<real, false, false, 2>(P, L, st);

template <typename real>
int launch_protocol(const PKParams& P, const PKLaunch& L, void* stream)
{
    cudaStream_t st = (cudaStream_t)stream;
    PK_DISPATCH(launch_protocol3, real)
}

template <typename real>
int launch_hmc(const PKParams& P, const PKLaunch& L, void* stream)
{
    cudaStream_t st = (cudaStream_t)stream;
    PK_DISPATCH(launch_hmc3, real)
}

template <typename real>
int launch_energy(const PKParams& P, const PKLaunch& L, void* stream)
{
    cudaStream_t st = (cudaStream_t)stream;
    PK_DISPATCH(launch_energy3, real)
}

}  // namespace pk

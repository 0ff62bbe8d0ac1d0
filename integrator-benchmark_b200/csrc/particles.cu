// K2/K3: whole-protocol kernel for particle systems (water cluster, alanine dipeptide,
// tiny water box): one CTA per replica, the replica's state resident in shared memory for
// the entire protocol (forward leg, midpoint, reverse leg).
//
// Reference semantics (paths relative to the reference repository):
//   substeps R / V / V{g} / O with constraints     benchmark/integrators/langevin.py:124-175
//   protocol, work = (dE - dheat)/kT               benchmark/testsystems/bookkeepers.py:184-295
//   systems                                        benchmark/testsystems/{watercluster,waterbox,alanine_dipeptide}.py
// OpenMM supplies forces, SETTLE/CCMA and the `gaussian` stream there; here:
//   * state (x, v) and every accumulator are fp64 in shared memory; forces are evaluated in
//     `real` = float ("mixed", as OpenMM's CUDA platform is configured by
//     testsystems/configuration.py:42-45) or double, from positions staged as real4 {x, y, z, q}
//   * thread-per-atom full-shell nonbonded loop: partner positions are shared-memory
//     broadcasts, each thread owns the force on its atom in registers -> no atomics, no
//     cross-thread force reduction, bit-reproducible
//   * bonded terms and 1-4 exceptions are gathered per atom (every participant evaluates its own
//     gradient of the term), same property
//   * per-force-group force caches with validity bits: a group is re-evaluated only after an
//     R substep moved the positions
//   * rigid waters: analytic SETTLE positions + exact 3x3 velocity projection, one thread per
//     water; other distance constraints: SHAKE/RATTLE per connected cluster, one thread per
//     cluster, both in fp64
//   * O substeps draw Philox-4x32-10 noise keyed by (replica, leg, draw, atom)
#include <cuda_runtime.h>

#include <algorithm>
#include <cmath>
#include <cstring>
#include <map>
#include <mutex>

#include "lsi_internal.h"
#include "rng.cuh"

#define LSI_COULOMB 138.935456  // kJ nm / (mol e^2), benchmark/integrators/kyle/constants.py:8

namespace {

enum { TERM_BOND = 0, TERM_ANGLE = 1, TERM_TORSION = 2, TERM_EXCEPTION = 3 };

struct TermInst {
    int32_t kind, role;
    int32_t a[4];
    double p[3];
};

struct DevTables {
    double* mass = nullptr;      // [N]
    void* nbp_f = nullptr;       // float4 [N]: q sqrt(K), sigma/2, 2 sqrt(eps), 0
    void* nbp_d = nullptr;       // double4 [N]
    uint32_t* excl = nullptr;    // [N][excl_words]
    TermInst* terms = nullptr;   // gathered per atom
    int* term_off = nullptr;     // [N+1]
    int* settle = nullptr;       // [n_settle][3]
    int* cl_off = nullptr;       // [n_clusters+1]
    int* cl_atoms = nullptr;     // [n_cons][2]
    double* cl_dist = nullptr;   // [n_cons]
    unsigned char* constrained = nullptr;  // [N] atom takes part in some constraint
};

}  // namespace

struct lsi_particles {
    int n_atoms = 0;
    std::vector<double> mass, charge, sigma, epsilon;
    int nb_method = 0;
    double cutoff = 0, box[3] = {0, 0, 0}, rf_dielectric = 78.3;
    double restraint_k = 0;
    int g_nb = 0, g_bond = 0, g_angle = 0, g_tors = 0, g_restr = 0;
    double oh = 0, hh = 0, tol = 1e-8;
    int excl_words = 1;
    std::vector<uint32_t> excl;
    std::vector<TermInst> terms;
    std::vector<int> term_off;
    std::vector<int> settle;
    std::vector<int> cl_off, cl_atoms;
    std::vector<double> cl_dist;
    std::vector<unsigned char> constrained;
    std::mutex mu;
    std::map<int, DevTables> dev;  // per CUDA device
};

namespace {

struct PKParams {
    // system
    int N, nb_method, excl_words, n_settle, n_clusters, has_constraints;
    double box[3], inv_box[3];
    double cutoff2, krf, crf, restraint_k;
    int g_nb, g_bond, g_angle, g_tors, g_restr;
    double oh, hh, tol;
    const double* mass;
    const void* nbp;
    const uint32_t* excl;
    const TermInst* terms;
    const int* term_off;
    const int* settle;
    const int* cl_off;
    const int* cl_atoms;
    const double* cl_dist;
    const unsigned char* constrained;
    // program (device arrays)
    int n_ops, n_O, n_slots;
    int slot_group[4];
    const unsigned char* op_kind;
    const unsigned char* op_slot;
    const double* op_c0;  // R: h   V: h   O: a
    const double* op_c1;  //               O: b
    // protocol
    uint64_t seed, replica0;
    int64_t n;
    int n_steps, n_legs, marginal, flags;
    double kT, inv_kT;
    const double* x_cache;
    int64_t n_cache;
    const double* x0;
    const double* v0;
    double *W, *heat, *W_direct, *x_final, *v_final, *trace_x, *trace_v, *trace_w;
};

template <typename real> struct R4;
template <> struct R4<float> { using type = float4; };
template <> struct R4<double> { using type = double4; };

__device__ __forceinline__ float rsqrt_(float x) { return rsqrtf(x); }
__device__ __forceinline__ double rsqrt_(double x) { return rsqrt(x); }
__device__ __forceinline__ float rint_(float x) { return rintf(x); }
__device__ __forceinline__ double rint_(double x) { return rint(x); }

// shared-memory view of one replica
template <typename real>
struct Smem {
    double* x;     // [3N]
    double* v;     // [3N]
    double* xref;  // [3N]
    double* red;   // [64]
    typename R4<real>::type* xq;  // [N]
    real* lj;      // [2N] sigma/2, 2 sqrt(eps)
    real* f;       // [n_slots][3N]
};

template <typename real>
__device__ __forceinline__ Smem<real> carve(unsigned char* base, int N, int n_slots)
{
    Smem<real> s;
    double* d = reinterpret_cast<double*>(base);
    s.x = d; d += 3 * N;
    s.v = d; d += 3 * N;
    s.xref = d; d += 3 * N;
    s.red = d; d += 64;
    size_t off = (size_t)(reinterpret_cast<unsigned char*>(d) - base);
    off = (off + 31) & ~(size_t)31;
    s.xq = reinterpret_cast<typename R4<real>::type*>(base + off);
    off += sizeof(typename R4<real>::type) * (size_t)N;
    s.lj = reinterpret_cast<real*>(base + off);
    off += sizeof(real) * 2 * (size_t)N;
    off = (off + 15) & ~(size_t)15;
    s.f = reinterpret_cast<real*>(base + off);
    (void)n_slots;
    return s;
}

template <typename real>
size_t smem_bytes(int N, int n_slots)
{
    size_t off = sizeof(double) * (9 * (size_t)N + 64);
    off = (off + 31) & ~(size_t)31;
    off += sizeof(typename R4<real>::type) * (size_t)N + sizeof(real) * 2 * (size_t)N;
    off = (off + 15) & ~(size_t)15;
    off += sizeof(real) * 3 * (size_t)N * (size_t)n_slots;
    return off;
}

__device__ __forceinline__ double block_sum(double val, double* red)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) val += __shfl_xor_sync(0xffffffffu, val, o);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
    __syncthreads();
    if (lane == 0) red[warp] = val;
    __syncthreads();
    double s = 0.0;
    for (int w = 0; w < nw; ++w) s += red[w];
    return s;
}

__device__ __forceinline__ double kinetic_energy(const PKParams& P, const double* v, double* red)
{
    double ke = 0.0;
    for (int i = threadIdx.x; i < P.N; i += blockDim.x)
        ke += 0.5 * P.mass[i] * (v[3 * i] * v[3 * i] + v[3 * i + 1] * v[3 * i + 1] + v[3 * i + 2] * v[3 * i + 2]);
    return block_sum(ke, red);
}

// x (fp64, smem) -> xq (force precision), wrapped into the box when periodic
template <typename real>
__device__ __forceinline__ void stage_positions(const PKParams& P, const Smem<real>& s)
{
    const typename R4<real>::type* nbp = reinterpret_cast<const typename R4<real>::type*>(P.nbp);
    for (int i = threadIdx.x; i < P.N; i += blockDim.x) {
        double px = s.x[3 * i], py = s.x[3 * i + 1], pz = s.x[3 * i + 2];
        if (P.nb_method == LSI_NB_CUTOFF_PERIODIC) {
            px -= P.box[0] * floor(px * P.inv_box[0]);
            py -= P.box[1] * floor(py * P.inv_box[1]);
            pz -= P.box[2] * floor(pz * P.inv_box[2]);
        }
        typename R4<real>::type q;
        q.x = (real)px; q.y = (real)py; q.z = (real)pz; q.w = nbp[i].x;
        s.xq[i] = q;
    }
}

template <typename real>
__device__ __forceinline__ void load_lj(const PKParams& P, const Smem<real>& s)
{
    const typename R4<real>::type* nbp = reinterpret_cast<const typename R4<real>::type*>(P.nbp);
    for (int i = threadIdx.x; i < P.N; i += blockDim.x) {
        s.lj[2 * i] = nbp[i].y;
        s.lj[2 * i + 1] = nbp[i].z;
    }
}

// minimum image of a displacement (positions are staged wrapped into the box, so every bonded
// displacement needs it too; equals the unwrapped displacement OpenMM's bonded forces use as long as
// bonded atoms are closer than half a box edge)
template <typename real, bool PERIODIC>
__device__ __forceinline__ void min_image(const PKParams& P, real& dx, real& dy, real& dz)
{
    if (PERIODIC) {
        dx -= (real)P.box[0] * rint_(dx * (real)P.inv_box[0]);
        dy -= (real)P.box[1] * rint_(dy * (real)P.inv_box[1]);
        dz -= (real)P.box[2] * rint_(dz * (real)P.inv_box[2]);
    }
}

// gradient pieces of one bonded term / exception for the atom in `role`; energy returned in full
template <typename real, bool PERIODIC>
__device__ __forceinline__ real term_force(const PKParams& P, const Smem<real>& s, const TermInst& t, real& fx,
                                           real& fy, real& fz)
{
    const typename R4<real>::type p0 = s.xq[t.a[0]], p1 = s.xq[t.a[1]];
    if (t.kind == TERM_BOND || t.kind == TERM_EXCEPTION) {
        real dx = p1.x - p0.x, dy = p1.y - p0.y, dz = p1.z - p0.z;  // from atom 0 to atom 1
        min_image<real, PERIODIC>(P, dx, dy, dz);
        const real r2 = dx * dx + dy * dy + dz * dz;
        real fr, e;
        if (t.kind == TERM_BOND) {
            const real inv_r = rsqrt_(r2), r = r2 * inv_r;
            const real dr = r - (real)t.p[0];
            e = (real)0.5 * (real)t.p[1] * dr * dr;
            fr = (real)t.p[1] * dr * inv_r;  // dE/dr / r ; force on atom 0 = +fr d
        } else {
            if (PERIODIC && r2 >= (real)P.cutoff2) return (real)0;
            const real inv_r = rsqrt_(r2), inv_r2 = inv_r * inv_r;
            const real s2 = (real)(t.p[1] * t.p[1]) * inv_r2, s6 = s2 * s2 * s2;
            const real qq = (real)(LSI_COULOMB * t.p[0]), eps4 = (real)(4.0 * t.p[2]);
            e = eps4 * (s6 * s6 - s6) + qq * inv_r;
            fr = -(eps4 * ((real)12 * s6 * s6 - (real)6 * s6) * inv_r2 + qq * inv_r * inv_r2);
        }
        const real sgn = t.role == 0 ? (real)1 : (real)-1;
        fx += sgn * fr * dx; fy += sgn * fr * dy; fz += sgn * fr * dz;
        return e;
    }
    const typename R4<real>::type p2 = s.xq[t.a[2]];
    if (t.kind == TERM_ANGLE) {
        real ax = p0.x - p1.x, ay = p0.y - p1.y, az = p0.z - p1.z;
        real bx = p2.x - p1.x, by = p2.y - p1.y, bz = p2.z - p1.z;
        min_image<real, PERIODIC>(P, ax, ay, az);
        min_image<real, PERIODIC>(P, bx, by, bz);
        const real ra2 = ax * ax + ay * ay + az * az, rb2 = bx * bx + by * by + bz * bz;
        const real ira = rsqrt_(ra2), irb = rsqrt_(rb2);
        real ct = (ax * bx + ay * by + az * bz) * ira * irb;
        ct = fmin(fmax(ct, (real)-1), (real)1);
        const real th = acos(ct);
        const real dth = th - (real)t.p[0];
        const real e = (real)0.5 * (real)t.p[1] * dth * dth;
        const real dE = (real)t.p[1] * dth;
        real st = sqrt(fmax((real)1 - ct * ct, (real)1e-24));
        const real c = dE / st;  // force = +c * (d cos(theta)/d r)
        // d cos/da = b/(ra rb) - ct a/ra^2 ; d cos/db = a/(ra rb) - ct b/rb^2
        const real gax = bx * ira * irb - ct * ax * ira * ira, gay = by * ira * irb - ct * ay * ira * ira,
                   gaz = bz * ira * irb - ct * az * ira * ira;
        const real gbx = ax * ira * irb - ct * bx * irb * irb, gby = ay * ira * irb - ct * by * irb * irb,
                   gbz = az * ira * irb - ct * bz * irb * irb;
        if (t.role == 0) { fx += c * gax; fy += c * gay; fz += c * gaz; }
        else if (t.role == 2) { fx += c * gbx; fy += c * gby; fz += c * gbz; }
        else { fx -= c * (gax + gbx); fy -= c * (gay + gby); fz -= c * (gaz + gbz); }
        return e;
    }
    // periodic torsion
    const typename R4<real>::type p3 = s.xq[t.a[3]];
    real b1x = p1.x - p0.x, b1y = p1.y - p0.y, b1z = p1.z - p0.z;
    real b2x = p2.x - p1.x, b2y = p2.y - p1.y, b2z = p2.z - p1.z;
    real b3x = p3.x - p2.x, b3y = p3.y - p2.y, b3z = p3.z - p2.z;
    min_image<real, PERIODIC>(P, b1x, b1y, b1z);
    min_image<real, PERIODIC>(P, b2x, b2y, b2z);
    min_image<real, PERIODIC>(P, b3x, b3y, b3z);
    const real n1x = b1y * b2z - b1z * b2y, n1y = b1z * b2x - b1x * b2z, n1z = b1x * b2y - b1y * b2x;
    const real n2x = b2y * b3z - b2z * b3y, n2y = b2z * b3x - b2x * b3z, n2z = b2x * b3y - b2y * b3x;
    const real b2sq = b2x * b2x + b2y * b2y + b2z * b2z;
    const real b2n = sqrt(b2sq);
    const real yy = b2n * (b1x * n2x + b1y * n2y + b1z * n2z);
    const real xx = n1x * n2x + n1y * n2y + n1z * n2z;
    const real phi = atan2(yy, xx);
    const real arg = (real)t.p[0] * phi - (real)t.p[1];
    const real e = (real)t.p[2] * ((real)1 + cos(arg));
    const real dE = -(real)t.p[2] * (real)t.p[0] * sin(arg);
    const real n1sq = n1x * n1x + n1y * n1y + n1z * n1z, n2sq = n2x * n2x + n2y * n2y + n2z * n2z;
    const real c1 = -b2n / n1sq, c4 = b2n / n2sq;
    const real g1x = c1 * n1x, g1y = c1 * n1y, g1z = c1 * n1z;
    const real g4x = c4 * n2x, g4y = c4 * n2y, g4z = c4 * n2z;
    const real u = (b1x * b2x + b1y * b2y + b1z * b2z) / b2sq, w = (b3x * b2x + b3y * b2y + b3z * b2z) / b2sq;
    real gx, gy, gz;
    if (t.role == 0) { gx = g1x; gy = g1y; gz = g1z; }
    else if (t.role == 3) { gx = g4x; gy = g4y; gz = g4z; }
    else if (t.role == 1) { gx = -g1x - u * g1x + w * g4x; gy = -g1y - u * g1y + w * g4y; gz = -g1z - u * g1z + w * g4z; }
    else { gx = -g4x + u * g1x - w * g4x; gy = -g4y + u * g1y - w * g4y; gz = -g4z + u * g1z - w * g4z; }
    fx -= dE * gx; fy -= dE * gy; fz -= dE * gz;
    return e;
}

// forces of force group `g` (-1 = all) into fout[3N] (smem), optional energy (block sum, fp64)
template <typename real, bool PERIODIC, bool ENERGY>
__device__ double compute_forces(const PKParams& P, const Smem<real>& s, int g, real* __restrict__ fout)
{
    const bool do_nb = g < 0 || g == P.g_nb;
    const bool do_restr = (g < 0 || g == P.g_restr) && P.restraint_k != 0.0;
    double e_thread = 0.0;
    for (int i = threadIdx.x; i < P.N; i += blockDim.x) {
        real fx = 0, fy = 0, fz = 0, en = 0;
        const typename R4<real>::type pi = s.xq[i];
        if (do_nb) {
            const real hs_i = s.lj[2 * i], se_i = s.lj[2 * i + 1];
            const uint32_t* mask = P.excl + (size_t)i * P.excl_words;
            uint32_t mword = 0;
            for (int j = 0; j < P.N; ++j) {
                if ((j & 31) == 0) mword = __ldg(mask + (j >> 5));
                const bool skip = (mword >> (j & 31)) & 1u;  // excluded, exception or self
                const typename R4<real>::type pj = s.xq[j];
                real dx = pj.x - pi.x, dy = pj.y - pi.y, dz = pj.z - pi.z;
                if (PERIODIC) {
                    dx -= (real)P.box[0] * rint_(dx * (real)P.inv_box[0]);
                    dy -= (real)P.box[1] * rint_(dy * (real)P.inv_box[1]);
                    dz -= (real)P.box[2] * rint_(dz * (real)P.inv_box[2]);
                }
                real r2 = dx * dx + dy * dy + dz * dz;
                if (skip || (PERIODIC && r2 >= (real)P.cutoff2)) continue;
                const real inv_r = rsqrt_(r2), inv_r2 = inv_r * inv_r;
                const real sig = hs_i + s.lj[2 * j];
                const real eps4 = se_i * s.lj[2 * j + 1];
                const real s2 = sig * sig * inv_r2, s6 = s2 * s2 * s2;
                const real qq = pi.w * pj.w;
                real fr = eps4 * ((real)12 * s6 * s6 - (real)6 * s6) * inv_r2 + qq * inv_r * inv_r2;
                if (PERIODIC) fr -= qq * (real)(2.0 * P.krf);
                fx -= fr * dx; fy -= fr * dy; fz -= fr * dz;
                if (ENERGY) {
                    real e = eps4 * (s6 * s6 - s6) + qq * inv_r;
                    if (PERIODIC) e += qq * ((real)P.krf * r2 - (real)P.crf);
                    en += e;
                }
            }
            en *= (real)0.5;  // every pair is visited from both ends
        }
        // bonded terms and 1-4 exceptions this atom takes part in
        for (int k = P.term_off[i]; k < P.term_off[i + 1]; ++k) {
            const TermInst t = P.terms[k];
            const int tg = t.kind == TERM_BOND ? P.g_bond : t.kind == TERM_ANGLE ? P.g_angle
                           : t.kind == TERM_TORSION ? P.g_tors : P.g_nb;
            if (g >= 0 && tg != g) continue;
            const real e = term_force<real, PERIODIC>(P, s, t, fx, fy, fz);
            if (ENERGY && t.role == 0) en += e;
        }
        if (do_restr) {
            // restraint acts on the true (unwrapped) coordinates
            const real k = (real)P.restraint_k;
            const real px = (real)s.x[3 * i], py = (real)s.x[3 * i + 1], pz = (real)s.x[3 * i + 2];
            fx -= k * px; fy -= k * py; fz -= k * pz;
            if (ENERGY) en += (real)0.5 * k * (px * px + py * py + pz * pz);
        }
        fout[3 * i] = fx; fout[3 * i + 1] = fy; fout[3 * i + 2] = fz;
        e_thread += (double)en;
    }
    if (ENERGY) return block_sum(e_thread, s.red);
    return 0.0;
}

// ---- constraints (fp64) ------------------------------------------------------------------
__device__ __forceinline__ double dot3(const double* a, const double* b) { return a[0] * b[0] + a[1] * b[1] + a[2] * b[2]; }

// analytic SETTLE (Miyamoto & Kollman 1992): x0 reference triangle, x1 unconstrained -> constrained
__device__ void settle_positions(const double* x0, double* x1, const int* at, double mO, double mH, double dOH,
                                 double dHH)
{
    const double *a0 = x0 + 3 * at[0], *b0 = x0 + 3 * at[1], *c0 = x0 + 3 * at[2];
    double *a1 = x1 + 3 * at[0], *b1 = x1 + 3 * at[1], *c1 = x1 + 3 * at[2];
    const double inv_mtot = 1.0 / (mO + 2.0 * mH);
    double xb0[3], xc0[3], xcom[3], xa1[3], xb1[3], xc1[3];
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        xb0[c] = b0[c] - a0[c];
        xc0[c] = c0[c] - a0[c];
        xcom[c] = (mO * a1[c] + mH * b1[c] + mH * c1[c]) * inv_mtot;
        xa1[c] = a1[c] - xcom[c];
        xb1[c] = b1[c] - xcom[c];
        xc1[c] = c1[c] - xcom[c];
    }
    double Z[3] = {xb0[1] * xc0[2] - xb0[2] * xc0[1], xb0[2] * xc0[0] - xb0[0] * xc0[2], xb0[0] * xc0[1] - xb0[1] * xc0[0]};
    double X[3] = {xa1[1] * Z[2] - xa1[2] * Z[1], xa1[2] * Z[0] - xa1[0] * Z[2], xa1[0] * Z[1] - xa1[1] * Z[0]};
    double Y[3] = {Z[1] * X[2] - Z[2] * X[1], Z[2] * X[0] - Z[0] * X[2], Z[0] * X[1] - Z[1] * X[0]};
    const double nz = rsqrt(dot3(Z, Z)), nx = rsqrt(dot3(X, X)), ny = rsqrt(dot3(Y, Y));
#pragma unroll
    for (int c = 0; c < 3; ++c) { X[c] *= nx; Y[c] *= ny; Z[c] *= nz; }
    const double xb0d = dot3(X, xb0), yb0d = dot3(Y, xb0), xc0d = dot3(X, xc0), yc0d = dot3(Y, xc0);
    const double za1d = dot3(Z, xa1);
    const double xb1d = dot3(X, xb1), yb1d = dot3(Y, xb1), zb1d = dot3(Z, xb1);
    const double xc1d = dot3(X, xc1), yc1d = dot3(Y, xc1), zc1d = dot3(Z, xc1);
    const double rc = 0.5 * dHH;
    const double hgt = sqrt(dOH * dOH - rc * rc);
    const double ra = hgt * 2.0 * mH * inv_mtot, rb = hgt - ra;
    const double sinphi = za1d / ra, cosphi = sqrt(1.0 - sinphi * sinphi);
    const double sinpsi = (zb1d - zc1d) / (2.0 * rc * cosphi), cospsi = sqrt(1.0 - sinpsi * sinpsi);
    const double ya2d = ra * cosphi;
    const double xb2d = -rc * cospsi;
    const double yb2d = -rb * cosphi - rc * sinpsi * sinphi;
    const double yc2d = -rb * cosphi + rc * sinpsi * sinphi;
    const double alpha = xb2d * (xb0d - xc0d) + yb0d * yb2d + yc0d * yc2d;
    const double beta = xb2d * (yc0d - yb0d) + xb0d * yb2d + xc0d * yc2d;
    const double gamma = xb0d * yb1d - xb1d * yb0d + xc0d * yc1d - xc1d * yc0d;
    const double al2be2 = alpha * alpha + beta * beta;
    const double sintheta = (alpha * gamma - beta * sqrt(al2be2 - gamma * gamma)) / al2be2;
    const double costheta = sqrt(1.0 - sintheta * sintheta);
    const double xa3d = -ya2d * sintheta, ya3d = ya2d * costheta, za3d = za1d;
    const double xb3d = xb2d * costheta - yb2d * sintheta, yb3d = xb2d * sintheta + yb2d * costheta, zb3d = zb1d;
    const double xc3d = -xb2d * costheta - yc2d * sintheta, yc3d = -xb2d * sintheta + yc2d * costheta, zc3d = zc1d;
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        a1[c] = xcom[c] + X[c] * xa3d + Y[c] * ya3d + Z[c] * za3d;
        b1[c] = xcom[c] + X[c] * xb3d + Y[c] * yb3d + Z[c] * zb3d;
        c1[c] = xcom[c] + X[c] * xc3d + Y[c] * yc3d + Z[c] * zc3d;
    }
}

// exact removal of the velocity components along the three edges of a rigid water
__device__ void settle_velocities(const double* x, double* v, const int* at, double mO, double mH)
{
    const double im[3] = {1.0 / mO, 1.0 / mH, 1.0 / mH};
    // edges: 0: a0->a1, 1: a1->a2, 2: a2->a0
    double u[3][3], b[3];
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        const int t = at[k], h = at[(k + 1) % 3];
        double n2 = 0.0;
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            u[k][c] = x[3 * h + c] - x[3 * t + c];
            n2 += u[k][c] * u[k][c];
        }
        const double inv = rsqrt(n2);
        b[k] = 0.0;
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            u[k][c] *= inv;
            b[k] += (v[3 * h + c] - v[3 * t + c]) * u[k][c];
        }
    }
    // A[k][l] = d(relative velocity of edge k along u_k) / d(impulse l); tails t_k = k, heads h_k = k+1
    const double u01 = dot3(u[0], u[1]), u02 = dot3(u[0], u[2]), u12 = dot3(u[1], u[2]);
    const double A00 = im[0] + im[1], A11 = im[1] + im[2], A22 = im[2] + im[0];
    const double A01 = -im[1] * u01, A12 = -im[2] * u12, A02 = -im[0] * u02;
    const double r0 = -b[0], r1 = -b[1], r2 = -b[2];
    const double det = A00 * (A11 * A22 - A12 * A12) - A01 * (A01 * A22 - A12 * A02) + A02 * (A01 * A12 - A11 * A02);
    const double inv_det = 1.0 / det;
    const double l0 = (r0 * (A11 * A22 - A12 * A12) - A01 * (r1 * A22 - A12 * r2) + A02 * (r1 * A12 - A11 * r2)) * inv_det;
    const double l1 = (A00 * (r1 * A22 - A12 * r2) - r0 * (A01 * A22 - A12 * A02) + A02 * (A01 * r2 - r1 * A02)) * inv_det;
    const double l2 = (A00 * (A11 * r2 - r1 * A12) - A01 * (A01 * r2 - r1 * A02) + r0 * (A01 * A12 - A11 * A02)) * inv_det;
    const double lam[3] = {l0, l1, l2};
#pragma unroll
    for (int l = 0; l < 3; ++l) {
        const int t = at[l], h = at[(l + 1) % 3];
        const double imt = im[l], imh = im[(l + 1) % 3];
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            v[3 * h + c] += lam[l] * u[l][c] * imh;
            v[3 * t + c] -= lam[l] * u[l][c] * imt;
        }
    }
}

// SHAKE on one cluster (Gauss-Seidel, displacements along the reference bond vectors)
__device__ void shake_cluster(const PKParams& P, const double* x0, double* x1, int c0, int c1)
{
    for (int it = 0; it < 500; ++it) {
        bool done = true;
        for (int k = c0; k < c1; ++k) {
            const int i = P.cl_atoms[2 * k], j = P.cl_atoms[2 * k + 1];
            const double d2 = P.cl_dist[k] * P.cl_dist[k];
            double r[3], s[3];
#pragma unroll
            for (int c = 0; c < 3; ++c) {
                r[c] = x1[3 * i + c] - x1[3 * j + c];
                s[c] = x0[3 * i + c] - x0[3 * j + c];
            }
            const double diff = dot3(r, r) - d2;
            if (fabs(diff) > 2.0 * P.tol * d2) done = false;
            const double imi = 1.0 / P.mass[i], imj = 1.0 / P.mass[j];
            const double gl = diff / (2.0 * dot3(r, s) * (imi + imj));
#pragma unroll
            for (int c = 0; c < 3; ++c) {
                x1[3 * i + c] -= gl * s[c] * imi;
                x1[3 * j + c] += gl * s[c] * imj;
            }
        }
        if (done) break;
    }
}

__device__ void rattle_cluster(const PKParams& P, const double* x, double* v, int c0, int c1)
{
    for (int it = 0; it < 500; ++it) {
        bool done = true;
        for (int k = c0; k < c1; ++k) {
            const int i = P.cl_atoms[2 * k], j = P.cl_atoms[2 * k + 1];
            const double d2 = P.cl_dist[k] * P.cl_dist[k];
            double r[3], w[3];
#pragma unroll
            for (int c = 0; c < 3; ++c) {
                r[c] = x[3 * i + c] - x[3 * j + c];
                w[c] = v[3 * i + c] - v[3 * j + c];
            }
            const double imi = 1.0 / P.mass[i], imj = 1.0 / P.mass[j];
            const double gl = dot3(r, w) / (d2 * (imi + imj));
            if (fabs(gl) > P.tol) done = false;
#pragma unroll
            for (int c = 0; c < 3; ++c) {
                v[3 * i + c] -= gl * r[c] * imi;
                v[3 * j + c] += gl * r[c] * imj;
            }
        }
        if (done) break;
    }
}

// positions: x holds the unconstrained update, xref the previous (constrained) positions
__device__ void constrain_positions(const PKParams& P, double* x, const double* xref)
{
    for (int w = threadIdx.x; w < P.n_settle; w += blockDim.x) {
        const int* at = P.settle + 3 * w;
        settle_positions(xref, x, at, P.mass[at[0]], P.mass[at[1]], P.oh, P.hh);
    }
    for (int c = threadIdx.x; c < P.n_clusters; c += blockDim.x) shake_cluster(P, xref, x, P.cl_off[c], P.cl_off[c + 1]);
}

__device__ void constrain_velocities(const PKParams& P, const double* x, double* v)
{
    for (int w = threadIdx.x; w < P.n_settle; w += blockDim.x) {
        const int* at = P.settle + 3 * w;
        settle_velocities(x, v, at, P.mass[at[0]], P.mass[at[1]]);
    }
    for (int c = threadIdx.x; c < P.n_clusters; c += blockDim.x) rattle_cluster(P, x, v, P.cl_off[c], P.cl_off[c + 1]);
}

// three normals for atom `unit`
template <bool NOISE32>
__device__ __forceinline__ void atom_normals(uint64_t seed, uint64_t rid, uint32_t draw, uint32_t tag, uint32_t unit,
                                             double& n0, double& n1, double& n2)
{
    const lsi::Philox4 r = lsi::philox_block(seed, rid, draw, tag, unit);
    if (NOISE32) {
        float a, b, c, d;
        lsi::box_muller(r.x, r.y, a, b);
        lsi::box_muller(r.z, r.w, c, d);
        n0 = (double)a; n1 = (double)b; n2 = (double)c;
    } else {
        double d;
        lsi::box_muller(r.x, r.y, n0, n1);
        lsi::box_muller(r.z, r.w, n2, d);
    }
}

template <typename real, bool PERIODIC, bool DIRECT, bool TRACE>
__global__ void __launch_bounds__(512) particles_protocol_kernel(const __grid_constant__ PKParams P)
{
    extern __shared__ __align__(32) unsigned char smem_raw[];
    const Smem<real> s = carve<real>(smem_raw, P.N, P.n_slots);
    constexpr bool NOISE32 = sizeof(real) == 4;
    const int N = P.N, D = 3 * N;
    const int64_t r = blockIdx.x;
    const uint64_t rid = P.replica0 + (uint64_t)r;
    const int64_t T = (int64_t)P.n_steps + 1;

    // ---- start state
    {
        const double* xs = P.x0 ? P.x0 + r * D : P.x_cache + lsi::cache_index(P.seed, rid, P.n_cache) * D;
        for (int k = threadIdx.x; k < D; k += blockDim.x) s.x[k] = xs[k];
        if (P.v0) {
            for (int k = threadIdx.x; k < D; k += blockDim.x) s.v[k] = P.v0[r * D + k];
        } else {
            for (int i = threadIdx.x; i < N; i += blockDim.x) {
                double n0, n1, n2;
                atom_normals<false>(P.seed, rid, 0u, LSI_TAG_V0, (uint32_t)i, n0, n1, n2);
                const double sg = sqrt(P.kT / P.mass[i]);
                s.v[3 * i] = sg * n0; s.v[3 * i + 1] = sg * n1; s.v[3 * i + 2] = sg * n2;
            }
        }
        load_lj<real>(P, s);
    }
    __syncthreads();

    double heat = 0.0;
    for (int leg = 0; leg < P.n_legs; ++leg) {
        if (leg > 0) {
            if (P.flags & LSI_FLAG_ROUND_MIDPOINT_X_FP32)
                for (int k = threadIdx.x; k < D; k += blockDim.x) s.x[k] = (double)(float)s.x[k];
            if (P.marginal == LSI_MARGINAL_CONFIGURATION)
                for (int i = threadIdx.x; i < N; i += blockDim.x) {
                    double n0, n1, n2;
                    atom_normals<false>(P.seed, rid, (uint32_t)(leg - 1), LSI_TAG_VMID, (uint32_t)i, n0, n1, n2);
                    const double sg = sqrt(P.kT / P.mass[i]);
                    s.v[3 * i] = sg * n0; s.v[3 * i + 1] = sg * n1; s.v[3 * i + 2] = sg * n2;
                }
            __syncthreads();
        }
        // Context.applyConstraints / applyVelocityConstraints (bookkeepers.py:206-208)
        if (P.has_constraints) {
            for (int k = threadIdx.x; k < D; k += blockDim.x) s.xref[k] = s.x[k];
            __syncthreads();
            constrain_positions(P, s.x, s.xref);
            __syncthreads();
            constrain_velocities(P, s.x, s.v);
            __syncthreads();
        }
        stage_positions<real>(P, s);
        __syncthreads();
        unsigned valid = 0u;  // bit per force-cache slot (block-uniform)
        double pe_cur = compute_forces<real, PERIODIC, true>(P, s, -1, s.f);  // slot 0 used as scratch
        __syncthreads();
        const double E0 = pe_cur + kinetic_energy(P, s.v, s.red);
        const double Q0 = heat;
        double wd = 0.0;
        uint32_t q = 0;
        if (TRACE) {
            const int64_t base = ((int64_t)leg * T) * P.n + r;
            if (P.trace_x) for (int k = threadIdx.x; k < D; k += blockDim.x) P.trace_x[base * D + k] = s.x[k];
            if (P.trace_v) for (int k = threadIdx.x; k < D; k += blockDim.x) P.trace_v[base * D + k] = s.v[k];
            if (P.trace_w && threadIdx.x == 0) P.trace_w[base] = 0.0;
        }

        for (int st = 0; st < P.n_steps; ++st) {
            for (int op = 0; op < P.n_ops; ++op) {
                const int kind = P.op_kind[op];
                const double c0 = P.op_c0[op];
                if (kind == LSI_OP_R) {
                    double e_old = 0.0;
                    if (DIRECT) e_old = pe_cur + kinetic_energy(P, s.v, s.red);
                    const double h = c0;
                    for (int k = threadIdx.x; k < D; k += blockDim.x) {
                        s.xref[k] = s.x[k];
                        s.x[k] = s.x[k] + h * s.v[k];
                    }
                    if (P.has_constraints) {
                        __syncthreads();
                        constrain_positions(P, s.x, s.xref);
                        __syncthreads();
                        // v += (x - x1)/h with x1 = the unconstrained update (langevin.py:131-133)
                        for (int k = threadIdx.x; k < D; k += blockDim.x)
                            if (P.constrained[k / 3]) {
                                const double x1 = s.xref[k] + h * s.v[k];
                                s.v[k] = s.v[k] + (s.x[k] - x1) / h;
                            }
                        __syncthreads();
                        constrain_velocities(P, s.x, s.v);
                    }
                    __syncthreads();
                    stage_positions<real>(P, s);
                    valid = 0u;
                    __syncthreads();
                    if (DIRECT) {
                        pe_cur = compute_forces<real, PERIODIC, true>(P, s, -1, s.f);
                        __syncthreads();
                        wd += (pe_cur + kinetic_energy(P, s.v, s.red)) - e_old;
                    }
                } else if (kind == LSI_OP_V) {
                    const int slot = P.op_slot[op];
                    real* fs = s.f + (size_t)slot * D;
                    double ke_old = 0.0;
                    if (DIRECT) ke_old = kinetic_energy(P, s.v, s.red);
                    if (!((valid >> slot) & 1u)) {
                        compute_forces<real, PERIODIC, false>(P, s, P.slot_group[slot], fs);
                        valid |= 1u << slot;
                    }
                    // each thread kicks the atoms whose forces it just wrote
                    for (int i = threadIdx.x; i < N; i += blockDim.x) {
                        const double hm = c0 / P.mass[i];
                        s.v[3 * i] += hm * (double)fs[3 * i];
                        s.v[3 * i + 1] += hm * (double)fs[3 * i + 1];
                        s.v[3 * i + 2] += hm * (double)fs[3 * i + 2];
                    }
                    if (P.has_constraints) {
                        __syncthreads();
                        constrain_velocities(P, s.x, s.v);
                    }
                    __syncthreads();
                    if (DIRECT) wd += kinetic_energy(P, s.v, s.red) - ke_old;
                } else {
                    const double a = c0, b = P.op_c1[op];
                    const double ke_old = kinetic_energy(P, s.v, s.red);
                    const uint32_t draw = ((uint32_t)leg << 26) | q;
                    ++q;
                    for (int i = threadIdx.x; i < N; i += blockDim.x) {
                        double n0, n1, n2;
                        atom_normals<NOISE32>(P.seed, rid, draw, LSI_TAG_O, (uint32_t)i, n0, n1, n2);
                        const double sg = b * sqrt(P.kT / P.mass[i]);
                        s.v[3 * i] = a * s.v[3 * i] + sg * n0;
                        s.v[3 * i + 1] = a * s.v[3 * i + 1] + sg * n1;
                        s.v[3 * i + 2] = a * s.v[3 * i + 2] + sg * n2;
                    }
                    if (P.has_constraints) {
                        __syncthreads();
                        constrain_velocities(P, s.x, s.v);
                        __syncthreads();  // projections are written per water / cluster, read per atom
                    }
                    heat += kinetic_energy(P, s.v, s.red) - ke_old;
                }
            }
            if (TRACE) {
                __syncthreads();
                const int64_t idx = ((int64_t)leg * T + st + 1) * P.n + r;
                if (P.trace_x) for (int k = threadIdx.x; k < D; k += blockDim.x) P.trace_x[idx * D + k] = s.x[k];
                if (P.trace_v) for (int k = threadIdx.x; k < D; k += blockDim.x) P.trace_v[idx * D + k] = s.v[k];
                if (P.trace_w) {
                    // needs a scratch force slot: the caches stay valid only if slot 0 is not clobbered
                    real* scratch = s.f + (size_t)(P.n_slots - 1) * D;
                    const double pe = compute_forces<real, PERIODIC, true>(P, s, -1, scratch);
                    valid &= ~(1u << (P.n_slots - 1));
                    __syncthreads();
                    const double E = pe + kinetic_energy(P, s.v, s.red);
                    if (threadIdx.x == 0) P.trace_w[idx] = ((E - E0) - (heat - Q0)) * P.inv_kT;
                }
            }
        }
        __syncthreads();
        {
            real* scratch = s.f + (size_t)(P.n_slots - 1) * D;
            const double pe1 = DIRECT ? pe_cur : compute_forces<real, PERIODIC, true>(P, s, -1, scratch);
            __syncthreads();
            const double E1 = pe1 + kinetic_energy(P, s.v, s.red);
            if (threadIdx.x == 0) {
                P.W[(int64_t)leg * P.n + r] = ((E1 - E0) - (heat - Q0)) * P.inv_kT;
                if (P.heat) P.heat[(int64_t)leg * P.n + r] = heat - Q0;
                if (DIRECT && P.W_direct) P.W_direct[(int64_t)leg * P.n + r] = wd * P.inv_kT;
            }
        }
        __syncthreads();
    }
    if (P.x_final) for (int k = threadIdx.x; k < D; k += blockDim.x) P.x_final[r * D + k] = s.x[k];
    if (P.v_final) for (int k = threadIdx.x; k < D; k += blockDim.x) P.v_final[r * D + k] = s.v[k];
}

template <typename real, bool PERIODIC>
__global__ void __launch_bounds__(512) particles_energy_kernel(const __grid_constant__ PKParams P, int group, const double* __restrict__ x,
                                        double* __restrict__ energy, double* __restrict__ forces)
{
    extern __shared__ __align__(32) unsigned char smem_raw[];
    const Smem<real> s = carve<real>(smem_raw, P.N, 1);
    const int D = 3 * P.N;
    const int64_t r = blockIdx.x;
    for (int k = threadIdx.x; k < D; k += blockDim.x) s.x[k] = x[r * D + k];
    load_lj<real>(P, s);
    __syncthreads();
    stage_positions<real>(P, s);
    __syncthreads();
    const double e = compute_forces<real, PERIODIC, true>(P, s, group, s.f);
    __syncthreads();
    if (energy && threadIdx.x == 0) energy[r] = e;
    if (forces) for (int k = threadIdx.x; k < D; k += blockDim.x) forces[r * D + k] = (double)s.f[k];
}

// ------------------------------------------------------------------------------------------
template <typename T>
cudaError_t upload(T** dst, const std::vector<T>& src)
{
    *dst = nullptr;
    if (src.empty()) return cudaSuccess;
    cudaError_t e = cudaMalloc((void**)dst, src.size() * sizeof(T));
    if (e != cudaSuccess) return e;
    return cudaMemcpy(*dst, src.data(), src.size() * sizeof(T), cudaMemcpyHostToDevice);
}

int get_tables(lsi_particles* p, DevTables** out)
{
    int dev = 0;
    LSI_CUDA_CHECK(cudaGetDevice(&dev));
    std::lock_guard<std::mutex> lock(p->mu);
    auto it = p->dev.find(dev);
    if (it != p->dev.end()) { *out = &it->second; return LSI_OK; }
    DevTables t;
    const int N = p->n_atoms;
    std::vector<float4> nf(N);
    std::vector<double4> nd(N);
    const double sk = std::sqrt(LSI_COULOMB);
    for (int i = 0; i < N; ++i) {
        nd[i] = make_double4(p->charge[i] * sk, 0.5 * p->sigma[i], 2.0 * std::sqrt(p->epsilon[i]), 0.0);
        nf[i] = make_float4((float)nd[i].x, (float)nd[i].y, (float)nd[i].z, 0.f);
    }
    float4* df = nullptr;
    double4* dd = nullptr;
    LSI_CUDA_CHECK(upload(&t.mass, p->mass));
    LSI_CUDA_CHECK(upload(&df, nf));
    LSI_CUDA_CHECK(upload(&dd, nd));
    t.nbp_f = df; t.nbp_d = dd;
    LSI_CUDA_CHECK(upload(&t.excl, p->excl));
    LSI_CUDA_CHECK(upload(&t.terms, p->terms));
    LSI_CUDA_CHECK(upload(&t.term_off, p->term_off));
    LSI_CUDA_CHECK(upload(&t.settle, p->settle));
    LSI_CUDA_CHECK(upload(&t.cl_off, p->cl_off));
    LSI_CUDA_CHECK(upload(&t.cl_atoms, p->cl_atoms));
    LSI_CUDA_CHECK(upload(&t.cl_dist, p->cl_dist));
    LSI_CUDA_CHECK(upload(&t.constrained, p->constrained));
    p->dev[dev] = t;
    *out = &p->dev[dev];
    return LSI_OK;
}

void fill_system_params(const lsi_particles* p, const DevTables* t, bool mixed, PKParams& P)
{
    P.N = p->n_atoms;
    P.nb_method = p->nb_method;
    P.excl_words = p->excl_words;
    P.n_settle = (int)p->settle.size() / 3;
    P.n_clusters = (int)p->cl_off.size() - 1;
    if (P.n_clusters < 0) P.n_clusters = 0;
    P.has_constraints = (P.n_settle + P.n_clusters) > 0;
    for (int c = 0; c < 3; ++c) {
        P.box[c] = p->box[c];
        P.inv_box[c] = p->box[c] > 0 ? 1.0 / p->box[c] : 0.0;
    }
    const double rc = p->cutoff, er = p->rf_dielectric;
    P.cutoff2 = rc * rc;
    P.krf = p->nb_method == LSI_NB_CUTOFF_PERIODIC ? (1.0 / (rc * rc * rc)) * (er - 1.0) / (2.0 * er + 1.0) : 0.0;
    P.crf = p->nb_method == LSI_NB_CUTOFF_PERIODIC ? (1.0 / rc) * (3.0 * er) / (2.0 * er + 1.0) : 0.0;
    P.restraint_k = p->restraint_k;
    P.g_nb = p->g_nb; P.g_bond = p->g_bond; P.g_angle = p->g_angle; P.g_tors = p->g_tors; P.g_restr = p->g_restr;
    P.oh = p->oh; P.hh = p->hh; P.tol = p->tol;
    P.mass = t->mass;
    P.nbp = mixed ? t->nbp_f : t->nbp_d;
    P.excl = t->excl; P.terms = t->terms; P.term_off = t->term_off;
    P.settle = t->settle; P.cl_off = t->cl_off; P.cl_atoms = t->cl_atoms; P.cl_dist = t->cl_dist;
    P.constrained = t->constrained;
}

int pick_threads(int N) { return std::min(512, std::max(32, ((N + 31) / 32) * 32)); }

template <typename K>
int set_smem(K kernel, size_t bytes)
{
    if (bytes > 48 * 1024) LSI_CUDA_CHECK(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
    return LSI_OK;
}

template <typename real, bool PERIODIC>
int launch_protocol(const PKParams& P, bool direct, bool trace, size_t bytes, int threads, cudaStream_t st)
{
    const unsigned grid = (unsigned)P.n;
#define LSI_LAUNCH(D_, T_)                                                                                  \
    {                                                                                                       \
        auto k = particles_protocol_kernel<real, PERIODIC, D_, T_>;                                         \
        int rc = set_smem(k, bytes);                                                                        \
        if (rc != LSI_OK) return rc;                                                                        \
        k<<<grid, threads, bytes, st>>>(P);                                                                 \
    }
    if (direct && trace) LSI_LAUNCH(true, true)
    else if (direct) LSI_LAUNCH(true, false)
    else if (trace) LSI_LAUNCH(false, true)
    else LSI_LAUNCH(false, false)
#undef LSI_LAUNCH
    LSI_CUDA_CHECK(cudaGetLastError());
    return LSI_OK;
}

}  // namespace

void lsi_particles_free(lsi_particles* p)
{
    if (!p) return;
    for (auto& kv : p->dev) {
        DevTables& t = kv.second;
        cudaFree(t.mass); cudaFree(t.nbp_f); cudaFree(t.nbp_d); cudaFree(t.excl); cudaFree(t.terms);
        cudaFree(t.term_off); cudaFree(t.settle); cudaFree(t.cl_off); cudaFree(t.cl_atoms); cudaFree(t.cl_dist);
        cudaFree(t.constrained);
    }
    delete p;
}

LSI_EXPORT int lsi_system_num_dof(const lsi_system* s)
{
    if (!s) return LSI_ERR_ARG;
    return s->kind == LSI_SYS_SCALAR ? 1 : 3 * s->particles->n_atoms;
}

LSI_EXPORT int lsi_system_create_particles(const lsi_particles_desc* d, lsi_system** out)
{
    if (!d || !out) return lsi_set_error(LSI_ERR_ARG, "lsi_system_create_particles: NULL argument");
    const int N = d->n_atoms;
    if (N <= 0 || N > 4096) return lsi_set_error(LSI_ERR_ARG, "n_atoms must be in 1..4096 (one CTA holds a replica)");
    if (!d->mass) return lsi_set_error(LSI_ERR_ARG, "mass is required");
    auto bad_atom = [N](int a) { return a < 0 || a >= N; };
    lsi_particles* p = new lsi_particles();
    p->n_atoms = N;
    p->mass.assign(d->mass, d->mass + N);
    for (double m : p->mass)
        if (!(m > 0.0)) { delete p; return lsi_set_error(LSI_ERR_ARG, "masses must be positive (massless sites are not supported)"); }
    p->charge.assign(N, 0.0); p->sigma.assign(N, 1.0); p->epsilon.assign(N, 0.0);
    if (d->charge) p->charge.assign(d->charge, d->charge + N);
    if (d->sigma) p->sigma.assign(d->sigma, d->sigma + N);
    if (d->epsilon) p->epsilon.assign(d->epsilon, d->epsilon + N);
    p->nb_method = d->nonbonded_method;
    if (p->nb_method != LSI_NB_NOCUTOFF && p->nb_method != LSI_NB_CUTOFF_PERIODIC) {
        delete p;
        return lsi_set_error(LSI_ERR_UNSUPPORTED, "nonbonded_method %d (only NoCutoff and CutoffPeriodic are built)", d->nonbonded_method);
    }
    p->cutoff = d->cutoff;
    for (int c = 0; c < 3; ++c) p->box[c] = d->box[c];
    if (p->nb_method == LSI_NB_CUTOFF_PERIODIC) {
        for (int c = 0; c < 3; ++c)
            if (!(p->box[c] >= 2.0 * p->cutoff) || !(p->cutoff > 0)) {
                delete p;
                return lsi_set_error(LSI_ERR_ARG, "periodic box edge must be at least twice the cutoff");
            }
    }
    p->rf_dielectric = d->rf_dielectric > 0 ? d->rf_dielectric : 78.3;
    p->restraint_k = d->restraint_k;
    p->g_nb = d->group_nonbonded; p->g_bond = d->group_bonds; p->g_angle = d->group_angles;
    p->g_tors = d->group_torsions; p->g_restr = d->group_restraint;
    p->oh = d->settle_oh; p->hh = d->settle_hh;
    p->tol = d->constraint_tolerance > 0 ? d->constraint_tolerance : 1e-8;

    // exclusion bitmasks (self, exclusions, exceptions)
    p->excl_words = (N + 31) / 32;
    p->excl.assign((size_t)N * p->excl_words, 0u);
    auto set_excl = [&](int i, int j) {
        p->excl[(size_t)i * p->excl_words + (j >> 5)] |= 1u << (j & 31);
        p->excl[(size_t)j * p->excl_words + (i >> 5)] |= 1u << (i & 31);
    };
    for (int i = 0; i < N; ++i) set_excl(i, i);
    for (int k = 0; k < d->n_exclusions; ++k) {
        const int i = d->exclusions[2 * k], j = d->exclusions[2 * k + 1];
        if (bad_atom(i) || bad_atom(j)) { delete p; return lsi_set_error(LSI_ERR_ARG, "exclusion %d: atom index out of range", k); }
        set_excl(i, j);
    }
    // per-atom gather lists of bonded terms and exceptions
    std::vector<std::vector<TermInst>> per_atom(N);
    auto add_term = [&](int kind, int n_at, const int32_t* at, const double* prm, int n_prm) -> bool {
        for (int r = 0; r < n_at; ++r)
            if (bad_atom(at[r])) return false;
        for (int r = 0; r < n_at; ++r) {
            TermInst t;
            t.kind = kind; t.role = r;
            for (int q = 0; q < 4; ++q) t.a[q] = q < n_at ? at[q] : 0;
            for (int q = 0; q < 3; ++q) t.p[q] = q < n_prm ? prm[q] : 0.0;
            per_atom[at[r]].push_back(t);
        }
        return true;
    };
    bool ok = true;
    for (int k = 0; k < d->n_exceptions && ok; ++k) {
        ok = add_term(TERM_EXCEPTION, 2, d->exception_atoms + 2 * k, d->exception_params + 3 * k, 3);
        if (ok) set_excl(d->exception_atoms[2 * k], d->exception_atoms[2 * k + 1]);
    }
    for (int k = 0; k < d->n_bonds && ok; ++k) ok = add_term(TERM_BOND, 2, d->bond_atoms + 2 * k, d->bond_params + 2 * k, 2);
    for (int k = 0; k < d->n_angles && ok; ++k) ok = add_term(TERM_ANGLE, 3, d->angle_atoms + 3 * k, d->angle_params + 2 * k, 2);
    for (int k = 0; k < d->n_torsions && ok; ++k) ok = add_term(TERM_TORSION, 4, d->torsion_atoms + 4 * k, d->torsion_params + 3 * k, 3);
    if (!ok) { delete p; return lsi_set_error(LSI_ERR_ARG, "bonded term / exception: atom index out of range"); }
    p->term_off.assign(N + 1, 0);
    for (int i = 0; i < N; ++i) {
        p->term_off[i + 1] = p->term_off[i] + (int)per_atom[i].size();
        p->terms.insert(p->terms.end(), per_atom[i].begin(), per_atom[i].end());
    }
    if (p->terms.empty()) p->terms.push_back(TermInst{});  // keep the device pointer valid

    // constraints
    p->constrained.assign(N, 0);
    for (int k = 0; k < d->n_settle; ++k)
        for (int q = 0; q < 3; ++q) {
            const int a = d->settle_atoms[3 * k + q];
            if (bad_atom(a)) { delete p; return lsi_set_error(LSI_ERR_ARG, "settle %d: atom index out of range", k); }
            p->settle.push_back(a);
            p->constrained[a] = 1;
        }
    if (d->n_settle > 0 && !(p->oh > 0 && p->hh > 0 && p->hh < 2 * p->oh)) {
        delete p;
        return lsi_set_error(LSI_ERR_ARG, "settle distances must satisfy 0 < hh < 2 oh");
    }
    // clusters = connected components of the pair-constraint graph (union-find), constraints kept in input order
    {
        std::vector<int> parent(N);
        for (int i = 0; i < N; ++i) parent[i] = i;
        auto find = [&](int a) { while (parent[a] != a) { parent[a] = parent[parent[a]]; a = parent[a]; } return a; };
        for (int k = 0; k < d->n_constraints; ++k) {
            const int i = d->constraint_atoms[2 * k], j = d->constraint_atoms[2 * k + 1];
            if (bad_atom(i) || bad_atom(j) || !(d->constraint_dist[k] > 0)) {
                delete p;
                return lsi_set_error(LSI_ERR_ARG, "constraint %d: bad atoms or distance", k);
            }
            parent[find(i)] = find(j);
            p->constrained[i] = p->constrained[j] = 1;
        }
        std::map<int, std::vector<int>> comps;
        for (int k = 0; k < d->n_constraints; ++k) comps[find(d->constraint_atoms[2 * k])].push_back(k);
        std::vector<std::pair<int, std::vector<int>>> ordered(comps.begin(), comps.end());
        std::sort(ordered.begin(), ordered.end(), [](const auto& a, const auto& b) { return a.second[0] < b.second[0]; });
        p->cl_off.push_back(0);
        for (auto& kv : ordered) {
            for (int k : kv.second) {
                p->cl_atoms.push_back(d->constraint_atoms[2 * k]);
                p->cl_atoms.push_back(d->constraint_atoms[2 * k + 1]);
                p->cl_dist.push_back(d->constraint_dist[k]);
            }
            p->cl_off.push_back((int)p->cl_dist.size());
        }
    }
    lsi_system* s = new lsi_system();
    s->kind = LSI_SYS_PARTICLES;
    s->potential = -1;
    s->mass = 0.0;
    for (int i = 0; i < 4; ++i) s->params[i] = 0.0;
    s->particles = p;
    *out = s;
    return LSI_OK;
}

int lsi_particles_run_protocol(const lsi_system* sys, const lsi_program* prog, const lsi_protocol_args* a, void* stream)
{
    cudaStream_t st = (cudaStream_t)stream;
    lsi_particles* p = sys->particles;
    if (a->n_replicas == 0) return LSI_OK;
    if (!a->W) return lsi_set_error(LSI_ERR_ARG, "lsi_run_protocol: W output is required");
    if (!a->x0 && !(a->x_cache && a->n_cache > 0)) return lsi_set_error(LSI_ERR_ARG, "lsi_run_protocol: need x0 or a non-empty x_cache");
    if (a->precision != LSI_PRECISION_DOUBLE && a->precision != LSI_PRECISION_MIXED)
        return lsi_set_error(LSI_ERR_ARG, "unknown precision %d", a->precision);
    if (a->W_direct && !prog->measure_shadow_work) return lsi_set_error(LSI_ERR_ARG, "W_direct needs a program compiled with measure_shadow_work");
    if (a->n_legs > 64) return lsi_set_error(LSI_ERR_ARG, "at most 64 legs");
    if ((int64_t)a->n_steps * (int64_t)(prog->n_O > 0 ? prog->n_O : 1) >= ((int64_t)1 << 26))
        return lsi_set_error(LSI_ERR_ARG, "n_steps * n_O exceeds the 2^26 draws a leg's noise stream holds");
    if (a->n_replicas > 0x7fffffff) return lsi_set_error(LSI_ERR_ARG, "at most 2^31 - 1 replicas per launch");
    if (a->moments) return lsi_set_error(LSI_ERR_ARG, "particle systems: reduce the returned works with lsi_reduce_moments");

    DevTables* t = nullptr;
    int rc = get_tables(p, &t);
    if (rc != LSI_OK) return rc;
    const bool mixed = a->precision == LSI_PRECISION_MIXED;
    PKParams P;
    std::memset(&P, 0, sizeof(P));
    fill_system_params(p, t, mixed, P);

    // program: force-cache slots per distinct group
    const int n_ops = (int)prog->ops.size();
    std::vector<unsigned char> kinds(n_ops), slots(n_ops, 0);
    std::vector<double> c0(n_ops), c1(n_ops, 0.0);
    int n_slots = 0;
    int slot_group[4] = {-1, -1, -1, -1};
    for (int i = 0; i < n_ops; ++i) {
        const lsi_op& op = prog->ops[i];
        kinds[i] = (unsigned char)op.kind;
        if (op.kind == LSI_OP_O) { c0[i] = op.a; c1[i] = op.b; }
        else c0[i] = prog->dt * op.fraction;
        if (op.kind == LSI_OP_V) {
            int sidx = -1;
            for (int q = 0; q < n_slots; ++q)
                if (slot_group[q] == op.group) sidx = q;
            if (sidx < 0) {
                if (n_slots == 4) return lsi_set_error(LSI_ERR_UNSUPPORTED, "at most 4 distinct force groups per splitting");
                sidx = n_slots;
                slot_group[n_slots++] = op.group;
            }
            slots[i] = (unsigned char)sidx;
        }
    }
    if (n_slots == 0) n_slots = 1;
    const bool trace = a->trace_x || a->trace_v || a->trace_w;
    if (trace && a->trace_w && n_slots < 4) ++n_slots;  // scratch slot so per-step energies keep the caches
    P.n_ops = n_ops; P.n_O = prog->n_O; P.n_slots = n_slots;
    for (int q = 0; q < 4; ++q) P.slot_group[q] = slot_group[q];

    unsigned char* dprog = nullptr;
    const size_t nb = (size_t)n_ops * (2 * sizeof(double)) + 2 * (size_t)n_ops;
    LSI_CUDA_CHECK(cudaMallocAsync((void**)&dprog, nb, st));
    double* d0 = (double*)dprog;
    double* d1 = d0 + n_ops;
    unsigned char* dk = (unsigned char*)(d1 + n_ops);
    unsigned char* ds = dk + n_ops;
    LSI_CUDA_CHECK(cudaMemcpyAsync(d0, c0.data(), n_ops * sizeof(double), cudaMemcpyHostToDevice, st));
    LSI_CUDA_CHECK(cudaMemcpyAsync(d1, c1.data(), n_ops * sizeof(double), cudaMemcpyHostToDevice, st));
    LSI_CUDA_CHECK(cudaMemcpyAsync(dk, kinds.data(), n_ops, cudaMemcpyHostToDevice, st));
    LSI_CUDA_CHECK(cudaMemcpyAsync(ds, slots.data(), n_ops, cudaMemcpyHostToDevice, st));
    LSI_CUDA_CHECK(cudaStreamSynchronize(st));  // host vectors die at return
    P.op_c0 = d0; P.op_c1 = d1; P.op_kind = dk; P.op_slot = ds;

    P.seed = a->seed; P.replica0 = a->replica0; P.n = a->n_replicas;
    P.n_steps = a->n_steps; P.n_legs = a->n_legs; P.marginal = a->marginal; P.flags = a->flags;
    P.kT = a->kT; P.inv_kT = 1.0 / a->kT;
    P.x_cache = a->x_cache; P.n_cache = a->n_cache; P.x0 = a->x0; P.v0 = a->v0;
    P.W = a->W; P.heat = a->heat; P.W_direct = a->W_direct; P.x_final = a->x_final; P.v_final = a->v_final;
    P.trace_x = a->trace_x; P.trace_v = a->trace_v; P.trace_w = a->trace_w;

    const bool direct = a->W_direct != nullptr;
    const bool periodic = p->nb_method == LSI_NB_CUTOFF_PERIODIC;
    const int threads = pick_threads(p->n_atoms);
    const size_t bytes = mixed ? smem_bytes<float>(p->n_atoms, n_slots) : smem_bytes<double>(p->n_atoms, n_slots);
    if (bytes > 227 * 1024) {
        cudaFreeAsync(dprog, st);
        return lsi_set_error(LSI_ERR_UNSUPPORTED, "system too large for the shared-memory-resident kernel (%zu bytes)", bytes);
    }
    if (mixed) rc = periodic ? launch_protocol<float, true>(P, direct, trace, bytes, threads, st)
                             : launch_protocol<float, false>(P, direct, trace, bytes, threads, st);
    else rc = periodic ? launch_protocol<double, true>(P, direct, trace, bytes, threads, st)
                       : launch_protocol<double, false>(P, direct, trace, bytes, threads, st);
    cudaFreeAsync(dprog, st);
    return rc;
}

LSI_EXPORT int lsi_particles_energy_forces(const lsi_system* sys, int32_t precision, int32_t group, int64_t n,
                                           const double* x, double* energy, double* forces, void* stream)
{
    if (!sys || sys->kind != LSI_SYS_PARTICLES || !x || n < 0) return lsi_set_error(LSI_ERR_ARG, "lsi_particles_energy_forces: bad argument");
    if (n == 0) return LSI_OK;
    if (lsi_device_count() <= 0) return lsi_set_error(LSI_ERR_CUDA, "no CUDA device: liblsi has no CPU fallback");
    lsi_particles* p = sys->particles;
    DevTables* t = nullptr;
    int rc = get_tables(p, &t);
    if (rc != LSI_OK) return rc;
    const bool mixed = precision == LSI_PRECISION_MIXED;
    PKParams P;
    std::memset(&P, 0, sizeof(P));
    fill_system_params(p, t, mixed, P);
    P.n_slots = 1;
    const bool periodic = p->nb_method == LSI_NB_CUTOFF_PERIODIC;
    const int threads = pick_threads(p->n_atoms);
    const size_t bytes = mixed ? smem_bytes<float>(p->n_atoms, 1) : smem_bytes<double>(p->n_atoms, 1);
    cudaStream_t st = (cudaStream_t)stream;
#define LSI_LAUNCH_E(R_, P_)                                                                    \
    {                                                                                           \
        auto k = particles_energy_kernel<R_, P_>;                                               \
        rc = set_smem(k, bytes);                                                                \
        if (rc != LSI_OK) return rc;                                                            \
        k<<<(unsigned)n, threads, bytes, st>>>(P, group, x, energy, forces);                    \
    }
    if (mixed) { if (periodic) LSI_LAUNCH_E(float, true) else LSI_LAUNCH_E(float, false) }
    else { if (periodic) LSI_LAUNCH_E(double, true) else LSI_LAUNCH_E(double, false) }
#undef LSI_LAUNCH_E
    LSI_CUDA_CHECK(cudaGetLastError());
    return LSI_OK;
}

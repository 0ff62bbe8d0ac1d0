/*
 * lsi_oracle.c -- CPU restatement (plain C, fp64) of the reference's
 * Langevin-splitting / shadow-work hot path.
 *
 * THIS IS TEST INFRASTRUCTURE.  Only tests/, __graft_entry__.smoke() and
 * bench.py's cpu_baseline / --impl reference legs may load it.  The product
 * (integrator-benchmark_b200/) never links, imports or calls anything here.
 *
 * What it restates (all paths relative to /root/reference):
 *   - substep semantics R / V / V{g} / O, heat and shadow-work bookkeeping:
 *       benchmark/integrators/langevin.py:124-175 (R_step, V_step, O_step),
 *       constants a, b: langevin.py:189-196
 *   - toy closed-form loops (n_steps-1 iterations, cumulative W_shads):
 *       benchmark/integrators/numba_integrators.py:17-63,71-114,126-171,180-225
 *   - toy potentials: benchmark/testsystems/low_dimensional_systems.py:52-76,132-136
 *   - protocol driver (F leg, midpoint, R leg):
 *       benchmark/testsystems/bookkeepers.py:184-295,
 *       benchmark/testsystems/low_dimensional_systems.py:160-184
 *
 * Parity pinning: tests/golden/ holds outputs of the reference's OWN
 * numba_integrators.py closures (run through .py_func with replayed noise) and
 * of the reference's own LangevinSplittingIntegrator step programs (recorded
 * by a stand-in CustomIntegrator and interpreted with numpy); see
 * tests/golden/make_golden.py.  The oracle is checked against those in
 * tests/test_oracle_golden.py.  Force-field values and constraint solvers for
 * the molecular systems live in OpenMM/openmmtools, which are absent here:
 * for those, parity is UNPINNED.
 *
 * RNG: Philox-4x32-10 (Salmon et al., SC'11; same constants as Random123 and
 * cuRAND's curand_philox4x32_x.h).  Counter layout is the repo's canonical one
 * (DESIGN.md "RNG layout"):
 *     key     = (seed_lo, seed_hi)
 *     counter = (replica_lo, replica_hi, draw, (tag << 28) | unit)
 * Box-Muller in fp64 on (r + 0.5) * 2^-32 uniforms.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define ORC_API __attribute__((visibility("default")))

/* ------------------------------------------------------------------ Philox */

static inline void philox_round(uint32_t c[4], const uint32_t k[2])
{
    const uint64_t p0 = (uint64_t)0xD2511F53u * c[0];
    const uint64_t p1 = (uint64_t)0xCD9E8D57u * c[2];
    const uint32_t hi0 = (uint32_t)(p0 >> 32), lo0 = (uint32_t)p0;
    const uint32_t hi1 = (uint32_t)(p1 >> 32), lo1 = (uint32_t)p1;
    const uint32_t n0 = hi1 ^ c[1] ^ k[0];
    const uint32_t n2 = hi0 ^ c[3] ^ k[1];
    c[0] = n0; c[1] = lo1; c[2] = n2; c[3] = lo0;
}

ORC_API void orc_philox4x32_10(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4])
{
    uint32_t c[4] = {ctr[0], ctr[1], ctr[2], ctr[3]};
    uint32_t k[2] = {key[0], key[1]};
    for (int r = 0; r < 10; ++r) {
        philox_round(c, k);
        k[0] += 0x9E3779B9u;
        k[1] += 0xBB67AE85u;
    }
    memcpy(out, c, sizeof(c));
}

enum { TAG_O = 0, TAG_V0 = 1, TAG_VMID = 2, TAG_X0 = 3, TAG_BOOT = 4, TAG_EQ = 5 };

static void philox_block(uint64_t seed, uint64_t replica, uint32_t draw, uint32_t tag, uint32_t unit,
                         uint32_t out[4])
{
    uint32_t ctr[4] = {(uint32_t)replica, (uint32_t)(replica >> 32), draw, (tag << 28) | unit};
    uint32_t key[2] = {(uint32_t)seed, (uint32_t)(seed >> 32)};
    orc_philox4x32_10(ctr, key, out);
}

static const double TWO_M32 = 2.3283064365386962890625e-10; /* 2^-32 */
static const double TWO_PI = 6.283185307179586476925286766559;

/* four standard normals from one Philox block: (r0,r1)->(n0,n1), (r2,r3)->(n2,n3) */
static void normals4(const uint32_t r[4], double n[4])
{
    for (int p = 0; p < 2; ++p) {
        const double u1 = ((double)r[2 * p] + 0.5) * TWO_M32;
        const double u2 = ((double)r[2 * p + 1] + 0.5) * TWO_M32;
        const double rad = sqrt(-2.0 * log(u1));
        n[2 * p] = rad * cos(TWO_PI * u2);
        n[2 * p + 1] = rad * sin(TWO_PI * u2);
    }
}

ORC_API void orc_normals4(uint64_t seed, uint64_t replica, uint32_t draw, uint32_t tag, uint32_t unit,
                          double out[4])
{
    uint32_t r[4];
    philox_block(seed, replica, draw, tag, unit, r);
    normals4(r, out);
}

/* scalar-stream normal number k of a (replica, tag): block k>>2, lane k&3 */
static double scalar_normal(uint64_t seed, uint64_t replica, uint32_t tag, uint32_t k)
{
    uint32_t r[4];
    double n[4];
    philox_block(seed, replica, k >> 2, tag, 0u, r);
    normals4(r, n);
    return n[k & 3u];
}

/* uniform index in [0, n): (r * n) >> 32, r = lane 0 of block (replica, 0, TAG_X0, 0) */
static int64_t cache_index(uint64_t seed, uint64_t replica, int64_t n)
{
    uint32_t r[4];
    philox_block(seed, replica, 0u, TAG_X0, 0u, r);
    return (int64_t)(((uint64_t)r[0] * (uint64_t)n) >> 32);
}

ORC_API int64_t orc_cache_index(uint64_t seed, uint64_t replica, int64_t n)
{
    return cache_index(seed, replica, n);
}

/* --------------------------------------------------------- toy potentials */
/* low_dimensional_systems.py:52-57 (quartic), :133-135 (double well); kind 2 is a
 * harmonic well (k/2 x^2, k = p[0]) used by analytic tests only. */
enum { POT_QUARTIC = 0, POT_DOUBLE_WELL = 1, POT_HARMONIC = 2 };

static double toy_potential(int kind, const double* p, double x)
{
    switch (kind) {
    case POT_QUARTIC: return x * x * x * x;
    case POT_DOUBLE_WELL: {
        const double x2 = x * x;
        return x2 * x2 * x2 + 2.0 * cos(5.0 * (x + 1.0));
    }
    default: return 0.5 * p[0] * x * x;
    }
}

static double toy_force(int kind, const double* p, double x)
{
    switch (kind) {
    case POT_QUARTIC: return -4.0 * x * x * x;
    case POT_DOUBLE_WELL: {
        const double x2 = x * x;
        return -(6.0 * x2 * x2 * x - 10.0 * sin(5.0 * (x + 1.0)));
    }
    default: return -p[0] * x;
    }
}

ORC_API double orc_toy_potential(int kind, const double* p, double x) { return toy_potential(kind, p, x); }
ORC_API double orc_toy_force(int kind, const double* p, double x) { return toy_force(kind, p, x); }

/* ------------------------------------------------------------ toy protocol */
/* op kinds */
enum { OP_R = 0, OP_V = 1, OP_O = 2 };

/* fp32 Box-Muller used by the engine's "mixed" precision (24-bit uniforms, as OpenMM's CUDA
 * platform generates its `gaussian` in single precision): log / sinpi / cospi are evaluated in
 * double and rounded to float, i.e. the correctly rounded values the GPU's logf / sincospif
 * approximate to within 1-2 ulp. */
static void normals4_f32(const uint32_t r[4], double n[4])
{
    for (int p = 0; p < 2; ++p) {
        const float u1 = ((float)(r[2 * p] >> 8) + 0.5f) * 5.9604644775390625e-8f;
        const float u2x2 = ((float)(r[2 * p + 1] >> 8) + 0.5f) * 1.1920928955078125e-7f;
        const float rad = sqrtf(-2.0f * (float)log((double)u1));
        const float c = (float)cos(3.14159265358979323846 * (double)u2x2);
        const float sn = (float)sin(3.14159265358979323846 * (double)u2x2);
        n[2 * p] = (double)(rad * c);
        n[2 * p + 1] = (double)(rad * sn);
    }
}

/*
 * Runs, for each replica, n_legs legs of `n_steps` applications of the
 * splitting program (bookkeepers.py:247-295 / low_dimensional_systems.py:160-184):
 *   leg 0: x0 from cache (or x0_in), v0 ~ MB (or v0_in)
 *   leg 1: x continues; v resampled (marginal 0 = "configuration") or kept (1 = "full")
 * W[leg] = ((E_end - E_start) - (Q_end - Q_start)) / kT     (bookkeepers.py:237-243)
 * Optional traces hold, per leg, n_steps+1 entries (entry 0 = leg start) of
 * x, v and the cumulative W in kT (numba_integrators.py:56-61).
 * If `wdirect` is non-NULL the direct shadow-work accumulator
 * (langevin.py:125-139,150-163) is returned per leg as well.
 * O-noise (DESIGN.md "RNG layout"): within leg l, the q-th O draw (q = step * n_O_tokens +
 * ordinal) is lane q & 3 of the Philox block with draw word (l << 26) | (q >> 2), tag TAG_O.
 */
ORC_API int orc_toy_protocol(int pot_kind, const double* pot_params, double mass, double kT,
                             int n_ops, const int32_t* op_kind, const double* op_frac,
                             const double* op_a, const double* op_b, double dt,
                             uint64_t seed, uint64_t replica0, int64_t n_replicas,
                             const double* x_cache, int64_t n_cache,
                             const double* x0_in, const double* v0_in,
                             int n_steps, int marginal, int n_legs, int noise_fp32,
                             double* W, double* Q, double* wdirect,
                             double* trace_x, double* trace_v, double* trace_w,
                             double* x_final, double* v_final)
{
    int n_O = 0;
    for (int i = 0; i < n_ops; ++i) n_O += (op_kind[i] == OP_O);
    const double sigma = sqrt(kT / mass);
    const int64_t T = (int64_t)n_steps + 1;

#pragma omp parallel for schedule(static)
    for (int64_t r = 0; r < n_replicas; ++r) {
        const uint64_t rid = replica0 + (uint64_t)r;
        double x = x0_in ? x0_in[r] : x_cache[cache_index(seed, rid, n_cache)];
        double v = v0_in ? v0_in[r] : sigma * scalar_normal(seed, rid, TAG_V0, 0u);
        double heat = 0.0;
        for (int leg = 0; leg < n_legs; ++leg) {
            if (leg > 0 && marginal == 0) v = sigma * scalar_normal(seed, rid, TAG_VMID, (uint32_t)(leg - 1));
            const double E0 = toy_potential(pot_kind, pot_params, x) + 0.5 * mass * v * v;
            const double Q0 = heat;
            double wd = 0.0;
            uint32_t blk_id = 0xFFFFFFFFu; /* O-stream block held in nblk (one Philox call per 4 draws) */
            double nblk[4] = {0, 0, 0, 0};
            uint32_t q = 0;
            const int64_t base = (r * n_legs + leg) * T;
            if (trace_x) trace_x[base] = x;
            if (trace_v) trace_v[base] = v;
            if (trace_w) trace_w[base] = 0.0;
            for (int s = 0; s < n_steps; ++s) {
                for (int i = 0; i < n_ops; ++i) {
                    const double h = dt * op_frac[i];
                    if (op_kind[i] == OP_R) {
                        const double pe_old = toy_potential(pot_kind, pot_params, x);
                        x = x + h * v;
                        wd += toy_potential(pot_kind, pot_params, x) - pe_old;
                    } else if (op_kind[i] == OP_V) {
                        const double ke_old = 0.5 * mass * v * v;
                        v = v + h * toy_force(pot_kind, pot_params, x) / mass;
                        wd += 0.5 * mass * v * v - ke_old;
                    } else {
                        const double ke_old = 0.5 * mass * v * v;
                        if ((q >> 2) != blk_id) {
                            uint32_t rr[4];
                            blk_id = q >> 2;
                            philox_block(seed, rid, ((uint32_t)leg << 26) | blk_id, TAG_O, 0u, rr);
                            if (noise_fp32) normals4_f32(rr, nblk);
                            else normals4(rr, nblk);
                        }
                        const double xi = nblk[q & 3u];
                        ++q;
                        v = op_a[i] * v + op_b[i] * sigma * xi;
                        heat += 0.5 * mass * v * v - ke_old;
                    }
                }
                if (trace_x) trace_x[base + s + 1] = x;
                if (trace_v) trace_v[base + s + 1] = v;
                if (trace_w) {
                    const double E = toy_potential(pot_kind, pot_params, x) + 0.5 * mass * v * v;
                    trace_w[base + s + 1] = ((E - E0) - (heat - Q0)) / kT;
                }
            }
            const double E1 = toy_potential(pot_kind, pot_params, x) + 0.5 * mass * v * v;
            W[r * n_legs + leg] = ((E1 - E0) - (heat - Q0)) / kT;
            if (Q) Q[r * n_legs + leg] = heat - Q0;
            if (wdirect) wdirect[r * n_legs + leg] = wd / kT;
        }
        if (x_final) x_final[r] = x;
        if (v_final) v_final[r] = v;
    }
    return 0;
}

/* Same protocol fed EXTERNAL noise instead of Philox (for checking the oracle
 * against the reference's own numba closures, whose np.random.randn() stream
 * tests/golden/make_golden.py replays): xi[r][k] is the k-th O draw. */
ORC_API int orc_toy_replay(int pot_kind, const double* pot_params, double mass, double kT,
                           int n_ops, const int32_t* op_kind, const double* op_frac,
                           const double* op_a, const double* op_b, double dt,
                           double x0, double v0, int n_steps, const double* xi,
                           double* trace_x, double* trace_v, double* trace_w, double* Q_out)
{
    const double sigma = sqrt(kT / mass);
    double x = x0, v = v0, heat = 0.0;
    const double E0 = toy_potential(pot_kind, pot_params, x) + 0.5 * mass * v * v;
    int64_t k = 0;
    trace_x[0] = x; trace_v[0] = v; trace_w[0] = 0.0;
    for (int s = 0; s < n_steps; ++s) {
        for (int i = 0; i < n_ops; ++i) {
            const double h = dt * op_frac[i];
            if (op_kind[i] == OP_R) x = x + h * v;
            else if (op_kind[i] == OP_V) v = v + h * toy_force(pot_kind, pot_params, x) / mass;
            else {
                const double ke_old = 0.5 * mass * v * v;
                v = op_a[i] * v + op_b[i] * sigma * xi[k++];
                heat += 0.5 * mass * v * v - ke_old;
            }
        }
        trace_x[s + 1] = x; trace_v[s + 1] = v;
        trace_w[s + 1] = ((toy_potential(pot_kind, pot_params, x) + 0.5 * mass * v * v - E0) - heat) / kT;
    }
    *Q_out = heat;
    return 0;
}

/* Independent random-walk Metropolis chains (numba_integrators.py:229-252 run n times for
 * n_burn steps from N(0,1) starts): chain c -> replica id replica0 + c, TAG_EQ, draw = step;
 * proposal = first Box-Muller normal of the block, acceptance uniform = (word 2 + 0.5) 2^-32. */
ORC_API int orc_sample_equilibrium_scalar(int pot_kind, const double* pot_params, double kT, int64_t n,
                                          int n_burn, uint64_t seed, uint64_t replica0, double* x_out)
{
    for (int64_t c = 0; c < n; ++c) {
        uint32_t r[4];
        double nn[4];
        philox_block(seed, replica0 + (uint64_t)c, 0u, TAG_EQ, 0u, r);
        normals4(r, nn);
        double x = nn[0];
        for (int i = 1; i <= n_burn; ++i) {
            philox_block(seed, replica0 + (uint64_t)c, (uint32_t)i, TAG_EQ, 0u, r);
            normals4(r, nn);
            const double accept_eps = ((double)r[2] + 0.5) * TWO_M32;
            const double x_prop = x + nn[0];
            const double ratio = exp(-toy_potential(pot_kind, pot_params, x_prop) / kT) /
                                 exp(-toy_potential(pot_kind, pot_params, x) / kT);
            if (ratio > accept_eps) x = x_prop;
        }
        x_out[c] = x;
    }
    return 0;
}

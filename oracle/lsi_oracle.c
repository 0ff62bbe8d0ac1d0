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

/* Scalar (toy) systems, DESIGN.md "RNG layout": ONE start block per replica, (replica, draw 0, TAG_V0, unit 0), feeds
 * everything a two-leg protocol draws outside its O substeps: words 0, 1 -> the Box-Muller pair (start velocity normal,
 * midpoint velocity normal), word 2 -> the equilibrium-cache index (w2 * n) >> 32, word 3 spare. */
static void toy_start_block(uint64_t seed, uint64_t replica, int64_t n_cache, double* n_start, double* n_mid, int64_t* index)
{
    uint32_t r[4];
    double n[4];
    philox_block(seed, replica, 0u, TAG_V0, 0u, r);
    normals4(r, n);
    *n_start = n[0];
    *n_mid = n[1];
    *index = (int64_t)(((uint64_t)r[2] * (uint64_t)n_cache) >> 32);
}

/* start-of-leg velocity normal of leg >= 2: normal leg - 2 of the TAG_V0 stream that begins at draw 1 */
static double toy_later_leg_normal(uint64_t seed, uint64_t replica, uint32_t leg)
{
    const uint32_t k = leg - 2u;
    uint32_t r[4];
    double n[4];
    philox_block(seed, replica, 1u + (k >> 2), TAG_V0, 0u, r);
    normals4(r, n);
    return n[k & 3u];
}

ORC_API int64_t orc_toy_cache_index(uint64_t seed, uint64_t replica, int64_t n)
{
    double a, b;
    int64_t idx;
    toy_start_block(seed, replica, n, &a, &b, &idx);
    return idx;
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
 *   leg 0: x0 from cache (or x0_in), v0 ~ MB (or v0_in): index and start normal from the replica's start block
 *   leg l: x continues; v resampled (marginal 0 = "configuration": leg 1 takes the second normal of the start block's
 *          Box-Muller pair, later legs the TAG_V0 stream from draw 1) or kept (1 = "full")
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
        double n_start, n_mid;
        int64_t idx0;
        toy_start_block(seed, rid, n_cache > 0 ? n_cache : 1, &n_start, &n_mid, &idx0);
        double x = x0_in ? x0_in[r] : x_cache[idx0];
        double v = v0_in ? v0_in[r] : sigma * n_start;
        double heat = 0.0;
        for (int leg = 0; leg < n_legs; ++leg) {
            if (leg > 0 && marginal == 0) v = sigma * (leg == 1 ? n_mid : toy_later_leg_normal(seed, rid, (uint32_t)leg));
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

/* =========================================================================================
 * Particle systems (water cluster, alanine dipeptide, tiny waterbox)
 *
 * Force field terms follow the OpenMM force classes the reference's systems are built from
 * ([ext]: OpenMM is not available here, the functional forms are restated from its
 * documentation): NonbondedForce (Lennard-Jones with Lorentz-Berthelot combination + Coulomb,
 * NoCutoff or CutoffPeriodic with reaction field, exclusions and 1-4 exceptions),
 * HarmonicBondForce, HarmonicAngleForce, PeriodicTorsionForce, and the CustomExternalForce
 * (K/2)(x^2+y^2+z^2) of benchmark/testsystems/watercluster.py:70-80.
 * Substeps with constraints: benchmark/integrators/langevin.py:124-175
 *   R: x += h v; constrain positions; v += (x - x_unconstrained)/h; constrain velocities
 *   V: v += h f_g/m; constrain velocities        O: v = a v + b sqrt(kT/m) xi; constrain velocities
 * Protocol: benchmark/testsystems/bookkeepers.py:184-295.
 * ========================================================================================= */

#define ONE_4PI_EPS0 138.935456 /* kJ nm / (mol e^2); benchmark/integrators/kyle/constants.py:8 */

typedef struct {
    int32_t n_atoms;
    const double* mass;
    const double* charge;
    const double* sigma;
    const double* epsilon;
    int32_t nonbonded_method; /* 0 NoCutoff, 1 CutoffPeriodic (reaction field) */
    double cutoff;
    double box[3];
    double rf_dielectric;
    int32_t n_exclusions;
    const int32_t* exclusions; /* [n][2] */
    int32_t n_exceptions;
    const int32_t* exception_atoms; /* [n][2] */
    const double* exception_params; /* [n][3] chargeProd, sigma, epsilon */
    int32_t n_bonds;
    const int32_t* bond_atoms;
    const double* bond_params; /* [n][2] r0, k */
    int32_t n_angles;
    const int32_t* angle_atoms;
    const double* angle_params; /* [n][2] theta0, k */
    int32_t n_torsions;
    const int32_t* torsion_atoms;
    const double* torsion_params; /* [n][3] periodicity, phase, k */
    double restraint_k;
    int32_t group_nonbonded, group_bonds, group_angles, group_torsions, group_restraint;
    int32_t n_settle;
    const int32_t* settle_atoms; /* [n][3] O, H1, H2 */
    double settle_oh, settle_hh;
    int32_t n_constraints;
    const int32_t* constraint_atoms; /* [n][2] */
    const double* constraint_dist;
    double constraint_tolerance;
    double switch_distance; /* LJ switching function starts here (<= 0: plain truncation); OpenMM NonbondedForce */
} OrcParticles;

static inline int in_group(int term_group, int g) { return g < 0 || term_group == g; }

static void min_image(const OrcParticles* S, double d[3])
{
    if (S->nonbonded_method == 1)
        for (int c = 0; c < 3; ++c) d[c] -= S->box[c] * floor(d[c] / S->box[c] + 0.5);
}

/* forces (3N, overwritten) and energy of the force classes selected by group g (-1 = all) */
static double particles_forces_f64(const OrcParticles* S, const double* x, int g, double* f)
{
    const int N = S->n_atoms;
    double E = 0.0;
    memset(f, 0, sizeof(double) * 3 * N);
    if (in_group(S->group_nonbonded, g)) {
        /* exclusion lookup (dense; N is at most a few hundred) */
        unsigned char* ex = (unsigned char*)calloc((size_t)N * N, 1);
        for (int k = 0; k < S->n_exclusions; ++k) {
            const int i = S->exclusions[2 * k], j = S->exclusions[2 * k + 1];
            ex[(size_t)i * N + j] = ex[(size_t)j * N + i] = 1;
        }
        for (int k = 0; k < S->n_exceptions; ++k) {
            const int i = S->exception_atoms[2 * k], j = S->exception_atoms[2 * k + 1];
            ex[(size_t)i * N + j] = ex[(size_t)j * N + i] = 1;
        }
        const double rc = S->cutoff, rc2 = rc * rc;
        const double eps_rf = S->rf_dielectric;
        const double krf = S->nonbonded_method == 1 ? (1.0 / (rc * rc * rc)) * (eps_rf - 1.0) / (2.0 * eps_rf + 1.0) : 0.0;
        const double crf = S->nonbonded_method == 1 ? (1.0 / rc) * (3.0 * eps_rf) / (2.0 * eps_rf + 1.0) : 0.0;
        for (int i = 0; i < N; ++i)
            for (int j = i + 1; j < N; ++j) {
                if (ex[(size_t)i * N + j]) continue;
                double d[3] = {x[3 * j] - x[3 * i], x[3 * j + 1] - x[3 * i + 1], x[3 * j + 2] - x[3 * i + 2]};
                min_image(S, d);
                const double r2 = d[0] * d[0] + d[1] * d[1] + d[2] * d[2];
                if (S->nonbonded_method == 1 && r2 >= rc2) continue;
                const double r = sqrt(r2), inv_r = 1.0 / r;
                const double sig = 0.5 * (S->sigma[i] + S->sigma[j]);
                const double eps = sqrt(S->epsilon[i] * S->epsilon[j]);
                const double sr2 = sig * sig / r2, sr6 = sr2 * sr2 * sr2;
                const double qq = ONE_4PI_EPS0 * S->charge[i] * S->charge[j];
                /* -dE/dr * (1/r): LJ 4 eps (12 sr12 - 6 sr6)/r^2 ; Coulomb qq (1/r^3 - 2 krf) */
                double e_lj = 4.0 * eps * (sr6 * sr6 - sr6);
                double f_lj = 4.0 * eps * (12.0 * sr6 * sr6 - 6.0 * sr6) / r2;
                if (S->nonbonded_method == 1 && S->switch_distance > 0.0 && r > S->switch_distance) {
                    /* OpenMM's switching function: E S(x), S = 1 - 6 x^5 + 15 x^4 - 10 x^3 */
                    const double w = rc - S->switch_distance, xs = (r - S->switch_distance) / w;
                    const double sw = 1.0 + xs * xs * xs * (-10.0 + xs * (15.0 - 6.0 * xs));
                    const double dsw = xs * xs * (-30.0 + xs * (60.0 - 30.0 * xs)) / w;
                    f_lj = f_lj * sw - e_lj * dsw * inv_r;
                    e_lj *= sw;
                }
                const double fr = f_lj + qq * (inv_r * inv_r * inv_r - 2.0 * krf);
                E += e_lj + qq * (inv_r + krf * r2 - crf);
                for (int c = 0; c < 3; ++c) {
                    f[3 * i + c] -= fr * d[c];
                    f[3 * j + c] += fr * d[c];
                }
            }
        /* exceptions: plain Coulomb + LJ with their own parameters (no reaction field) */
        for (int k = 0; k < S->n_exceptions; ++k) {
            const int i = S->exception_atoms[2 * k], j = S->exception_atoms[2 * k + 1];
            const double qq = ONE_4PI_EPS0 * S->exception_params[3 * k], sig = S->exception_params[3 * k + 1],
                         eps = S->exception_params[3 * k + 2];
            double d[3] = {x[3 * j] - x[3 * i], x[3 * j + 1] - x[3 * i + 1], x[3 * j + 2] - x[3 * i + 2]};
            min_image(S, d);
            const double r2 = d[0] * d[0] + d[1] * d[1] + d[2] * d[2];
            if (S->nonbonded_method == 1 && r2 >= rc2) continue;
            const double inv_r = 1.0 / sqrt(r2);
            const double sr2 = sig * sig / r2, sr6 = sr2 * sr2 * sr2;
            const double fr = 4.0 * eps * (12.0 * sr6 * sr6 - 6.0 * sr6) / r2 + qq * inv_r * inv_r * inv_r;
            E += 4.0 * eps * (sr6 * sr6 - sr6) + qq * inv_r;
            for (int c = 0; c < 3; ++c) {
                f[3 * i + c] -= fr * d[c];
                f[3 * j + c] += fr * d[c];
            }
        }
        free(ex);
    }
    if (in_group(S->group_bonds, g))
        for (int k = 0; k < S->n_bonds; ++k) {
            const int i = S->bond_atoms[2 * k], j = S->bond_atoms[2 * k + 1];
            const double r0 = S->bond_params[2 * k], kb = S->bond_params[2 * k + 1];
            double d[3] = {x[3 * j] - x[3 * i], x[3 * j + 1] - x[3 * i + 1], x[3 * j + 2] - x[3 * i + 2]};
            const double r = sqrt(d[0] * d[0] + d[1] * d[1] + d[2] * d[2]);
            E += 0.5 * kb * (r - r0) * (r - r0);
            const double fr = kb * (r - r0) / r; /* dE/dr / r */
            for (int c = 0; c < 3; ++c) {
                f[3 * i + c] += fr * d[c];
                f[3 * j + c] -= fr * d[c];
            }
        }
    if (in_group(S->group_angles, g))
        for (int k = 0; k < S->n_angles; ++k) {
            const int i = S->angle_atoms[3 * k], j = S->angle_atoms[3 * k + 1], l = S->angle_atoms[3 * k + 2];
            const double th0 = S->angle_params[2 * k], ka = S->angle_params[2 * k + 1];
            double a[3], b[3];
            for (int c = 0; c < 3; ++c) {
                a[c] = x[3 * i + c] - x[3 * j + c];
                b[c] = x[3 * l + c] - x[3 * j + c];
            }
            const double ra2 = a[0] * a[0] + a[1] * a[1] + a[2] * a[2], rb2 = b[0] * b[0] + b[1] * b[1] + b[2] * b[2];
            const double ra = sqrt(ra2), rb = sqrt(rb2);
            double ct = (a[0] * b[0] + a[1] * b[1] + a[2] * b[2]) / (ra * rb);
            if (ct > 1.0) ct = 1.0;
            if (ct < -1.0) ct = -1.0;
            const double th = acos(ct);
            E += 0.5 * ka * (th - th0) * (th - th0);
            const double dEdth = ka * (th - th0);
            double st = sqrt(1.0 - ct * ct);
            if (st < 1e-12) st = 1e-12;
            /* dtheta/da = -(b/(ra rb) - ct a/ra^2)/sin(theta) */
            for (int c = 0; c < 3; ++c) {
                const double dth_da = -(b[c] / (ra * rb) - ct * a[c] / ra2) / st;
                const double dth_db = -(a[c] / (ra * rb) - ct * b[c] / rb2) / st;
                f[3 * i + c] -= dEdth * dth_da;
                f[3 * l + c] -= dEdth * dth_db;
                f[3 * j + c] += dEdth * (dth_da + dth_db);
            }
        }
    if (in_group(S->group_torsions, g))
        for (int k = 0; k < S->n_torsions; ++k) {
            const int a1 = S->torsion_atoms[4 * k], a2 = S->torsion_atoms[4 * k + 1], a3 = S->torsion_atoms[4 * k + 2],
                      a4 = S->torsion_atoms[4 * k + 3];
            const double n = S->torsion_params[3 * k], ph = S->torsion_params[3 * k + 1], kt = S->torsion_params[3 * k + 2];
            double b1[3], b2[3], b3[3];
            for (int c = 0; c < 3; ++c) {
                b1[c] = x[3 * a2 + c] - x[3 * a1 + c];
                b2[c] = x[3 * a3 + c] - x[3 * a2 + c];
                b3[c] = x[3 * a4 + c] - x[3 * a3 + c];
            }
            /* n1 = b1 x b2, n2 = b2 x b3; phi = atan2(|b2| b1.n2, n1.n2)  (IUPAC sign) */
            const double n1[3] = {b1[1] * b2[2] - b1[2] * b2[1], b1[2] * b2[0] - b1[0] * b2[2], b1[0] * b2[1] - b1[1] * b2[0]};
            const double n2[3] = {b2[1] * b3[2] - b2[2] * b3[1], b2[2] * b3[0] - b2[0] * b3[2], b2[0] * b3[1] - b2[1] * b3[0]};
            const double b2n = sqrt(b2[0] * b2[0] + b2[1] * b2[1] + b2[2] * b2[2]);
            const double yy = b2n * (b1[0] * n2[0] + b1[1] * n2[1] + b1[2] * n2[2]);
            const double xx = n1[0] * n2[0] + n1[1] * n2[1] + n1[2] * n2[2];
            const double phi = atan2(yy, xx);
            E += kt * (1.0 + cos(n * phi - ph));
            const double dEdphi = -kt * n * sin(n * phi - ph);
            /* gradient of phi (Blondel & Karplus 1996) */
            const double n1sq = n1[0] * n1[0] + n1[1] * n1[1] + n1[2] * n1[2];
            const double n2sq = n2[0] * n2[0] + n2[1] * n2[1] + n2[2] * n2[2];
            const double b1b2 = b1[0] * b2[0] + b1[1] * b2[1] + b1[2] * b2[2];
            const double b3b2 = b3[0] * b2[0] + b3[1] * b2[1] + b3[2] * b2[2];
            for (int c = 0; c < 3; ++c) {
                const double g1 = -b2n / n1sq * n1[c];
                const double g4 = b2n / n2sq * n2[c];
                const double g2 = -g1 - (b1b2 / (b2n * b2n)) * g1 + (b3b2 / (b2n * b2n)) * g4;
                const double g3 = -g4 + (b1b2 / (b2n * b2n)) * g1 - (b3b2 / (b2n * b2n)) * g4;
                f[3 * a1 + c] -= dEdphi * g1;
                f[3 * a2 + c] -= dEdphi * g2;
                f[3 * a3 + c] -= dEdphi * g3;
                f[3 * a4 + c] -= dEdphi * g4;
            }
        }
    if (in_group(S->group_restraint, g) && S->restraint_k != 0.0)
        for (int i = 0; i < 3 * N; ++i) {
            E += 0.5 * S->restraint_k * x[i] * x[i];
            f[i] -= S->restraint_k * x[i];
        }
    return E;
}

ORC_API double orc_particles_energy_forces(const OrcParticles* S, const double* x, int group, double* f)
{
    return particles_forces_f64(S, x, group, f);
}

/* ---- "mixed" precision restatement of the same forces (how the reference configures OpenMM's CUDA platform,
 * benchmark/testsystems/configuration.py:42-45: fp32 pair / bonded arithmetic, fp64 state and accumulation of
 * energies).  It follows the engine's mixed path (integrator-benchmark_b200/csrc/particles_kernel.cuh) operation
 * by operation where that matters for rounding: positions wrapped into the box in fp64 and THEN rounded to float,
 * charges pre-multiplied by sqrt(K), every atom accumulates its own force over a full shell of partners in float,
 * pair energies are float values summed in double, bonded terms are evaluated once per term in float, the restraint
 * uses the unwrapped position rounded to float.  It is not bit-identical to the GPU (rsqrt.approx vs 1/sqrtf, order
 * of the partner loop), which is why tests compare with a tolerance of 1e-5 of the energy scale. */
static double particles_forces_f32(const OrcParticles* S, const double* x, int g, double* f)
{
    const int N = S->n_atoms;
    const int periodic = S->nonbonded_method == 1;
    double E = 0.0;
    float* xs = (float*)malloc(sizeof(float) * 3 * N);
    float* ff = (float*)calloc((size_t)3 * N, sizeof(float));
    for (int i = 0; i < N; ++i)
        for (int c = 0; c < 3; ++c) {
            double p = x[3 * i + c];
            if (periodic) p -= S->box[c] * floor(p / S->box[c]);
            xs[3 * i + c] = (float)p;
        }
    const float bx[3] = {(float)S->box[0], (float)S->box[1], (float)S->box[2]};
    const float ibx[3] = {periodic ? (float)(1.0 / S->box[0]) : 0.f, periodic ? (float)(1.0 / S->box[1]) : 0.f,
                          periodic ? (float)(1.0 / S->box[2]) : 0.f};
#define MIN_IMAGE_F(d)                                                         \
    do {                                                                       \
        if (periodic)                                                          \
            for (int c_ = 0; c_ < 3; ++c_) (d)[c_] = (d)[c_] - bx[c_] * rintf((d)[c_] * ibx[c_]); \
    } while (0)
    if (in_group(S->group_nonbonded, g)) {
        unsigned char* ex = (unsigned char*)calloc((size_t)N * N, 1);
        for (int k = 0; k < S->n_exclusions; ++k) {
            const int i = S->exclusions[2 * k], j = S->exclusions[2 * k + 1];
            ex[(size_t)i * N + j] = ex[(size_t)j * N + i] = 1;
        }
        for (int k = 0; k < S->n_exceptions; ++k) {
            const int i = S->exception_atoms[2 * k], j = S->exception_atoms[2 * k + 1];
            ex[(size_t)i * N + j] = ex[(size_t)j * N + i] = 1;
        }
        const double rc = S->cutoff, eps_rf = S->rf_dielectric;
        const float rc2 = (float)(rc * rc);
        const double krf_d = periodic ? (1.0 / (rc * rc * rc)) * (eps_rf - 1.0) / (2.0 * eps_rf + 1.0) : 0.0;
        const float krf = (float)krf_d, krf2 = (float)(2.0 * krf_d);
        const float crf = periodic ? (float)((1.0 / rc) * (3.0 * eps_rf) / (2.0 * eps_rf + 1.0)) : 0.f;
        const int use_switch = periodic && S->switch_distance > 0.0 && S->switch_distance < rc;
        const float sw_r = (float)S->switch_distance, sw_iw = use_switch ? (float)(1.0 / (rc - S->switch_distance)) : 0.f;
        const double sk = sqrt(ONE_4PI_EPS0);
        double e_pairs = 0.0;
        for (int i = 0; i < N; ++i) {
            const float qi = (float)(S->charge[i] * sk), hsi = (float)(0.5 * S->sigma[i]), sei = (float)(2.0 * sqrt(S->epsilon[i]));
            float fx = 0.f, fy = 0.f, fz = 0.f;
            double en = 0.0;
            for (int j = 0; j < N; ++j) {
                if (j == i || ex[(size_t)i * N + j]) continue;
                float d[3] = {xs[3 * j] - xs[3 * i], xs[3 * j + 1] - xs[3 * i + 1], xs[3 * j + 2] - xs[3 * i + 2]};
                MIN_IMAGE_F(d);
                const float r2 = d[0] * d[0] + d[1] * d[1] + d[2] * d[2];
                if (periodic && !(r2 < rc2)) continue;
                const float inv_r = 1.0f / sqrtf(r2), inv_r2 = inv_r * inv_r;
                const float qq = qi * (float)(S->charge[j] * sk);
                float fr = periodic ? qq * (inv_r * inv_r2 - krf2) : qq * inv_r * inv_r2;
                float e = periodic ? qq * (inv_r + krf * r2 - crf) : qq * inv_r;
                const float sej = (float)(2.0 * sqrt(S->epsilon[j]));
                if (sei != 0.f && sej != 0.f) {
                    const float sig = hsi + (float)(0.5 * S->sigma[j]);
                    const float s2 = sig * sig * inv_r2, s6 = s2 * s2 * s2;
                    const float t = sei * sej * s6; /* 4 eps s6 */
                    float f_lj = t * (12.0f * s6 - 6.0f) * inv_r2;
                    float e_lj = t * (s6 - 1.0f);
                    if (use_switch) {
                        const float xs_ = fmaxf((r2 * inv_r - sw_r) * sw_iw, 0.f), xs2 = xs_ * xs_;
                        const float sw = 1.0f + xs2 * xs_ * (-10.0f + xs_ * (15.0f - 6.0f * xs_));
                        const float dsw = xs2 * (-30.0f + xs_ * (60.0f - 30.0f * xs_)) * sw_iw;
                        f_lj = f_lj * sw - e_lj * dsw * inv_r;
                        e_lj *= sw;
                    }
                    fr += f_lj;
                    e += e_lj;
                }
                fx -= fr * d[0]; fy -= fr * d[1]; fz -= fr * d[2];
                en += (double)e;
            }
            ff[3 * i] += fx; ff[3 * i + 1] += fy; ff[3 * i + 2] += fz;
            e_pairs += 0.5 * en; /* every pair is visited from both ends */
        }
        E += e_pairs;
        for (int k = 0; k < S->n_exceptions; ++k) {
            const int i = S->exception_atoms[2 * k], j = S->exception_atoms[2 * k + 1];
            const float kqq = (float)(ONE_4PI_EPS0 * S->exception_params[3 * k]), sig = (float)S->exception_params[3 * k + 1],
                        eps4 = (float)(4.0 * S->exception_params[3 * k + 2]);
            float d[3] = {xs[3 * j] - xs[3 * i], xs[3 * j + 1] - xs[3 * i + 1], xs[3 * j + 2] - xs[3 * i + 2]};
            MIN_IMAGE_F(d);
            const float r2 = d[0] * d[0] + d[1] * d[1] + d[2] * d[2];
            if (periodic && r2 >= rc2) continue;
            const float inv_r = 1.0f / sqrtf(r2), inv_r2 = inv_r * inv_r;
            const float s2 = sig * sig * inv_r2, s6 = s2 * s2 * s2;
            const float e = eps4 * (s6 * s6 - s6) + kqq * inv_r;
            const float fr = eps4 * (12.0f * s6 * s6 - 6.0f * s6) * inv_r2 + kqq * inv_r * inv_r2;
            E += (double)e;
            for (int c = 0; c < 3; ++c) {
                ff[3 * i + c] -= fr * d[c];
                ff[3 * j + c] += fr * d[c];
            }
        }
        free(ex);
    }
    if (in_group(S->group_bonds, g))
        for (int k = 0; k < S->n_bonds; ++k) {
            const int i = S->bond_atoms[2 * k], j = S->bond_atoms[2 * k + 1];
            const float r0 = (float)S->bond_params[2 * k], kb = (float)S->bond_params[2 * k + 1];
            float d[3] = {xs[3 * j] - xs[3 * i], xs[3 * j + 1] - xs[3 * i + 1], xs[3 * j + 2] - xs[3 * i + 2]};
            MIN_IMAGE_F(d);
            const float r2 = d[0] * d[0] + d[1] * d[1] + d[2] * d[2];
            const float inv_r = 1.0f / sqrtf(r2), dr = r2 * inv_r - r0;
            E += (double)(0.5f * kb * dr * dr);
            const float fr = kb * dr * inv_r;
            for (int c = 0; c < 3; ++c) {
                ff[3 * i + c] += fr * d[c];
                ff[3 * j + c] -= fr * d[c];
            }
        }
    if (in_group(S->group_angles, g))
        for (int k = 0; k < S->n_angles; ++k) {
            const int i = S->angle_atoms[3 * k], j = S->angle_atoms[3 * k + 1], l = S->angle_atoms[3 * k + 2];
            const float th0 = (float)S->angle_params[2 * k], ka = (float)S->angle_params[2 * k + 1];
            float a[3], b[3];
            for (int c = 0; c < 3; ++c) {
                a[c] = xs[3 * i + c] - xs[3 * j + c];
                b[c] = xs[3 * l + c] - xs[3 * j + c];
            }
            MIN_IMAGE_F(a);
            MIN_IMAGE_F(b);
            const float ira = 1.0f / sqrtf(a[0] * a[0] + a[1] * a[1] + a[2] * a[2]);
            const float irb = 1.0f / sqrtf(b[0] * b[0] + b[1] * b[1] + b[2] * b[2]);
            float ct = (a[0] * b[0] + a[1] * b[1] + a[2] * b[2]) * ira * irb;
            ct = fminf(fmaxf(ct, -1.0f), 1.0f);
            const float dth = acosf(ct) - th0;
            const float st = sqrtf(fmaxf(1.0f - ct * ct, 1e-24f));
            const float cf = ka * dth / st; /* force = +cf d cos(theta)/d r */
            const float iab = ira * irb, ia2 = ira * ira, ib2 = irb * irb;
            E += (double)(0.5f * ka * dth * dth);
            for (int c = 0; c < 3; ++c) {
                const float ga = b[c] * iab - ct * a[c] * ia2, gb = a[c] * iab - ct * b[c] * ib2;
                ff[3 * i + c] += cf * ga;
                ff[3 * l + c] += cf * gb;
                ff[3 * j + c] -= cf * (ga + gb);
            }
        }
    if (in_group(S->group_torsions, g))
        for (int k = 0; k < S->n_torsions; ++k) {
            const int a1 = S->torsion_atoms[4 * k], a2 = S->torsion_atoms[4 * k + 1], a3 = S->torsion_atoms[4 * k + 2],
                      a4 = S->torsion_atoms[4 * k + 3];
            const int nper = (int)floor(S->torsion_params[3 * k] + 0.5);
            const float cph = (float)cos(S->torsion_params[3 * k + 1]), sph = (float)sin(S->torsion_params[3 * k + 1]);
            const float kt = (float)S->torsion_params[3 * k + 2];
            float b1[3], b2[3], b3[3];
            for (int c = 0; c < 3; ++c) {
                b1[c] = xs[3 * a2 + c] - xs[3 * a1 + c];
                b2[c] = xs[3 * a3 + c] - xs[3 * a2 + c];
                b3[c] = xs[3 * a4 + c] - xs[3 * a3 + c];
            }
            MIN_IMAGE_F(b1);
            MIN_IMAGE_F(b2);
            MIN_IMAGE_F(b3);
            const float n1[3] = {b1[1] * b2[2] - b1[2] * b2[1], b1[2] * b2[0] - b1[0] * b2[2], b1[0] * b2[1] - b1[1] * b2[0]};
            const float n2[3] = {b2[1] * b3[2] - b2[2] * b3[1], b2[2] * b3[0] - b2[0] * b3[2], b2[0] * b3[1] - b2[1] * b3[0]};
            const float b2sq = b2[0] * b2[0] + b2[1] * b2[1] + b2[2] * b2[2];
            const float ib2 = 1.0f / sqrtf(b2sq), b2n = b2sq * ib2;
            const float yy = b2n * (b1[0] * n2[0] + b1[1] * n2[1] + b1[2] * n2[2]);
            const float xx = n1[0] * n2[0] + n1[1] * n2[1] + n1[2] * n2[2];
            const float n1sq = n1[0] * n1[0] + n1[1] * n1[1] + n1[2] * n1[2];
            const float n2sq = n2[0] * n2[0] + n2[1] * n2[1] + n2[2] * n2[2];
            /* cos(n phi), sin(n phi) from cos phi, sin phi by the multiple-angle recurrence (as the kernel) */
            const float inorm = 1.0f / sqrtf(fmaxf(n1sq * n2sq, 1e-36f));
            const float c1 = xx * inorm, s1 = yy * inorm;
            float cn = c1, sn = s1;
            for (int q = 1; q < nper; ++q) {
                const float c2 = cn * c1 - sn * s1;
                sn = sn * c1 + cn * s1;
                cn = c2;
            }
            if (nper == 0) { cn = 1.0f; sn = 0.0f; }
            E += (double)(kt * (1.0f + cn * cph + sn * sph));
            const float dE = -kt * (float)nper * (sn * cph - cn * sph);
            const float k1 = -b2n / n1sq, k4 = b2n / n2sq, ib2sq = ib2 * ib2;
            const float u = (b1[0] * b2[0] + b1[1] * b2[1] + b1[2] * b2[2]) * ib2sq;
            const float w = (b3[0] * b2[0] + b3[1] * b2[1] + b3[2] * b2[2]) * ib2sq;
            for (int c = 0; c < 3; ++c) {
                const float g1 = k1 * n1[c], g4 = k4 * n2[c];
                ff[3 * a1 + c] += -dE * g1;
                ff[3 * a2 + c] += -dE * (-g1 - u * g1 + w * g4);
                ff[3 * a3 + c] += -dE * (-g4 + u * g1 - w * g4);
                ff[3 * a4 + c] += -dE * g4;
            }
        }
    if (in_group(S->group_restraint, g) && S->restraint_k != 0.0) {
        const float kr = (float)S->restraint_k;
        for (int i = 0; i < N; ++i) {
            const float px = (float)x[3 * i], py = (float)x[3 * i + 1], pz = (float)x[3 * i + 2];
            ff[3 * i] -= kr * px; ff[3 * i + 1] -= kr * py; ff[3 * i + 2] -= kr * pz;
            E += (double)(0.5f * kr * (px * px + py * py + pz * pz));
        }
    }
#undef MIN_IMAGE_F
    for (int i = 0; i < 3 * N; ++i) f[i] = (double)ff[i];
    free(xs);
    free(ff);
    return E;
}

ORC_API double orc_particles_energy_forces_f32(const OrcParticles* S, const double* x, int group, double* f)
{
    return particles_forces_f32(S, x, group, f);
}

/* ---- constraints ----------------------------------------------------------------------- */
/* SETTLE (Miyamoto & Kollman 1992) for one rigid three-site water: x0 = reference (constrained)
 * positions, x1 = unconstrained new positions (overwritten with the constrained ones). */
static void settle_positions(const double* x0, double* x1, const int32_t at[3], double mO, double mH,
                             double dOH, double dHH)
{
    const double *a0 = x0 + 3 * at[0], *b0 = x0 + 3 * at[1], *c0 = x0 + 3 * at[2];
    double *a1 = x1 + 3 * at[0], *b1 = x1 + 3 * at[1], *c1 = x1 + 3 * at[2];
    const double mtot = mO + 2.0 * mH;
    double xb0[3], xc0[3], xcom[3], xa1[3], xb1[3], xc1[3];
    for (int c = 0; c < 3; ++c) {
        xb0[c] = b0[c] - a0[c];
        xc0[c] = c0[c] - a0[c];
        xcom[c] = (mO * a1[c] + mH * b1[c] + mH * c1[c]) / mtot;
        xa1[c] = a1[c] - xcom[c];
        xb1[c] = b1[c] - xcom[c];
        xc1[c] = c1[c] - xcom[c];
    }
    /* orthonormal frame: Z normal to the reference triangle, X = xa1 x Z, Y = Z x X */
    double Z[3] = {xb0[1] * xc0[2] - xb0[2] * xc0[1], xb0[2] * xc0[0] - xb0[0] * xc0[2], xb0[0] * xc0[1] - xb0[1] * xc0[0]};
    double X[3] = {xa1[1] * Z[2] - xa1[2] * Z[1], xa1[2] * Z[0] - xa1[0] * Z[2], xa1[0] * Z[1] - xa1[1] * Z[0]};
    double Y[3] = {Z[1] * X[2] - Z[2] * X[1], Z[2] * X[0] - Z[0] * X[2], Z[0] * X[1] - Z[1] * X[0]};
    const double nz = 1.0 / sqrt(Z[0] * Z[0] + Z[1] * Z[1] + Z[2] * Z[2]);
    const double nx = 1.0 / sqrt(X[0] * X[0] + X[1] * X[1] + X[2] * X[2]);
    const double ny = 1.0 / sqrt(Y[0] * Y[0] + Y[1] * Y[1] + Y[2] * Y[2]);
    for (int c = 0; c < 3; ++c) { X[c] *= nx; Y[c] *= ny; Z[c] *= nz; }
#define DOT3(u, w) ((u)[0] * (w)[0] + (u)[1] * (w)[1] + (u)[2] * (w)[2])
    const double xb0d = DOT3(X, xb0), yb0d = DOT3(Y, xb0), xc0d = DOT3(X, xc0), yc0d = DOT3(Y, xc0);
    const double za1d = DOT3(Z, xa1);
    const double xb1d = DOT3(X, xb1), yb1d = DOT3(Y, xb1), zb1d = DOT3(Z, xb1);
    const double xc1d = DOT3(X, xc1), yc1d = DOT3(Y, xc1), zc1d = DOT3(Z, xc1);
    /* canonical triangle: ra = distance O - COM, rb = distance COM - HH midpoint, rc = half HH */
    const double rc = 0.5 * dHH;
    const double hgt = sqrt(dOH * dOH - rc * rc);
    const double ra = hgt * 2.0 * mH / mtot, rb = hgt - ra;
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
    for (int c = 0; c < 3; ++c) {
        a1[c] = xcom[c] + X[c] * xa3d + Y[c] * ya3d + Z[c] * za3d;
        b1[c] = xcom[c] + X[c] * xb3d + Y[c] * yb3d + Z[c] * zb3d;
        c1[c] = xcom[c] + X[c] * xc3d + Y[c] * yc3d + Z[c] * zc3d;
    }
}

/* exact velocity projection for a rigid triangle: solve the 3x3 system for the impulses along
 * the three edges so that every edge's relative velocity along itself vanishes */
static void settle_velocities(const double* x, double* v, const int32_t at[3], double mO, double mH)
{
    const int ia[3] = {at[0], at[1], at[2]};
    const double im[3] = {1.0 / mO, 1.0 / mH, 1.0 / mH};
    const int e[3][2] = {{0, 1}, {1, 2}, {2, 0}};
    double u[3][3], b[3], A[3][3];
    for (int k = 0; k < 3; ++k) {
        double n2 = 0.0;
        for (int c = 0; c < 3; ++c) {
            u[k][c] = x[3 * ia[e[k][1]] + c] - x[3 * ia[e[k][0]] + c];
            n2 += u[k][c] * u[k][c];
        }
        const double inv = 1.0 / sqrt(n2);
        b[k] = 0.0;
        for (int c = 0; c < 3; ++c) {
            u[k][c] *= inv;
            b[k] += (v[3 * ia[e[k][1]] + c] - v[3 * ia[e[k][0]] + c]) * u[k][c];
        }
    }
    /* impulse lambda_l along edge l changes v_p by  -lambda_l u_l / m_p (p = tail), +... (p = head) */
    for (int k = 0; k < 3; ++k)
        for (int l = 0; l < 3; ++l) {
            double coef = 0.0; /* d(relative velocity of edge k along u_k)/d lambda_l */
            const double ukul = DOT3(u[k], u[l]);
            if (e[k][1] == e[l][1]) coef += im[e[k][1]];
            if (e[k][1] == e[l][0]) coef -= im[e[k][1]];
            if (e[k][0] == e[l][1]) coef -= im[e[k][0]];
            if (e[k][0] == e[l][0]) coef += im[e[k][0]];
            A[k][l] = coef * ukul;
        }
    /* solve A lambda = -b (Cramer) */
    const double det = A[0][0] * (A[1][1] * A[2][2] - A[1][2] * A[2][1]) - A[0][1] * (A[1][0] * A[2][2] - A[1][2] * A[2][0]) +
                       A[0][2] * (A[1][0] * A[2][1] - A[1][1] * A[2][0]);
    const double r0 = -b[0], r1 = -b[1], r2 = -b[2];
    const double l0 = (r0 * (A[1][1] * A[2][2] - A[1][2] * A[2][1]) - A[0][1] * (r1 * A[2][2] - A[1][2] * r2) +
                       A[0][2] * (r1 * A[2][1] - A[1][1] * r2)) / det;
    const double l1 = (A[0][0] * (r1 * A[2][2] - A[1][2] * r2) - r0 * (A[1][0] * A[2][2] - A[1][2] * A[2][0]) +
                       A[0][2] * (A[1][0] * r2 - r1 * A[2][0])) / det;
    const double l2 = (A[0][0] * (A[1][1] * r2 - r1 * A[2][1]) - A[0][1] * (A[1][0] * r2 - r1 * A[2][0]) +
                       r0 * (A[1][0] * A[2][1] - A[1][1] * A[2][0])) / det;
    const double lam[3] = {l0, l1, l2};
    for (int l = 0; l < 3; ++l)
        for (int c = 0; c < 3; ++c) {
            v[3 * ia[e[l][1]] + c] += lam[l] * u[l][c] * im[e[l][1]];
            v[3 * ia[e[l][0]] + c] -= lam[l] * u[l][c] * im[e[l][0]];
        }
}
#undef DOT3

/* SHAKE: Gauss-Seidel sweeps over the distance constraints, displacements along the reference
 * bond vectors, until every |r^2 - d^2| <= 2 tol d^2 */
static int shake_positions(const OrcParticles* S, const double* x0, double* x1)
{
    const double tol = S->constraint_tolerance;
    for (int it = 0; it < 500; ++it) {
        int done = 1;
        for (int k = 0; k < S->n_constraints; ++k) {
            const int i = S->constraint_atoms[2 * k], j = S->constraint_atoms[2 * k + 1];
            const double d2 = S->constraint_dist[k] * S->constraint_dist[k];
            double r[3], s[3];
            for (int c = 0; c < 3; ++c) {
                r[c] = x1[3 * i + c] - x1[3 * j + c];
                s[c] = x0[3 * i + c] - x0[3 * j + c];
            }
            const double diff = r[0] * r[0] + r[1] * r[1] + r[2] * r[2] - d2;
            if (fabs(diff) > 2.0 * tol * d2) done = 0;
            const double rs = r[0] * s[0] + r[1] * s[1] + r[2] * s[2];
            const double gl = diff / (2.0 * rs * (1.0 / S->mass[i] + 1.0 / S->mass[j]));
            for (int c = 0; c < 3; ++c) {
                x1[3 * i + c] -= gl * s[c] / S->mass[i];
                x1[3 * j + c] += gl * s[c] / S->mass[j];
            }
        }
        if (done) return it;
    }
    return -1;
}

static int rattle_velocities(const OrcParticles* S, const double* x, double* v)
{
    const double tol = S->constraint_tolerance;
    for (int it = 0; it < 500; ++it) {
        int done = 1;
        for (int k = 0; k < S->n_constraints; ++k) {
            const int i = S->constraint_atoms[2 * k], j = S->constraint_atoms[2 * k + 1];
            const double d2 = S->constraint_dist[k] * S->constraint_dist[k];
            double r[3], w[3];
            for (int c = 0; c < 3; ++c) {
                r[c] = x[3 * i + c] - x[3 * j + c];
                w[c] = v[3 * i + c] - v[3 * j + c];
            }
            const double rv = r[0] * w[0] + r[1] * w[1] + r[2] * w[2];
            const double gl = rv / (d2 * (1.0 / S->mass[i] + 1.0 / S->mass[j]));
            if (fabs(gl) > tol) done = 0; /* impulse per unit mass-distance, in 1/ps */
            for (int c = 0; c < 3; ++c) {
                v[3 * i + c] -= gl * r[c] / S->mass[i];
                v[3 * j + c] += gl * r[c] / S->mass[j];
            }
        }
        if (done) return it;
    }
    return -1;
}

static void constrain_positions(const OrcParticles* S, const double* x0, double* x1)
{
    for (int w = 0; w < S->n_settle; ++w) {
        const int32_t* at = S->settle_atoms + 3 * w;
        settle_positions(x0, x1, at, S->mass[at[0]], S->mass[at[1]], S->settle_oh, S->settle_hh);
    }
    if (S->n_constraints) shake_positions(S, x0, x1);
}

static void constrain_velocities(const OrcParticles* S, const double* x, double* v)
{
    for (int w = 0; w < S->n_settle; ++w) {
        const int32_t* at = S->settle_atoms + 3 * w;
        settle_velocities(x, v, at, S->mass[at[0]], S->mass[at[1]]);
    }
    if (S->n_constraints) rattle_velocities(S, x, v);
}

ORC_API void orc_constrain_positions(const OrcParticles* S, const double* x0, double* x1) { constrain_positions(S, x0, x1); }
ORC_API void orc_constrain_velocities(const OrcParticles* S, const double* x, double* v) { constrain_velocities(S, x, v); }

static double kinetic_energy(const OrcParticles* S, const double* v)
{
    double ke = 0.0;
    for (int i = 0; i < S->n_atoms; ++i)
        ke += 0.5 * S->mass[i] * (v[3 * i] * v[3 * i] + v[3 * i + 1] * v[3 * i + 1] + v[3 * i + 2] * v[3 * i + 2]);
    return ke;
}

static void atom_normals(uint64_t seed, uint64_t rid, uint32_t draw, uint32_t tag, int N, int noise_fp32, double* out)
{
    for (int i = 0; i < N; ++i) {
        uint32_t r[4];
        double n[4];
        philox_block(seed, rid, draw, tag, (uint32_t)i, r);
        if (noise_fp32) normals4_f32(r, n);
        else normals4(r, n);
        out[3 * i] = n[0]; out[3 * i + 1] = n[1]; out[3 * i + 2] = n[2];
    }
}

/* flags */
#define ORC_FLAG_ROUND_MIDPOINT_X_FP32 1 /* bookkeepers.py:266,298-305: x passes through a float32 frame */
#define ORC_FLAG_FORCE_FP32 2            /* forces and potential energies from particles_forces_f32 ("mixed") */

/*
 * Protocol for particle systems (bookkeepers.py:184-295).  x_cache: [n_cache][N][3];
 * x0_in / v0_in: [n][N][3].  Per leg: set state, apply position then velocity constraints
 * (:206-208), E0 = PE + KE, run n_steps, W = ((E1 - E0) - (heat1 - heat0)) / kT.
 * O noise: atom i, q-th O draw of leg l -> Philox(replica, draw = (l << 26) | q, TAG_O, unit i),
 * lanes 0..2 = x, y, z.  Start / midpoint velocities: TAG_V0 draw 0 / TAG_VMID draw (l - 1).
 * traces (optional): trace_x/v [n][n_legs][n_steps+1][N][3], trace_pe/ke/heat/wdir [n][n_legs][n_steps+1]
 */
ORC_API int orc_particles_protocol(const OrcParticles* S, double kT, int n_ops, const int32_t* op_kind,
                                   const int32_t* op_group, const double* op_frac, const double* op_a,
                                   const double* op_b, double dt, uint64_t seed, uint64_t replica0,
                                   int64_t n_replicas, const double* x_cache, int64_t n_cache,
                                   const double* x0_in, const double* v0_in, int n_steps, int marginal,
                                   int n_legs, int noise_fp32, int flags, double* W, double* Q, double* wdirect,
                                   double* x_final, double* v_final, double* trace_x, double* trace_v,
                                   double* trace_pe, double* trace_ke, double* trace_heat, double* trace_wdir)
{
    const int N = S->n_atoms, D = 3 * N;
    int n_O = 0;
    for (int i = 0; i < n_ops; ++i) n_O += (op_kind[i] == OP_O);
    const int64_t T = (int64_t)n_steps + 1;
    int rc_all = 0;
    double (*const particles_forces)(const OrcParticles*, const double*, int, double*) =
        (flags & ORC_FLAG_FORCE_FP32) ? particles_forces_f32 : particles_forces_f64;

#pragma omp parallel for schedule(dynamic, 1)
    for (int64_t r = 0; r < n_replicas; ++r) {
        const uint64_t rid = replica0 + (uint64_t)r;
        double* x = (double*)malloc(sizeof(double) * D);
        double* v = (double*)malloc(sizeof(double) * D);
        double* f = (double*)malloc(sizeof(double) * D);
        double* xu = (double*)malloc(sizeof(double) * D);
        double* xi = (double*)malloc(sizeof(double) * D);
        const double* xs = x0_in ? x0_in + r * D : x_cache + cache_index(seed, rid, n_cache) * D;
        memcpy(x, xs, sizeof(double) * D);
        if (v0_in) memcpy(v, v0_in + r * D, sizeof(double) * D);
        else {
            atom_normals(seed, rid, 0u, TAG_V0, N, 0, xi);
            for (int i = 0; i < D; ++i) v[i] = sqrt(kT / S->mass[i / 3]) * xi[i];
        }
        double heat = 0.0;
        double* fcache[34] = {0};
        char fvalid[34] = {0};
        const int want_direct = wdirect != NULL || trace_wdir != NULL;
        for (int leg = 0; leg < n_legs; ++leg) {
            memset(fvalid, 0, sizeof(fvalid));
            if (leg > 0) {
                if (flags & ORC_FLAG_ROUND_MIDPOINT_X_FP32)
                    for (int i = 0; i < D; ++i) x[i] = (double)(float)x[i];
                if (marginal == 0) {
                    atom_normals(seed, rid, (uint32_t)(leg - 1), TAG_VMID, N, 0, xi);
                    for (int i = 0; i < D; ++i) v[i] = sqrt(kT / S->mass[i / 3]) * xi[i];
                }
            }
            /* context.applyConstraints / applyVelocityConstraints (bookkeepers.py:206-208) */
            memcpy(xu, x, sizeof(double) * D);
            constrain_positions(S, xu, x);
            constrain_velocities(S, x, v);
            double pe = particles_forces(S, x, -1, f);
            const double E0 = pe + kinetic_energy(S, v);
            const double Q0 = heat;
            double wd = 0.0;
            uint32_t q = 0;
            const int64_t tb = (r * n_legs + leg) * T;
            if (trace_x) memcpy(trace_x + tb * D, x, sizeof(double) * D);
            if (trace_v) memcpy(trace_v + tb * D, v, sizeof(double) * D);
            if (trace_pe) trace_pe[tb] = pe;
            if (trace_ke) trace_ke[tb] = kinetic_energy(S, v);
            if (trace_heat) trace_heat[tb] = 0.0;
            if (trace_wdir) trace_wdir[tb] = 0.0;
            for (int s = 0; s < n_steps; ++s) {
                for (int i = 0; i < n_ops; ++i) {
                    const double h = dt * op_frac[i];
                    if (op_kind[i] == OP_R) {
                        const double e_old = want_direct ? particles_forces(S, x, -1, f) + kinetic_energy(S, v) : 0.0;
                        memset(fvalid, 0, sizeof(fvalid));
                        memcpy(xu, x, sizeof(double) * D); /* reference positions for the constraints */
                        for (int k = 0; k < D; ++k) x[k] = x[k] + h * v[k];
                        memcpy(xi, x, sizeof(double) * D); /* x1: unconstrained */
                        constrain_positions(S, xu, x);
                        for (int k = 0; k < D; ++k) v[k] = v[k] + (x[k] - xi[k]) / h;
                        constrain_velocities(S, x, v);
                        if (want_direct) wd += particles_forces(S, x, -1, f) + kinetic_energy(S, v) - e_old;
                    } else if (op_kind[i] == OP_V) {
                        const double ke_old = kinetic_energy(S, v);
                        /* forces of a group are reused until the positions change (as OpenMM caches them) */
                        const int gi = op_group[i] + 1;
                        if (!fcache[gi]) fcache[gi] = (double*)malloc(sizeof(double) * D);
                        if (!fvalid[gi]) { particles_forces(S, x, op_group[i], fcache[gi]); fvalid[gi] = 1; }
                        const double* fg = fcache[gi];
                        for (int k = 0; k < D; ++k) v[k] = v[k] + h * fg[k] / S->mass[k / 3];
                        constrain_velocities(S, x, v);
                        wd += kinetic_energy(S, v) - ke_old;
                    } else {
                        const double ke_old = kinetic_energy(S, v);
                        atom_normals(seed, rid, ((uint32_t)leg << 26) | q, TAG_O, N, noise_fp32, xi);
                        ++q;
                        for (int k = 0; k < D; ++k) v[k] = op_a[i] * v[k] + op_b[i] * sqrt(kT / S->mass[k / 3]) * xi[k];
                        constrain_velocities(S, x, v);
                        heat += kinetic_energy(S, v) - ke_old;
                    }
                }
                const int64_t ti = tb + s + 1;
                if (trace_x) memcpy(trace_x + ti * D, x, sizeof(double) * D);
                if (trace_v) memcpy(trace_v + ti * D, v, sizeof(double) * D);
                if (trace_pe) trace_pe[ti] = particles_forces(S, x, -1, f);
                if (trace_ke) trace_ke[ti] = kinetic_energy(S, v);
                if (trace_heat) trace_heat[ti] = heat - Q0;
                if (trace_wdir) trace_wdir[ti] = wd;
            }
            pe = particles_forces(S, x, -1, f);
            const double E1 = pe + kinetic_energy(S, v);
            W[r * n_legs + leg] = ((E1 - E0) - (heat - Q0)) / kT;
            if (Q) Q[r * n_legs + leg] = heat - Q0;
            if (wdirect) wdirect[r * n_legs + leg] = wd / kT;
        }
        if (x_final) memcpy(x_final + r * D, x, sizeof(double) * D);
        if (v_final) memcpy(v_final + r * D, v, sizeof(double) * D);
        free(x); free(v); free(f); free(xu); free(xi);
        for (int k = 0; k < 34; ++k) free(fcache[k]);
    }
    return rc_all;
}

/*
 * Extra-chance HMC equilibrium sampler (benchmark/integrators/kyle/xchmc.py:84-116 with collision_rate = None,
 * hamiltonian step of kyle/ghmc.py:103-114 in RATTLE form: every substep is followed by its projection).
 * Same streams as lsi_sample_equilibrium_particles: velocities (replica, draw = 2 round, TAG_EQ, unit = atom),
 * uniform (replica, 2 round + 1, TAG_EQ, 0) lane 0.  accept: [n][2] = {accepted rounds, accepted at first chance} / rounds.
 */
ORC_API int orc_particles_hmc(const OrcParticles* S, double kT, int64_t n, int n_rounds, int steps_per_hmc, int extra_chances,
                              double dt, uint64_t seed, uint64_t replica0, const double* x_in, double* x_out, double* v_out,
                              double* accept)
{
    const int N = S->n_atoms, D = 3 * N;
#pragma omp parallel for schedule(dynamic, 1)
    for (int64_t r = 0; r < n; ++r) {
        const uint64_t rid = replica0 + (uint64_t)r;
        double* x = (double*)malloc(sizeof(double) * D);
        double* v = (double*)calloc(D, sizeof(double));
        double* f = (double*)malloc(sizeof(double) * D);
        double* xu = (double*)malloc(sizeof(double) * D);
        double* xi = (double*)malloc(sizeof(double) * D);
        double* xold = (double*)malloc(sizeof(double) * D);
        double* vold = (double*)malloc(sizeof(double) * D);
        double* fold = (double*)malloc(sizeof(double) * D);
        memcpy(x, x_in + r * D, sizeof(double) * D);
        memcpy(xu, x, sizeof(double) * D);
        constrain_positions(S, xu, x);
        double pe = particles_forces_f64(S, x, -1, f);
        int n_acc = 0, n_first = 0;
        for (int round = 0; round < n_rounds; ++round) {
            atom_normals(seed, rid, 2u * (uint32_t)round, TAG_EQ, N, 0, xi);
            for (int i = 0; i < D; ++i) v[i] = sqrt(kT / S->mass[i / 3]) * xi[i];
            memcpy(xu, x, sizeof(double) * D);
            constrain_positions(S, xu, x);
            constrain_velocities(S, x, v);
            const double e_old = pe + kinetic_energy(S, v);
            memcpy(xold, x, sizeof(double) * D);
            memcpy(vold, v, sizeof(double) * D);
            memcpy(fold, f, sizeof(double) * D);
            uint32_t ur[4];
            philox_block(seed, rid, 2u * (uint32_t)round + 1u, TAG_EQ, 0u, ur);
            const double uni = ((double)ur[0] + 0.5) * 2.3283064365386962890625e-10;
            double mu1 = 0.0, pe_new = pe;
            for (int c = 0; c <= extra_chances; ++c) {
                if (!(uni > mu1)) break;
                for (int st = 0; st < steps_per_hmc; ++st) {
                    for (int k = 0; k < D; ++k) v[k] = v[k] + 0.5 * dt * f[k] / S->mass[k / 3];
                    constrain_velocities(S, x, v);
                    memcpy(xu, x, sizeof(double) * D);
                    for (int k = 0; k < D; ++k) x[k] = x[k] + dt * v[k];
                    memcpy(xi, x, sizeof(double) * D);
                    constrain_positions(S, xu, x);
                    for (int k = 0; k < D; ++k) v[k] = v[k] + (x[k] - xi[k]) / dt;
                    constrain_velocities(S, x, v);
                    pe_new = particles_forces_f64(S, x, -1, f);
                    for (int k = 0; k < D; ++k) v[k] = v[k] + 0.5 * dt * f[k] / S->mass[k / 3];
                    constrain_velocities(S, x, v);
                }
                double e_new = pe_new + kinetic_energy(S, v);
                if (!(e_new == e_new)) e_new = INFINITY;
                const double mu = fmin(1.0, exp(-(e_new - e_old) / kT));
                if (mu > mu1) mu1 = mu;
                if (c == 0 && uni <= mu1) ++n_first;
            }
            if (uni > mu1) {
                memcpy(x, xold, sizeof(double) * D);
                memcpy(v, vold, sizeof(double) * D);
                memcpy(f, fold, sizeof(double) * D);
            } else {
                pe = pe_new;
                ++n_acc;
            }
        }
        memcpy(x_out + r * D, x, sizeof(double) * D);
        if (v_out) memcpy(v_out + r * D, v, sizeof(double) * D);
        if (accept) {
            accept[2 * r] = (double)n_acc / (double)(n_rounds > 0 ? n_rounds : 1);
            accept[2 * r + 1] = (double)n_first / (double)(n_rounds > 0 ? n_rounds : 1);
        }
        free(x); free(v); free(f); free(xu); free(xi); free(xold); free(vold); free(fold);
    }
    return 0;
}

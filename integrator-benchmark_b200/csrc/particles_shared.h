// Tables and launch parameters shared by the host side (particles.cu) and the kernel
// translation units (particles_f32.cu / particles_f64.cu) of the particle-system path.
#pragma once
#include <cstdint>

#define LSI_COULOMB 138.935456  // kJ nm / (mol e^2), benchmark/integrators/kyle/constants.py:8

enum { TERM_BOND = 0, TERM_ANGLE = 1, TERM_TORSION = 2, TERM_EXCEPTION = 3 };
enum { UNIT_FREE = 0, UNIT_WATER = 1, UNIT_STAR = 2, UNIT_GENERIC = 3 };

constexpr int kMaxSets = 8;    // term sets: [0] = all terms, [1..] = one per distinct force group
constexpr int kMaxSlots = 4;   // force caches (distinct V tokens of a splitting)
constexpr int kMaxInlineOps = 96;

// One bonded term or 1-4 exception, evaluated by one lane; its role gradients go to the
// term-force buffer at slot .. slot + n_roles - 1.
//   bond:      p = {r0, k}                 angle: p = {theta0, k}
//   torsion:   p = {n, cos(phase), sin(phase), k}
//   exception: p = {K qq, sigma, 4 eps}
template <typename real>
struct PKTerm {
    int32_t kind, a0, a1, a2, a3, slot;  // a0..a3: position SLOTS of the atoms
    real p[4];
};

// One unit of the owner-mapped substeps: a free atom, a rigid water (SETTLE), a star of pair
// constraints around a[0] (<= 3 of them), or a generic constraint cluster (a[0] = cluster id).
struct PKUnit {
    int32_t kind, n;
    int32_t a[4];
    double d[3];
};

struct PKParams {
    // ---- system
    int N, nb_method, excl_words, n_units, n_generic;
    double box[3], inv_box[3];
    double cutoff2, krf, crf, restraint_k;
    double switch_r, inv_switch_w;  // LJ switching function over [switch_r, cutoff]; inv_switch_w = 0: none
    int g_nb, g_restr;
    double oh, hh, tol;
    double settle_c[8];       // SETTLE positions: mO/M, mH/M, ra, rb, rc, 1/ra, 1/(2 rc)
    double settle_m[9];       // SETTLE velocities: impulses = M b (row-major)
    const double* mass;       // [N]
    const void* nbp;          // real4 [N]: q sqrt(K), sigma/2, 2 sqrt(eps), 0
    const uint32_t* excl;     // [N][excl_words] in SLOT space (self, exclusions, exceptions)
    int n_lj;                 // slots [0, n_lj) hold the Lennard-Jones-active atoms
    const int32_t* iperm;     // [N] atom -> slot
    const int32_t* perm;      // [N] slot -> atom (LJ-active atoms first)
    const PKUnit* units;      // [n_units], constrained units first
    const int32_t* gc_off;    // generic clusters: constraints [gc_off[c], gc_off[c+1])
    const int32_t* gc_pairs;  // [n][2]
    const double* gc_dist;    // [n]
    const int32_t* gc_atom_off;  // atoms of cluster c: gc_atoms[gc_atom_off[c] .. gc_atom_off[c+1])
    const int32_t* gc_atoms;
    // bonded terms (sorted by force group) and the per-atom gather lists of every term set
    const void* terms;        // PKTerm<real> [n_terms]
    int n_terms, n_tf, n_sets;
    int set_group[kMaxSets], set_lo[kMaxSets], set_hi[kMaxSets];
    const int32_t* gather_off;  // [n_sets][N + 1]
    const int32_t* gather_idx;  // concatenated slot lists; offsets are absolute
    // ---- program (device arrays)
    int n_ops, n_slots;
    int slot_group[kMaxSlots];
    // programs of up to kMaxInlineOps substeps travel in the kernel parameters (constant bank, no allocation,
    // no host synchronisation); longer multi-timestep strings are staged in device memory (ext_* != NULL)
    unsigned char op_kind[kMaxInlineOps];
    unsigned char op_slot[kMaxInlineOps];
    double op_c0[kMaxInlineOps];  // R: h   V: h   O: a
    double op_c1[kMaxInlineOps];  //               O: b
    const unsigned char* ext_kind;
    const unsigned char* ext_slot;
    const double* ext_c0;
    const double* ext_c1;
    // ---- protocol
    uint64_t seed, replica0;
    int64_t n;
    int n_steps, n_legs, marginal, flags, diag;
    int team_lanes;           // warp-team mode: lanes per replica (32, or 16 when n_atoms <= 32)
    double kT, inv_kT;
    const double* x_cache;
    int64_t n_cache;
    const double* x0;
    const double* v0;
    double *W, *heat, *W_direct, *x_final, *v_final, *trace_x, *trace_v, *trace_w;
    // ---- equilibrium sampler (particles_hmc_kernel): input P.x0, output P.x_final / P.v_final
    int hmc_rounds, hmc_steps, hmc_extra;
    double hmc_dt;
    double* hmc_accept;       // [n][2]: fraction of rounds accepted, fraction accepted at the first chance
    // ---- energy / force query (particles_energy_kernel)
    int q_group;
    const double* q_x;
    double* q_energy;
    double* q_forces;
};

struct PKLaunch {
    bool periodic, warp_team, diag;
    int apl;          // atoms per lane in the force loop: 1 or 2
    int threads;      // CTA size
    int teams;        // replicas per CTA (warp-team mode), else 1
    int team_lanes;   // lanes per replica in warp-team mode
    size_t smem;      // dynamic shared memory per CTA
    int64_t n;        // replicas (protocol) or configurations (energy query)
};

size_t lsi_pk_smem_per_team(bool mixed, int N, int n_slots, int n_tf, bool generic, bool warp_team, bool hmc);
int lsi_pk_launch_hmc_f32(const PKParams& P, const PKLaunch& L, void* stream);
int lsi_pk_launch_hmc_f64(const PKParams& P, const PKLaunch& L, void* stream);
int lsi_pk_launch_protocol_f32(const PKParams& P, const PKLaunch& L, void* stream);
int lsi_pk_launch_protocol_f64(const PKParams& P, const PKLaunch& L, void* stream);
int lsi_pk_launch_energy_f32(const PKParams& P, const PKLaunch& L, void* stream);
int lsi_pk_launch_energy_f64(const PKParams& P, const PKLaunch& L, void* stream);

/*
 * lsi.h -- C ABI of the B200 Langevin-splitting ensemble engine (liblsi.so).
 *
 * This is the drop-in boundary for ONE hot path of choderalab/integrator-benchmark:
 * running ensembles of independent replicas through a Langevin splitting string
 * and reducing the shadow work into the near-equilibrium DeltaF_neq estimate.
 * Every entry point names the reference interface it replaces (paths relative
 * to the reference repository).  The reference is pure Python on top of
 * OpenMM / numba; its "FFI" for this path is therefore a ctypes binding, shown
 * in INTEGRATION.md and implemented in integrator-benchmark_b200/_cabi.py.
 *
 * Conventions
 *   - plain C types only; device buffers are raw CUDA device pointers
 *     (e.g. torch.Tensor.data_ptr()); `stream` is a cudaStream_t passed as void*
 *     (NULL = legacy default stream)
 *   - every call returns LSI_OK or a negative code; lsi_last_error() returns a
 *     thread-local message for the last failure on the calling thread
 *   - handles are opaque, owned by the caller, freed with the matching
 *     *_destroy; one handle must not be used from two threads at once
 *   - no call synchronises the stream or the device unless stated.  Programs of more than 96 substeps (particle
 *     systems) / 160 (scalar systems) do not fit the kernel parameters: their tables are copied to device memory
 *     (one blocking copy) the FIRST time lsi_run_protocol sees them on a device and are cached on the lsi_program
 *     handle; later launches of the same program make no allocation, copy or synchronisation
 *   - there is NO CPU fallback: compute calls fail with LSI_ERR_CUDA when no
 *     device is usable
 *   - units are OpenMM's: nm, ps, amu, kJ/mol, K, e   (toy systems: reduced units)
 */
#ifndef LSI_H_
#define LSI_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* 3: lsi_run_protocol_batch added; scalar systems draw the start velocity, the midpoint velocity and the cache index of a
 *    replica from ONE start block (results of version-2 libraries are not reproduced for scalar systems);
 *    LSI_PRECISION_MIXED evaluates the fp32 O noise of particle systems with the special-function unit */
#define LSI_ABI_VERSION 3

/* error codes; the Python shim maps them to the exception types the reference raises */
#define LSI_OK 0
#define LSI_ERR_ASSERT (-1)      /* AssertionError   (langevin.py:110-115, 289-302)          */
#define LSI_ERR_VALUE (-2)       /* ValueError       (langevin.py:119-121)                   */
#define LSI_ERR_ARG (-3)         /* bad argument at this boundary (NULL, negative size ...)  */
#define LSI_ERR_CUDA (-4)        /* CUDA runtime failure or no device                        */
#define LSI_ERR_UNSUPPORTED (-5) /* NotImplementedError (bookkeepers.py:252-253 etc.)        */
#define LSI_ERR_KEY (-6)         /* KeyError         (langevin.py:300-302, no bare "V" entry) */

/* substep kinds of a compiled program */
#define LSI_OP_R 0
#define LSI_OP_V 1
#define LSI_OP_O 2

/* scalar (1-DOF) potentials: low_dimensional_systems.py:52-57,132-136 */
#define LSI_POT_QUARTIC 0     /* U = x^4                      */
#define LSI_POT_DOUBLE_WELL 1 /* U = x^6 + 2 cos(5 (x + 1))   */
#define LSI_POT_HARMONIC 2    /* U = params[0]/2 x^2  (analytic checks) */

/* marginal of the reverse leg: bookkeepers.py:267-270 */
#define LSI_MARGINAL_CONFIGURATION 0
#define LSI_MARGINAL_FULL 1

/* arithmetic */
#define LSI_PRECISION_DOUBLE 0 /* everything fp64 (OpenMM Reference platform, numba toy path)            */
#define LSI_PRECISION_MIXED 1  /* fp64 state and accumulators, fp32 force evaluation (OpenMM "mixed")     */

/* RNG stream tags (counter word 3 = tag << 28 | unit); DESIGN.md "RNG layout" */
#define LSI_TAG_O 0
#define LSI_TAG_V0 1
#define LSI_TAG_VMID 2
#define LSI_TAG_X0 3
#define LSI_TAG_BOOT 4
#define LSI_TAG_EQ 5

typedef struct lsi_program lsi_program;
typedef struct lsi_system lsi_system;

const char* lsi_last_error(void);
int lsi_abi_version(void);
/* number of usable CUDA devices (0 if none); never fails */
int lsi_device_count(void);

/* ------------------------------------------------------------------ splitting compiler
 * Replaces LangevinSplittingIntegrator.__init__ (benchmark/integrators/langevin.py:47-121,
 * 189-220): tokenises `splitting`, counts substeps, detects multi-timestep force groups,
 * applies the reference's checks (same accept/reject set; LSI_ERR_ASSERT / LSI_ERR_VALUE
 * where the reference raises AssertionError / ValueError) and bakes the O coefficients
 * a = exp(-gamma h), b = sqrt(1 - exp(-2 gamma h)), h = dt / max(1, n_O).
 * Deviation: a string whose only V token is a single V<g>
 * (reference emits "dt / 0") is rejected with LSI_ERR_VALUE.
 * Host only: makes no CUDA call. */
int lsi_compile(const char* splitting, double dt_ps, double gamma_per_ps,
                int measure_shadow_work, int measure_heat, int override_splitting_checks,
                lsi_program** out);

/* Replaces ContinuousLangevinSplittingIntegrator.__init__ (langevin.py:234-408):
 * explicit (token, fraction) pairs; zero-length substeps are dropped (:406-408). */
int lsi_compile_continuous(const char* const* tokens, const double* fractions, int n_tokens,
                           double dt_ps, double gamma_per_ps,
                           int measure_shadow_work, int measure_heat, int override_splitting_checks,
                           lsi_program** out);

int lsi_program_num_ops(const lsi_program* p);
int lsi_program_is_mts(const lsi_program* p);
int lsi_program_num_O(const lsi_program* p);
/* op i: kind (LSI_OP_*), force group (-1 = all forces), fraction of dt, and for O the baked a, b */
int lsi_program_get_op(const lsi_program* p, int i, int* kind, int* group, double* fraction,
                       double* a, double* b);
/* CustomIntegrator.getStepSize / setStepSize.  As in the reference changing the
 * step size rescales R and V substeps only; a and b keep their construction-time values. */
double lsi_program_get_step_size(const lsi_program* p);
int lsi_program_set_step_size(lsi_program* p, double dt_ps);
int lsi_program_flags(const lsi_program* p, int* measure_shadow_work, int* measure_heat);
void lsi_program_destroy(lsi_program* p);

/* ------------------------------------------------------------------ systems */
/* 1-DOF toy systems (NumbaBookkeepingSimulator, low_dimensional_systems.py:52-76) */
int lsi_system_create_scalar(int potential_kind, const double params[4], double mass, lsi_system** out);

/* Particle systems.  The description mirrors the OpenMM System objects the reference builds in
 * benchmark/testsystems/{watercluster.py:10-84, waterbox.py:36-48, alanine_dipeptide.py:12-21}:
 * NonbondedForce (Lennard-Jones, Lorentz-Berthelot combination, Coulomb constant 138.935456;
 * NoCutoff, or CutoffPeriodic = minimum image + reaction field with `rf_dielectric`, LJ truncated or
 * switched off over [switch_distance, cutoff]), its exclusions and 1-4 exceptions, HarmonicBondForce 1/2 k (r - r0)^2,
 * HarmonicAngleForce 1/2 k (theta - theta0)^2, PeriodicTorsionForce k (1 + cos(n phi - phase)),
 * the restraint (K/2) |x|^2 on every atom (watercluster.py:70-80), force-group ids per force
 * class ("Nonbonded -> 0, else -> 1", experiments/submission_scripts/4_mts_integrator_splittings.py:36-42),
 * rigid three-site waters (SETTLE) and pair distance constraints (HBonds), tolerance as
 * CustomIntegrator.setConstraintTolerance (langevin.py:202).  Masses are plain numbers, so
 * hydrogen-mass repartitioning (utilities/openmm_utilities.py:223-269) is a different `mass`.
 * All arrays are HOST pointers and are copied; device tables are built on first use. */
#define LSI_NB_NOCUTOFF 0
#define LSI_NB_CUTOFF_PERIODIC 1
typedef struct lsi_particles_desc {
    int32_t n_atoms;
    const double* mass;    /* [n_atoms] amu */
    const double* charge;  /* [n_atoms] e */
    const double* sigma;   /* [n_atoms] nm */
    const double* epsilon; /* [n_atoms] kJ/mol */
    int32_t nonbonded_method;
    double cutoff; /* nm */
    double box[3]; /* orthorhombic box edge lengths, nm */
    double rf_dielectric;
    int32_t n_exclusions;
    const int32_t* exclusions; /* [n][2] */
    int32_t n_exceptions;
    const int32_t* exception_atoms;  /* [n][2] */
    const double* exception_params;  /* [n][3] chargeProd (e^2), sigma, epsilon */
    int32_t n_bonds;
    const int32_t* bond_atoms;  /* [n][2] */
    const double* bond_params;  /* [n][2] r0, k */
    int32_t n_angles;
    const int32_t* angle_atoms; /* [n][3] */
    const double* angle_params; /* [n][2] theta0 (rad), k */
    int32_t n_torsions;
    const int32_t* torsion_atoms; /* [n][4] */
    const double* torsion_params; /* [n][3] periodicity, phase (rad), k */
    double restraint_k;
    int32_t group_nonbonded, group_bonds, group_angles, group_torsions, group_restraint;
    int32_t n_settle;
    const int32_t* settle_atoms; /* [n][3] O, H1, H2 */
    double settle_oh, settle_hh;
    int32_t n_constraints;
    const int32_t* constraint_atoms; /* [n][2] */
    const double* constraint_dist;   /* [n] nm */
    double constraint_tolerance;
    /* Lennard-Jones switching function of NonbondedForce (setUseSwitchingFunction / setSwitchingDistance, which
     * openmmtools' WaterBox -- the reference's tiny_waterbox, benchmark/testsystems/waterbox.py:36-48 -- turns on with
     * switch_width = 1.5 A): for switch_distance < r < cutoff the LJ energy is multiplied by
     * S(x) = 1 - 6 x^5 + 15 x^4 - 10 x^3, x = (r - switch_distance) / (cutoff - switch_distance), so that energy and
     * force go to zero continuously at the cutoff.  <= 0: plain truncation.  CutoffPeriodic only. */
    double switch_distance;
} lsi_particles_desc;

int lsi_system_create_particles(const lsi_particles_desc* desc, lsi_system** out);
/* potential energy (kJ/mol) and forces of force group `group` (-1 = all) for n configurations:
 * x [n][n_atoms][3] -> energy [n], forces [n][n_atoms][3] (either output may be NULL); device
 * pointers.  Replaces Context.getState(getEnergy / getForces) as used by
 * utilities/openmm_utilities.py:10-18. */
int lsi_particles_energy_forces(const lsi_system* sys, int32_t precision, int32_t group, int64_t n,
                                const double* x, double* energy, double* forces, void* stream);

void lsi_system_destroy(lsi_system* s);
int lsi_system_num_dof(const lsi_system* s);

/* ------------------------------------------------------------------ protocol
 * Replaces NumbaNonequilibriumSimulator.collect_protocol_samples
 * (low_dimensional_systems.py:160-184) / NonequilibriumSimulator.collect_protocol_samples
 * and accumulate_shadow_work (bookkeepers.py:184-295) for a whole ensemble in one launch.
 *
 * Per replica r (global id replica0 + r, which alone keys its RNG streams):
 *   leg 0: x0 = x0[r] if given, else x_cache[index]; v0 = v0[r] if given, else sqrt(kT/m) * N(0,1).
 *          Particle systems: index from lane 0 of the block (draw 0, TAG_X0), velocities from TAG_V0 (unit = atom).
 *          Scalar systems: ONE start block (draw 0, TAG_V0, unit 0) per replica: words 0, 1 -> the Box-Muller pair
 *          (start-velocity normal, midpoint-velocity normal), word 2 -> index = (word * n_cache) >> 32
 *   leg l>0: x continues; v is redrawn (configuration) or kept (full).  Particle systems redraw from TAG_VMID
 *          (draw l - 1, unit = atom); scalar systems take the second normal of the start block for leg 1 and normal l - 2 of
 *          the TAG_V0 stream that begins at draw 1 for legs l >= 2
 *   each leg applies the program n_steps times and yields
 *          W = ((E_end - E_start) - (heat_end - heat_start)) / kT
 * Particle systems: states are [replica][atom][3] (x_cache [n_cache][atom][3]); before each leg
 * positions then velocities are projected onto the constraints (bookkeepers.py:206-208).
 * Outputs are leg-major so that warps write coalesced: W[leg * n_replicas + r].
 * Optional outputs may be NULL.  Traces hold n_steps + 1 entries per leg (entry 0 is the
 * leg start): trace[(leg * (n_steps + 1) + s) * n_replicas + r]; trace_w is cumulative
 * work in kT (numba_integrators.py:59-61).
 * `moments` (device, 8 doubles, optional) receives, over replicas whose works are all
 * finite: {count, sum wF, sum wR, sum wF^2, sum wR^2, sum wF wR, count_nonfinite, 0};
 * it needs n_legs == 2.  Scratch for the block partials is taken from `workspace`
 * (device, at least lsi_protocol_workspace_bytes(n_replicas) bytes).  Scalar systems fuse the block
 * partials into the protocol kernel; particle systems reduce W with the kernel of lsi_reduce_moments
 * on the same stream. */
/* flags: positions pass through a float32 frame at the midpoint, as the reference's mdtraj
 * round trip does (bookkeepers.py:266,298-305).  Particle systems only. */
#define LSI_FLAG_ROUND_MIDPOINT_X_FP32 1
/* scalar systems: traces in the reference's own layout (low_dimensional_systems.py:160-184 returns arrays of shape
 * (n, n_steps) and (n, n_steps, 2)): trace_w is [leg][replica][step], trace_x holds (x, v) pairs
 * [leg][replica][step][2] and trace_v is not used.  Default layout: [leg][step][replica] for each of the three. */
#define LSI_FLAG_TRACE_REPLICA_MAJOR 2
/* scalar systems: default [leg][step][replica] order, but trace_x holds (x, v) pairs [leg][step][replica][2]
 * (trace_v not used): coalesced stores, and a transposed VIEW of it has the reference's shape (n, n_steps, 2). */
#define LSI_FLAG_TRACE_XV_PAIRS 4
typedef struct lsi_protocol_args {
    uint64_t seed;
    uint64_t replica0;
    int64_t n_replicas;
    int32_t n_steps;
    int32_t n_legs;
    int32_t marginal;
    int32_t precision;
    int32_t flags; /* LSI_FLAG_* */
    int32_t reserved;
    double kT;
    /* inputs (device) */
    const double* x_cache;
    int64_t n_cache;
    const double* x0;
    const double* v0;
    /* outputs (device) */
    double* W;
    double* heat;
    double* W_direct; /* needs a program compiled with measure_shadow_work */
    double* x_final;
    double* v_final;
    double* trace_x;
    double* trace_v;
    double* trace_w;
    double* moments;
    void* workspace;
    int64_t workspace_bytes;
} lsi_protocol_args;

int64_t lsi_protocol_workspace_bytes(int64_t n_replicas);
int lsi_run_protocol(const lsi_system* sys, const lsi_program* prog, const lsi_protocol_args* args,
                     void* stream);

/* The same for several programs on one system: works, heats and final states are identical, bit for bit, to calling
 * lsi_run_protocol(sys, progs[p], &args[p]) for p = 0 .. n_progs - 1 in that order on `stream` (moment sums hold the same
 * terms; their summation order, hence their last bits, may differ).  What the reference does when it compares schemes
 * (experiments/toy/quartic_kl_validation.py:56-70, tests/test_shadow_work.py:47-55: a loop over splittings around the same
 * equilibrium cache).  Scalar systems run the whole batch as ONE kernel launch when the programs are distinct members of
 * one pre-compiled family -- the six 5-letter palindromes over {O, R, V}, or the three 6-letter strings the reference's
 * numba closures perform -- and all args share the ensemble (n_replicas, n_steps, n_legs, marginal, precision, kT, seed,
 * replica0, x_cache / x0 / v0) and ask for no traces: tables are staged once and the GPU never drains between programs.
 * When the batch is a WHOLE family whose programs were compiled with the same time step and collision rate, every thread
 * carries one replica through all programs at once: programs that share an ensemble share the replica's start state and its
 * O-noise stream (the RNG layout keys noise by seed, replica, leg and draw, not by program), so the start block, the cache
 * gather and every Philox block / Box-Muller pair are computed once per replica instead of once per program -- common random
 * numbers across the splittings, at about half the arithmetic of separate launches.
 * Every args[p] needs its own outputs and its own workspace.  Anything else falls back to one launch per program. */
int lsi_run_protocol_batch(const lsi_system* sys, const lsi_program* const* progs, int32_t n_progs,
                           const lsi_protocol_args* args, void* stream);

/* Equilibrium configurations for a scalar system: n independent random-walk Metropolis chains
 * (proposal N(0,1), acceptance q(x')/q(x) > u as in metropolis_hastings_factory,
 * numba_integrators.py:229-252), each started from N(0,1) and run for n_burn steps; the final
 * states are written to x_out (device, n doubles).  Replaces the single 10^6-step chain the
 * reference caches in {name}_samples.npy (low_dimensional_systems.py:95-98) by independent
 * samples.  Chain c uses Philox replica id replica0 + c, tag LSI_TAG_EQ, draw = step:
 * lane 0/1 -> proposal normal (Box-Muller cos branch), lane 2 -> acceptance uniform. */
int lsi_sample_equilibrium_scalar(const lsi_system* sys, double kT, int64_t n, int32_t n_burn,
                                  uint64_t seed, uint64_t replica0, double* x_out, void* stream);

/* Equilibrium configurations for a particle system by extra-chance HMC: the sampler behind the reference's
 * {name}_samples.h5 caches (EquilibriumSimulator.collect_equilibrium_samples, benchmark/testsystems/
 * bookkeepers.py:66-113, with XCGHMCIntegrator(steps_per_hmc = 10, extra_chances = 15, collision_rate = None),
 * benchmark/integrators/kyle/xchmc.py:84-116).  n independent chains; chain c starts from x_in[c] (projected onto
 * the constraints) and runs n_rounds rounds: fresh Maxwell-Boltzmann velocities, up to 1 + extra_chances
 * proposals of steps_per_hmc velocity-Verlet steps, Metropolis test with one uniform per round, restore on
 * failure.  x_in / x_out [n][n_atoms][3] (may alias), v_out optional, accept_out optional [n][2] =
 * {fraction of rounds accepted, fraction accepted at the first chance}; device pointers.
 * Streams: velocities Philox(replica0 + c, draw = 2 round, LSI_TAG_EQ, unit = atom), uniform
 * Philox(replica0 + c, 2 round + 1, LSI_TAG_EQ, 0) lane 0. */
int lsi_sample_equilibrium_particles(const lsi_system* sys, int32_t precision, double kT, int64_t n, int32_t n_rounds,
                                     int32_t steps_per_hmc, int32_t extra_chances, double dt_ps, uint64_t seed,
                                     uint64_t replica0, const double* x_in, double* x_out, double* v_out,
                                     double* accept_out, void* stream);

/* ------------------------------------------------------------------ estimators
 * lsi_reduce_moments: the six sums estimate_nonequilibrium_free_energy needs
 * (benchmark/evaluation/analysis.py:8-17) from device arrays W_F, W_R of length n; same
 * layout of `moments` as above.  The host-side formula (np.var ddof=0, np.cov ddof=1) is
 * lsi_neq_free_energy. */
int lsi_reduce_moments(const double* W_F, const double* W_R, int64_t n, double* moments,
                       void* workspace, int64_t workspace_bytes, void* stream);
/* host arithmetic on a (possibly all-reduced) moment vector; N_eff <= 0 means "len(W)" */
int lsi_neq_free_energy(const double moments[8], double N_eff, double* DeltaF_neq, double* squared_uncertainty);

/* Row-wise log-mean-exp(-w) with the reference's stderr heuristic
 * (compare_near_eq_and_exact.py:62-97): for each row i over w[row_offsets[i] : row_offsets[i+1]]
 *   estimate[i] = log(mean(exp(-w)))       (computed stably; equal in exact arithmetic)
 *   stdev[i]    = std(exp(-w)) / (mean(exp(-w)) sqrt(n))     (ddof = 0)
 * All pointers are device pointers; row_offsets has n_rows + 1 entries. */
int lsi_reduce_logmeanexp(const double* w, const int64_t* row_offsets, int64_t n_rows,
                          double* estimate, double* stdev, void* stream);

/* Two-level bootstrap of the nested estimator (compare_near_eq_and_exact.py:293-320):
 * for each of n_boot replicates resample rows, then each row's columns, uniformly with
 * replacement (Philox, TAG_BOOT), and write mean over rows of log-mean-exp(-w) to out[b]. */
int lsi_bootstrap_nested(const double* w, const int64_t* row_offsets, int64_t n_rows, int32_t n_boot,
                         uint64_t seed, double* out, void* stream);

/* ------------------------------------------------------------------ measurement helper
 * FMA-pipe microbenchmark (no reference counterpart): blocks x 256 threads, each running
 * iters x 8 independent fused multiply-adds in fp64 (LSI_PRECISION_DOUBLE) or fp32;
 * flops = blocks * 256 * iters * 16.  bench.py times it with CUDA events to obtain the
 * measured non-tensor FMA peak that the FMA-bound kernels are held against.
 * precision = 10 + k (k = 0..4): dispatch test, each iteration runs the 8 DFMAs interleaved with 8 k independent
 * 32-bit integer multiply-adds; the time against k shows whether integer work hides under fp64 instructions
 * (scripts/microbench_dispatch.py, profiles/README.md). */
int lsi_microbench_fma(int precision, int64_t iters, int blocks, void* sink, void* stream);

/* Test hook (no reference counterpart): evaluates the lean fp64 functions the toy kernels use for Box-Muller noise
 * and the double-well force (csrc/fastmath64.cuh) element-wise on x[n]: which = 0 log, 1 sincos(pi x) for 0 <= x < 4,
 * 2 sincos(x), 3 sqrt, 4 reciprocal, 5 the two uniforms made from the 32-bit word x, 6 lean minus library Box-Muller
 * normals made from the word x; the table-driven versions the kernels run: 7 -2 log(x), 8 sincos(2 pi (x + 1/2) / 2^32)
 * for the word x, 9 table-driven minus library Box-Muller normals, 10 sin(x), 11 five-instruction sqrt.
 * out1 may be NULL.  Device pointers. */
int lsi_test_fastmath(int which, int64_t n, const double* x, double* out0, double* out1, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* LSI_H_ */

// Host side of the particle-system path: system tables, launch configuration, C-ABI entry points.
// The kernels live in particles_kernel.cuh (instantiated by particles_f32.cu / particles_f64.cu).
#include <cuda_runtime.h>

#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <map>
#include <memory>
#include <mutex>
#include <set>

#include "lsi_internal.h"
#include "particles_shared.h"

namespace {

struct HostTerm {
    int kind, group, n_roles, slot;
    int a[4];
    double p[4];
};

struct DevTables {
    double* mass = nullptr;
    void* nbp_f = nullptr;
    void* nbp_d = nullptr;
    uint32_t* excl = nullptr;
    int32_t *perm = nullptr, *iperm = nullptr;
    PKUnit* units = nullptr;
    int32_t *gc_off = nullptr, *gc_pairs = nullptr, *gc_atom_off = nullptr, *gc_atoms = nullptr;
    double* gc_dist = nullptr;
    void* terms_f = nullptr;
    void* terms_d = nullptr;
    int32_t *gather_off = nullptr, *gather_idx = nullptr;
};

}  // namespace

struct lsi_particles {
    int n_atoms = 0;
    std::vector<double> mass, charge, sigma, epsilon;
    int nb_method = 0;
    double cutoff = 0, box[3] = {0, 0, 0}, rf_dielectric = 78.3;
    double restraint_k = 0;
    double switch_distance = 0;
    int g_nb = 0, g_bond = 0, g_angle = 0, g_tors = 0, g_restr = 0;
    double oh = 0, hh = 0, tol = 1e-8;
    double settle_c[8] = {0}, settle_m[9] = {0};
    int excl_words = 1;
    std::vector<uint32_t> excl, excl_slots;
    int n_lj = 0;
    std::vector<int32_t> perm, iperm;
    std::vector<HostTerm> terms;  // sorted by force group
    int n_tf = 0;
    int n_sets = 1;
    int set_group[kMaxSets], set_lo[kMaxSets], set_hi[kMaxSets];
    std::vector<int32_t> gather_off, gather_idx;
    std::vector<PKUnit> units;
    int n_generic = 0;
    std::vector<int32_t> gc_off, gc_pairs, gc_atom_off, gc_atoms;
    std::vector<double> gc_dist;
    std::mutex mu;
    std::map<int, DevTables> dev;  // per CUDA device
};

namespace {

template <typename T>
cudaError_t upload(T** dst, const std::vector<T>& src)
{
    *dst = nullptr;
    if (src.empty()) return cudaSuccess;
    cudaError_t e = cudaMalloc((void**)dst, src.size() * sizeof(T));
    if (e != cudaSuccess) return e;
    return cudaMemcpy(*dst, src.data(), src.size() * sizeof(T), cudaMemcpyHostToDevice);
}

template <typename real>
std::vector<PKTerm<real>> pack_terms(const std::vector<HostTerm>& terms, const std::vector<int32_t>& iperm)
{
    std::vector<PKTerm<real>> out(terms.size());
    for (size_t k = 0; k < terms.size(); ++k) {
        const HostTerm& t = terms[k];
        PKTerm<real>& o = out[k];
        o.kind = t.kind; o.slot = t.slot;
        o.a0 = iperm[t.a[0]]; o.a1 = iperm[t.a[1]]; o.a2 = iperm[t.a[2]]; o.a3 = iperm[t.a[3]];  // position slots
        for (int q = 0; q < 4; ++q) o.p[q] = (real)t.p[q];
    }
    return out;
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
    LSI_CUDA_CHECK(upload(&t.excl, p->excl_slots));
    LSI_CUDA_CHECK(upload(&t.iperm, p->iperm));
    LSI_CUDA_CHECK(upload(&t.perm, p->perm));
    LSI_CUDA_CHECK(upload(&t.units, p->units));
    LSI_CUDA_CHECK(upload(&t.gc_off, p->gc_off));
    LSI_CUDA_CHECK(upload(&t.gc_pairs, p->gc_pairs));
    LSI_CUDA_CHECK(upload(&t.gc_atom_off, p->gc_atom_off));
    LSI_CUDA_CHECK(upload(&t.gc_atoms, p->gc_atoms));
    LSI_CUDA_CHECK(upload(&t.gc_dist, p->gc_dist));
    {
        std::vector<PKTerm<float>> tf = pack_terms<float>(p->terms, p->iperm);
        std::vector<PKTerm<double>> td = pack_terms<double>(p->terms, p->iperm);
        PKTerm<float>* pf = nullptr;
        PKTerm<double>* pd = nullptr;
        LSI_CUDA_CHECK(upload(&pf, tf));
        LSI_CUDA_CHECK(upload(&pd, td));
        t.terms_f = pf; t.terms_d = pd;
    }
    LSI_CUDA_CHECK(upload(&t.gather_off, p->gather_off));
    LSI_CUDA_CHECK(upload(&t.gather_idx, p->gather_idx));
    p->dev[dev] = t;
    *out = &p->dev[dev];
    return LSI_OK;
}

void fill_system_params(const lsi_particles* p, const DevTables* t, bool mixed, PKParams& P)
{
    P.N = p->n_atoms;
    P.nb_method = p->nb_method;
    P.excl_words = p->excl_words;
    P.n_units = (int)p->units.size();
    P.n_generic = p->n_generic;
    for (int c = 0; c < 3; ++c) {
        P.box[c] = p->box[c];
        P.inv_box[c] = p->box[c] > 0 ? 1.0 / p->box[c] : 0.0;
    }
    const double rc = p->cutoff, er = p->rf_dielectric;
    const bool periodic = p->nb_method == LSI_NB_CUTOFF_PERIODIC;
    P.cutoff2 = rc * rc;
    P.krf = periodic ? (1.0 / (rc * rc * rc)) * (er - 1.0) / (2.0 * er + 1.0) : 0.0;
    P.crf = periodic ? (1.0 / rc) * (3.0 * er) / (2.0 * er + 1.0) : 0.0;
    P.restraint_k = p->restraint_k;
    const bool sw = periodic && p->switch_distance > 0.0 && p->switch_distance < rc;
    P.switch_r = sw ? p->switch_distance : 0.0;
    P.inv_switch_w = sw ? 1.0 / (rc - p->switch_distance) : 0.0;
    P.g_nb = p->g_nb; P.g_restr = p->g_restr;
    P.oh = p->oh; P.hh = p->hh; P.tol = p->tol;
    for (int k = 0; k < 8; ++k) P.settle_c[k] = p->settle_c[k];
    for (int k = 0; k < 9; ++k) P.settle_m[k] = p->settle_m[k];
    P.mass = t->mass;
    P.nbp = mixed ? t->nbp_f : t->nbp_d;
    P.excl = t->excl;
    P.n_lj = p->n_lj; P.iperm = t->iperm;
    P.perm = t->perm;
    P.units = t->units;
    P.gc_off = t->gc_off; P.gc_pairs = t->gc_pairs; P.gc_dist = t->gc_dist;
    P.gc_atom_off = t->gc_atom_off; P.gc_atoms = t->gc_atoms;
    P.terms = mixed ? t->terms_f : t->terms_d;
    P.n_terms = (int)p->terms.size(); P.n_tf = p->n_tf; P.n_sets = p->n_sets;
    for (int q = 0; q < kMaxSets; ++q) { P.set_group[q] = p->set_group[q]; P.set_lo[q] = p->set_lo[q]; P.set_hi[q] = p->set_hi[q]; }
    P.gather_off = t->gather_off; P.gather_idx = t->gather_idx;
}

int env_int(const char* name, int dflt)
{
    const char* v = std::getenv(name);
    return v && *v ? std::atoi(v) : dflt;
}

// team shape: one warp per replica up to 64 atoms (several replicas per CTA), else one CTA per replica
int configure(const lsi_particles* p, bool mixed, int n_slots, bool generic_smem, int64_t n, bool diag, PKLaunch& L, bool hmc = false)
{
    const int N = p->n_atoms;
    L.periodic = p->nb_method == LSI_NB_CUTOFF_PERIODIC;
    L.diag = diag;
    L.n = n;
    L.warp_team = N <= 64 && env_int("LSI_PK_WARP_TEAM", 1) != 0;
    const size_t per_team = lsi_pk_smem_per_team(mixed, N, n_slots, p->n_tf, generic_smem, L.warp_team, hmc);
    L.team_lanes = 32;
    if (L.warp_team) {
        // up to 32 atoms: two replicas per warp (half-warp teams, 2 atoms per lane) keep more lanes busy
        if (N <= 32 && env_int("LSI_PK_HALF_WARP", 1) != 0) L.team_lanes = 16;
        L.apl = N <= L.team_lanes ? 1 : 2;
        const int warps = std::max(1, std::min(4, env_int("LSI_PK_TEAMS", 4)));
        L.teams = warps * (32 / L.team_lanes);
        while (L.teams > 32 / L.team_lanes && per_team * L.teams > 200 * 1024) L.teams -= 32 / L.team_lanes;
        L.threads = L.team_lanes * L.teams;
        L.smem = per_team * L.teams;
    } else {
        int apl = env_int("LSI_PK_APL", 2);
        if (apl != 1 && apl != 2) apl = 2;
        if ((N + apl - 1) / apl > 384) apl = 2;
        L.apl = apl;
        L.teams = 1;
        L.threads = std::max(32, (((N + apl - 1) / apl + 31) / 32) * 32);
        L.smem = per_team;
    }
    if (L.smem > 227 * 1024 || L.threads > 384)
        return lsi_set_error(LSI_ERR_UNSUPPORTED, "system too large for the shared-memory-resident kernel (%zu bytes, %d threads)",
                             L.smem, L.threads);
    return LSI_OK;
}

}  // namespace

void lsi_particles_free(lsi_particles* p)
{
    if (!p) return;
    for (auto& kv : p->dev) {
        DevTables& t = kv.second;
        cudaFree(t.mass); cudaFree(t.nbp_f); cudaFree(t.nbp_d); cudaFree(t.excl); cudaFree(t.perm); cudaFree(t.iperm); cudaFree(t.units);
        cudaFree(t.gc_off); cudaFree(t.gc_pairs); cudaFree(t.gc_atom_off); cudaFree(t.gc_atoms); cudaFree(t.gc_dist);
        cudaFree(t.terms_f); cudaFree(t.terms_d); cudaFree(t.gather_off); cudaFree(t.gather_idx);
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
    if (N <= 0 || N > 768) return lsi_set_error(LSI_ERR_ARG, "n_atoms must be in 1..768 (one CTA of at most 384 threads holds a replica)");
    if (!d->mass) return lsi_set_error(LSI_ERR_ARG, "mass is required");
    auto bad_atom = [N](int a) { return a < 0 || a >= N; };
    std::unique_ptr<lsi_particles> p(new lsi_particles());
    p->n_atoms = N;
    p->mass.assign(d->mass, d->mass + N);
    for (double m : p->mass)
        if (!(m > 0.0)) return lsi_set_error(LSI_ERR_ARG, "masses must be positive (massless sites are not supported)");
    p->charge.assign(N, 0.0); p->sigma.assign(N, 1.0); p->epsilon.assign(N, 0.0);
    if (d->charge) p->charge.assign(d->charge, d->charge + N);
    if (d->sigma) p->sigma.assign(d->sigma, d->sigma + N);
    if (d->epsilon) p->epsilon.assign(d->epsilon, d->epsilon + N);
    p->nb_method = d->nonbonded_method;
    if (p->nb_method != LSI_NB_NOCUTOFF && p->nb_method != LSI_NB_CUTOFF_PERIODIC)
        return lsi_set_error(LSI_ERR_UNSUPPORTED, "nonbonded_method %d (only NoCutoff and CutoffPeriodic are built)", d->nonbonded_method);
    p->cutoff = d->cutoff;
    for (int c = 0; c < 3; ++c) p->box[c] = d->box[c];
    if (p->nb_method == LSI_NB_CUTOFF_PERIODIC)
        for (int c = 0; c < 3; ++c)
            if (!(p->box[c] >= 2.0 * p->cutoff) || !(p->cutoff > 0))
                return lsi_set_error(LSI_ERR_ARG, "periodic box edge must be at least twice the cutoff");
    p->rf_dielectric = d->rf_dielectric > 0 ? d->rf_dielectric : 78.3;
    p->restraint_k = d->restraint_k;
    p->switch_distance = d->switch_distance;
    if (p->switch_distance > 0.0 && !(p->nb_method == LSI_NB_CUTOFF_PERIODIC && p->switch_distance < p->cutoff))
        return lsi_set_error(LSI_ERR_ARG, "switch_distance needs CutoffPeriodic and must be below the cutoff");
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
        if (bad_atom(i) || bad_atom(j)) return lsi_set_error(LSI_ERR_ARG, "exclusion %d: atom index out of range", k);
        set_excl(i, j);
    }
    // force-loop order: Lennard-Jones-active atoms first (stable), so that whole warps skip the LJ branch
    for (int pass = 0; pass < 2; ++pass)
        for (int i = 0; i < N; ++i)
            if ((p->epsilon[i] != 0.0) == (pass == 0)) p->perm.push_back(i);

    // ---- bonded terms and exceptions, sorted by force group
    std::vector<HostTerm> terms;
    auto add_term = [&](int kind, int group, int n_at, const int32_t* at, double p0, double p1, double p2, double p3) -> bool {
        HostTerm t;
        t.kind = kind; t.group = group; t.n_roles = n_at; t.slot = 0;
        for (int q = 0; q < 4; ++q) {
            t.a[q] = q < n_at ? at[q] : 0;
            if (q < n_at && bad_atom(at[q])) return false;
        }
        t.p[0] = p0; t.p[1] = p1; t.p[2] = p2; t.p[3] = p3;
        terms.push_back(t);
        return true;
    };
    bool ok = true;
    for (int k = 0; k < d->n_exceptions && ok; ++k) {
        const double* e = d->exception_params + 3 * k;
        ok = add_term(TERM_EXCEPTION, p->g_nb, 2, d->exception_atoms + 2 * k, LSI_COULOMB * e[0], e[1], 4.0 * e[2], 0.0);
        if (ok) set_excl(d->exception_atoms[2 * k], d->exception_atoms[2 * k + 1]);
    }
    for (int k = 0; k < d->n_bonds && ok; ++k)
        ok = add_term(TERM_BOND, p->g_bond, 2, d->bond_atoms + 2 * k, d->bond_params[2 * k], d->bond_params[2 * k + 1], 0.0, 0.0);
    for (int k = 0; k < d->n_angles && ok; ++k)
        ok = add_term(TERM_ANGLE, p->g_angle, 3, d->angle_atoms + 3 * k, d->angle_params[2 * k], d->angle_params[2 * k + 1], 0.0, 0.0);
    for (int k = 0; k < d->n_torsions && ok; ++k) {
        const double* t = d->torsion_params + 3 * k;
        ok = add_term(TERM_TORSION, p->g_tors, 4, d->torsion_atoms + 4 * k, std::floor(t[0] + 0.5), std::cos(t[1]), std::sin(t[1]), t[2]);
    }
    if (!ok) return lsi_set_error(LSI_ERR_ARG, "bonded term / exception: atom index out of range");
    std::stable_sort(terms.begin(), terms.end(), [](const HostTerm& a, const HostTerm& b) { return a.group < b.group; });
    int slot = 0;
    for (HostTerm& t : terms) { t.slot = slot; slot += t.n_roles; }
    p->n_tf = slot;
    p->terms = terms;
    // term sets: [0] all, then one per distinct group; per-atom gather lists (fixed order -> deterministic sums)
    p->n_sets = 1;
    p->set_group[0] = -1; p->set_lo[0] = 0; p->set_hi[0] = (int)terms.size();
    for (int q = 1; q < kMaxSets; ++q) { p->set_group[q] = -2; p->set_lo[q] = p->set_hi[q] = 0; }
    for (size_t k = 0; k < terms.size(); ++k) {
        if (k == 0 || terms[k].group != terms[k - 1].group) {
            if (p->n_sets == kMaxSets) return lsi_set_error(LSI_ERR_UNSUPPORTED, "too many distinct force groups among bonded terms");
            p->set_group[p->n_sets] = terms[k].group;
            p->set_lo[p->n_sets] = (int)k;
            ++p->n_sets;
        }
        p->set_hi[p->n_sets - 1] = (int)k + 1;
    }
    p->gather_off.assign((size_t)p->n_sets * (N + 1), 0);
    for (int s = 0; s < p->n_sets; ++s) {
        std::vector<std::vector<int32_t>> per_atom(N);
        for (int k = p->set_lo[s]; k < p->set_hi[s]; ++k)
            for (int r = 0; r < terms[k].n_roles; ++r) per_atom[terms[k].a[r]].push_back(terms[k].slot + r);
        for (int i = 0; i < N; ++i) {
            p->gather_off[(size_t)s * (N + 1) + i] = (int32_t)p->gather_idx.size();
            p->gather_idx.insert(p->gather_idx.end(), per_atom[i].begin(), per_atom[i].end());
        }
        p->gather_off[(size_t)s * (N + 1) + N] = (int32_t)p->gather_idx.size();
    }
    if (p->gather_idx.empty()) p->gather_idx.push_back(0);
    if (p->terms.empty()) p->terms.push_back(HostTerm{});  // keeps the device pointers valid
    if (p->terms.size() == 1 && terms.empty()) p->set_hi[0] = 0;

    // slot space: positions are staged with the LJ-active atoms first; exclusion bits follow
    p->iperm.assign(N, 0);
    for (int k = 0; k < N; ++k) p->iperm[p->perm[k]] = k;
    for (int i = 0; i < N; ++i) p->n_lj += p->epsilon[i] != 0.0;
    p->excl_slots.assign((size_t)N * p->excl_words, 0u);
    for (int i = 0; i < N; ++i)
        for (int j = 0; j < N; ++j)
            if ((p->excl[(size_t)i * p->excl_words + (j >> 5)] >> (j & 31)) & 1u) {
                const int ki = p->iperm[i], kj = p->iperm[j];
                p->excl_slots[(size_t)ki * p->excl_words + (kj >> 5)] |= 1u << (kj & 31);
            }

    // ---- units of the owner-mapped substeps: waters, constraint stars, generic clusters, free atoms
    std::vector<int> owner(N, -1);
    for (int k = 0; k < d->n_settle; ++k) {
        PKUnit u;
        std::memset(&u, 0, sizeof(u));
        u.kind = UNIT_WATER; u.n = 3;
        for (int q = 0; q < 3; ++q) {
            const int a = d->settle_atoms[3 * k + q];
            if (bad_atom(a) || owner[a] >= 0) return lsi_set_error(LSI_ERR_ARG, "settle %d: atom index out of range or used twice", k);
            u.a[q] = a;
            owner[a] = (int)p->units.size();
        }
        u.a[3] = -1;
        if (p->mass[u.a[1]] != p->mass[u.a[2]]) return lsi_set_error(LSI_ERR_ARG, "settle %d: the two hydrogens must have equal mass", k);
        p->units.push_back(u);
    }
    if (d->n_settle > 0 && !(p->oh > 0 && p->hh > 0 && p->hh < 2 * p->oh))
        return lsi_set_error(LSI_ERR_ARG, "settle distances must satisfy 0 < hh < 2 oh");
    if (d->n_settle > 0) {
        // constants of the rigid geometry; every water must share them (one set per system)
        const double mO = p->mass[p->units[0].a[0]], mH = p->mass[p->units[0].a[1]];
        for (const PKUnit& u : p->units)
            if (p->mass[u.a[0]] != mO || p->mass[u.a[1]] != mH)
                return lsi_set_error(LSI_ERR_UNSUPPORTED, "all SETTLE waters must have the same masses");
        const double M = mO + 2.0 * mH, rc = 0.5 * p->hh, hgt = std::sqrt(p->oh * p->oh - rc * rc);
        const double ra = hgt * 2.0 * mH / M, rb = hgt - ra;
        const double sc[8] = {mO / M, mH / M, ra, rb, rc, 1.0 / ra, 1.0 / (2.0 * rc), 0.0};
        for (int k = 0; k < 8; ++k) p->settle_c[k] = sc[k];
        // C lambda = -b for edges 0: O->H1, 1: H1->H2, 2: H2->O (e0 + e1 + e2 = 0 fixes the dot products)
        const double im[3] = {1.0 / mO, 1.0 / mH, 1.0 / mH};
        const double d2[3] = {p->oh * p->oh, p->hh * p->hh, p->oh * p->oh};
        const double e01 = 0.5 * (d2[2] - d2[0] - d2[1]), e02 = 0.5 * (d2[1] - d2[0] - d2[2]), e12 = 0.5 * (d2[0] - d2[1] - d2[2]);
        const double C[3][3] = {{d2[0] * (im[1] + im[0]), -im[1] * e01, -im[0] * e02},
                                {-im[1] * e01, d2[1] * (im[2] + im[1]), -im[2] * e12},
                                {-im[0] * e02, -im[2] * e12, d2[2] * (im[0] + im[2])}};
        const double det = C[0][0] * (C[1][1] * C[2][2] - C[1][2] * C[2][1]) - C[0][1] * (C[1][0] * C[2][2] - C[1][2] * C[2][0]) +
                           C[0][2] * (C[1][0] * C[2][1] - C[1][1] * C[2][0]);
        for (int i = 0; i < 3; ++i)
            for (int j = 0; j < 3; ++j) {
                const int i1 = (i + 1) % 3, i2 = (i + 2) % 3, j1 = (j + 1) % 3, j2 = (j + 2) % 3;
                // inverse of a symmetric matrix: cofactor(j, i) / det
                p->settle_m[3 * i + j] = -(C[j1][i1] * C[j2][i2] - C[j1][i2] * C[j2][i1]) / det;
            }
    }
    {
        // connected components of the pair-constraint graph (union-find), constraints kept in input order
        std::vector<int> parent(N);
        for (int i = 0; i < N; ++i) parent[i] = i;
        auto find = [&](int a) { while (parent[a] != a) { parent[a] = parent[parent[a]]; a = parent[a]; } return a; };
        for (int k = 0; k < d->n_constraints; ++k) {
            const int i = d->constraint_atoms[2 * k], j = d->constraint_atoms[2 * k + 1];
            if (bad_atom(i) || bad_atom(j) || i == j || !(d->constraint_dist[k] > 0))
                return lsi_set_error(LSI_ERR_ARG, "constraint %d: bad atoms or distance", k);
            if (owner[i] >= 0 || owner[j] >= 0) return lsi_set_error(LSI_ERR_ARG, "constraint %d touches a SETTLE water", k);
            parent[find(i)] = find(j);
        }
        std::map<int, std::vector<int>> comps;
        for (int k = 0; k < d->n_constraints; ++k) comps[find(d->constraint_atoms[2 * k])].push_back(k);
        std::vector<std::pair<int, std::vector<int>>> ordered(comps.begin(), comps.end());
        std::sort(ordered.begin(), ordered.end(), [](const auto& a, const auto& b) { return a.second[0] < b.second[0]; });
        std::vector<PKUnit> stars, generics;
        p->gc_off.push_back(0);
        p->gc_atom_off.push_back(0);
        for (auto& kv : ordered) {
            const std::vector<int>& ks = kv.second;
            const int m = (int)ks.size();
            std::set<int> atoms;
            for (int k : ks) { atoms.insert(d->constraint_atoms[2 * k]); atoms.insert(d->constraint_atoms[2 * k + 1]); }
            int center = -1;
            if (m <= 3 && (int)atoms.size() == m + 1) {
                for (int cand : {d->constraint_atoms[2 * ks[0]], d->constraint_atoms[2 * ks[0] + 1]}) {
                    bool all = true;
                    for (int k : ks) all = all && (d->constraint_atoms[2 * k] == cand || d->constraint_atoms[2 * k + 1] == cand);
                    if (all && center < 0) center = cand;
                }
            }
            PKUnit u;
            std::memset(&u, 0, sizeof(u));
            if (center >= 0) {
                u.kind = UNIT_STAR; u.n = m + 1;
                u.a[0] = center;
                for (int q = 0; q < m; ++q) {
                    const int i = d->constraint_atoms[2 * ks[q]], j = d->constraint_atoms[2 * ks[q] + 1];
                    u.a[q + 1] = i == center ? j : i;
                    u.d[q] = d->constraint_dist[ks[q]];
                }
                for (int q = m + 1; q < 4; ++q) u.a[q] = -1;
                stars.push_back(u);
            } else {
                u.kind = UNIT_GENERIC; u.n = 0;
                u.a[0] = p->n_generic++;
                for (int k : ks) {
                    p->gc_pairs.push_back(d->constraint_atoms[2 * k]);
                    p->gc_pairs.push_back(d->constraint_atoms[2 * k + 1]);
                    p->gc_dist.push_back(d->constraint_dist[k]);
                }
                p->gc_off.push_back((int32_t)p->gc_dist.size());
                p->gc_atoms.insert(p->gc_atoms.end(), atoms.begin(), atoms.end());
                p->gc_atom_off.push_back((int32_t)p->gc_atoms.size());
                generics.push_back(u);
            }
            for (int a : atoms) owner[a] = 1;
        }
        p->units.insert(p->units.end(), stars.begin(), stars.end());
        p->units.insert(p->units.end(), generics.begin(), generics.end());
    }
    for (int i = 0; i < N; ++i)
        if (owner[i] < 0) {
            PKUnit u;
            std::memset(&u, 0, sizeof(u));
            u.kind = UNIT_FREE; u.n = 1;
            u.a[0] = i; u.a[1] = u.a[2] = u.a[3] = -1;
            p->units.push_back(u);
        }
    if (p->gc_pairs.empty()) { p->gc_pairs.push_back(0); p->gc_dist.push_back(1.0); p->gc_atoms.push_back(0); }

    lsi_system* s = new lsi_system();
    s->kind = LSI_SYS_PARTICLES;
    s->potential = -1;
    s->mass = 0.0;
    for (int i = 0; i < 4; ++i) s->params[i] = 0.0;
    s->particles = p.release();
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
    if (a->moments) {
        if (a->n_legs != 2) return lsi_set_error(LSI_ERR_ARG, "moments need n_legs == 2");
        if (!a->workspace || a->workspace_bytes < lsi_protocol_workspace_bytes(a->n_replicas))
            return lsi_set_error(LSI_ERR_ARG, "workspace too small: need %lld bytes", (long long)lsi_protocol_workspace_bytes(a->n_replicas));
    }

    DevTables* t = nullptr;
    int rc = get_tables(p, &t);
    if (rc != LSI_OK) return rc;
    const bool mixed = a->precision == LSI_PRECISION_MIXED;
    PKParams P;
    std::memset(&P, 0, sizeof(P));
    fill_system_params(p, t, mixed, P);

    // program: one force cache per distinct V token
    const int n_ops = (int)prog->ops.size();
    std::vector<unsigned char> kinds(n_ops), slots(n_ops, 0);
    std::vector<double> c0(n_ops), c1(n_ops, 0.0);
    int n_slots = 0;
    int slot_group[kMaxSlots] = {-1, -1, -1, -1};
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
                if (n_slots == kMaxSlots) return lsi_set_error(LSI_ERR_UNSUPPORTED, "at most 4 distinct force groups per splitting");
                sidx = n_slots;
                slot_group[n_slots++] = op.group;
            }
            slots[i] = (unsigned char)sidx;
        }
    }
    if (n_slots == 0) n_slots = 1;
    if (a->flags & (LSI_FLAG_TRACE_REPLICA_MAJOR | LSI_FLAG_TRACE_XV_PAIRS))
        return lsi_set_error(LSI_ERR_UNSUPPORTED, "replica-major / (x, v)-pair traces are layouts of scalar systems");
    const bool diag = a->W_direct || a->trace_x || a->trace_v || a->trace_w;
    if (diag) ++n_slots;  // scratch slot for the per-substep energy evaluations
    P.n_ops = n_ops; P.n_slots = n_slots; P.diag = diag ? 1 : 0;
    for (int q = 0; q < kMaxSlots; ++q) P.slot_group[q] = slot_group[q];

    PKLaunch L;
    rc = configure(p, mixed, n_slots, p->n_generic > 0, a->n_replicas, diag, L);
    if (rc != LSI_OK) return rc;

    unsigned char* dprog = nullptr;
    if (n_ops <= kMaxInlineOps) {
        for (int i = 0; i < n_ops; ++i) { P.op_kind[i] = kinds[i]; P.op_slot[i] = slots[i]; P.op_c0[i] = c0[i]; P.op_c1[i] = c1[i]; }
    } else {
        // long multi-timestep strings: the tables live in device memory, cached on the program handle
        // (lsi_program_stage: no allocation, copy or synchronisation after the first launch)
        const size_t nb = (size_t)n_ops * (2 * sizeof(double)) + 2 * (size_t)n_ops;
        std::vector<unsigned char> bytes(nb);
        std::memcpy(bytes.data(), c0.data(), n_ops * sizeof(double));
        std::memcpy(bytes.data() + n_ops * sizeof(double), c1.data(), n_ops * sizeof(double));
        std::memcpy(bytes.data() + 2 * n_ops * sizeof(double), kinds.data(), n_ops);
        std::memcpy(bytes.data() + 2 * n_ops * sizeof(double) + n_ops, slots.data(), n_ops);
        void* ext = nullptr;
        rc = lsi_program_stage(prog, bytes.data(), nb, &ext);
        if (rc != LSI_OK) return rc;
        dprog = (unsigned char*)ext;
        double* d0 = (double*)dprog;
        double* d1 = d0 + n_ops;
        unsigned char* dk = (unsigned char*)(d1 + n_ops);
        unsigned char* ds = dk + n_ops;
        P.ext_c0 = d0; P.ext_c1 = d1; P.ext_kind = dk; P.ext_slot = ds;
    }

    P.seed = a->seed; P.replica0 = a->replica0; P.n = a->n_replicas;
    P.n_steps = a->n_steps; P.n_legs = a->n_legs; P.marginal = a->marginal; P.flags = a->flags;
    P.kT = a->kT; P.inv_kT = 1.0 / a->kT;
    P.x_cache = a->x_cache; P.n_cache = a->n_cache; P.x0 = a->x0; P.v0 = a->v0;
    P.W = a->W; P.heat = a->heat; P.W_direct = a->W_direct; P.x_final = a->x_final; P.v_final = a->v_final;
    P.trace_x = a->trace_x; P.trace_v = a->trace_v; P.trace_w = a->trace_w;

    P.team_lanes = L.team_lanes;
    rc = mixed ? lsi_pk_launch_protocol_f32(P, L, stream) : lsi_pk_launch_protocol_f64(P, L, stream);
    if (rc == LSI_OK && a->moments)  // works are leg-major: W_F = W[0 .. n), W_R = W[n .. 2n)
        rc = lsi_reduce_moments(a->W, a->W + a->n_replicas, a->n_replicas, a->moments, a->workspace, a->workspace_bytes, stream);
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
    P.n = n;
    P.kT = 1.0; P.inv_kT = 1.0;
    P.q_group = group; P.q_x = x; P.q_energy = energy; P.q_forces = forces;
    PKLaunch L;
    rc = configure(p, mixed, 1, false, n, false, L);
    if (rc != LSI_OK) return rc;
    P.team_lanes = L.team_lanes;
    return mixed ? lsi_pk_launch_energy_f32(P, L, stream) : lsi_pk_launch_energy_f64(P, L, stream);
}

LSI_EXPORT int lsi_sample_equilibrium_particles(const lsi_system* sys, int32_t precision, double kT, int64_t n, int32_t n_rounds,
                                                int32_t steps_per_hmc, int32_t extra_chances, double dt_ps, uint64_t seed,
                                                uint64_t replica0, const double* x_in, double* x_out, double* v_out,
                                                double* accept_out, void* stream)
{
    if (!sys || sys->kind != LSI_SYS_PARTICLES || !x_in || !x_out || n < 0 || n_rounds < 0 || steps_per_hmc < 1 || extra_chances < 0 ||
        !(kT > 0.0) || !(dt_ps > 0.0))
        return lsi_set_error(LSI_ERR_ARG, "lsi_sample_equilibrium_particles: bad argument");
    if (precision != LSI_PRECISION_DOUBLE && precision != LSI_PRECISION_MIXED) return lsi_set_error(LSI_ERR_ARG, "unknown precision %d", precision);
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
    P.n_slots = 2;
    P.n = n;
    P.seed = seed; P.replica0 = replica0;
    P.kT = kT; P.inv_kT = 1.0 / kT;
    P.x0 = x_in; P.x_final = x_out; P.v_final = v_out;
    P.hmc_rounds = n_rounds; P.hmc_steps = steps_per_hmc; P.hmc_extra = extra_chances; P.hmc_dt = dt_ps; P.hmc_accept = accept_out;
    PKLaunch L;
    rc = configure(p, mixed, 2, p->n_generic > 0, n, false, L, true);
    if (rc != LSI_OK) return rc;
    P.team_lanes = L.team_lanes;
    return mixed ? lsi_pk_launch_hmc_f32(P, L, stream) : lsi_pk_launch_hmc_f64(P, L, stream);
}

// Internal declarations shared by the translation units of liblsi.so.
#pragma once
#include <cstdarg>
#include <cstdint>
#include <cstdio>
#include <mutex>
#include <string>
#include <vector>

#include "../../include/lsi.h"

#define LSI_EXPORT extern "C" __attribute__((visibility("default")))

struct lsi_op {
    int kind;         // LSI_OP_*
    int group;        // -1 = all forces
    double fraction;  // substep length / dt
    double a, b;      // O coefficients (baked at compile time)
};

// device copy of a long program's substep tables (more substeps than fit the kernel parameters), cached on the
// program handle: the next launch of the same tables on the same device finds it and needs no allocation, no copy
// and no host synchronisation
struct lsi_staged_program {
    int device;
    std::vector<unsigned char> host;  // the staged bytes (the key)
    void* dev;
};

struct lsi_program {
    std::vector<lsi_op> ops;
    double dt;
    double gamma;
    bool mts;
    bool measure_shadow_work;
    bool measure_heat;
    int n_O;
    mutable std::mutex mu;
    mutable std::vector<lsi_staged_program> staged;
};

enum lsi_system_kind { LSI_SYS_SCALAR = 0, LSI_SYS_PARTICLES = 1 };

struct lsi_particles;  // particles.cu

struct lsi_system {
    lsi_system_kind kind;
    // scalar
    int potential;
    double params[4];
    double mass;
    // particles
    lsi_particles* particles;
};

int lsi_set_error(int code, const char* fmt, ...);

#define LSI_CUDA_CHECK(expr)                                                                      \
    do {                                                                                          \
        cudaError_t err__ = (expr);                                                               \
        if (err__ != cudaSuccess)                                                                 \
            return lsi_set_error(LSI_ERR_CUDA, "%s failed: %s (%s:%d)", #expr,                    \
                                 cudaGetErrorString(err__), __FILE__, __LINE__);                  \
    } while (0)

// api.cu: device pointer of `bytes` staged for the current device (cached on the program; the first call for a given
// content copies synchronously, later calls return at once)
int lsi_program_stage(const lsi_program* prog, const void* bytes, size_t n_bytes, void** dev_out);
void lsi_program_free_staged(lsi_program* prog);
// toy.cu
int lsi_toy_run_protocol(const lsi_system* sys, const lsi_program* prog, const lsi_protocol_args* args,
                         void* stream);
// particles.cu
int lsi_particles_run_protocol(const lsi_system* sys, const lsi_program* prog, const lsi_protocol_args* args,
                               void* stream);
void lsi_particles_free(lsi_particles* p);
// reduce.cu
int lsi_finalize_moments(const double* partials, int n_partials, double* moments, void* stream);
// n (<= 8) independent finalizations in one launch: partials[i] ([n_partials][8]) -> moments[i] ([8])
int lsi_finalize_moments_batch(const double* const* partials, int n_partials, double* const* moments, int n, void* stream);
int lsi_toy_run_protocol_batch(const lsi_system* sys, const lsi_program* const* progs, int n_progs, const lsi_protocol_args* args,
                               void* stream);

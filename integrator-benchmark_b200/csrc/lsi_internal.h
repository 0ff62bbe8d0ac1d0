// Internal declarations shared by the translation units of liblsi.so.
#pragma once
#include <cstdarg>
#include <cstdint>
#include <cstdio>
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

struct lsi_program {
    std::vector<lsi_op> ops;
    double dt;
    double gamma;
    bool mts;
    bool measure_shadow_work;
    bool measure_heat;
    int n_O;
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

// toy.cu
int lsi_toy_run_protocol(const lsi_system* sys, const lsi_program* prog, const lsi_protocol_args* args,
                         void* stream);
// particles.cu
int lsi_particles_run_protocol(const lsi_system* sys, const lsi_program* prog, const lsi_protocol_args* args,
                               void* stream);
void lsi_particles_free(lsi_particles* p);
// reduce.cu
int lsi_finalize_moments(const double* partials, int n_partials, double* moments, void* stream);

// C-ABI glue: system handles and protocol dispatch (include/lsi.h).
#include <cuda_runtime.h>

#include "lsi_internal.h"

LSI_EXPORT int lsi_device_count(void)
{
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) {
        cudaGetLastError();  // clear the sticky "no device" error
        return 0;
    }
    return n;
}

LSI_EXPORT int lsi_system_create_scalar(int potential_kind, const double params[4], double mass, lsi_system** out)
{
    if (!out) return lsi_set_error(LSI_ERR_ARG, "lsi_system_create_scalar: NULL out");
    if (potential_kind < LSI_POT_QUARTIC || potential_kind > LSI_POT_HARMONIC)
        return lsi_set_error(LSI_ERR_UNSUPPORTED, "unknown scalar potential kind %d", potential_kind);
    if (!(mass > 0.0)) return lsi_set_error(LSI_ERR_ARG, "mass must be positive");
    lsi_system* s = new lsi_system();
    s->kind = LSI_SYS_SCALAR;
    s->potential = potential_kind;
    for (int i = 0; i < 4; ++i) s->params[i] = params ? params[i] : 0.0;
    s->mass = mass;
    s->particles = nullptr;
    *out = s;
    return LSI_OK;
}

LSI_EXPORT void lsi_system_destroy(lsi_system* s)
{
    if (!s) return;
    if (s->particles) lsi_particles_free(s->particles);
    delete s;
}

LSI_EXPORT int lsi_run_protocol(const lsi_system* sys, const lsi_program* prog, const lsi_protocol_args* args, void* stream)
{
    if (!sys || !prog || !args) return lsi_set_error(LSI_ERR_ARG, "lsi_run_protocol: NULL argument");
    if (args->n_replicas < 0 || args->n_steps < 0 || args->n_legs < 1)
        return lsi_set_error(LSI_ERR_ARG, "lsi_run_protocol: negative size");
    if (args->marginal != LSI_MARGINAL_CONFIGURATION && args->marginal != LSI_MARGINAL_FULL)
        return lsi_set_error(LSI_ERR_UNSUPPORTED, "`marginal` must be either 'configuration' or 'full'");
    if (!(args->kT > 0.0)) return lsi_set_error(LSI_ERR_ARG, "kT must be positive");
    if (prog->ops.empty()) return lsi_set_error(LSI_ERR_ARG, "empty program");
    if (lsi_device_count() <= 0)
        return lsi_set_error(LSI_ERR_CUDA, "no CUDA device: liblsi has no CPU fallback");
    if (sys->kind == LSI_SYS_SCALAR) return lsi_toy_run_protocol(sys, prog, args, stream);
    return lsi_particles_run_protocol(sys, prog, args, stream);
}

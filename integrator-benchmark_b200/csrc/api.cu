// C-ABI glue: system handles and protocol dispatch (include/lsi.h).
#include <cuda_runtime.h>

#include <cstring>

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

LSI_EXPORT int lsi_run_protocol_batch(const lsi_system* sys, const lsi_program* const* progs, int32_t n_progs,
                                      const lsi_protocol_args* args, void* stream)
{
    if (n_progs == 0) return LSI_OK;
    if (!sys || !progs || !args || n_progs < 0) return lsi_set_error(LSI_ERR_ARG, "lsi_run_protocol_batch: bad argument");
    for (int p = 0; p < n_progs; ++p)
        if (!progs[p]) return lsi_set_error(LSI_ERR_ARG, "lsi_run_protocol_batch: NULL program");
    if (sys->kind != LSI_SYS_SCALAR) {
        for (int p = 0; p < n_progs; ++p) {
            const int rc = lsi_run_protocol(sys, progs[p], &args[p], stream);
            if (rc != LSI_OK) return rc;
        }
        return LSI_OK;
    }
    for (int p = 0; p < n_progs; ++p) {
        const lsi_protocol_args* a = &args[p];
        if (a->n_replicas < 0 || a->n_steps < 0 || a->n_legs < 1) return lsi_set_error(LSI_ERR_ARG, "lsi_run_protocol_batch: negative size");
        if (a->marginal != LSI_MARGINAL_CONFIGURATION && a->marginal != LSI_MARGINAL_FULL)
            return lsi_set_error(LSI_ERR_UNSUPPORTED, "`marginal` must be either 'configuration' or 'full'");
        if (!(a->kT > 0.0)) return lsi_set_error(LSI_ERR_ARG, "kT must be positive");
        if (progs[p]->ops.empty()) return lsi_set_error(LSI_ERR_ARG, "empty program");
    }
    if (lsi_device_count() <= 0) return lsi_set_error(LSI_ERR_CUDA, "no CUDA device: liblsi has no CPU fallback");
    return lsi_toy_run_protocol_batch(sys, progs, n_progs, args, stream);
}

int lsi_program_stage(const lsi_program* prog, const void* bytes, size_t n_bytes, void** dev_out)
{
    int dev = 0;
    LSI_CUDA_CHECK(cudaGetDevice(&dev));
    std::lock_guard<std::mutex> lock(prog->mu);
    for (const lsi_staged_program& e : prog->staged)
        if (e.device == dev && e.host.size() == n_bytes && std::memcmp(e.host.data(), bytes, n_bytes) == 0) {
            *dev_out = e.dev;
            return LSI_OK;
        }
    if (prog->staged.size() >= 16) {  // setStepSize sweeps: drop the oldest copy (cudaFree waits for its users)
        cudaFree(prog->staged.front().dev);
        prog->staged.erase(prog->staged.begin());
    }
    void* d = nullptr;
    LSI_CUDA_CHECK(cudaMalloc(&d, n_bytes));
    const cudaError_t err = cudaMemcpy(d, bytes, n_bytes, cudaMemcpyHostToDevice);
    if (err != cudaSuccess) {
        cudaFree(d);
        return lsi_set_error(LSI_ERR_CUDA, "staging the program failed: %s", cudaGetErrorString(err));
    }
    lsi_staged_program e;
    e.device = dev;
    e.host.assign((const unsigned char*)bytes, (const unsigned char*)bytes + n_bytes);
    e.dev = d;
    prog->staged.push_back(e);
    *dev_out = d;
    return LSI_OK;
}

void lsi_program_free_staged(lsi_program* prog)
{
    std::lock_guard<std::mutex> lock(prog->mu);
    for (lsi_staged_program& e : prog->staged) cudaFree(e.dev);
    prog->staged.clear();
}

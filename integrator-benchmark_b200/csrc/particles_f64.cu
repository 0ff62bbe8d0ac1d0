// all-fp64 instantiations of the particle-system kernels (+ the shared-memory size query).
#include "particles_kernel.cuh"

int lsi_pk_launch_protocol_f64(const PKParams& P, const PKLaunch& L, void* stream) { return pk::launch_protocol<double>(P, L, stream); }
int lsi_pk_launch_hmc_f64(const PKParams& P, const PKLaunch& L, void* stream) { return pk::launch_hmc<double>(P, L, stream); }
int lsi_pk_launch_energy_f64(const PKParams& P, const PKLaunch& L, void* stream) { return pk::launch_energy<double>(P, L, stream); }

size_t lsi_pk_smem_per_team(bool mixed, int N, int n_slots, int n_tf, bool generic, bool warp_team, bool hmc)
{
    return mixed ? pk::layout<float>(N, n_slots, n_tf, generic, warp_team, hmc).total
                 : pk::layout<double>(N, n_slots, n_tf, generic, warp_team, hmc).total;
}

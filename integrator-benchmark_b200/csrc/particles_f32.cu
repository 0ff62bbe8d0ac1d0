// fp32-force ("mixed") instantiations of the particle-system kernels.
#include "particles_kernel.cuh"

int lsi_pk_launch_protocol_f32(const PKParams& P, const PKLaunch& L, void* stream) { return pk::launch_protocol<float>(P, L, stream); }
int lsi_pk_launch_hmc_f32(const PKParams& P, const PKLaunch& L, void* stream) { return pk::launch_hmc<float>(P, L, stream); }
int lsi_pk_launch_energy_f32(const PKParams& P, const PKLaunch& L, void* stream) { return pk::launch_energy<float>(P, L, stream); }

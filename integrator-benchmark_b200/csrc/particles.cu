// placeholder until the particle engine lands (K2/K3)
#include "lsi_internal.h"
struct lsi_particles { int n_atoms; };
void lsi_particles_free(lsi_particles* p) { delete p; }
LSI_EXPORT int lsi_system_num_dof(const lsi_system* s)
{
    if (!s) return LSI_ERR_ARG;
    return s->kind == LSI_SYS_SCALAR ? 1 : 3 * s->particles->n_atoms;
}
int lsi_particles_run_protocol(const lsi_system*, const lsi_program*, const lsi_protocol_args*, void*)
{
    return lsi_set_error(LSI_ERR_UNSUPPORTED, "particle systems not built yet");
}

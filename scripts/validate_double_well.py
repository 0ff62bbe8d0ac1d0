#!/usr/bin/env python
"""1-D double well (m = 10, beta = 1, gamma = 10): near-equilibrium estimate of the configuration-space KL divergence
for the six 5-letter splittings over a range of time steps, next to the values read off the reference's figure
notebooks/conf_summary.jpg (histogram D_KL at dt ~ 0.7, 1e7 steps per point: OVRVO ~0.23, ORVRO ~0.11, RVOVR ~0.08,
VRORV ~1.5e-3).  The work estimator is first order in the perturbation, so agreement is expected
for small values and only in order of magnitude / ranking for D_KL ~ 0.1."""
import json
import os
import sys
from functools import partial

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
READ_OFF = {"OVRVO": 0.23, "ORVRO": 0.11, "RVOVR": 0.08, "VRORV": 1.5e-3}


def main():
    import integrator_benchmark_b200 as ib
    from integrator_benchmark_b200.evaluation import estimate_from_moments
    from integrator_benchmark_b200.integrators import splitting_factory
    from integrator_benchmark_b200.testsystems import NumbaNonequilibriumSimulator, double_well
    ib.engine.require_cuda()
    ib.set_seed(20261020)
    out = []
    for name in ("OVRVO", "ORVRO", "RVOVR", "ROVOR", "VRORV", "VOROV"):
        scheme = splitting_factory(" ".join(name), double_well.potential, double_well.force, double_well.velocity_scale, double_well.mass)
        for dt in (0.1, 0.3, 0.5, 0.7):
            sim = NumbaNonequilibriumSimulator(double_well, partial(scheme, gamma=10.0, dt=dt))
            sim.collect_protocol_samples(2000000, 200, "configuration", traces=False, as_numpy=False)
            m = sim.last_moments.cpu().numpy()
            dF, var = estimate_from_moments(m)
            out.append({"scheme": name, "dt": dt, "DeltaF_neq_configuration": dF, "stderr": float(np.sqrt(max(var, 0))),
                        "n_crashed": int(m[6]), "figure_read_off_at_0.7": READ_OFF.get(name) if dt == 0.7 else None})
            print(out[-1], flush=True)
    json.dump(out, open(os.path.join(ROOT, "gpurun_out", "validate_double_well.json"), "w"), indent=1)


if __name__ == "__main__":
    main()

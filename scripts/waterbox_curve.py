"""DeltaF_neq(dt) of the tiny water box (config C5: 102 TIP3P waters, 1.5 nm box, cutoff 0.6 nm + reaction field, SETTLE) for
OVRVO and VRORV, configuration marginal -- the sweep of evaluation/compare_near_eq_and_exact.py for `tiny_waterbox`, for which
the reference publishes no numbers.  Reduced ensemble by default (8192 protocols x 1000 steps per point).

    python scripts/waterbox_curve.py [n_protocols] [n_cache]
"""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402

import integrator_benchmark_b200 as ib  # noqa: E402
from integrator_benchmark_b200 import unit  # noqa: E402
from integrator_benchmark_b200.evaluation import estimate_from_moments  # noqa: E402
from integrator_benchmark_b200.integrators import LangevinSplittingIntegrator  # noqa: E402
from integrator_benchmark_b200.testsystems import NonequilibriumSimulator, tiny_waterbox  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
n_cache = int(sys.argv[2]) if len(sys.argv) > 2 else 256
ib.set_seed(5)
eq = tiny_waterbox
eq.n_samples, eq.burn_in_length, eq.metropolis_rounds = n_cache, int(os.environ.get("LSI_BURN", 50000)), 2000
t0 = time.time()
eq.collect_equilibrium_samples()
out = {"n_protocols": n, "protocol_length": 1000, "cache": {"n": n_cache, "seconds": time.time() - t0, "hmc_acceptance": [float(a) for a in eq.acceptance]},
       "points": []}
t0 = time.time()
for splitting in ("O V R V O", "V R O R V"):
    for dt in (0.5, 1.0, 2.0, 3.0, 4.0):
        integ = LangevinSplittingIntegrator(splitting, temperature=298 * unit.kelvin, collision_rate=1.0 / unit.picoseconds,
                                            timestep=dt * unit.femtosecond)
        sim = NonequilibriumSimulator(eq, integ)
        sim.collect_protocol_samples(n, 1000, "configuration", as_numpy=False)
        m = sim.last_moments.cpu().numpy()
        dF, var = estimate_from_moments(m)
        out["points"].append({"splitting": splitting, "dt_fs": dt, "DeltaF_neq": dF, "stderr": float(np.sqrt(max(var, 0.0))),
                              "per_water": dF / 102.0, "n_crashed": int(m[6])})
        print(out["points"][-1], flush=True)
out["sweep_seconds"] = time.time() - t0
print(json.dumps(out))

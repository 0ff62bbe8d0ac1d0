"""DeltaF_neq(dt) of config C4 (alanine dipeptide in vacuum, H-bonds constrained, 298 K, gamma = 1/ps) for OVRVO, VRORV and the
MTS string of experiments/submission_scripts/4_mts_integrator_splittings.py ("V0 V1 R O R V1 V1 R O R V1 V0": nonbonded forces in group
0, valence terms in group 1, two inner steps), configuration marginal, protocol length 2000 as submission_scripts/5_hmr_alanine_vacuum.py:28-29
-- the sweep of evaluation/compare_near_eq_and_exact.py for `alanine_constrained`, for which the reference publishes no numbers.
With `--hmr` the same sweep for hydrogen masses scaled by 3 (utilities/openmm_utilities.py:223-269).

    python scripts/alanine_curve.py [n_protocols] [n_cache] [--hmr]
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
from integrator_benchmark_b200.testsystems import NonequilibriumSimulator, alanine_constrained, systems  # noqa: E402

args = [a for a in sys.argv[1:] if not a.startswith("--")]
n = int(args[0]) if len(args) > 0 else 65536
n_cache = int(args[1]) if len(args) > 1 else 1024
hmr = "--hmr" in sys.argv
ib.set_seed(7)
eq = alanine_constrained
if hmr:
    eq.system = systems.alanine_dipeptide(constrained=True, mass_scale=3.0)[0]
eq.n_samples, eq.burn_in_length, eq.metropolis_rounds = n_cache, 50000, 2000
t0 = time.time()
eq.collect_equilibrium_samples()
out = {"system": "alanine_constrained" + ("_hmr3" if hmr else ""), "n_protocols": n, "protocol_length": 2000,
       "cache": {"n": n_cache, "seconds": time.time() - t0, "hmc_acceptance": [float(a) for a in eq.acceptance]}, "points": []}
t0 = time.time()
for splitting in ("O V R V O", "V R O R V", "V0 V1 R O R V1 V1 R O R V1 V0"):
    for dt in (0.5, 1.0, 1.5, 2.0, 2.5, 3.0, 3.5, 4.0):
        integ = LangevinSplittingIntegrator(splitting, temperature=298 * unit.kelvin, collision_rate=1.0 / unit.picoseconds,
                                            timestep=dt * unit.femtosecond)
        sim = NonequilibriumSimulator(eq, integ, precision="mixed")
        sim.collect_protocol_samples(n, 2000, "configuration", as_numpy=False)
        m = sim.last_moments.cpu().numpy()
        dF, var = estimate_from_moments(m)
        out["points"].append({"splitting": splitting, "dt_fs": dt, "DeltaF_neq": dF, "stderr": float(np.sqrt(max(var, 0.0))), "n_crashed": int(m[6])})
        print(out["points"][-1], flush=True)
out["sweep_seconds"] = time.time() - t0
print(json.dumps(out))

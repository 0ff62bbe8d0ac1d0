"""One point of the nested ("exact") estimator with the reference's settings (compare_near_eq_and_exact.py:32-59, 322-383):
water_cluster_rigid, OVRVO, inner loop 50 + 1 per batch until stdev 0.01 (max 50 000), outer loop 50 + 100 per batch until
stdev 0.01 (max 1000), n_steps = n_steps_(dt, n_collisions=2).  Prints wall time and the replica-steps spent.

    python scripts/bench_nested.py [dt_fs] [marginal] [outer_max]
"""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402

import integrator_benchmark_b200 as ib  # noqa: E402
from integrator_benchmark_b200 import unit  # noqa: E402
from integrator_benchmark_b200.evaluation import compare_near_eq_and_exact as nested  # noqa: E402
from integrator_benchmark_b200.integrators import LangevinSplittingIntegrator  # noqa: E402
from integrator_benchmark_b200.testsystems import NonequilibriumSimulator, water_cluster_rigid  # noqa: E402
from functools import partial  # noqa: E402

dt = float(sys.argv[1]) if len(sys.argv) > 1 else 4.0
marginal = sys.argv[2] if len(sys.argv) > 2 else "full"
outer_max = int(sys.argv[3]) if len(sys.argv) > 3 else 1000
ib.set_seed(2026)
eq = water_cluster_rigid
t0 = time.time()
eq.load_or_simulate_samples()
t_cache = time.time() - t0
integ = LangevinSplittingIntegrator("O V R V O", temperature=298 * unit.kelvin, collision_rate=1.0 / unit.picoseconds,
                                    timestep=dt * unit.femtosecond)
sim = NonequilibriumSimulator(eq, integ)
n_steps = nested.n_steps_(dt * unit.femtosecond, n_collisions=2)
outer = partial(nested.outer_sample_adaptive, initial_size=50, batch_size=1, threshold=0.01, max_samples=50000)
t0 = time.time()
fn = getattr(nested, sys.argv[4]) if len(sys.argv) > 4 else nested.estimate_kl_div_adaptive_outer_loop
res = fn(sim, marginal, outer, n_steps=n_steps, initial_size=50, batch_size=100, threshold=0.01, max_samples=outer_max)
wall = time.time() - t0
n_inner = [len(w) for w in res["Ws"]]
steps = int(np.sum(n_inner)) * n_steps + len(n_inner) * nested.n_steps_(dt * unit.femtosecond, n_collisions=2)
print(json.dumps({"dt_fs": dt, "marginal": marginal, "n_steps": n_steps, "estimate": float(res["new_estimate"]),
                  "n_outer": len(n_inner), "inner_mean": float(np.mean(n_inner)), "inner_max": int(np.max(n_inner)),
                  "replica_steps": steps, "wall_s": wall, "replica_steps_per_s": steps / wall, "cache_s": t_cache}))

"""Nested estimator with the outer samples sharded over the GPUs of one node:

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 \
        scripts/run_nested_sharded.py [dt_fs] [outer_max]

Every rank builds the same equilibrium cache (same seed), collects its share of each round of outer samples with the
batched inner loops, and all ranks stop together on the statistics of the union (one 24-byte all-reduce per round)."""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch.distributed as dist  # noqa: E402

import integrator_benchmark_b200 as ib  # noqa: E402
from integrator_benchmark_b200 import parallel, unit  # noqa: E402
from integrator_benchmark_b200.evaluation import compare_near_eq_and_exact as nested  # noqa: E402
from integrator_benchmark_b200.integrators import LangevinSplittingIntegrator  # noqa: E402
from integrator_benchmark_b200.testsystems import NonequilibriumSimulator, water_cluster_rigid  # noqa: E402

dt = float(sys.argv[1]) if len(sys.argv) > 1 else 4.0
outer_max = int(sys.argv[2]) if len(sys.argv) > 2 else 300
rank, world, local = parallel.init_process_group()
ib.set_seed(2026)
eq = water_cluster_rigid
eq.burn_in_length = int(os.environ.get("LSI_BURN", 20000))
eq.load_or_simulate_samples()
ib.set_seed(2026 + 1000 * (rank + 1))  # different trajectories on every rank from here on
import numpy as np  # noqa: E402
np.random.seed(rank + 1)
integ = LangevinSplittingIntegrator("O V R V O", temperature=298 * unit.kelvin, collision_rate=1.0 / unit.picoseconds,
                                    timestep=dt * unit.femtosecond)
sim = NonequilibriumSimulator(eq, integ)
n_steps = nested.n_steps_(dt * unit.femtosecond, n_collisions=2)
t0 = time.time()
res = nested.estimate_kl_div_adaptive_outer_loop_sharded(sim, "full", None, n_steps=n_steps, initial_size=50, batch_size=50, threshold=0.01,
                                                         max_samples=outer_max,
                                                         inner=dict(initial_size=50, batch_size=1, threshold=0.01, max_samples=50000))
wall = time.time() - t0
print(json.dumps({"rank": rank, "world": world, "estimate": res["new_estimate"], "stdev": res["stdev"], "n_outer_global": res["n_outer_samples"],
                  "n_outer_here": len(res["Ws"]), "wall_s": wall}), flush=True)
if world > 1:
    dist.barrier()
    dist.destroy_process_group()

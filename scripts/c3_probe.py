#!/usr/bin/env python
"""Probe for the water-cluster (C3) parity experiment: (1) rate of lone extra-chance-HMC chains (one warp each), which sets
the cost of building equilibrium caches the reference's way (ONE chain, burn-in, thinning; bookkeepers.py:89-113);
(2) DeltaF_neq at 4 fs (OVRVO, full) in mixed and in double precision on the same cache and streams."""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import torch  # noqa: E402

import integrator_benchmark_b200 as ib  # noqa: E402
from integrator_benchmark_b200 import engine, unit  # noqa: E402
from integrator_benchmark_b200.evaluation import estimate_from_moments  # noqa: E402
from integrator_benchmark_b200.integrators import LangevinSplittingIntegrator  # noqa: E402
from integrator_benchmark_b200.testsystems import NonequilibriumSimulator, water_cluster_rigid  # noqa: E402

eq = water_cluster_rigid
eq.n_samples, eq.burn_in_length = 1000, 100000
ib.set_seed(20261019)
t0 = time.time()
eq.collect_equilibrium_samples(seed=5)
print(json.dumps({"cache_s": time.time() - t0, "acceptance": [float(a) for a in eq.acceptance]}), flush=True)
system = eq.engine_system()
x = eq.samples_device()
for n_chains in (32, 592):
    for rounds in (2000,):
        torch.cuda.synchronize()
        t0 = time.time()
        xo, _, acc = engine.sample_equilibrium_particles(system, x[:n_chains], eq.kT_value, rounds, 10, 15, 0.001, seed=1, replica0=0)
        torch.cuda.synchronize()
        dt = time.time() - t0
        print(json.dumps({"chains": n_chains, "rounds": rounds, "s": dt, "us_per_round": dt / rounds * 1e6,
                          "accept": acc.mean(dim=0).tolist()}), flush=True)
for precision in ("mixed", "double"):
    for marginal, dt_fs in (("full", 4.0), ("configuration", 2.0)):
        ib.set_seed(777)
        integ = LangevinSplittingIntegrator("O V R V O", temperature=298 * unit.kelvin, collision_rate=1.0 / unit.picoseconds,
                                            timestep=dt_fs * unit.femtosecond)
        sim = NonequilibriumSimulator(eq, integ, precision=precision)
        torch.cuda.synchronize()
        t0 = time.time()
        sim.collect_protocol_samples(100000, 1000, marginal, as_numpy=False)
        m = sim.last_moments.cpu().numpy()
        el = time.time() - t0
        dF, var = estimate_from_moments(m)
        print(json.dumps({"precision": precision, "marginal": marginal, "dt_fs": dt_fs, "DeltaF_neq": dF, "stderr": var ** 0.5,
                          "s": el, "replica_steps_per_s": 2e8 / el}), flush=True)

"""How much of the error of a DeltaF_neq estimate comes from the finite equilibrium cache?  Regenerates the cache of
water_cluster_rigid with different seeds (1000 configurations as in the reference, and 8192) and repeats two points of the
published curves with a fixed protocol seed.  Result: profiles/r1_cache_variance.log."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import integrator_benchmark_b200 as ib
from integrator_benchmark_b200 import unit
from integrator_benchmark_b200.evaluation import estimate_from_moments
from integrator_benchmark_b200.integrators import LangevinSplittingIntegrator
from integrator_benchmark_b200.testsystems import NonequilibriumSimulator, water_cluster_rigid
eq = water_cluster_rigid
eq.burn_in_length = 100000
def point(marginal, dt):
    ib.set_seed(777)
    integ = LangevinSplittingIntegrator("O V R V O", temperature=298 * unit.kelvin, collision_rate=1.0 / unit.picoseconds, timestep=dt * unit.femtosecond)
    sim = NonequilibriumSimulator(eq, integ, precision="mixed")
    sim.collect_protocol_samples(100000, 1000, marginal, as_numpy=False)
    dF, var = estimate_from_moments(sim.last_moments.cpu().numpy())
    return dF, var ** 0.5
for n_cache, seeds in ((1000, (1, 2, 3, 4, 5)), (8192, (11, 12))):
    for seed in seeds:
        eq.n_samples = n_cache
        eq.cached = False
        t0 = time.time()
        eq.collect_equilibrium_samples(seed=seed)
        t1 = time.time() - t0
        a, b = point("configuration", 2.0), point("full", 4.0)
        print("n_cache %d seed %d (%.1f s): conf 2 fs %.5f +/- %.5f (pub -0.00053) | full 4 fs %.5f +/- %.5f (pub 0.0201)" % (n_cache, seed, t1, a[0], a[1], b[0], b[1]), flush=True)

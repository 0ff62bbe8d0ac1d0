"""Drop-in for benchmark/evaluation/thresholds.py:15-147 and the acceptance-rate probe of
notebooks/"Water Cluster GHMC accept rates.ipynb" cells 4 and 7 (scope row 8f-4).

Both are short-trajectory ensembles over the protocol kernel: the reference loops `n_samples` times over one
OpenMM Context; here all `n_samples` trajectories advance together, one launch per energy check.
"""
from copy import deepcopy

import numpy as np

from .. import _rngstate, engine, unit
from ..integrators import LangevinSplittingIntegrator
from .analysis import estimate_nonequilibrium_free_energy


def integrator_factory(dt):
    return LangevinSplittingIntegrator(collision_rate=91.0 / unit.picosecond, timestep=dt)


def _fs(dt):
    return dt.value_in_unit(unit.femtosecond) if unit.is_quantity(dt) else float(dt)


# ------------------------------------------------------------------------------------------- stability
def test_stability(test_system, integrator, n_samples, trajectory_length, check_interval=10, precision=None,
                   seed=None):
    """True iff all `n_samples` trajectories of `trajectory_length` steps, started from x ~ cache and v ~ MB, keep a
    potential energy that is neither NaN nor positive at every `check_interval`-th step (thresholds.py:19-35: the
    reference checks the energy, not the positions).  `integrator` replaces the reference's `simulation` argument.
    Stops at the first failed check."""
    engine.require_cuda()
    eq = test_system
    n = int(n_samples)
    precision = eq.precision if precision is None else precision  # the platform's arithmetic ("Reference" -> double)
    system = eq.engine_system(integrator.getConstraintTolerance())
    seed = _rngstate.get_seed() if seed is None else int(seed)
    rid = _rngstate.reserve_replicas(n)
    t = engine.torch()
    cache = eq.samples_device()
    gen = t.Generator(device="cuda")
    gen.manual_seed(seed)
    x = cache[t.randint(len(cache), (n,), device="cuda", generator=gen)].contiguous()
    v = None
    for k in range(int(round(trajectory_length / check_interval))):
        res = engine.run_protocol(system, integrator.program, n, int(check_interval), eq.kT_value, n_legs=1,
                                  seed=seed + 15485863 * (k + 1), replica0=rid, x0=x, v0=v, precision=precision,
                                  want_final=True, want_moments=False, out={})
        x, v = res["x_final"], res["v_final"]
        U, _ = system.energy_forces(x, precision=precision, want_forces=False)
        if not bool((U <= 0).all()):
            return False
    return True


test_stability.__test__ = False  # not a pytest test, whatever its name


def find_stability_threshold_deterministic_search(test_system, integrator_factory, max_dt=10 * unit.femtoseconds,
                                                  max_bisections=15, n_samples=10, trajectory_length=10000,
                                                  verbose=True):
    """Bisect [0, max_dt] for the largest stable time step (thresholds.py:38-62); returns (min_dt, max_dt)."""
    lo, hi = 0.0, _fs(max_dt)
    say = print if verbose else (lambda *a: None)
    say("Searching for maximum stable timestep in range: [{} fs, {} fs]".format(lo, hi))
    for _ in range(max_bisections):
        mid = 0.5 * (lo + hi)
        stable = test_stability(test_system, integrator_factory(mid * unit.femtosecond), n_samples, trajectory_length)
        if stable:
            lo = mid
        else:
            hi = mid
        say("\t{} fs {}! Current range: [{} fs, {} fs]".format(mid, "stable" if stable else "unstable", lo, hi))
        if lo == hi:
            break
    say("Final range for stability threshold: [{} fs, {} fs]\n\n".format(lo, hi))
    return lo * unit.femtosecond, hi * unit.femtosecond


def probabilistic_bisection(noisy_oracle, initial_limits=(0, 1), p=0.6, n_iterations=1000, resolution=100000):
    """Probabilistic bisection with a numerically represented belief density (thresholds.py:68-103): query the
    belief's median; a positive answer ("threshold is above") multiplies the density at and above the median
    by p and below by 1 - p, a non-positive answer the other way round.  Returns (grid, answers, densities)."""
    grid = np.linspace(initial_limits[0], initial_limits[1], resolution)
    dx = np.diff(grid)

    def normalised(f):
        return f / np.sum(0.5 * (f[1:] + f[:-1]) * dx)  # trapezoid rule (np.trapz left numpy 2)
    densities = [normalised(np.ones(resolution))]
    answers = []
    for _ in range(n_iterations):
        f = densities[-1]
        cdf = np.cumsum(f) / np.sum(f)
        median = grid[np.argmin(np.abs(cdf - 0.5))]
        z = noisy_oracle(median)
        answers.append(z)
        upper = grid >= median
        weight = np.where(upper, p, 1.0 - p) if z > 0 else np.where(upper, 1.0 - p, p)
        densities.append(normalised(f * weight))
    return grid, answers, densities


def get_hmr_stability_threshold_curve(test_system, splitting="V R O R V", traj_length=1000, h_masses=None,
                                      n_iterations=100, n_samples=1):
    """Stability threshold (fs) against hydrogen-mass scale factor (thresholds.py:106-130).  The equilibrium cache of
    `test_system` is reused for every mass: configurational equilibrium does not depend on the masses."""
    from ..utilities.openmm_utilities import repartition_hydrogen_mass_amber
    topology = test_system.topology
    default_system = deepcopy(test_system.system)
    h_masses = np.arange(0.5, 4.51, 0.25) if h_masses is None else np.asarray(h_masses, dtype=float)
    thresholds = []
    try:
        for scale in h_masses:
            test_system.system = repartition_hydrogen_mass_amber(topology, default_system, scale_factor=scale)
            lsi = LangevinSplittingIntegrator(splitting)

            def stability_oracle(dt):
                lsi.setStepSize(dt * unit.femtoseconds)
                return 1 if (dt > 0 and test_stability(test_system, lsi, n_samples, traj_length)) else -1
            grid, _, densities = probabilistic_bisection(stability_oracle, initial_limits=(0, 10), p=0.6,
                                                         n_iterations=n_iterations)
            thresholds.append(grid[np.argmax(densities[-1])])
    finally:
        test_system.system = default_system
    return h_masses, thresholds


def error_oracle_factory(test_system, error_threshold, scheme="O V R V O", n_protocol_samples=100, protocol_length=1000):
    """+1 while |DeltaF_neq| at the queried time step (fs) stays below `error_threshold`, else -1 (:133-147)"""
    from ..testsystems.bookkeepers import NonequilibriumSimulator
    integrator = LangevinSplittingIntegrator(scheme, timestep=1.0 * unit.femtosecond)
    sim = NonequilibriumSimulator(test_system, integrator)

    def error_oracle(dt):
        integrator.setStepSize(dt * unit.femtosecond)
        out = sim.collect_protocol_samples(n_protocol_samples, protocol_length)
        DeltaF_neq, sq_unc = estimate_nonequilibrium_free_energy(out["W_shads_F"], out["W_shads_R"])
        return 1 if error_threshold > np.abs(DeltaF_neq) else -1
    return error_oracle


# ------------------------------------------------------------------------------------------- acceptance
def estimate_acceptance_rate(test_system, scheme, dt, n_samples=10000, gamma=1.0 / unit.picosecond,
                             temperature=298.0 * unit.kelvin, precision=None, seed=None):
    """GHMC acceptance ratios min(1, exp(-W_shad)) of ONE step of `scheme` from equilibrium, for `n_samples`
    draws (notebook cell 4); dt in fs.  One launch of the protocol kernel."""
    from ..testsystems.bookkeepers import NonequilibriumSimulator
    integrator = LangevinSplittingIntegrator(splitting=" ".join(scheme.replace(" ", "")), temperature=temperature,
                                             timestep=_fs(dt) * unit.femtosecond, collision_rate=gamma)
    sim = NonequilibriumSimulator(test_system, integrator, precision=precision)
    eq = test_system
    n = int(n_samples)
    seed = _rngstate.get_seed() if seed is None else int(seed)
    rid = _rngstate.reserve_replicas(n)
    res = engine.run_protocol(sim._system(), integrator.program, n, 1, eq.kT_value, n_legs=1, seed=seed,
                              replica0=rid, x_cache=eq.samples_device(), precision=sim.precision, want_moments=False,
                              out={})
    W = res["W"][0]
    return engine.torch().exp(-W).clamp(max=1.0).nan_to_num(0.0).cpu().numpy()


def get_acceptance_rate_curve(test_system, scheme, dts, n_samples=10000, **kwargs):
    """Acceptance rate and its 95 % half-width, 1.96 std / sqrt(n), per time step (notebook cell 7)."""
    rates, unc = np.zeros(len(dts)), np.zeros(len(dts))
    for i, dt in enumerate(dts):
        ratios = estimate_acceptance_rate(test_system, scheme, dt, n_samples, **kwargs)
        rates[i] = np.mean(ratios)
        unc[i] = 1.96 * np.std(ratios) / np.sqrt(len(ratios))
    return rates, unc

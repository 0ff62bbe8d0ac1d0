"""Scope row 8f-4 on the GPU: GHMC acceptance-rate probe, stability search (benchmark/evaluation/thresholds.py,
notebooks/"Water Cluster GHMC accept rates.ipynb") and the CustomizableGHMC ensemble (benchmark/integrators/ghmc.py)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ib():
    import integrator_benchmark_b200 as ib
    ib.engine.require_cuda()
    return ib


@pytest.fixture(scope="module")
def orc():
    from oracle import oracle
    return oracle


def _eq(builder, name, n_samples=64, burn=300, **kw):
    from integrator_benchmark_b200 import unit
    from integrator_benchmark_b200.testsystems import EquilibriumSimulator, configure_platform
    desc, pos = builder(**kw)
    eq = EquilibriumSimulator(platform=configure_platform("Reference"), topology=None, system=desc, positions=pos,
                              temperature=298 * unit.kelvin, timestep=1.0 * unit.femtosecond, burn_in_length=burn,
                              n_samples=n_samples, thinning_interval=10, name=name)
    eq.metropolis_rounds = 50
    return eq


@pytest.fixture(scope="module")
def cluster(ib):
    from integrator_benchmark_b200.testsystems import systems
    eq = _eq(systems.water_cluster, "wc", constrained=True)
    eq.load_or_simulate_samples()
    return eq


def test_acceptance_rate_probe_matches_one_step_works(ib, orc, cluster):
    """min(1, exp(-W)) of one step from equilibrium: against the oracle's works for the same replicas, and
    decreasing with the time step as in the notebook's reject-rate figure"""
    from integrator_benchmark_b200.evaluation import thresholds
    ib.set_seed(2024, next_replica=1000)
    r0 = 1000
    ratios = thresholds.estimate_acceptance_rate(cluster, "OVRVO", 2.0, n_samples=64, precision="double", seed=77)
    assert ratios.shape == (64,) and np.all((ratios > 0) & (ratios <= 1))
    ref = orc.particles_protocol(cluster.system, "O V R V O", 0.002, 1.0, cluster.kT_value, 64, 1, n_legs=1, seed=77,
                                 replica0=r0, x_cache=cluster.samples)
    np.testing.assert_allclose(ratios, np.minimum(1.0, np.exp(-ref["W"][:, 0])), rtol=1e-7, atol=1e-9)
    rates, unc = thresholds.get_acceptance_rate_curve(cluster, "VRORV", [0.5, 4.0, 8.0], n_samples=4096)
    assert rates[0] > 0.995 and rates[0] > rates[1] > rates[2]
    assert np.all(unc > 0) and np.all(unc < 0.05)


def test_stability_and_bisection(ib, cluster):
    from integrator_benchmark_b200 import unit
    from integrator_benchmark_b200.evaluation import thresholds
    from integrator_benchmark_b200.integrators import LangevinSplittingIntegrator

    def factory(dt):
        return LangevinSplittingIntegrator("V R O R V", collision_rate=1.0 / unit.picosecond, timestep=dt)
    assert thresholds.test_stability(cluster, factory(1.0 * unit.femtosecond), 16, 100)
    assert not thresholds.test_stability(cluster, factory(40.0 * unit.femtosecond), 16, 100)
    lo, hi = thresholds.find_stability_threshold_deterministic_search(cluster, factory, max_dt=32 * unit.femtoseconds,
                                                                     max_bisections=4, n_samples=16,
                                                                     trajectory_length=200, verbose=False)
    lo, hi = lo.value_in_unit(unit.femtosecond), hi.value_in_unit(unit.femtosecond)
    assert hi - lo == pytest.approx(2.0) and 2.0 <= lo and hi <= 24.0, (lo, hi)
    oracle = thresholds.error_oracle_factory(cluster, 0.05, "V R O R V", n_protocol_samples=256, protocol_length=50)
    assert oracle(0.5) == 1


def test_ghmc_ensemble(ib):
    """acceptance ~ 1 at a small step; rejected chains keep x and flip v; with the O substep applied the chains keep
    the kinetic temperature of the constrained system"""
    from integrator_benchmark_b200 import unit
    from integrator_benchmark_b200.integrators import CustomizableGHMC
    from integrator_benchmark_b200.testsystems import systems
    t = ib.engine.torch()
    eq = _eq(systems.alanine_dipeptide, "ala", n_samples=256, burn=300, constrained=True)
    eq.load_or_simulate_samples()
    x0 = eq.samples_device()
    small = CustomizableGHMC("V R V", timestep=0.5 * unit.femtosecond, constraint_tolerance=1e-8)
    x, v, acc = small.run(eq, x0, n_trials=4, seed=3)
    assert acc.shape == (4, len(x0)) and acc.float().mean().item() > 0.97
    assert small.getGlobalVariableByName("ntrials") == 4 * len(x0)
    assert small.acceptance_rate == pytest.approx(acc.float().mean().item())

    big = CustomizableGHMC("V R V", timestep=4.0 * unit.femtosecond, constraint_tolerance=1e-8)
    res = ib.engine.run_protocol(eq.engine_system(), big._segments[0][1], len(x0), 0, eq.kT_value, n_legs=1, seed=11,
                                 replica0=0, x0=x0, precision="double", want_final=True, want_moments=False, out={})
    v0 = res["v_final"].clone()
    x1, v1, acc = big.run(eq, x0, v0, n_trials=1, seed=5)
    rej = ~acc[0]
    assert 0 < rej.sum().item() < len(x0)
    assert t.equal(x1[rej], x0[rej]) and t.equal(v1[rej], -v0[rej])
    assert not t.equal(x1[~rej], x0[~rej])
    # detailed energy check of the accepted moves: the work the test used is e_new - e_old
    sys_ = eq.engine_system()
    m = t.as_tensor(np.asarray(eq.system["mass"], dtype=np.float64), device="cuda")
    e_old = sys_.energy_forces(x0, precision="double", want_forces=False)[0] + 0.5 * (m[None, :, None] * v0 ** 2).sum((1, 2))
    e_new = sys_.energy_forces(x1, precision="double", want_forces=False)[0] + 0.5 * (m[None, :, None] * v1 ** 2).sum((1, 2))
    assert t.allclose((e_new - e_old)[rej], t.zeros_like(e_old[rej]), atol=1e-9)
    assert (e_new - e_old)[~rej].abs().max().item() < 40.0 * eq.kT_value

    thermo = CustomizableGHMC("O V R V", timestep=1.0 * unit.femtosecond, collision_rate=50.0 / unit.picosecond,
                              constraint_tolerance=1e-8, apply_O=True)
    x2, v2, acc = thermo.run(eq, x0, n_trials=100, seed=9)
    ke = 0.5 * (m[None, :, None] * v2 ** 2).sum((1, 2)).mean().item()
    n_dof = 3 * 22 - 12
    assert ke == pytest.approx(0.5 * n_dof * eq.kT_value, rel=0.06)
    assert acc.float().mean().item() > 0.9


def test_hmr_threshold_curve_runs_and_restores_the_system(ib):
    """thresholds.py:106-130 with a short (hence coarse) search: thresholds land inside the search interval, and the
    test system gets its own masses back"""
    from integrator_benchmark_b200.evaluation import thresholds
    from integrator_benchmark_b200.testsystems import systems
    eq = _eq(systems.alanine_dipeptide, "ala_hmr", n_samples=32, burn=200, constrained=True)
    eq.load_or_simulate_samples()
    masses = np.array(eq.system["mass"], dtype=float).copy()
    h, thr = thresholds.get_hmr_stability_threshold_curve(eq, "V R O R V", traj_length=300, h_masses=[1.0, 3.0], n_iterations=14,
                                                          n_samples=8)
    assert list(h) == [1.0, 3.0] and len(thr) == 2
    assert 1.0 < thr[0] < 9.9 and 1.0 < thr[1] < 9.9, thr
    np.testing.assert_array_equal(np.array(eq.system["mass"], dtype=float), masses)

"""The reference-facing molecular API (LangevinSplittingIntegrator + Equilibrium/NonequilibriumSimulator,
benchmark/integrators/langevin.py, benchmark/testsystems/bookkeepers.py) on the GPU, checked against
the oracle.  Mirrors the reference's tests/test_shadow_work.py and tests/test_bookkeeper.py."""
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


def _eq(ib, builder, name, n_samples=48, burn=300, **kw):
    from integrator_benchmark_b200 import unit
    from integrator_benchmark_b200.testsystems import EquilibriumSimulator, configure_platform
    desc, pos = builder(**kw)
    return EquilibriumSimulator(platform=configure_platform("Reference"), topology=None, system=desc, positions=pos,
                                temperature=298 * unit.kelvin, timestep=1.0 * unit.femtosecond, burn_in_length=burn,
                                n_samples=n_samples, thinning_interval=10, name=name)


def test_shadow_work_identity_like_reference_test(ib, orc):
    """tests/test_shadow_work.py:45-55: alanine dipeptide (constrained and not) x 8 splittings, 2 fs (1 fs
    unconstrained), 10 warm-up + 100 steps: |W_shad(direct global) - (dE - dQ)| <= 1e-6 kJ/mol."""
    from integrator_benchmark_b200 import unit
    from integrator_benchmark_b200.integrators import LangevinSplittingIntegrator
    from integrator_benchmark_b200.testsystems import NonequilibriumSimulator, systems
    for constrained in (True, False):
        eq = _eq(ib, systems.alanine_dipeptide, "ala", n_samples=4, burn=200, constrained=constrained)
        for splitting in ["R V O", "R O V", "V R O R V", "R V O V R", "O R V R O", "R O V O R", "V O R O V",
                          "V R R O R R V"]:
            integ = LangevinSplittingIntegrator(splitting, timestep=(2.0 if constrained else 1.0) * unit.femtosecond,
                                                measure_shadow_work=True)
            names = [integ.getGlobalVariableName(i) for i in range(integ.getNumGlobalVariables())]
            assert "shadow_work" in names and "heat" in names
            sim = NonequilibriumSimulator(eq, integ, precision="double")
            x0 = sim.sample_x_from_equilibrium()
            v0 = sim.sample_v_given_x(x0)
            sim.accumulate_shadow_work(x0, v0, 10)
            sw0, q0 = integ.getGlobalVariableByName("shadow_work"), integ.getGlobalVariableByName("heat")
            x1, v1 = sim._last
            out = sim.accumulate_shadow_work(x1, v1, 100)
            sw = integ.getGlobalVariableByName("shadow_work") - sw0
            kT = eq.kT_value
            assert abs(sw - out["W_shad"] * kT) <= 1e-6, (splitting, constrained)
            assert integ.getGlobalVariableByName("heat") != q0


def test_accumulate_shadow_work_matches_oracle(ib, orc):
    from integrator_benchmark_b200 import unit
    from integrator_benchmark_b200.integrators import LangevinSplittingIntegrator
    from integrator_benchmark_b200.testsystems import NonequilibriumSimulator, systems
    eq = _eq(ib, systems.water_cluster, "wc", n_samples=8, burn=200, constrained=True)
    integ = LangevinSplittingIntegrator("O V R V O", collision_rate=1.0 / unit.picosecond, timestep=2.0 * unit.femtosecond)
    sim = NonequilibriumSimulator(eq, integ, precision="double")
    ib.set_seed(99, next_replica=500)
    x0 = sim.sample_x_from_equilibrium()
    v0 = sim.sample_v_given_x(x0)            # replica id 500
    res = sim.accumulate_shadow_work(x0, v0, 25, store_potential_energy=True, store_W_shad_trace=True)  # id 501
    ref = orc.particles_protocol(eq.system, "O V R V O", 0.002, 1.0, eq.kT_value, 1, 25, seed=99, replica0=501,
                                 x0=x0[None], v0=np.asarray(v0.value_in_unit(unit.nanometer / unit.picosecond))[None],
                                 n_legs=1, traces=True)
    assert res["W_shad"] == pytest.approx(ref["W"][0, 0], rel=1e-7, abs=2e-8)
    assert res["W_shad_trace"].shape == (25,) and res["W_shad_trace"].sum() == pytest.approx(res["W_shad"], abs=1e-9)
    pe = res["potential_energies"].value_in_unit(unit.kilojoule_per_mole)
    np.testing.assert_allclose(pe, ref["trace_pe"][0, 0], rtol=1e-9, atol=1e-7)
    assert integ.getGlobalVariableByName("heat") == pytest.approx(ref["Q"][0, 0], rel=1e-7, abs=1e-7)
    # velocities given x: Maxwell-Boltzmann, projected on the rigid-water constraints
    v = np.asarray(v0.value_in_unit(unit.nanometer / unit.picosecond)).reshape(20, 3, 3)
    x = x0.reshape(20, 3, 3)
    for a, b in ((0, 1), (0, 2), (1, 2)):
        assert np.abs(np.einsum("wc,wc->w", x[:, a] - x[:, b], v[:, a] - v[:, b])).max() < 1e-10
    # a non-finite start gives the reference's empty dictionary
    bad = x0.copy()
    bad[0, 0] = np.nan
    assert sim.accumulate_shadow_work(bad, v0, 5) == {}


def test_collect_protocol_samples_ensemble_matches_oracle_and_estimator(ib, orc):
    from integrator_benchmark_b200 import unit
    from integrator_benchmark_b200.evaluation import estimate_nonequilibrium_free_energy
    from integrator_benchmark_b200.integrators import LangevinSplittingIntegrator
    from integrator_benchmark_b200.testsystems import NonequilibriumSimulator, systems
    eq = _eq(ib, systems.water_cluster, "wc", n_samples=32, burn=300, constrained=True)
    samples = eq.collect_equilibrium_samples(seed=4)
    assert samples.shape == (32, 60, 3) and np.isfinite(samples).all()
    # the cache is equilibrated: rigid geometry holds and the energy dropped well below the lattice start
    oh = np.linalg.norm(samples[:, 1::3] - samples[:, 0::3], axis=-1)
    assert np.abs(oh - 0.09572).max() < 1e-7
    e_start = orc.particles_energy_forces(eq.system, eq.positions)[0]
    e_eq = np.mean([orc.particles_energy_forces(eq.system, s)[0] for s in samples[:8]])
    assert e_eq < e_start - 100.0
    integ = LangevinSplittingIntegrator("V R O R V", collision_rate=1.0 / unit.picosecond, timestep=4.0 * unit.femtosecond)
    for marginal in ("configuration", "full"):
        sim = NonequilibriumSimulator(eq, integ, precision="double")
        ib.set_seed(7, next_replica=0)
        n, length = 24, 12
        out = sim.collect_protocol_samples(n, length, marginal, store_W_shad_traces=True)
        ref = orc.particles_protocol(eq.system, "V R O R V", 0.004, 1.0, eq.kT_value, n, length, marginal=marginal, seed=7,
                                     replica0=0, x_cache=samples, round_midpoint=True)
        np.testing.assert_allclose(out["W_shads_F"], ref["W"][:, 0], rtol=1e-7, atol=2e-8)
        np.testing.assert_allclose(out["W_shads_R"], ref["W"][:, 1], rtol=1e-6, atol=1e-6)
        assert len(out["W_shad_traces"]) == n and out["W_shad_traces"][3][0].shape == (length,)
        assert out["W_shad_traces"][3][1].sum() == pytest.approx(out["W_shads_R"][3], abs=1e-8)
        dF, var = estimate_nonequilibrium_free_energy(out["W_shads_F"], out["W_shads_R"])
        wf, wr = ref["W"][:, 0], ref["W"][:, 1]
        assert dF == pytest.approx(0.5 * (wf.mean() - wr.mean()), abs=1e-6)
        assert var == pytest.approx((np.var(wf) + np.var(wr) - 2 * np.cov(wf, wr)[0, 1]) / (4 * n), rel=1e-4)
    # potential-energy traces of the reverse leg (store_potential_energy_traces, bookkeepers.py:272-274,288-293)
    sim = NonequilibriumSimulator(eq, integ, precision="double")
    ib.set_seed(7, next_replica=0)
    out_pe = sim.collect_protocol_samples(6, 5, "full", store_potential_energy_traces=True)
    ref_pe = orc.particles_protocol(eq.system, "V R O R V", 0.004, 1.0, eq.kT_value, 6, 5, marginal="full", seed=7, replica0=0,
                                    x_cache=samples, round_midpoint=True, traces=True)
    assert len(out_pe["potential_energies"]) == 6 and out_pe["potential_energies"][2].shape == (6,)
    np.testing.assert_allclose(np.array(out_pe["potential_energies"]), ref_pe["trace_pe"][:, 1], rtol=1e-8, atol=1e-6)
    with pytest.raises(NotImplementedError):
        sim.collect_protocol_samples(4, 4, "momentum")
    # mixed precision (the reference's CUDA configuration) stays close to the fp64 run
    sim32 = NonequilibriumSimulator(eq, integ)
    ib.set_seed(7, next_replica=0)
    out32 = sim32.collect_protocol_samples(24, 12, "full")
    assert np.abs(out32["W_shads_F"] - out["W_shads_F"]).max() < 5e-3


def test_named_reference_systems_run(ib):
    """water_cluster_rigid / alanine_constrained / tiny_waterbox by their reference names."""
    import os
    from integrator_benchmark_b200 import testsystems, unit
    from integrator_benchmark_b200.integrators import LangevinSplittingIntegrator
    from integrator_benchmark_b200.testsystems import NonequilibriumSimulator
    golden = os.path.join(os.path.dirname(__file__), "golden", "waterbox_frames.npz")
    testsystems.tiny_waterbox.set_samples(np.load(golden)["xyz"].astype(np.float64))
    for eq, string in ((testsystems.tiny_waterbox, "O V R V O"), (testsystems.alanine_constrained, "V0 V1 R O R V1 V0")):
        if not eq.cached:
            eq.n_samples, eq.burn_in_length = 16, 200
        integ = LangevinSplittingIntegrator(string, collision_rate=1.0 / unit.picosecond, timestep=2.0 * unit.femtosecond)
        out = NonequilibriumSimulator(eq, integ).collect_protocol_samples(64, 20)
        assert np.isfinite(out["W_shads_F"]).all() and np.isfinite(out["W_shads_R"]).all()
        assert np.abs(out["W_shads_F"]).max() < 20.0


def test_hmc_equilibrium_sampler_matches_oracle_and_target(ib, orc):
    """lsi_sample_equilibrium_particles (extra-chance HMC, the sampler behind the reference's sample caches,
    benchmark/integrators/kyle/xchmc.py:84-116): same chains as the oracle, and the right distribution."""
    from integrator_benchmark_b200.testsystems import systems
    KT = 8.3144621e-3 * 298.0
    for builder, kw, dt in ((systems.water_cluster, dict(n_waters=20, constrained=True, seed=1), 0.002),
                            (systems.alanine_dipeptide, dict(constrained=True), 0.002)):
        desc, x = builder(**kw)
        desc = dict(desc, constraint_tolerance=1e-13)
        system = ib.engine.ParticleSystem(desc)
        # start from a relaxed configuration so that proposals are accepted and rejected
        x0, _, _ = orc.particles_hmc(desc, KT, x[None], 30, 5, 3, 0.0005, seed=1)
        x0 = np.repeat(x0, 5, axis=0)
        seen = []
        for extra, big_dt in ((0, 4.0 * dt), (0, 2.5 * dt), (3, 3.0 * dt), (15, dt)):
            xg, vg, ag = ib.engine.sample_equilibrium_particles(system, x0, KT, 8, 6, extra, big_dt, seed=9, replica0=40,
                                                                precision="double", want_velocities=True)
            xo, vo, ao = orc.particles_hmc(desc, KT, x0, 8, 6, extra, big_dt, seed=9, replica0=40)
            np.testing.assert_allclose(ag.cpu().numpy(), ao, atol=1e-12)
            np.testing.assert_allclose(xg.cpu().numpy(), xo, rtol=1e-7, atol=2e-9)
            seen.append(ao[:, 0])
        assert 0.0 < np.concatenate(seen).mean() < 1.0  # the large steps make the test see both branches
    # exact target: independent particles in the harmonic restraint -> Gaussian with variance kT / K per coordinate
    desc = dict(n_atoms=32, mass=np.linspace(1.0, 40.0, 32), restraint_k=50.0, groups=dict(restraint=0))
    system = ib.engine.ParticleSystem(desc)
    n = 4096
    x, _, acc = ib.engine.sample_equilibrium_particles(system, np.random.RandomState(0).randn(n, 32, 3) * 0.3, KT, 40, 10, 2, 0.05, seed=2)
    x = x.cpu().numpy()
    var = KT / 50.0
    assert acc[:, 0].mean().item() > 0.6
    assert x.mean() == pytest.approx(0.0, abs=5 * np.sqrt(var / x.size))
    assert x.var() == pytest.approx(var, rel=5 * np.sqrt(2.0 / x.size) * 3)
    # per-atom variances do not depend on the mass
    pv = x.reshape(n, 32, 3).var(axis=(0, 2))
    assert np.abs(pv / var - 1.0).max() < 0.08

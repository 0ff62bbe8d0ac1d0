"""GPU parity of the particle-system path (configs C3-C5: water cluster, alanine dipeptide,
tiny water box) against the CPU oracle.  Everything goes through the C ABI
(integrator_benchmark_b200 -> ctypes -> liblsi.so); the oracle is only the checker.

Tolerances: both sides are fp64 with the same Philox words.  They differ in summation order
(the kernel gathers per atom, the oracle loops over pairs), in libm-vs-CUDA log/sincos/acos last
ulps and in fma contraction; over the 8-12 steps run here that stays below 1e-8 relative on
coordinates -- three orders inside BASELINE.json's 1e-5 bar.
"""
import json
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

KT = 8.3144621e-3 * 298.0
XTOL = dict(rtol=1e-7, atol=1e-9)
WTOL = dict(rtol=1e-7, atol=2e-8)  # works in kT; energies of the box are ~ -4e3 kJ/mol


@pytest.fixture(scope="module")
def ib():
    import integrator_benchmark_b200 as ib
    ib.engine.require_cuda()
    return ib


@pytest.fixture(scope="module")
def orc():
    from oracle import oracle
    return oracle


@pytest.fixture(scope="module")
def systems(ib, golden_dir):
    from integrator_benchmark_b200.testsystems import systems as ts
    out = {}
    out["cluster_rigid"] = ts.water_cluster(20, constrained=True, seed=1)
    out["cluster_flexible"] = ts.water_cluster(20, constrained=False, seed=2)
    out["alanine_constrained"] = ts.alanine_dipeptide(constrained=True)
    out["alanine_unconstrained"] = ts.alanine_dipeptide(constrained=False)
    out["alanine_hmr"] = ts.alanine_dipeptide(constrained=True, mass_scale=3.0)
    # the kernel solves the H-bond stars by Newton iteration (residual ~1e-16), the oracle by Gauss-Seidel
    # SHAKE to `constraint_tolerance`: compare at a tolerance where both are converged
    for k in ("alanine_constrained", "alanine_hmr"):
        out[k][0]["constraint_tolerance"] = 1e-13
    desc, _ = ts.tiny_waterbox(constrained=True)
    frames = np.load(os.path.join(golden_dir, "waterbox_frames.npz"))["xyz"].astype(np.float64)
    out["waterbox"] = (desc, frames[0])
    out["waterbox_frames"] = frames
    return out


def _jitter(x, n, scale, seed):
    rng = np.random.RandomState(seed)
    return x[None] + scale * rng.randn(n, *x.shape)


MIXED_DESC = dict(
    n_atoms=9, mass=np.ones(9), charge=np.tile([-0.834, 0.417, 0.417], 3), sigma=np.tile([0.315075, 0.1, 0.1], 3),
    epsilon=np.tile([0.635968, 0.05, 0.05], 3),
    exclusions=[(0, 1), (0, 2), (1, 2), (3, 4), (3, 5), (4, 5), (6, 7), (6, 8), (7, 8)], restraint_k=1.5,
    bond_atoms=[(0, 1), (0, 2)], bond_params=[(0.09572, 462750.4)] * 2, angle_atoms=[(1, 0, 2)],
    angle_params=[(1.82421813, 836.8)], torsion_atoms=[(1, 0, 3, 4), (2, 0, 3, 5)],
    torsion_params=[(3, 0.3, 5.0), (2, 3.14159, 2.5)], exception_atoms=[(1, 4)], exception_params=[(0.1, 0.25, 0.3)],
    groups=dict(nonbonded=0, bonds=1, angles=1, torsions=2, restraint=3))


def test_energy_forces_match_oracle_every_term_and_group(ib, orc):
    rng = np.random.RandomState(3)
    x = rng.randn(5, 9, 3) * 0.12 + np.repeat(np.array([[0, 0, 0], [0.35, 0, 0.1], [0, 0.4, -0.1]]), 3, axis=0)
    for extra in [{}, dict(nonbonded_method="CutoffPeriodic", cutoff=0.45, box=(1.1, 1.2, 1.3))]:
        d = dict(MIXED_DESC, **extra)
        system = ib.engine.ParticleSystem(d)
        for group in (-1, 0, 1, 2, 3):
            e, f = system.energy_forces(x, group=group)
            e, f = e.cpu().numpy(), f.cpu().numpy()
            for k in range(len(x)):
                e_ref, f_ref = orc.particles_energy_forces(d, x[k], group=group)
                assert e[k] == pytest.approx(e_ref, rel=1e-11, abs=1e-10), (extra, group)
                np.testing.assert_allclose(f[k], f_ref, rtol=1e-9, atol=1e-8, err_msg=str((extra, group)))
        # group forces add up to the total
        tot = system.energy_forces(x)[1]
        parts = sum(system.energy_forces(x, group=g)[1] for g in range(4))
        np.testing.assert_allclose(parts.cpu().numpy(), tot.cpu().numpy(), rtol=1e-10, atol=1e-8)


@pytest.mark.parametrize("name", ["cluster_rigid", "cluster_flexible", "alanine_constrained", "alanine_unconstrained",
                                  "waterbox"])
def test_energy_forces_of_reference_systems(ib, orc, systems, name):
    desc, x0 = systems[name]
    x = _jitter(x0, 3, 0.002, 5)
    system = ib.engine.ParticleSystem(desc)
    e, f = system.energy_forces(x)
    em, fm = system.energy_forces(x, precision="mixed")
    for k in range(len(x)):
        e_ref, f_ref = orc.particles_energy_forces(desc, x[k])
        scale = np.abs(f_ref).max()
        assert e[k].item() == pytest.approx(e_ref, rel=1e-11, abs=1e-9)
        np.testing.assert_allclose(f[k].cpu().numpy(), f_ref, rtol=1e-9, atol=1e-9 * scale)
        # fp32 force evaluation ("mixed"): single-precision round-off of the pair sums
        assert em[k].item() == pytest.approx(e_ref, rel=2e-5, abs=2e-3)
        np.testing.assert_allclose(fm[k].cpu().numpy(), f_ref, rtol=2e-4, atol=2e-5 * scale)


def test_waterbox_fixture_energies_are_sane(ib, orc, systems, golden_dir):
    """The four frames of data/tiny_waterbox_constrained_samples.h5 carry PME + dispersion-correction energies
    (-4136.6 ... -4029.0 kJ/mol).  The variant built here (reaction-field electrostatics, the reference's
    switched Lennard-Jones) lacks the LJ tail correction (about -74 kJ/mol beyond 0.6 nm) and the long-range
    electrostatics: it recovers 94.7-95.3 % of each stored energy; lowest and highest frame agree."""
    desc, _ = systems["waterbox"]
    assert desc["switch_distance"] == pytest.approx(0.45)  # openmmtools WaterBox: switch_width = 1.5 A below the cutoff
    e, _ = ib.engine.ParticleSystem(desc).energy_forces(systems["waterbox_frames"], want_forces=False)
    e = e.cpu().numpy()
    stored = np.load(os.path.join(golden_dir, "waterbox_frames.npz"))["potential_energy"].astype(np.float64)
    ratio = e / stored
    assert np.all(ratio > 0.94) and np.all(ratio < 0.96), ratio
    assert np.argmin(e) == np.argmin(stored) and np.argmax(e) == np.argmax(stored)
    # the switching function lowers the magnitude of every frame's energy by 30-40 kJ/mol against plain truncation
    plain = dict(desc, switch_distance=0.0)
    e0, _ = ib.engine.ParticleSystem(plain).energy_forces(systems["waterbox_frames"], want_forces=False)
    assert np.all(e - e0.cpu().numpy() > 20.0) and np.all(e - e0.cpu().numpy() < 60.0)


def test_switched_lennard_jones_is_continuous_and_conservative(ib, orc):
    """Two LJ sites crossing the switching region of a periodic system: energy and force of the kernel equal the
    oracle's, go to zero at the cutoff, and the force is the derivative of the energy (OpenMM's switching function)."""
    d = dict(n_atoms=2, mass=[16.0, 16.0], charge=[0.0, 0.0], sigma=[0.3150752406575124] * 2, epsilon=[0.635968] * 2,
             nonbonded_method="CutoffPeriodic", cutoff=0.6, switch_distance=0.45, box=(1.5, 1.5, 1.5))
    r = np.linspace(0.40, 0.62, 221)
    x = np.zeros((len(r), 2, 3))
    x[:, 0] = 0.2
    x[:, 1] = 0.2
    x[:, 1, 0] += r
    system = ib.engine.ParticleSystem(d)
    e, f = system.energy_forces(x)
    e, f = e.cpu().numpy(), f.cpu().numpy()
    for k in range(0, len(r), 10):
        e_ref, f_ref = orc.particles_energy_forces(d, x[k])
        assert e[k] == pytest.approx(e_ref, rel=1e-11, abs=1e-13)
        np.testing.assert_allclose(f[k], f_ref, rtol=1e-9, atol=1e-10)
    assert np.all(e[r >= 0.6] == 0.0) and np.all(f[r >= 0.6] == 0.0)
    assert abs(e[np.searchsorted(r, 0.5999)]) < 1e-9  # continuous at the cutoff (plain truncation jumps by -0.052 kJ/mol)
    dEdr = np.gradient(e, r)
    inside = (r > 0.41) & (r < 0.59)
    np.testing.assert_allclose(f[inside, 1, 0], -dEdr[inside], atol=2e-3)
    em, fm = system.energy_forces(x, precision="mixed")
    np.testing.assert_allclose(em.cpu().numpy(), e, atol=2e-6)
    np.testing.assert_allclose(fm.cpu().numpy(), f, atol=2e-5)


def _compare_protocol(ib, orc, desc, x0, splitting, dt, gamma, n_steps, marginal, seed, want_direct=True, v0=None,
                      round_midpoint=False, tol_scale=1.0):
    n = len(x0)
    system = ib.engine.ParticleSystem(desc)
    prog = ib.engine.Program(splitting, dt, gamma, measure_shadow_work=want_direct)
    res = ib.engine.run_protocol(system, prog, n, n_steps, KT, marginal=marginal, seed=seed, replica0=900, x0=x0, v0=v0,
                                 want_heat=True, want_direct=want_direct, want_final=True, traces=True,
                                 round_midpoint=round_midpoint)
    ref = orc.particles_protocol(desc, splitting, dt, gamma, KT, n, n_steps, marginal=marginal, seed=seed, replica0=900,
                                 x0=x0, v0=v0, traces=True, want_direct=want_direct, round_midpoint=round_midpoint)
    xt = dict(rtol=XTOL["rtol"] * tol_scale, atol=XTOL["atol"] * tol_scale)
    wt = dict(rtol=WTOL["rtol"] * tol_scale, atol=WTOL["atol"] * tol_scale)
    np.testing.assert_allclose(res["x_final"].cpu().numpy(), ref["x_final"], err_msg=splitting, **xt)
    np.testing.assert_allclose(res["v_final"].cpu().numpy(), ref["v_final"], err_msg=splitting, rtol=xt["rtol"], atol=xt["atol"] * 100)
    np.testing.assert_allclose(res["trace_x"].cpu().numpy().transpose(2, 0, 1, 3, 4), ref["trace_x"], **xt)
    np.testing.assert_allclose(res["W"].cpu().numpy().T, ref["W"], err_msg=splitting, **wt)
    np.testing.assert_allclose(res["heat"].cpu().numpy().T, ref["Q"], rtol=wt["rtol"], atol=wt["atol"] * KT)
    E = ref["trace_pe"] + ref["trace_ke"]
    tw = ((E - E[:, :, :1]) - ref["trace_heat"]) / KT
    np.testing.assert_allclose(res["trace_w"].cpu().numpy().transpose(2, 0, 1), tw, **wt)
    if want_direct:
        np.testing.assert_allclose(res["W_direct"].cpu().numpy().T, ref["Wdirect"], **wt)
        # reference tests/test_shadow_work.py:30-42: direct shadow work == dE - heat to 1e-6 kJ/mol
        assert np.abs(res["W_direct"].cpu().numpy() - res["W"].cpu().numpy()).max() * KT < 1e-6
    # the production variant (no traces, no direct work) gives the same works
    res2 = ib.engine.run_protocol(system, ib.engine.Program(splitting, dt, gamma), n, n_steps, KT, marginal=marginal,
                                  seed=seed, replica0=900, x0=x0, v0=v0, round_midpoint=round_midpoint)
    np.testing.assert_allclose(res2["W"].cpu().numpy().T, ref["W"], err_msg=splitting, **wt)
    m = res2["moments"].cpu().numpy()
    w = ref["W"]
    np.testing.assert_allclose(m[:6], [n, w[:, 0].sum(), w[:, 1].sum(), (w[:, 0] ** 2).sum(), (w[:, 1] ** 2).sum(),
                                       (w[:, 0] * w[:, 1]).sum()], rtol=1e-6, atol=1e-7)
    return res, ref


EIGHT = ["R V O", "R O V", "V R O R V", "R V O V R", "O R V R O", "R O V O R", "V O R O V", "V R R O R R V"]


@pytest.mark.parametrize("name", ["cluster_rigid", "cluster_flexible"])
def test_water_cluster_protocol_matches_oracle(ib, orc, systems, name):
    desc, x = systems[name]
    x0 = _jitter(x, 3, 0.0 if "rigid" in name else 0.001, 7)
    dt = 0.002 if "rigid" in name else 0.0005
    for j, splitting in enumerate(EIGHT + ["O V R V O"]):
        _compare_protocol(ib, orc, desc, x0, splitting, dt, 1.0 if j % 2 else 91.0, 8, ["configuration", "full"][j % 2],
                          seed=40 + j)


@pytest.mark.parametrize("name", ["alanine_constrained", "alanine_unconstrained", "alanine_hmr"])
def test_alanine_protocol_matches_oracle_incl_mts(ib, orc, systems, name):
    """the 8 splittings x {constrained, unconstrained} grid of the reference's
    tests/test_shadow_work.py:45-55 plus multi-timestep strings over force groups
    (experiments/submission_scripts/4_mts_integrator_splittings.py:10-32)."""
    from integrator_benchmark_b200.utilities import generate_baoab_mts_string, generate_gbaoab_solvent_solute_string
    desc, x = systems[name]
    x0 = _jitter(x, 3, 0.001 if "unconstrained" in name else 0.0, 9)
    dt = 0.001 if "unconstrained" in name else 0.002
    mts = ["V0 V1 R O R V1 V0", "V1 V0 R R O R R V0 V1", generate_baoab_mts_string([([0], 1), ([1], 2)]),
           generate_gbaoab_solvent_solute_string(K_r=2, slow_group=[0], fast_group=[1]),
           generate_baoab_mts_string([([0], 1), ([1], 30)])]  # 152 substeps: the program travels in device memory
    for j, splitting in enumerate(EIGHT + mts):
        _compare_protocol(ib, orc, desc, x0, splitting, dt, 91.0 if j % 2 else 1.0, 10, ["configuration", "full"][j % 2],
                          seed=60 + j)


def test_waterbox_protocol_matches_oracle(ib, orc, systems):
    desc, _ = systems["waterbox"]
    x0 = systems["waterbox_frames"][:2]
    for j, splitting in enumerate(["O V R V O", "V R O R V", "R V O V R"]):
        _compare_protocol(ib, orc, desc, x0, splitting, 0.002, 1.0, 6, ["configuration", "full"][j % 2], seed=80 + j,
                          round_midpoint=(j == 1))


def test_given_velocities_cache_draws_and_replica_offsets(ib, orc, systems):
    desc, x = systems["cluster_rigid"]
    cache = _jitter(x, 7, 0.0, 1)
    cache[:, :, 0] += np.arange(7)[:, None] * 1e-3  # distinguishable frames
    system = ib.engine.ParticleSystem(desc)
    prog = ib.engine.Program("V R O R V", 0.002, 1.0)
    n = 12
    res = ib.engine.run_protocol(system, prog, n, 4, KT, seed=5, replica0=100, x_cache=cache, want_final=True)
    ref = orc.particles_protocol(desc, "V R O R V", 0.002, 1.0, KT, n, 4, seed=5, replica0=100, x_cache=cache)
    np.testing.assert_allclose(res["W"].cpu().numpy().T, ref["W"], **WTOL)
    # the global replica id alone keys the streams: a shard reproduces its slice bit for bit
    part = ib.engine.run_protocol(system, prog, 5, 4, KT, seed=5, replica0=104, x_cache=cache)
    assert np.array_equal(part["W"].cpu().numpy(), res["W"].cpu().numpy()[:, 4:9])
    # explicit start velocities
    v0 = np.random.RandomState(2).randn(n, 60, 3) * 0.3
    res = ib.engine.run_protocol(system, prog, n, 3, KT, seed=5, x0=cache[np.arange(n) % 7], v0=v0)
    ref = orc.particles_protocol(desc, "V R O R V", 0.002, 1.0, KT, n, 3, seed=5, x0=cache[np.arange(n) % 7], v0=v0)
    np.testing.assert_allclose(res["W"].cpu().numpy().T, ref["W"], **WTOL)
    # empty ensemble
    res = ib.engine.run_protocol(system, prog, 0, 3, KT, seed=5, x_cache=cache)
    assert res["W"].shape == (2, 0)


def test_mixed_precision_tracks_double_within_bar(ib, systems):
    """fp32 force evaluation with fp64 state/accumulators (OpenMM "mixed") stays within 1e-5 relative of
    the all-fp64 trajectory over a short protocol (BASELINE.json's trajectory tolerance)."""
    for name, dt in (("cluster_rigid", 0.002), ("alanine_constrained", 0.002), ("waterbox", 0.002)):
        desc, x = systems[name]
        x0 = _jitter(x, 4, 0.0, 3)
        system = ib.engine.ParticleSystem(desc)
        prog = ib.engine.Program("V R O R V", dt, 1.0)
        v0 = np.random.RandomState(1).randn(*x0.shape) * np.sqrt(KT / np.asarray(desc["mass"]))[None, :, None]
        a = ib.engine.run_protocol(system, prog, 4, 5, KT, seed=1, x0=x0, v0=v0, want_final=True, marginal="full")
        b = ib.engine.run_protocol(system, prog, 4, 5, KT, seed=1, x0=x0, v0=v0, want_final=True, marginal="full",
                                   precision="mixed", out={})
        xa, xb = a["x_final"].cpu().numpy(), b["x_final"].cpu().numpy()
        # noise: mixed draws fp32 normals (different last bits) -> compare displacement scale
        assert np.abs(xa - xb).max() < 1e-4, name


MIXED_CASES = [("cluster_rigid", "O V R V O", 0.002), ("cluster_rigid", "V R O R V", 0.004), ("cluster_flexible", "V R O R V", 0.0005),
               ("alanine_constrained", "V R O R V", 0.002), ("alanine_constrained", "V0 V1 R O R V1 V0", 0.002),
               ("alanine_hmr", "O V R V O", 0.003), ("waterbox", "O V R V O", 0.002), ("waterbox", "V R O R V", 0.002)]


@pytest.mark.parametrize("name,splitting,dt", MIXED_CASES)
def test_mixed_precision_matches_mixed_oracle(ib, orc, systems, name, splitting, dt):
    """LSI_PRECISION_MIXED (the path every C3-C5 throughput number runs on; how the reference configures OpenMM's CUDA
    platform, testsystems/configuration.py:42-45) against the oracle's mixed restatement fed the same Philox words:
    fp32 Box-Muller for the O noise, fp32 pair / bonded arithmetic with energies summed in fp64, fp64 state and
    constraints.  The two differ in summation order, rsqrt.approx vs 1/sqrtf and libm-vs-CUDA logf / sincospif last
    ulps -- all at the fp32 rounding level.  Start: thermally equilibrated configurations.  Bar: works and heats within
    1e-5 of the energy scale |PE| + 3/2 N kT, coordinates within 1e-5 of their own scale (BASELINE.json's trajectory
    tolerance)."""
    desc, x = systems[name]
    n, n_steps = 4, 8
    x0 = _equilibrated(ib, desc, systems["waterbox_frames"] if name == "waterbox" else x, n, dt, seed=21)
    system = ib.engine.ParticleSystem(desc)
    prog = ib.engine.Program(splitting, dt, 1.0)
    res = ib.engine.run_protocol(system, prog, n, n_steps, KT, seed=11, replica0=300, x0=x0, precision="mixed", want_heat=True,
                                 want_final=True, out={})
    ref = orc.particles_protocol(desc, splitting, dt, 1.0, KT, n, n_steps, seed=11, replica0=300, x0=x0, mixed=True)
    e_scale = abs(orc.particles_energy_forces(desc, x0[0])[0]) + 1.5 * desc["n_atoms"] * KT  # kJ/mol
    dW = np.abs(res["W"].cpu().numpy().T - ref["W"]).max() * KT
    assert dW < 1e-5 * e_scale, (name, splitting, dW, e_scale)
    dQ = np.abs(res["heat"].cpu().numpy().T - ref["Q"]).max()
    assert dQ < 1e-5 * e_scale, (name, splitting, dQ)
    np.testing.assert_allclose(res["x_final"].cpu().numpy(), ref["x_final"], rtol=1e-5, atol=2e-6, err_msg=name)
    np.testing.assert_allclose(res["v_final"].cpu().numpy(), ref["v_final"], rtol=1e-4, atol=5e-4, err_msg=name)
    # and the mixed energies / forces themselves against the oracle's fp32 restatement
    em, fm = system.energy_forces(x0, precision="mixed")
    for k in range(n):
        e_ref, f_ref = orc.particles_energy_forces(desc, x0[k], force_fp32=True)
        assert abs(em[k].item() - e_ref) < 2e-6 * e_scale
        np.testing.assert_allclose(fm[k].cpu().numpy(), f_ref, rtol=1e-4, atol=3e-6 * np.abs(f_ref).max())


def _equilibrated(ib, desc, x, n, dt, seed, steps=3000):
    """n decorrelated start configurations: VRORV Langevin dynamics (gamma = 5 / ps) from jittered copies of x."""
    system = ib.engine.ParticleSystem(desc)
    x0 = x[np.arange(n) % len(x)] if x.ndim == 3 else np.repeat(x[None], n, axis=0)
    for k, (h, g, ns) in enumerate([(0.25 * dt, 50.0, 400), (dt, 5.0, steps)]):
        res = ib.engine.run_protocol(system, ib.engine.Program("V R O R V", h, g), n, ns, KT, n_legs=1, seed=seed + k, x0=x0,
                                     want_final=True, want_moments=False, out={})
        x0 = res["x_final"].cpu().numpy()
    assert np.isfinite(x0).all()
    return x0


LONG_CASES = [("cluster_rigid", "O V R V O", 0.004, 512, 1000), ("alanine_constrained", "V R O R V", 0.002, 512, 1000),
              ("alanine_constrained", "V0 V1 R O R V1 V0", 0.002, 256, 1000), ("waterbox", "V R O R V", 0.002, 256, 1000)]


@pytest.mark.parametrize("name,splitting,dt,n,n_steps", LONG_CASES)
def test_work_distribution_at_production_length_matches_oracle(ib, orc, systems, name, splitting, dt, n, n_steps):
    """One test per molecular config at the protocol length of the reference's production runs (1000 steps per leg,
    evaluation/collect_near_eq_samples.py:24) and at the production constraint tolerance 1e-8: trajectories of the kernel
    and of the oracle decorrelate after a few hundred steps (chaos amplifies last-ulp differences; Newton star-SHAKE vs
    Gauss-Seidel), so what must agree are the work DISTRIBUTIONS.  Same start configurations and Philox streams in
    both; DeltaF_neq, <W_F>, <W_R> within 3 standard errors, work variances within 25 %; for fp64 and for mixed."""
    from integrator_benchmark_b200.evaluation import estimate_nonequilibrium_free_energy
    desc, x = systems[name]
    desc = dict(desc, constraint_tolerance=1e-8)
    orc.set_threads()
    x0 = _equilibrated(ib, desc, systems["waterbox_frames"] if name == "waterbox" else x, n, dt, seed=21)
    ref = orc.particles_protocol(desc, splitting, dt, 1.0, KT, n, n_steps, seed=77, replica0=5000, x0=x0)
    wF, wR = ref["W"][:, 0], ref["W"][:, 1]
    assert np.isfinite(ref["W"]).all()
    dF_ref, var_ref = estimate_nonequilibrium_free_energy(wF, wR)
    system = ib.engine.ParticleSystem(desc)
    prog = ib.engine.Program(splitting, dt, 1.0)
    for precision in ("double", "mixed"):
        res = ib.engine.run_protocol(system, prog, n, n_steps, KT, seed=77, replica0=5000, x0=x0, precision=precision, out={})
        W = res["W"].cpu().numpy()
        assert np.isfinite(W).all()
        dF, var = estimate_nonequilibrium_free_energy(W[0], W[1])
        assert abs(dF - dF_ref) < 3.0 * np.sqrt(var + var_ref), (name, precision, dF, dF_ref, var, var_ref)
        for a, b in ((W[0], wF), (W[1], wR)):
            se = np.sqrt(a.var() / n + b.var() / n)
            assert abs(a.mean() - b.mean()) < 3.0 * se, (name, precision, a.mean(), b.mean(), se)
            assert 0.75 < a.var() / b.var() < 1.0 / 0.75, (name, precision, a.var(), b.var())
        # the first steps are still the same trajectory: work after 1/50 of the protocol agrees replica by replica
        short = ib.engine.run_protocol(system, prog, n, n_steps // 50, KT, seed=77, replica0=5000, x0=x0, precision=precision, out={})
        ref_s = orc.particles_protocol(desc, splitting, dt, 1.0, KT, n, n_steps // 50, seed=77, replica0=5000, x0=x0,
                                       mixed=(precision == "mixed"))
        # fp64: solver differences at tolerance 1e-8 (Newton vs Gauss-Seidel) times forces of 1e3-1e4 kJ/mol/nm;
        # mixed: 1e-5 of the energy scale, as in test_mixed_precision_matches_mixed_oracle
        e_scale = max(abs(orc.particles_energy_forces(desc, x0[0])[0]), 100.0)
        tol = 5e-5 if precision == "double" else 1e-5 * e_scale / KT
        assert np.abs(short["W"].cpu().numpy().T - ref_s["W"]).max() < tol, (name, precision)


def test_constraints_hold_after_protocol(ib, systems):
    desc, x = systems["cluster_rigid"]
    system = ib.engine.ParticleSystem(desc)
    res = ib.engine.run_protocol(system, ib.engine.Program("O V R V O", 0.004, 1.0), 64, 50, KT, seed=3,
                                 x0=_jitter(x, 64, 0.0, 0), want_final=True)
    xf = res["x_final"].cpu().numpy().reshape(64, 20, 3, 3)
    vf = res["v_final"].cpu().numpy().reshape(64, 20, 3, 3)
    for a, b, d in ((0, 1, 0.09572), (0, 2, 0.09572), (1, 2, desc["settle_hh"])):
        r = xf[:, :, a] - xf[:, :, b]
        assert np.abs(np.linalg.norm(r, axis=-1) - d).max() < 1e-10
        assert np.abs(np.einsum("rwc,rwc->rw", r, vf[:, :, a] - vf[:, :, b])).max() < 1e-9
    desc, x = systems["alanine_constrained"]
    system = ib.engine.ParticleSystem(desc)
    res = ib.engine.run_protocol(system, ib.engine.Program("V R O R V", 0.002, 1.0), 32, 50, KT, seed=3,
                                 x0=_jitter(x, 32, 0.0, 0), want_final=True)
    xf = res["x_final"].cpu().numpy()
    for (i, j), d in zip(desc["constraint_atoms"], desc["constraint_dist"]):
        rel = np.abs(np.linalg.norm(xf[:, i] - xf[:, j], axis=-1) - d) / d
        assert rel.max() < 1e-7  # tolerance 1e-8 on r^2/d^2 - 1, as OpenMM defines it


def test_continuous_splitting_and_xml_round_trip_on_gpu(ib, orc, systems):
    """ContinuousLangevinSplittingIntegrator (explicit substep fractions, langevin.py:223-408) through both
    kernels, and a system that went through the OpenMM System XML layout gives identical energies."""
    from oracle import splitting as sp
    from integrator_benchmark_b200.integrators import ContinuousLangevinSplittingIntegrator
    from integrator_benchmark_b200.serialization import load_system_xml, system_to_xml
    from integrator_benchmark_b200 import unit
    frac = [("V", 0.25), ("R", 0.5), ("O", 0.3), ("V", 0.5), ("O", 0.7), ("R", 0.5), ("V", 0.25)]
    integ = ContinuousLangevinSplittingIntegrator(frac, collision_rate=5.0 / unit.picosecond, timestep=2.0 * unit.femtosecond,
                                                  measure_shadow_work=True)
    ref_prog = sp.compile_continuous(frac, 0.002, 5.0)
    desc, x = systems["cluster_rigid"]
    x0 = _jitter(x, 3, 0.0, 1)
    res = ib.engine.run_protocol(ib.engine.ParticleSystem(desc), integ.program, 3, 7, KT, seed=3, replica0=5, x0=x0,
                                 want_direct=True, want_final=True)
    ref = orc.particles_protocol(desc, None, 0.002, 5.0, KT, 3, 7, seed=3, replica0=5, x0=x0, want_direct=True, program=ref_prog)
    np.testing.assert_allclose(res["W"].cpu().numpy().T, ref["W"], **WTOL)
    np.testing.assert_allclose(res["W_direct"].cpu().numpy().T, ref["Wdirect"], **WTOL)
    np.testing.assert_allclose(res["x_final"].cpu().numpy(), ref["x_final"], **XTOL)
    # toy kernel (interpreter path: fractions differ per substep)
    cache = orc.sample_equilibrium_scalar("double_well", 512, 50, seed=3)
    toy_prog = ib.engine.Program(frac, 0.3, 10.0)
    tres = ib.engine.run_protocol(ib.engine.ScalarSystem("double_well", 10.0), toy_prog, 777, 9, 1.0, seed=4, x_cache=cache)
    tref = orc.toy_protocol("double_well", None, 0.3, 10.0, 777, 9, seed=4, x_cache=cache, program=sp.compile_continuous(frac, 0.3, 10.0))
    np.testing.assert_allclose(tres["W"].cpu().numpy().T, tref["W"], rtol=1e-7, atol=1e-9)
    # XML round trip
    for name in ("cluster_rigid", "alanine_constrained", "waterbox"):
        d, xx = systems[name]
        back = load_system_xml(system_to_xml(d))
        back["constraint_tolerance"] = d.get("constraint_tolerance", 1e-8)
        xs = _jitter(xx, 2, 0.001, 2)
        e0, f0 = ib.engine.ParticleSystem(d).energy_forces(xs)
        e1, f1 = ib.engine.ParticleSystem(back).energy_forces(xs)
        np.testing.assert_allclose(e1.cpu().numpy(), e0.cpu().numpy(), rtol=1e-13)
        np.testing.assert_allclose(f1.cpu().numpy(), f0.cpu().numpy(), rtol=1e-12, atol=1e-9)

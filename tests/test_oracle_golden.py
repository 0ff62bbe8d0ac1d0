"""The CPU oracle against fixtures produced by the reference's own code
(tests/golden/make_golden.py) and against published known answers."""
import json
import os

import numpy as np
import pytest

from oracle import oracle as orc
from oracle import splitting as sp

# numba closures (numba_integrators.py) expressed as splitting strings:
# two half-O draws per step in baoab/aboba, two half-R in vvvr, one full V in orvro
SCHEME_STRING = {"vvvr": "O V R R V O", "orvro": "O R V R O", "baoab": "V R O O R V", "aboba": "R V O O V R"}


def test_philox_known_answers():
    # Random123 kat_vectors, philox4x32 10 rounds
    assert orc.philox4x32_10([0, 0, 0, 0], [0, 0]) == [0x6627E8D5, 0xE169C58D, 0xBC57AC4C, 0x9B00DBD8]
    assert orc.philox4x32_10([0xFFFFFFFF] * 4, [0xFFFFFFFF] * 2) == [0x408F276D, 0x41C83B0E, 0xA20BC7C6, 0x6D5451FD]
    assert orc.philox4x32_10([0x243F6A88, 0x85A308D3, 0x13198A2E, 0x03707344],
                             [0xA4093822, 0x299F31D0]) == [0xD16CFE09, 0x94FDCCEB, 0x5001E420, 0x24126EA1]


def test_normals_are_standard():
    z = np.concatenate([orc.normals4(7, r, 0, 0, 0) for r in range(5000)])
    assert abs(z.mean()) < 0.03 and abs(z.std() - 1.0) < 0.03
    assert abs(np.mean(z ** 4) - 3.0) < 0.2


def test_toy_potentials_known_values():
    # partition function of the double well on [-2.25, 2.25]: notebooks/1D figures, updated.ipynb:415
    xs = np.linspace(-2.25, 2.25, 200001)
    u = np.array([orc.toy_potential("double_well", x) for x in xs[::20]])
    Z = np.trapezoid(np.exp(-u), xs[::20])
    assert abs(Z - 4.4848163908) < 1e-6
    for x in [-1.3, -0.2, 0.0, 0.7, 1.9]:
        for kind in ["quartic", "double_well"]:
            h = 1e-6
            fd = -(orc.toy_potential(kind, x + h) - orc.toy_potential(kind, x - h)) / (2 * h)
            assert abs(fd - orc.toy_force(kind, x)) < 1e-6 * max(1.0, abs(fd))


def _toy_cases(golden_dir):
    z = np.load(os.path.join(golden_dir, "toy_numba.npz"))
    return z, json.loads(str(z["meta"]))


def test_oracle_matches_reference_numba_closures(golden_dir):
    """oracle (replayed noise) == reference numba_integrators.py bodies, all 4 schemes x 2 systems."""
    z, meta = _toy_cases(golden_dir)
    assert len(meta) == 32
    for c in meta:
        k = c["key"]
        n = c["n_steps"]
        noise = []
        for blk in range((2 * (n - 1) + 3) // 4):
            noise.extend(orc.normals4(c["seed"], c["replica"], blk, 0, 0))
        tx, tv, Q, tw = orc.toy_replay(c["system"], SCHEME_STRING[c["scheme"]], c["dt"], c["gamma"], c["x0"], c["v0"],
                                       n - 1, noise, mass=c["mass"], kT=1.0 / c["beta"])
        np.testing.assert_allclose(tx, z[k + "_xs"], rtol=1e-12, atol=1e-13)
        np.testing.assert_allclose(tv, z[k + "_vs"], rtol=1e-12, atol=1e-13)
        np.testing.assert_allclose(tw, z[k + "_W"], rtol=1e-10, atol=1e-11)
        np.testing.assert_allclose(Q, float(z[k + "_Q"]), rtol=1e-10, atol=1e-11)


def test_oracle_philox_protocol_equals_replay(golden_dir):
    """the Philox-driven protocol consumes the scalar stream exactly as the replay does."""
    z, meta = _toy_cases(golden_dir)
    for c in meta[::3]:
        n = c["n_steps"]
        out = orc.toy_protocol(c["system"], SCHEME_STRING[c["scheme"]], c["dt"], c["gamma"], 1, n - 1,
                               mass=c["mass"], kT=1.0 / c["beta"], seed=c["seed"], replica0=c["replica"],
                               x0=[c["x0"]], v0=[c["v0"]], n_legs=1, traces=True)
        np.testing.assert_allclose(out["trace_x"][0, 0], z[c["key"] + "_xs"], rtol=1e-12, atol=1e-13)
        np.testing.assert_allclose(out["trace_w"][0, 0], z[c["key"] + "_W"], rtol=1e-10, atol=1e-11)
        np.testing.assert_allclose(out["W"][0, 0], z[c["key"] + "_W"][-1], rtol=1e-10, atol=1e-11)


# ----------------------------------------------------------------------------- recorded step programs
def _expected_ops_from_recorded(entry):
    """Decode the recorded CustomIntegrator program back into (kind, group, frac, a, b)."""
    import re
    ops = []
    g = entry["globals"]
    for op in entry["program"]:
        if op[0] != "perdof":
            continue
        var, expr = op[1], op[2]
        if var == "x" and "v" in expr:
            m = re.match(r"x \+ \(\(dt ([/*]) ([0-9.e+-]+)\) \* v\)", expr)
            val = float(m.group(2))
            ops.append((sp.R, -1, 1.0 / val if m.group(1) == "/" else val, 0.0, 0.0))
        elif var == "v" and "gaussian" in expr:
            m = re.match(r"\(([0-9.e+-]+|a) \* v\) \+ \(([0-9.e+-]+|b) \* sqrt", expr)
            a = g["a"] if m.group(1) == "a" else float(m.group(1))
            b = g["b"] if m.group(2) == "b" else float(m.group(2))
            ops.append((sp.O, -1, None, a, b))
        elif var == "v" and re.search(r"\* f\d* / m", expr):
            m = re.match(r"v \+ \(?\(dt ([/*]) ([0-9.e+-]+)\) \* f(\d*) / m\)?", expr)
            val = float(m.group(2))
            grp = int(m.group(3)) if m.group(3) else -1
            ops.append((sp.V, grp, (1.0 / val if val != 0 else float("inf")) if m.group(1) == "/" else val, 0.0, 0.0))
    return ops


def test_splitting_compiler_matches_reference_constructor(golden_dir):
    data = json.load(open(os.path.join(golden_dir, "programs.json")))
    n_checked = 0
    for e in data["entries"]:
        kw = e["kwargs"]
        args = dict(dt=kw["timestep"], gamma=kw["collision_rate"],
                    override_splitting_checks=kw.get("override_splitting_checks", False),
                    measure_shadow_work=kw.get("measure_shadow_work", False),
                    measure_heat=kw.get("measure_heat", True))
        fn = sp.compile_splitting if e["kind"] == "discrete" else sp.compile_continuous
        splitting = e["splitting"] if e["kind"] == "discrete" else [tuple(t) for t in e["splitting"]]
        if "error" in e:
            with pytest.raises(Exception) as ei:
                fn(splitting, **args)
            assert type(ei.value).__name__ == e["error"], (splitting, kw)
            continue
        expected = _expected_ops_from_recorded(e)
        if any(op[0] == sp.V and op[2] == float("inf") for op in expected):
            # reference quirk Q2 ("dt / 0" for a lone V<g> token): the oracle refuses instead
            with pytest.raises(ZeroDivisionError):
                fn(splitting, **args)
            continue
        prog = fn(splitting, **args)
        assert len(prog.ops) == len(expected), (splitting, prog.ops, expected)
        for got, exp in zip(prog.ops, expected):
            assert got[0] == exp[0] and got[1] == exp[1], (splitting, got, exp)
            if exp[2] is not None:
                assert got[2] == pytest.approx(exp[2], rel=1e-15)
            if exp[0] == sp.O:
                assert got[3] == pytest.approx(exp[3], rel=1e-15) and got[4] == pytest.approx(exp[4], rel=1e-15)
        n_checked += 1
    assert n_checked > 60


# ----------------------------------------------------------------------------- particle oracle
def chain_desc(z, constrained, meta):
    bonds = z["bonds"]
    cons = meta["constraints"] if constrained else []
    keep = [b for b in bonds if (int(b[0]), int(b[1])) not in [(c[0], c[1]) for c in cons]]
    return dict(n_atoms=len(z["masses"]), mass=z["masses"], restraint_k=float(z["K"]),
                bond_atoms=[(int(b[0]), int(b[1])) for b in keep], bond_params=[(b[2], b[3]) for b in keep],
                groups=dict(restraint=0, bonds=1, nonbonded=0, angles=1, torsions=1),
                constraint_atoms=[(c[0], c[1]) for c in cons], constraint_dist=[c[2] for c in cons],
                constraint_tolerance=1e-13)


def test_particle_oracle_matches_interpreted_reference_programs(golden_dir):
    """The reference's own CustomIntegrator step programs (recorded from langevin.py and interpreted
    with numpy in make_golden.py) against the oracle's substep loop: R/V/V{g}/O order, constraint
    placement, heat and shadow-work bookkeeping (langevin.py:124-175)."""
    z = np.load(os.path.join(golden_dir, "lsi_trajectories.npz"))
    meta = json.loads(str(z["meta"]))
    assert len(meta) == 12
    for m in meta:
        k = m["key"]
        desc = chain_desc(z, m["constrained"], m)
        out = orc.particles_protocol(desc, m["splitting"], m["timestep"], m["collision_rate"], m["kT"], 1, m["n_steps"],
                                     seed=m["seed"], replica0=m["replica"], x0=z[k + "_x0"][None], v0=z[k + "_v0"][None],
                                     n_legs=1, traces=True, want_direct=True)
        np.testing.assert_allclose(out["trace_x"][0, 0, 1:], z[k + "_x"], rtol=1e-9, atol=1e-11, err_msg=m["splitting"])
        np.testing.assert_allclose(out["trace_v"][0, 0, 1:], z[k + "_v"], rtol=1e-8, atol=1e-10, err_msg=m["splitting"])
        np.testing.assert_allclose(out["trace_pe"][0, 0, 1:], z[k + "_pe"], rtol=1e-9, atol=1e-9)
        np.testing.assert_allclose(out["trace_ke"][0, 0, 1:], z[k + "_ke"], rtol=1e-9, atol=1e-9)
        np.testing.assert_allclose(out["trace_heat"][0, 0, 1:], z[k + "_heat"], rtol=1e-8, atol=1e-8)
        np.testing.assert_allclose(out["trace_wdir"][0, 0, 1:], z[k + "_shadow_work"], rtol=1e-8, atol=1e-8)
        # energy bookkeeping identity of the reference's tests/test_shadow_work.py:30-42
        dE = (out["trace_pe"][0, 0] + out["trace_ke"][0, 0]) - (out["trace_pe"][0, 0, 0] + out["trace_ke"][0, 0, 0])
        np.testing.assert_allclose(dE - out["trace_heat"][0, 0], out["trace_wdir"][0, 0], atol=1e-6)


def test_particle_oracle_forces_are_gradients():
    rng = np.random.RandomState(3)
    N = 9
    x = rng.randn(N, 3) * 0.12 + np.repeat(np.array([[0, 0, 0], [0.35, 0, 0.1], [0, 0.4, -0.1]]), 3, axis=0)
    desc = dict(n_atoms=N, mass=np.ones(N), charge=np.tile([-0.834, 0.417, 0.417], 3),
                sigma=np.tile([0.315075, 0.1, 0.1], 3), epsilon=np.tile([0.635968, 0.05, 0.05], 3),
                exclusions=[(0, 1), (0, 2), (1, 2), (3, 4), (3, 5), (4, 5), (6, 7), (6, 8), (7, 8)], restraint_k=1.5,
                bond_atoms=[(0, 1), (0, 2)], bond_params=[(0.09572, 462750.4)] * 2,
                angle_atoms=[(1, 0, 2)], angle_params=[(1.82421813, 836.8)],
                torsion_atoms=[(1, 0, 3, 4), (2, 0, 3, 5)], torsion_params=[(3, 0.3, 5.0), (2, 3.14159, 2.5)],
                exception_atoms=[(1, 4)], exception_params=[(0.1, 0.25, 0.3)])
    for extra in [{}, dict(nonbonded_method="CutoffPeriodic", cutoff=0.45, box=(1.1, 1.2, 1.3))]:
        d = dict(desc, **extra)
        e, f = orc.particles_energy_forces(d, x)
        h = 1e-6
        for i in range(N):
            for c in range(3):
                xp, xm = x.copy(), x.copy()
                xp[i, c] += h
                xm[i, c] -= h
                fd = -(orc.particles_energy_forces(d, xp)[0] - orc.particles_energy_forces(d, xm)[0]) / (2 * h)
                assert abs(fd - f[i, c]) < 2e-5 * max(1.0, abs(fd)), (extra, i, c, fd, f[i, c])


def test_settle_equals_converged_shake():
    """analytic SETTLE (positions) and the triangle velocity projection solve the same equations as
    SHAKE / RATTLE iterated to machine precision on the three edges."""
    rng = np.random.RandomState(5)
    oh, hh, mO, mH = 0.09572, 0.15139, 15.99943, 1.007947
    s, c = hh / 2, np.sqrt(oh ** 2 - (hh / 2) ** 2)
    tri = np.array([[0, 0, 0], [s, c, 0], [-s, c, 0]])
    xs = []
    for w in range(4):
        R = np.linalg.qr(rng.randn(3, 3))[0]
        xs.append(tri @ R.T + rng.randn(3) * 0.4)
    x0 = np.concatenate(xs)
    n = len(x0)
    mass = np.tile([mO, mH, mH], 4)
    settle = dict(n_atoms=n, mass=mass, settle_atoms=[(3 * w, 3 * w + 1, 3 * w + 2) for w in range(4)], settle_oh=oh, settle_hh=hh)
    cons = []
    for w in range(4):
        cons += [(3 * w, 3 * w + 1, oh), (3 * w, 3 * w + 2, oh), (3 * w + 1, 3 * w + 2, hh)]
    shake = dict(n_atoms=n, mass=mass, constraint_atoms=[(a, b) for a, b, _ in cons], constraint_dist=[d for _, _, d in cons],
                 constraint_tolerance=1e-15)
    x1 = x0 + 0.006 * rng.randn(n, 3)
    a, b = orc.constrain_positions(settle, x0, x1), orc.constrain_positions(shake, x0, x1)
    np.testing.assert_allclose(a, b, atol=1e-13)
    v = rng.randn(n, 3)
    np.testing.assert_allclose(orc.constrain_velocities(settle, a, v), orc.constrain_velocities(shake, a, v), atol=1e-12)


def test_waterbox_fixture_geometry(golden_dir):
    """data/tiny_waterbox_constrained_samples.h5: 4 frames x 102 rigid TIP3P waters."""
    z = np.load(os.path.join(golden_dir, "waterbox_frames.npz"))
    xyz = z["xyz"].astype(np.float64)
    assert xyz.shape == (4, 306, 3)
    oh1 = np.linalg.norm(xyz[:, 1::3] - xyz[:, 0::3], axis=-1)
    hh = np.linalg.norm(xyz[:, 1::3] - xyz[:, 2::3], axis=-1)
    assert np.abs(oh1 - 0.09572).max() < 2e-6 and np.abs(hh - 0.15139).max() < 2e-6
    np.testing.assert_allclose(z["potential_energy"], [-4136.6406, -4029.0005, -4030.8572, -4068.7595], atol=1e-3)


def test_waterbox_fixture_energies_track_the_stored_pme_energies(golden_dir):
    """oracle energies (cutoff 0.6 nm + reaction field, switched Lennard-Jones, no LJ tail correction) of the reference's
    four stored frames against the PME + dispersion-correction energies stored with them: 94.7-95.3 %, same order."""
    import sys
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    from integrator_benchmark_b200.testsystems import systems as ts
    z = np.load(os.path.join(golden_dir, "waterbox_frames.npz"))
    desc, _ = ts.tiny_waterbox()
    e = np.array([orc.particles_energy_forces(desc, z["xyz"][k].astype(np.float64))[0] for k in range(4)])
    ratio = e / z["potential_energy"]
    assert np.all(ratio > 0.94) and np.all(ratio < 0.96), ratio
    assert np.argsort(e).tolist() == np.argsort(z["potential_energy"]).tolist()


def test_switched_lennard_jones_of_the_oracle(golden_dir):
    """OpenMM's switching function (openmmtools WaterBox: the last 1.5 A before the cutoff): S(0) = 1, S(1) = 0 with
    zero slope at both ends, and the oracle's forces are the gradient of its switched energy on a real frame."""
    import sys
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    from integrator_benchmark_b200.testsystems import systems as ts
    d = dict(n_atoms=2, mass=[16.0, 16.0], charge=[0.0, 0.0], sigma=[0.3150752406575124] * 2, epsilon=[0.635968] * 2,
             nonbonded_method="CutoffPeriodic", cutoff=0.6, switch_distance=0.45, box=(1.5, 1.5, 1.5))

    def e_of(r, dd=d):
        return orc.particles_energy_forces(dd, np.array([[0.2, 0.2, 0.2], [0.2 + r, 0.2, 0.2]]))

    lj = lambda r: 4 * 0.635968 * ((0.3150752406575124 / r) ** 12 - (0.3150752406575124 / r) ** 6)  # noqa: E731
    assert e_of(0.44)[0] == pytest.approx(lj(0.44), rel=1e-13)          # untouched below the switching distance
    x = (0.525 - 0.45) / 0.15
    assert e_of(0.525)[0] == pytest.approx(lj(0.525) * (1 - 6 * x ** 5 + 15 * x ** 4 - 10 * x ** 3), rel=1e-12)
    assert abs(e_of(0.6 - 1e-9)[0]) < 1e-12 and e_of(0.6 + 1e-12)[0] == 0.0
    assert e_of(0.6 - 1e-9, dict(d, switch_distance=0.0))[0] == pytest.approx(lj(0.6), rel=1e-6)  # the truncation's jump
    for r in (0.47, 0.53, 0.59):
        h = 1e-6
        assert e_of(r)[1][1, 0] == pytest.approx(-(e_of(r + h)[0] - e_of(r - h)[0]) / (2 * h), rel=1e-6, abs=1e-9)
    z = np.load(os.path.join(golden_dir, "waterbox_frames.npz"))
    desc, _ = ts.tiny_waterbox()
    x0 = z["xyz"][1].astype(np.float64)
    e0, f = orc.particles_energy_forces(desc, x0)
    u = np.random.RandomState(0).randn(*x0.shape)
    u /= np.linalg.norm(u)
    h = 1e-6
    fd = (orc.particles_energy_forces(desc, x0 + h * u)[0] - orc.particles_energy_forces(desc, x0 - h * u)[0]) / (2 * h)
    assert -(f * u).sum() == pytest.approx(fd, rel=1e-6)


def test_mixed_precision_oracle_tracks_the_fp64_oracle():
    """particles_forces_f32 (fp32 pair / bonded arithmetic, energies summed in fp64) against the fp64 restatement:
    energies within 2e-6 relative of the energy scale, forces within 1e-5 of the largest force, every reference system."""
    import sys
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    from integrator_benchmark_b200.testsystems import systems as ts
    cases = {"cluster": ts.water_cluster(20, seed=1), "flexible": ts.water_cluster(20, constrained=False, seed=1),
             "alanine": ts.alanine_dipeptide(), "box": ts.tiny_waterbox()}
    for name, (desc, x) in cases.items():
        x = x + 0.003 * np.random.RandomState(0).randn(*x.shape)
        for group in (-1, 0, 1):
            e64, f64 = orc.particles_energy_forces(desc, x, group=group)
            e32, f32 = orc.particles_energy_forces(desc, x, group=group, force_fp32=True)
            assert abs(e64 - e32) < 4e-6 * max(abs(e64), 100.0), (name, group, e64, e32)
            assert np.abs(f64 - f32).max() < 1e-5 * max(np.abs(f64).max(), 1.0), (name, group)
    # and a short mixed protocol stays within 1e-5 of the energy scale of the fp64 one (same fp64 noise)
    desc, x = cases["cluster"]
    kT = 8.3144621e-3 * 298.0
    a = orc.particles_protocol(desc, "V R O R V", 0.002, 1.0, kT, 2, 6, seed=3, x0=np.repeat(x[None], 2, axis=0))
    b = orc.particles_protocol(desc, "V R O R V", 0.002, 1.0, kT, 2, 6, seed=3, x0=np.repeat(x[None], 2, axis=0), force_fp32=True)
    assert np.abs(a["W"] - b["W"]).max() * kT < 1e-5 * abs(orc.particles_energy_forces(desc, x)[0]) + 1e-3


def test_tip3p_dimer_minimum_matches_the_published_value():
    """An external pin for the water force field (the TIP3P constants are restated from the literature, there is
    no OpenMM here to compare with): the global minimum of the water dimer for the constants in testsystems/systems.py is the one
    Jorgensen et al. publish for TIP3P (J. Chem. Phys. 79, 926 (1983), table II: -6.50 kcal/mol at r_OO = 2.74 A).  A wrong charge,
    sigma, epsilon or geometry shows here: the energy scales with q^2 (q_H 0.417 -> 0.41 moves it by 3 %)."""
    from importlib import import_module
    from scipy.optimize import minimize
    ts = import_module("integrator-benchmark_b200.testsystems.systems")
    desc, _ = ts.water_cluster(2, K=0.0, constrained=True, seed=0)
    geo = ts._water_geometry()

    def rot(a, b, c):
        ca, sa, cb, sb, cc, sc = np.cos(a), np.sin(a), np.cos(b), np.sin(b), np.cos(c), np.sin(c)
        Rz = np.array([[ca, -sa, 0], [sa, ca, 0], [0, 0, 1]])
        Ry = np.array([[cb, 0, sb], [0, 1, 0], [-sb, 0, cb]])
        Rx = np.array([[1, 0, 0], [0, cc, -sc], [0, sc, cc]])
        return Rz @ Ry @ Rx

    def energy(p):
        x = np.vstack([geo @ rot(*p[0:3]).T, geo @ rot(*p[3:6]).T + np.array([p[6], 0.0, 0.0])])
        return orc.particles_energy_forces(desc, x)[0]
    rng = np.random.RandomState(0)
    best = min((minimize(energy, np.concatenate([rng.uniform(-np.pi, np.pi, 6), [0.28]]), method="Nelder-Mead",
                         options={"xatol": 1e-6, "fatol": 1e-9, "maxiter": 6000}) for _ in range(24)), key=lambda r: r.fun)
    e_kcal, r_oo = best.fun / 4.184, abs(best.x[6])
    assert -6.60 < e_kcal < -6.45, e_kcal      # published -6.50 (this parameter set: -6.54)
    assert 0.272 < r_oo < 0.277, r_oo          # published 2.74 A (here 2.747 A)

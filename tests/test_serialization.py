"""OpenMM System XML reader / writer (scope row 8f-3) and the Experiment pickle layout.  CPU only."""
import os
import pickle
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

HAND_WRITTEN = """<?xml version="1.0" ?>
<System openmmVersion="7.2" type="System" version="1">
	<PeriodicBoxVectors>
		<A x="1.5" y="0" z="0"/>
		<B x="0" y="1.5" z="0"/>
		<C x="0" y="0" z="1.5"/>
	</PeriodicBoxVectors>
	<Particles>
		<Particle mass="15.99943"/>
		<Particle mass="1.007947"/>
		<Particle mass="1.007947"/>
		<Particle mass="12.01"/>
		<Particle mass="1.008"/>
	</Particles>
	<Constraints>
		<Constraint d=".09572" p1="0" p2="1"/>
		<Constraint d=".09572" p1="0" p2="2"/>
		<Constraint d=".15139006545247014" p1="1" p2="2"/>
		<Constraint d=".109" p1="3" p2="4"/>
	</Constraints>
	<Forces>
		<Force forceGroup="1" type="HarmonicBondForce" usesPeriodic="0" version="2">
			<Bonds>
				<Bond d=".1526" k="259408" p1="0" p2="3"/>
			</Bonds>
		</Force>
		<Force forceGroup="1" type="HarmonicAngleForce" usesPeriodic="0" version="2">
			<Angles>
				<Angle a="1.911" k="418.4" p1="0" p2="3" p3="4"/>
			</Angles>
		</Force>
		<Force forceGroup="1" type="PeriodicTorsionForce" usesPeriodic="0" version="2">
			<Torsions>
				<Torsion k=".6508" p1="1" p2="0" p3="3" p4="4" periodicity="3" phase="0"/>
			</Torsions>
		</Force>
		<Force alpha="0" cutoff=".6" dispersionCorrection="1" ewaldTolerance=".0005" forceGroup="0" method="2" nx="0" ny="0" nz="0" recipForceGroup="-1" rfDielectric="78.3" switchingDistance="-1" type="NonbondedForce" useSwitchingFunction="0" version="2">
			<GlobalParameters/>
			<ParticleOffsets/>
			<ExceptionOffsets/>
			<Particles>
				<Particle eps=".635968" q="-.834" sig=".3150752406575124"/>
				<Particle eps="0" q=".417" sig="1"/>
				<Particle eps="0" q=".417" sig="1"/>
				<Particle eps=".4577296" q="-.1" sig=".3399669508423535"/>
				<Particle eps=".0656888" q=".1" sig=".2649532787749369"/>
			</Particles>
			<Exceptions>
				<Exception eps="0" p1="0" p2="1" q="0" sig="1"/>
				<Exception eps="0" p1="0" p2="2" q="0" sig="1"/>
				<Exception eps="0" p1="1" p2="2" q="0" sig="1"/>
				<Exception eps=".1" p1="1" p2="4" q=".02" sig=".25"/>
			</Exceptions>
		</Force>
		<Force energy="(K/2.0) * (x^2 + y^2 + z^2);K = 1.000000;" forceGroup="0" type="CustomExternalForce" version="1">
			<PerParticleParameters/>
			<GlobalParameters/>
			<Particles>
				<Particle index="0"/>
				<Particle index="1"/>
				<Particle index="2"/>
				<Particle index="3"/>
				<Particle index="4"/>
			</Particles>
		</Force>
		<Force forceGroup="0" frequency="1" type="CMMotionRemover" version="1"/>
		<Force defaultPressure="1.01325" defaultTemperature="298" forceGroup="0" frequency="25" type="MonteCarloBarostat" version="1"/>
	</Forces>
</System>
"""


def test_reads_an_openmm_system_xml():
    from integrator_benchmark_b200.serialization import load_system_xml
    d = load_system_xml(HAND_WRITTEN)
    assert d["n_atoms"] == 5 and d["nonbonded_method"] == "CutoffPeriodic" and d["cutoff"] == 0.6 and d["box"] == (1.5, 1.5, 1.5)
    assert d["settle_atoms"] == [(0, 1, 2)] and d["settle_oh"] == pytest.approx(0.09572) and d["settle_hh"] == pytest.approx(0.15139006545)
    assert d["constraint_atoms"] == [(3, 4)] and d["constraint_dist"] == [0.109]
    assert d["exclusions"] == [(0, 1), (0, 2), (1, 2)] and d["exception_atoms"] == [(1, 4)] and d["exception_params"] == [(0.02, 0.25, 0.1)]
    assert d["bond_atoms"] == [(0, 3)] and d["angle_params"] == [(1.911, 418.4)] and d["torsion_params"] == [(3.0, 0.0, 0.6508)]
    assert d["restraint_k"] == 1.0
    assert d["groups"] == dict(bonds=1, angles=1, torsions=1, nonbonded=0, restraint=0)
    np.testing.assert_allclose(d["charge"], [-0.834, 0.417, 0.417, -0.1, 0.1])
    with pytest.raises(NotImplementedError):
        load_system_xml(HAND_WRITTEN.replace('method="2"', 'method="4"'))
    assert load_system_xml(HAND_WRITTEN.replace('method="2"', 'method="4"'), pme_as_reaction_field=True)["nonbonded_method"] == "CutoffPeriodic"
    with pytest.raises(NotImplementedError):
        load_system_xml(HAND_WRITTEN.replace('type="CMMotionRemover"', 'type="GBSAOBCForce"'))


@pytest.mark.parametrize("name", ["water_cluster", "water_cluster_flexible", "alanine", "alanine_unconstrained", "waterbox"])
def test_round_trip_of_the_reference_systems(name):
    from integrator_benchmark_b200.serialization import load_system_xml, system_to_xml
    from integrator_benchmark_b200.testsystems import systems as ts
    desc = {"water_cluster": lambda: ts.water_cluster(20, constrained=True)[0],
            "water_cluster_flexible": lambda: ts.water_cluster(20, constrained=False)[0],
            "alanine": lambda: ts.alanine_dipeptide(constrained=True)[0],
            "alanine_unconstrained": lambda: ts.alanine_dipeptide(constrained=False)[0],
            "waterbox": lambda: ts.tiny_waterbox()[0]}[name]()
    back = load_system_xml(system_to_xml(desc))
    for key in ("mass", "charge", "sigma", "epsilon"):
        np.testing.assert_array_equal(np.asarray(back[key]), np.asarray(desc[key]))
    for key in ("exclusions", "exception_atoms", "bond_atoms", "angle_atoms", "torsion_atoms", "settle_atoms", "constraint_atoms"):
        assert [tuple(int(v) for v in t) for t in back.get(key, []) or []] == [tuple(int(v) for v in t) for t in desc.get(key, []) or []], key
    for key in ("exception_params", "bond_params", "angle_params", "torsion_params", "constraint_dist"):
        np.testing.assert_allclose(np.asarray(back.get(key, []) or [], dtype=float).reshape(-1),
                                   np.asarray(desc.get(key, []) or [], dtype=float).reshape(-1), rtol=1e-15)
    assert back["nonbonded_method"] == desc["nonbonded_method"]
    assert back.get("restraint_k", 0.0) == desc.get("restraint_k", 0.0)
    if desc.get("settle_atoms"):
        assert back["settle_oh"] == pytest.approx(desc["settle_oh"], rel=1e-12) and back["settle_hh"] == pytest.approx(desc["settle_hh"], rel=1e-9)
    present = {"nonbonded": True, "bonds": bool(desc.get("bond_atoms")), "angles": bool(desc.get("angle_atoms")),
               "torsions": bool(desc.get("torsion_atoms")), "restraint": bool(desc.get("restraint_k"))}
    for k, there in present.items():  # the group of a force class without terms is not part of the file
        if there:
            assert back["groups"][k] == desc["groups"][k], k


def test_experiment_pickle_layout(tmp_path):
    """experiments/driver.py:56-61: {"result": ..., "descriptor": everything but the simulator}"""
    from integrator_benchmark_b200.experiments import Experiment, ExperimentDescriptor
    d = ExperimentDescriptor("exp", "sys", object(), "OVRVO", "O V R V O", 2.0, "full", "low", 1.0, 10, 5, 1)
    e = Experiment(d, str(tmp_path / "out.pkl"))
    e.result = {"W_shads_F": np.zeros(3), "W_shads_R": np.ones(3)}
    e.save()
    got = pickle.load(open(str(tmp_path / "out.pkl"), "rb"))
    assert set(got) == {"result", "descriptor"} and "equilibrium_simulator" not in got["descriptor"]
    assert got["descriptor"]["splitting_string"] == "O V R V O" and got["descriptor"]["h_mass_factor"] == 1
    assert "dt=2.0fs" in str(e)


# ------------------------------------------------------------------ equilibrium-sample caches
def _shuffled_zlib(a):
    import zlib
    raw = np.ascontiguousarray(a, "<f4").view(np.uint8).reshape(-1, 4).T.copy().tobytes()
    return zlib.compress(raw, 6)


def test_mdtraj_hdf5_chunks_synthetic(tmp_path):
    """byte-shuffled zlib chunks between unrelated bytes (some of them 0x78) are found, unshuffled and trimmed"""
    from importlib import import_module
    h5 = import_module("integrator-benchmark_b200.serialization.mdtraj_hdf5")
    rng = np.random.default_rng(5)
    n_atoms, n_frames, chunk_frames = 7, 5, 8
    xyz = np.zeros((chunk_frames, n_atoms, 3), np.float32)
    xyz[:n_frames] = rng.normal(size=(n_frames, n_atoms, 3))
    lengths = np.zeros((64, 3), np.float32)
    lengths[:n_frames] = 1.5
    angles = np.zeros((64, 3), np.float32)
    angles[:n_frames] = 90.0
    blob = (b"\x89HDF\r\n\x1a\n" + b"\x78\x9cnot a stream" + bytes(rng.integers(0, 256, 300, dtype=np.uint8))
            + _shuffled_zlib(lengths) + b"\x00" * 17 + _shuffled_zlib(angles) + b"x\x01x\xda"
            + _shuffled_zlib(np.arange(50, dtype=np.float32)) + _shuffled_zlib(xyz) + b"tail")
    path = tmp_path / "fake_samples.h5"
    path.write_bytes(blob)
    out = h5.load_mdtraj_hdf5(str(path), n_atoms)
    assert out["xyz"].shape == (n_frames, n_atoms, 3)
    np.testing.assert_array_equal(out["xyz"], xyz[:n_frames])
    np.testing.assert_array_equal(out["cell_lengths"], lengths[:n_frames])
    np.testing.assert_array_equal(out["cell_angles"], angles[:n_frames])
    with pytest.raises(ValueError):
        h5.load_mdtraj_hdf5(str(path), n_atoms + 1000)
    (tmp_path / "not.h5").write_bytes(b"hello")
    with pytest.raises(ValueError):
        h5.load_mdtraj_hdf5(str(tmp_path / "not.h5"), n_atoms)


REF_H5 = "/root/reference/data/tiny_waterbox_constrained_samples.h5"


@pytest.mark.skipif(not os.path.exists(REF_H5), reason="reference checkout not present")
def test_mdtraj_hdf5_reference_cache():
    """the one cache the reference ships decodes to the committed fixture frames"""
    from importlib import import_module
    h5 = import_module("integrator-benchmark_b200.serialization.mdtraj_hdf5")
    out = h5.load_mdtraj_hdf5(REF_H5, 306)
    gold = np.load(os.path.join(ROOT, "tests", "golden", "waterbox_frames.npz"))
    assert out["xyz"].shape[1:] == (306, 3) and out["xyz"].shape[0] >= 4
    np.testing.assert_array_equal(out["xyz"][:4], gold["xyz"])
    np.testing.assert_allclose(out["cell_lengths"][:4], 1.5)
    np.testing.assert_allclose(out["cell_angles"][:4], 90.0)


def test_equilibrium_cache_on_disk(tmp_path, monkeypatch):
    """get_path_to_samples / check_for_cached_samples / load_or_simulate_samples (bookkeepers.py:66-125) without a GPU:
    a stored cache is picked up and never re-simulated"""
    from importlib import import_module
    pkg = import_module("integrator-benchmark_b200")
    bk = import_module("integrator-benchmark_b200.testsystems.bookkeepers")
    monkeypatch.setattr(pkg, "DATA_PATH", str(tmp_path))
    eq = bk.EquilibriumSimulator(platform=None, topology=None, system={}, positions=np.zeros((5, 3)), temperature=298.0,
                                 timestep=0.001, burn_in_length=10, n_samples=3, thinning_interval=1, name="five_atoms")
    eq.use_disk_cache = True
    assert eq.get_path_to_samples() == os.path.join(str(tmp_path), "five_atoms_samples.npz")
    assert not eq.check_for_cached_samples()
    frames = np.random.default_rng(1).normal(size=(3, 5, 3)).astype(np.float32)
    eq.samples = frames
    eq.save_samples()
    assert eq.check_for_cached_samples()
    eq2 = bk.EquilibriumSimulator(platform=None, topology=None, system={}, positions=np.zeros((5, 3)), temperature=298.0,
                                  timestep=0.001, burn_in_length=10, n_samples=3, thinning_interval=1, name="five_atoms")
    eq2.use_disk_cache = True
    eq2.collect_equilibrium_samples = lambda *a, **k: pytest.fail("cache ignored")
    eq2.load_or_simulate_samples()
    assert eq2.cached and eq2.samples.dtype == np.float64
    np.testing.assert_array_equal(eq2.samples, frames.astype(np.float64))
    assert eq2.sample_x_from_equilibrium().shape == (5, 3)
    # wrong atom count in the file is an error, not a silent reshape
    eq3 = bk.EquilibriumSimulator(platform=None, topology=None, system={}, positions=np.zeros((6, 3)), temperature=298.0,
                                  timestep=0.001, burn_in_length=10, n_samples=3, thinning_interval=1, name="five_atoms")
    eq3.use_disk_cache = True
    with pytest.raises(ValueError):
        eq3.load_or_simulate_samples()

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
REF_H5 = "/root/reference/data/tiny_waterbox_constrained_samples.h5"


def _h5():
    from importlib import import_module
    return (import_module("integrator-benchmark_b200.serialization.mdtraj_hdf5"),
            import_module("integrator-benchmark_b200.serialization.hdf5_min"))


@pytest.mark.parametrize("n_frames,n_atoms", [(5, 7), (1000, 60), (1, 306), (70, 3)])
def test_mdtraj_hdf5_round_trip(tmp_path, n_frames, n_atoms):
    """save_mdtraj_hdf5 -> load_mdtraj_hdf5 returns the same frames, cell, energies and topology; the file has the
    structures mdtraj / PyTables write (checked field by field against the reference's own cache below)"""
    h5, hmin = _h5()
    rng = np.random.default_rng(5)
    xyz = rng.normal(size=(n_frames, n_atoms, 3)).astype(np.float32)
    pe = rng.normal(size=n_frames).astype(np.float32)
    topo = {"chains": [{"index": 0, "residues": [{"index": 0, "name": "MOL", "resSeq": 1,
                                                 "atoms": [{"index": i, "name": "C%d" % i, "element": "C"} for i in range(n_atoms)]}]}], "bonds": []}
    path = str(tmp_path / "t_samples.h5")
    h5.save_mdtraj_hdf5(path, xyz, cell_lengths=(1.5, 1.5, 1.5), potential_energy=pe, topology=topo)
    out = h5.load_mdtraj_hdf5(path, n_atoms)
    np.testing.assert_array_equal(out["xyz"], xyz)
    np.testing.assert_array_equal(out["cell_lengths"], np.full((n_frames, 3), 1.5, np.float32))
    np.testing.assert_array_equal(out["cell_angles"], np.full((n_frames, 3), 90.0, np.float32))
    np.testing.assert_array_equal(out["potential_energy"], pe)
    assert out["topology"] == topo
    raw = hmin.read(path)
    assert raw["/"].attrs["conventions"] == "Pande" and raw["/"].attrs["conventionVersion"] == "1.1"
    assert raw["coordinates"].attrs["units"] == "nanometers" and raw["coordinates"].attrs["CLASS"] == "EARRAY"
    assert raw["coordinates"].maxshape == (None, n_atoms, 3) and raw["potentialEnergy"].attrs["units"] == "kilojoules_per_mole"
    with pytest.raises(ValueError):
        h5.load_mdtraj_hdf5(path, n_atoms + 1)
    (tmp_path / "not.h5").write_bytes(b"hello")
    with pytest.raises(ValueError):
        h5.load_mdtraj_hdf5(str(tmp_path / "not.h5"), n_atoms)
    # truncated and corrupted files are errors, not garbage frames
    blob = open(path, "rb").read()
    for cut in (50, 200, len(blob) // 2):
        (tmp_path / "cut.h5").write_bytes(blob[:cut])
        with pytest.raises(ValueError):
            h5.load_mdtraj_hdf5(str(tmp_path / "cut.h5"), n_atoms)
    # without a cell and without energies (the water cluster, alanine dipeptide)
    h5.save_mdtraj_hdf5(path, xyz)
    out = h5.load_mdtraj_hdf5(path)
    np.testing.assert_array_equal(out["xyz"], xyz)
    assert out["cell_lengths"] is None and out["potential_energy"] is None


def test_mdtraj_hdf5_empty_and_large(tmp_path):
    """edge cases: a trajectory without frames; more frames than 64 chunks of the default size (chunks grow, one leaf node)"""
    h5, hmin = _h5()
    path = str(tmp_path / "e.h5")
    h5.save_mdtraj_hdf5(path, np.zeros((0, 4, 3), np.float32), cell_lengths=(1.0, 1.0, 1.0))
    out = h5.load_mdtraj_hdf5(path, 4)
    assert out["xyz"].shape == (0, 4, 3) and out["cell_lengths"].shape == (0, 3)
    xyz = np.random.default_rng(0).normal(size=(5000, 700, 3)).astype(np.float32)  # 8400 B per frame: 7 frames per 64 KB chunk
    h5.save_mdtraj_hdf5(path, xyz)
    raw = hmin.read(path)
    assert raw["coordinates"].chunks[0] >= 5000 / 64 and raw["coordinates"].chunks[1:] == (700, 3)
    np.testing.assert_array_equal(h5.load_mdtraj_hdf5(path, 700)["xyz"], xyz)


def test_mdtraj_hdf5_file_structure(tmp_path):
    """the bytes follow the HDF5 specification the way libhdf5 lays these files out: fixed-size B-tree nodes, 8-byte alignment,
    sorted symbol table, end-of-file address, chunk keys (size, mask, offsets) with the closing key"""
    import struct
    h5, hmin = _h5()
    xyz = np.arange(40 * 5 * 3, dtype=np.float32).reshape(40, 5, 3)
    path = str(tmp_path / "s.h5")
    h5.save_mdtraj_hdf5(path, xyz, cell_lengths=(2.0, 2.0, 2.0), potential_energy=np.zeros(40))
    b = open(path, "rb").read()
    assert b[:8] == b"\x89HDF\r\n\x1a\n" and b[8] == 0 and b[13] == 8 and b[14] == 8
    assert struct.unpack_from("<HH", b, 16) == (4, 16)                       # group leaf / internal node K
    assert struct.unpack_from("<Q", b, 40)[0] == len(b)                       # end-of-file address
    root, = struct.unpack_from("<Q", b, 64)
    btree, heap = struct.unpack_from("<QQ", b, 80)
    assert root == 96 and b[btree:btree + 4] == b"TREE" and b[heap:heap + 4] == b"HEAP"
    assert all(a % 8 == 0 for a in (root, btree, heap))
    r = hmin._Reader(b)
    names = [n for n, _ in r.group_entries(btree, heap)]
    assert names == sorted(names) == ["cell_angles", "cell_lengths", "coordinates", "potentialEnergy"]
    for name, ohdr in r.group_entries(btree, heap):
        layout = [body for t, body in r.messages(ohdr) if t == 0x08][0]
        assert layout[0] == 3 and layout[1] == 2
        rank = layout[2] - 1
        tree, = struct.unpack_from("<Q", layout, 3)
        assert b[tree:tree + 4] == b"TREE" and b[tree + 4] == 1 and tree % 8 == 0
        chunks = r.chunks_of(tree, rank)
        cshape = struct.unpack_from("<%dI" % rank, layout, 11)
        assert [c[0][0] for c in chunks] == list(range(0, 40, cshape[0])) and len(chunks) <= 64
        # closing key: one chunk past the last offset, the chunk dims and the element size
        key = 8 + 8 * (rank + 1)
        p = tree + 24 + len(chunks) * (key + 8)
        assert struct.unpack_from("<II%dQ" % (rank + 1), b, p) == (0, 0, chunks[-1][0][0] + cshape[0]) + tuple(cshape[1:]) + (4,)


@pytest.mark.skipif(not os.path.exists(REF_H5), reason="reference checkout not present")
def test_mdtraj_hdf5_reference_cache(tmp_path):
    """the one cache the reference ships is parsed structurally and decodes to the committed fixture frames; a file written
    here carries, message for message, the same dataset descriptions as the reference's (PyTables / libhdf5-written) file"""
    h5, hmin = _h5()
    out = h5.load_mdtraj_hdf5(REF_H5, 306)
    gold = np.load(os.path.join(ROOT, "tests", "golden", "waterbox_frames.npz"))
    assert out["xyz"].shape == (4, 306, 3)
    np.testing.assert_array_equal(out["xyz"], gold["xyz"])
    np.testing.assert_allclose(out["cell_lengths"], 1.5)
    np.testing.assert_allclose(out["cell_angles"], 90.0)
    np.testing.assert_allclose(out["potential_energy"], [-4136.6406, -4029.0005, -4030.8572, -4068.7595], rtol=1e-6)
    assert len(out["topology"]["chains"][0]["residues"]) == 102 and out["topology"]["chains"][0]["residues"][0]["name"] == "HOH"
    # re-write the same content and compare the per-dataset header messages with the reference file's
    path = str(tmp_path / "rewritten.h5")
    h5.save_mdtraj_hdf5(path, out["xyz"], cell_lengths=out["cell_lengths"], cell_angles=out["cell_angles"],
                        potential_energy=out["potential_energy"], topology=out["topology"])
    ref, new = hmin._Reader(open(REF_H5, "rb").read()), hmin._Reader(open(path, "rb").read())

    def headers(r):
        sym = [struct_body for t, struct_body in r.messages(r.root_ohdr) if t == 0x11][0]
        import struct
        return dict(r.group_entries(*struct.unpack_from("<QQ", sym, 0)))
    h_ref, h_new = headers(ref), headers(new)
    assert sorted(h_ref) == sorted(h_new)
    for name in ("coordinates", "cell_lengths", "cell_angles", "potentialEnergy"):
        m_ref = {t: body for t, body in ref.messages(h_ref[name]) if t in (0x01, 0x03, 0x05, 0x0B)}
        m_new = {t: body for t, body in new.messages(h_new[name]) if t in (0x01, 0x03, 0x05, 0x0B)}
        assert m_ref == m_new, name                                   # dataspace, datatype, fill value, filter pipeline: identical bytes
        a_ref = {body for t, body in ref.messages(h_ref[name]) if t == 0x0C}
        a_new = {body for t, body in new.messages(h_new[name]) if t == 0x0C}
        assert a_ref == a_new, name                                   # CLASS / VERSION / TITLE / EXTDIM / units attributes: identical bytes
    back = h5.load_mdtraj_hdf5(path, 306)
    np.testing.assert_array_equal(back["xyz"], out["xyz"])
    assert back["topology"] == out["topology"]


def test_topology_from_description():
    from importlib import import_module
    h5, _ = _h5()
    ts = import_module("integrator-benchmark_b200.testsystems.systems")
    desc, _ = ts.water_cluster(4, constrained=True)
    topo = h5.topology_from_description(desc)
    res = topo["chains"][0]["residues"]
    assert len(res) == 4 and all(r["name"] == "HOH" and [a["element"] for a in r["atoms"]] == ["O", "H", "H"] for r in res)
    assert len(topo["bonds"]) == 8
    desc, _ = ts.alanine_dipeptide(constrained=True, mass_scale=3.0)
    topo = h5.topology_from_description(desc)
    els = [a["element"] for r in topo["chains"][0]["residues"] for a in r["atoms"]]
    assert len(els) == 22 and els.count("H") == 12 and els.count("C") == 6 and els.count("N") == 2 and els.count("O") == 2


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
    assert eq.get_path_to_samples() == os.path.join(str(tmp_path), "five_atoms_samples.h5")
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

"""Host-side mirrors of the reference's utilities against fixtures generated from the reference
(tests/golden/make_golden.py): MTS string generators, hydrogen-mass repartitioning, unit handling."""
import json
import os

import numpy as np
import pytest

from integrator_benchmark_b200 import unit
from integrator_benchmark_b200.utilities import mts_utilities as mts
from integrator_benchmark_b200.utilities import openmm_utilities as omu


def test_mts_string_generators_match_reference(golden_dir):
    cases = json.load(open(os.path.join(golden_dir, "mts_strings.json")))
    assert len(cases) >= 20
    for c in cases:
        kw = dict(c["kwargs"])
        if c["fn"] == "generate_baoab_mts_string":
            kw["groups"] = [(g[0], g[1]) for g in kw["groups"]]
        got = getattr(mts, c["fn"])(**kw)
        if c["fn"] == "generate_all_BAOAB_permutation_strings":
            got = [[list(p), s] for p, s in got]
        assert got == c["value"], (c["fn"], kw)


def test_hmr_matches_reference(golden_dir):
    d = json.load(open(os.path.join(golden_dir, "hmr.json")))
    elements, bonds, masses = d["symbols"], [tuple(b) for b in d["bonds"]], d["masses"]
    n = 0
    for c in d["cases"]:
        if c["fn"] == "amber":
            fn = lambda: omu.repartition_hydrogen_mass_amber_masses(elements, bonds, masses, c["scale_factor"])  # noqa: E731
        else:
            fn = lambda: omu.repartition_hydrogen_mass_masses(elements, bonds, masses, c["h_mass"], c["mode"], c["atoms"])  # noqa: E731
        if "error" in c:
            with pytest.raises(Exception) as ei:
                fn()
            assert type(ei.value).__name__ == c["error"]
        else:
            np.testing.assert_allclose(fn(), c["masses"], rtol=1e-13)
            n += 1
    assert n >= 12


def test_hmr_invariants_like_reference_tests():
    """benchmark/tests/test_openmm_utilities.py:8-94: total mass conserved, masses positive, H mass set."""
    from integrator_benchmark_b200.testsystems.systems import alanine_dipeptide
    desc, _ = alanine_dipeptide()
    total = omu.get_sum_of_masses(desc)
    hs = [i for i, e in enumerate(desc["elements"]) if e == "H"]
    for mode in ["decrement", "scale"]:
        for atoms in ["connected", "all"]:
            for h_mass in [0.1, 1.0, 4.0]:
                out = omu.repartition_hydrogen_mass(None, desc, h_mass, mode, atoms)
                assert np.isclose(omu.get_sum_of_masses(out), total)
                assert omu.get_masses(out).min() > 0
                assert np.all(omu.get_masses(out)[hs] == h_mass)
    out = omu.repartition_hydrogen_mass_amber(None, desc)
    assert np.isclose(omu.get_sum_of_masses(out), total) and omu.get_masses(out).min() > 0
    assert np.isclose(omu.get_sum_of_masses(out, hs) / omu.get_sum_of_masses(desc, hs), 3.0)


def test_units():
    kT = unit.MOLAR_GAS_CONSTANT_R * (298.0 * unit.kelvin)
    assert kT.value_in_unit(unit.kilojoule_per_mole) == pytest.approx(2.4777098, rel=1e-6)
    assert (2.0 * unit.femtoseconds).value_in_unit(unit.picoseconds) == pytest.approx(0.002)
    assert unit.md_value(91.0 / unit.picoseconds) == 91.0 and unit.md_value(0.5) == 0.5
    assert int((2 / (1.0 / unit.picoseconds)) / (4.0 * unit.femtosecond)) == 500
    with pytest.raises(TypeError):
        (1.0 * unit.kelvin).value_in_unit(unit.picosecond)


def test_alanine_description_is_consistent():
    from integrator_benchmark_b200.testsystems.systems import alanine_dipeptide
    from oracle import oracle as orc
    desc, pos = alanine_dipeptide(constrained=True)
    assert desc["n_atoms"] == 22 and len(desc["constraint_atoms"]) == 12 and len(desc["bond_atoms"]) == 9
    assert len(desc["angle_atoms"]) == 36 and abs(sum(desc["charge"])) < 1e-9
    for (a, b), d in zip(desc["constraint_atoms"], desc["constraint_dist"]):
        assert abs(np.linalg.norm(pos[a] - pos[b]) - d) < 1e-3
    e, f = orc.particles_energy_forces(desc, pos)
    assert np.isfinite(e) and np.abs(f).max() < 5e3  # a sane, unstrained start geometry


def test_fastmath_tables_header_is_what_the_generator_writes(tmp_path):
    """csrc/fastmath64_tables.h is generated (scripts/gen_fastmath_tables.py, mpmath at 200 bits); the committed copy must
    not drift from the generator, and a few entries are checked against their definitions directly"""
    import importlib.util
    import math
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    spec = importlib.util.spec_from_file_location("gen_tables", os.path.join(root, "scripts", "gen_fastmath_tables.py"))
    gen = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(gen)
    out = tmp_path / "tables.h"
    gen.main(str(out))
    committed = open(os.path.join(root, "integrator-benchmark_b200", "csrc", "fastmath64_tables.h")).read()
    assert out.read_text() == committed
    rows = [tuple(float.fromhex(v) for v in line.strip(" {},\n").split(", "))
            for line in committed.splitlines() if line.startswith("    {")]
    log_rows, trig_rows = rows[:128], rows[128:]
    assert len(trig_rows) == 1024
    assert log_rows[80] == (1.0, 0.0)                         # the interval centred on z = 1
    c0 = gen.bits_to_double(0x3FE60000)
    assert log_rows[0][0] == pytest.approx(1.0 / c0, rel=1e-16) and log_rows[0][1] == pytest.approx(-2.0 * math.log(c0), rel=1e-15)
    assert trig_rows[0] == (0.0, 1.0) and trig_rows[256] == pytest.approx((1.0, 0.0), abs=1e-16)
    assert trig_rows[100] == pytest.approx((math.sin(2 * math.pi * 100 / 1024), math.cos(2 * math.pi * 100 / 1024)), rel=1e-15)

"""Host-side parts of scope row 8f-4: probabilistic bisection and the GHMC splitting checks.  CPU only."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def test_probabilistic_bisection_finds_noisy_threshold():
    from integrator_benchmark_b200.evaluation import thresholds
    rng = np.random.default_rng(0)
    true = 6.3

    def oracle(x):  # right 80 % of the time
        z = 1 if x < true else -1
        return z if rng.random() < 0.8 else -z
    grid, answers, densities = thresholds.probabilistic_bisection(oracle, initial_limits=(0, 10), p=0.6,
                                                                  n_iterations=300, resolution=20000)
    assert len(answers) == 300 and len(densities) == 301
    assert abs(grid[np.argmax(densities[-1])] - true) < 0.1
    for f in (densities[0], densities[-1]):
        assert np.sum(0.5 * (f[1:] + f[:-1]) * np.diff(grid)) == pytest.approx(1.0)
    # a noiseless oracle and p close to 1 is plain bisection
    grid, _, densities = thresholds.probabilistic_bisection(lambda x: 1 if x < 0.3 else -1, p=0.999, n_iterations=20,
                                                            resolution=100001)
    assert abs(grid[np.argmax(densities[-1])] - 0.3) < 1e-3


def test_ghmc_constructor_checks_and_globals():
    """ghmc.py:74-86: same AssertionError / ValueError cases; an O is prepended to a deterministic splitting"""
    from integrator_benchmark_b200 import unit
    from integrator_benchmark_b200.integrators import CustomizableGHMC
    g = CustomizableGHMC()
    assert g._tokens == ["O", "V", "R", "V"]
    assert [k for k, _ in g._segments] == ["D"]          # reference quirk: the O token emits nothing
    assert [op[0] for op in g._segments[0][1].ops()] == [1, 0, 1]
    assert [op[2] for op in g._segments[0][1].ops()] == [0.5, 1.0, 0.5]
    names = [g.getGlobalVariableName(i) for i in range(g.getNumGlobalVariables())]
    assert names == ["kT", "a", "b", "ke_old", "ke_new", "heat", "shadow_work", "e_old", "e_new", "acc_ratio", "accept",
                     "naccept", "ntrials"]
    assert g.getGlobalVariableByName("a") == pytest.approx(np.exp(-91.0 * 0.001))
    assert g.getConstraintTolerance() == 1e-4
    g.setStepSize(2.0 * unit.femtosecond)
    assert g.getStepSize().value_in_unit(unit.femtosecond) == pytest.approx(2.0)
    assert g.getGlobalVariableByName("a") == pytest.approx(np.exp(-91.0 * 0.001))  # globals keep their values
    o = CustomizableGHMC("V R O R V", apply_O=True)
    assert [k for k, _ in o._segments] == ["D", "O", "D"]
    m = CustomizableGHMC("V0 V1 R V1 V0")
    assert [op[1] for op in m._segments[0][1].ops()] == [0, 1, -1, 1, 0]
    with pytest.raises(AssertionError):
        CustomizableGHMC("V V")          # no R
    with pytest.raises(AssertionError):
        CustomizableGHMC("R O R")        # no V
    with pytest.raises(AssertionError):
        CustomizableGHMC("V R Q")        # invalid token
    with pytest.raises(ValueError):
        CustomizableGHMC("V V0 R V1")    # all forces and subsets of forces

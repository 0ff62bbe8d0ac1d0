"""Drop-in for benchmark/integrators/langevin.py:9-408.

The reference classes subclass OpenMM's CustomIntegrator and emit a per-DOF step program;
these shims keep the constructor, the accept/reject behaviour of the splitting checks and the
handful of CustomIntegrator methods the reference's callers use, but compile the splitting
with lsi_compile / lsi_compile_continuous (include/lsi.h) into the opcode list that the fused
GPU kernels execute.  Nothing here steps a simulation: NonequilibriumSimulator
(testsystems/bookkeepers.py) hands the compiled program to lsi_run_protocol.
"""
import numpy as np

from .. import engine, unit
from ..constants import kB


def _kT_value(temperature):
    """kB * T in kJ/mol, T a Quantity in kelvin or a plain number of kelvin."""
    T = temperature.value_in_unit(unit.kelvin) if unit.is_quantity(temperature) else float(temperature)
    return kB.value_in_unit(unit.kilojoule_per_mole / unit.kelvin) * T, T


def _per_ps(x):
    return x.value_in_unit(unit.picosecond ** -1) if unit.is_quantity(x) else float(x)


def _ps(x):
    return x.value_in_unit(unit.picosecond) if unit.is_quantity(x) else float(x)


class _SplittingIntegratorBase:
    """The CustomIntegrator surface the reference's callers touch:
    getStepSize / setStepSize (evaluation/compare_near_eq_and_exact.py:147, evaluation/thresholds.py:117,135),
    getConstraintTolerance (testsystems/bookkeepers.py:165), getGlobalVariableByName
    (bookkeepers.py:195, tests/test_shadow_work.py:61,74), getNumGlobalVariables / getGlobalVariableName
    (tests/test_shadow_work.py:140-141)."""

    def _finish_init(self, temperature, constraint_tolerance, measure_shadow_work, measure_heat):
        self.kT, self.temperature_kelvin = _kT_value(temperature)
        self._constraint_tolerance = float(constraint_tolerance)
        self._measure_shadow_work = bool(measure_shadow_work)
        self._measure_heat = bool(measure_heat)
        # globals in the order the reference adds them (langevin.py:189-215)
        ops = self.program.ops()
        first_o = next((op for op in ops if op[0] == 2), None)
        self._globals = {"kT": self.kT}
        if first_o is not None and not getattr(self, "_continuous", False):
            self._globals["a"], self._globals["b"] = first_o[3], first_o[4]
        if self._measure_heat:
            self._globals["heat"] = 0.0
        if self._measure_heat or self._measure_shadow_work:
            self._globals["old_ke"] = 0.0
            self._globals["new_ke"] = 0.0
        if self._measure_shadow_work:
            self._globals["old_pe"] = 0.0
            self._globals["new_pe"] = 0.0
            self._globals["shadow_work"] = 0.0

    # --- step size (a, b keep their construction-time values)
    def getStepSize(self):
        return self.program.dt * unit.picosecond

    def setStepSize(self, timestep):
        self.program.dt = _ps(timestep)

    def rebuild(self, timestep=None):
        """Not in the reference: recompile so that the O coefficients follow the current step size."""
        dt = self.program.dt if timestep is None else _ps(timestep)
        self.program = engine.Program(self._splitting_arg, dt, self._gamma, self._measure_shadow_work,
                                      self._measure_heat, self._override)
        return self

    def getConstraintTolerance(self):
        return self._constraint_tolerance

    def setConstraintTolerance(self, tol):
        self._constraint_tolerance = float(tol)

    # --- globals
    def getNumGlobalVariables(self):
        return len(self._globals)

    def getGlobalVariableName(self, index):
        return list(self._globals)[index]

    def getGlobalVariableByName(self, name):
        if name not in self._globals:
            raise Exception("getGlobalVariableByName: unknown variable '%s'" % name)
        return self._globals[name]

    def getGlobalVariable(self, index):
        return self._globals[self.getGlobalVariableName(index)]

    def setGlobalVariableByName(self, name, value):
        if name not in self._globals:
            raise Exception("setGlobalVariableByName: unknown variable '%s'" % name)
        self._globals[name] = float(value)

    def _accumulate(self, heat=None, shadow_work=None):
        """bookkeeping after a GPU launch: `heat` / `shadow_work` are never reset (bookkeepers.py:193-195)"""
        if heat is not None and "heat" in self._globals:
            self._globals["heat"] += float(heat)
        if shadow_work is not None and "shadow_work" in self._globals:
            self._globals["shadow_work"] += float(shadow_work)

    @property
    def splitting(self):
        return self._splitting_arg


class LangevinSplittingIntegrator(_SplittingIntegratorBase):
    """Langevin dynamics with a prescribed operator splitting over {R, V, O, V0..V31}
    (reference: langevin.py:9-221).  Same constructor, same AssertionError / ValueError cases."""

    def __init__(self, splitting="V R O R V", temperature=298.0 * unit.kelvin,
                 collision_rate=91.0 / unit.picoseconds, timestep=1.0 * unit.femtoseconds,
                 constraint_tolerance=1e-8, override_splitting_checks=False, measure_shadow_work=False,
                 measure_heat=True):
        self._splitting_arg = splitting
        self._gamma = _per_ps(collision_rate)
        self._override = bool(override_splitting_checks)
        self.program = engine.Program(splitting, _ps(timestep), self._gamma, measure_shadow_work, measure_heat,
                                      override_splitting_checks)
        self._finish_init(temperature, constraint_tolerance, measure_shadow_work, measure_heat)


class ContinuousLangevinSplittingIntegrator(_SplittingIntegratorBase):
    """Splitting given as [(token, fraction of dt), ...] (reference: langevin.py:223-408): the
    O coefficients follow each substep's own length, zero-length substeps are dropped, the
    fractions of each letter must sum to 1."""
    _continuous = True

    def __init__(self, splitting=(("V", 0.5), ("R", 0.5), ("O", 1.0), ("R", 0.5), ("V", 0.5)),
                 temperature=298.0 * unit.kelvin, collision_rate=91.0 / unit.picoseconds,
                 timestep=1.0 * unit.femtoseconds, constraint_tolerance=1e-8, override_splitting_checks=False,
                 measure_shadow_work=False, measure_heat=True):
        self._splitting_arg = [(str(t), float(f)) for t, f in splitting]
        self._gamma = _per_ps(collision_rate)
        self._override = bool(override_splitting_checks)
        self.program = engine.Program(self._splitting_arg, _ps(timestep), self._gamma, measure_shadow_work,
                                      measure_heat, override_splitting_checks)
        self._finish_init(temperature, constraint_tolerance, measure_shadow_work, measure_heat)


def splitting_fractions(splitting):
    """Fractions of dt each substep of a splitting string advances (for callers that convert a
    string into the continuous form); numpy only, mirrors langevin.py:92-105,130,155-157,192."""
    toks = splitting.upper().split()
    n_R, n_V, n_O = toks.count("R"), toks.count("V"), toks.count("O")
    n_Vs = {t: toks.count(t) for t in set(toks) if t.startswith("V")}
    out = []
    for t in toks:
        if t == "R":
            out.append((t, 1.0 / n_R))
        elif t == "O":
            out.append((t, 1.0 / max(1, n_O)))
        else:
            out.append((t, 1.0 / (n_Vs[t] if len(n_Vs) > 1 else n_V)))
    return np.array([f for _, f in out]), [t for t, _ in out]

"""Drop-in for benchmark/integrators/ghmc.py:9-194 (scope row 8f-4).

`CustomizableGHMC` is generalized hybrid Monte Carlo whose proposal is the deterministic part of a
splitting string: one trial = [partial velocity refresh] -> R / V substeps -> Metropolis test on
exp(-(e_new - e_old) / kT) -> on rejection restore x and FLIP the momenta.

The reference emits an OpenMM step program for ONE Context.  Here a trial advances an ENSEMBLE of
chains at once: the proposal is one launch of the fused protocol kernel (lsi_run_protocol, n_steps = 1;
its work output IS (e_new - e_old) / kT because a proposal has no O substep), the accept / restore /
flip is a masked copy on the device.

Reference quirk, kept by default: `substep_function` (ghmc.py:129-133) has no branch for "O", so the
O token -- including the one the constructor prepends to a deterministic splitting (:58-59) -- emits
nothing and velocities are never refreshed; `a` and `b` are defined (:143-147) but unused.  Pass
`apply_O=True` for the behaviour the docstring describes (an O substep of the full time step where
the token stands, its heat excluded from the Metropolis ratio).
"""
import numpy as np

from .. import _rngstate, engine, unit
from .langevin import _kT_value, _per_ps, _ps


class CustomizableGHMC:
    def __init__(self, splitting="V R V", temperature=298.0 * unit.kelvin, collision_rate=91.0 / unit.picoseconds,
                 timestep=1.0 * unit.femtoseconds, constraint_tolerance=1e-4, apply_O=False):
        tokens = splitting.upper().split()
        n_V = sum(t == "V" for t in tokens)
        n_O = sum(t == "O" for t in tokens)
        if n_O == 0:
            tokens = ["O"] + tokens
        v_tokens = set(t for t in tokens if t[0] == "V")
        mts = len(v_tokens) > 1
        # the same sanity checks, in the same order, as ghmc.py:74-86
        assert "R" in tokens
        assert "V" in [t[0] for t in tokens]
        assert "O" in tokens
        assert set(tokens).issubset(set("RVO").union("V{}".format(i) for i in range(32)))
        if mts and n_V > 0:
            raise ValueError("Splitting string includes an evaluation of all forces and "
                             "evaluation of subsets of forces.")
        self._tokens = tokens
        self._apply_O = bool(apply_O)
        self._gamma = _per_ps(collision_rate)
        self._dt = _ps(timestep)
        self.kT, self.temperature_kelvin = _kT_value(temperature)
        self._constraint_tolerance = float(constraint_tolerance)
        self._globals = {"kT": self.kT, "a": float(np.exp(-self._gamma * self._dt)),
                         "b": float(np.sqrt(1.0 - np.exp(-2.0 * self._gamma * self._dt))),
                         "ke_old": 0.0, "ke_new": 0.0, "heat": 0.0, "shadow_work": 0.0, "e_old": 0.0, "e_new": 0.0,
                         "acc_ratio": 0.0, "accept": 0.0, "naccept": 0.0, "ntrials": 0.0}
        self._compile()

    # ---- compiled pieces: maximal runs of R / V tokens, separated by O substeps (when applied)
    def _compile(self):
        """Substep lengths follow the counts over the WHOLE string (dt / n_R, dt / n_V or dt / n_Vs[g]), so each
        deterministic run is compiled as an explicit-fraction program."""
        toks = self._tokens
        count = {}
        for t in toks:
            count[t] = count.get(t, 0) + 1
        self._segments = []  # ("D", Program) or ("O", Program)
        run = []

        def flush():
            if run:
                frac = [(t, 1.0 / count[t]) for t in run]
                self._segments.append(("D", engine.Program(frac, self._dt, self._gamma, False, True, True)))
                del run[:]
        for t in toks:
            if t == "O":
                flush()
                if self._apply_O:
                    self._segments.append(("O", engine.Program([("O", 1.0)], self._dt, self._gamma, False, True, True)))
            else:
                run.append(t)
        flush()

    # ---- CustomIntegrator surface
    def getStepSize(self):
        return self._dt * unit.picosecond

    def setStepSize(self, timestep):
        """a, b keep their construction-time values, as OpenMM globals do; substep lengths follow"""
        self._dt = _ps(timestep)
        self._compile()

    def getConstraintTolerance(self):
        return self._constraint_tolerance

    def setConstraintTolerance(self, tol):
        self._constraint_tolerance = float(tol)

    def getNumGlobalVariables(self):
        return len(self._globals)

    def getGlobalVariableName(self, index):
        return list(self._globals)[index]

    def getGlobalVariableByName(self, name):
        if name not in self._globals:
            raise Exception("getGlobalVariableByName: unknown variable '%s'" % name)
        return self._globals[name]

    def setGlobalVariableByName(self, name, value):
        if name not in self._globals:
            raise Exception("setGlobalVariableByName: unknown variable '%s'" % name)
        self._globals[name] = float(value)

    @property
    def acceptance_rate(self):
        return self._globals["naccept"] / max(1.0, self._globals["ntrials"])

    # ---- the ensemble of chains
    def run(self, equilibrium_simulator, x, v=None, n_trials=1, seed=None, precision=None):
        """`n_trials` GHMC trials for every chain.  x: (n, n_atoms, 3) nm; v: same shape in nm/ps or None for
        Maxwell-Boltzmann velocities.  Returns (x, v, accept) as device tensors, accept = (n_trials, n) bool.
        naccept / ntrials accumulate over chains and trials."""
        engine.require_cuda()
        t = engine.torch()
        eq = equilibrium_simulator
        precision = eq.precision if precision is None else precision
        system = eq.engine_system(self._constraint_tolerance)
        seed = _rngstate.get_seed() if seed is None else int(seed)
        x = t.as_tensor(np.asarray(x, dtype=np.float64) if not t.is_tensor(x) else x).cuda().to(t.float64).contiguous()
        n = x.shape[0]
        kT = eq.kT_value
        rid = _rngstate.reserve_replicas(n)

        def launch(prog, x0, v0, s):
            return engine.run_protocol(system, prog, n, 1, kT, n_legs=1, seed=s, replica0=rid, x0=x0, v0=v0,
                                       precision=precision, want_heat=True, want_final=True, want_moments=False, out={})
        if v is None:  # a 0-step launch draws and projects Maxwell-Boltzmann velocities
            res = engine.run_protocol(system, self._segments[0][1], n, 0, kT, n_legs=1, seed=seed, replica0=rid, x0=x,
                                      precision=precision, want_final=True, want_moments=False, out={})
            v = res["v_final"].clone()
        else:
            v = t.as_tensor(np.asarray(v, dtype=np.float64) if not t.is_tensor(v) else v).cuda().to(t.float64).contiguous()
        gen = t.Generator(device="cuda")
        gen.manual_seed(seed ^ 0x6A09E667)
        accepts = t.empty((int(n_trials), n), dtype=t.bool, device="cuda")
        for trial in range(int(n_trials)):
            x_new, v_new = x, v
            w = t.zeros(n, dtype=t.float64, device="cuda")
            for k, (kind, prog) in enumerate(self._segments):
                res = launch(prog, x_new, v_new, seed + 104729 * (trial * len(self._segments) + k + 1))
                x_new, v_new = res["x_final"].clone(), res["v_final"].clone()
                if kind == "D":
                    w = w + res["W"][0]  # (e_new - e_old) / kT of the deterministic run
            ratio = t.exp(-w)
            u = t.rand(n, dtype=t.float64, device="cuda", generator=gen)
            acc = (ratio - u) >= 0  # step(acc_ratio - uniform); NaN rejects
            m = acc[:, None, None]
            x = t.where(m, x_new, x)
            v = t.where(m, v_new, -v)
            accepts[trial] = acc
        if int(n_trials) > 0:  # the per-trial globals show the last chain's last trial (one host sync, after the loop)
            self._globals["shadow_work"] = float(w[-1].item()) * kT
            self._globals["acc_ratio"] = float(ratio[-1].item())
            self._globals["accept"] = float(acc[-1].item())
        self._globals["naccept"] += float(accepts.sum().item())
        self._globals["ntrials"] += float(accepts.numel())
        return x, v, accepts

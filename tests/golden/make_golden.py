#!/usr/bin/env python
"""Generate tests/golden/* from the REFERENCE's own Python code.

Run in the build container only (needs /root/reference):

    python tests/golden/make_golden.py

Nothing under tests/ reads /root/reference at test time; the fixtures written
here are committed.  What is produced and how the reference is involved:

  toy_numba.npz      xs, vs, Q, W_shads of the reference's four closed-form
                     integrators (benchmark/integrators/numba_integrators.py,
                     loaded by file path, run through `.py_func`, i.e. the
                     unmodified Python body) on the quartic and double-well
                     potentials of benchmark/testsystems/low_dimensional_systems.py:132-136.
                     The only intervention is that the module-level name `np` is
                     replaced by a proxy whose `random.randn()` replays a noise
                     array (the repo's Philox normals), so the run is reproducible.
  programs.json      the step programs emitted by the reference's own
                     LangevinSplittingIntegrator / ContinuousLangevinSplittingIntegrator
                     constructors (benchmark/integrators/langevin.py), recorded by a
                     stand-in for simtk.openmm.CustomIntegrator, plus the exception
                     type raised for rejected strings.
  lsi_trajectories.npz  trajectories obtained by interpreting those recorded
                     programs (numpy, OpenMM CustomIntegrator semantics) on small
                     particle systems, with and without distance constraints.
  estimators.json    outputs of the reference's estimator functions
                     (evaluation/analysis.py:8-17, compare_near_eq_and_exact.py:32-46,62-97)
                     extracted from source by `ast` (their modules import matplotlib/simtk).
  mts_strings.json   outputs of the multi-timestep string generators
                     (utilities/mts_utilities.py).
  hmr.json           masses produced by repartition_hydrogen_mass_amber and friends
                     (utilities/openmm_utilities.py:134-269) on a duck-typed topology/system.
  waterbox_frames.npz  the 4 valid frames + energies of data/tiny_waterbox_constrained_samples.h5.
"""
import ast
import importlib.util
import json
import os
import sys
import types
import zlib

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = "/root/reference"
sys.path.insert(0, ROOT)

from oracle import oracle as orc  # noqa: E402  (Philox normals used as the replayed noise)

KB = 8.31446261815324e-3  # kJ/mol/K, value given to the stand-in openmmtools.constants.kB


# --------------------------------------------------------------------------- stand-ins
class Recorder:
    """Stand-in for simtk.openmm.CustomIntegrator that records the program."""

    def __init__(self, timestep):
        self.timestep = timestep
        self.program = []
        self.globals = {}
        self.perdof = {}
        self.tolerance = None

    def addGlobalVariable(self, name, value):
        self.globals[name] = float(value)

    def addPerDofVariable(self, name, value):
        self.perdof[name] = float(value)

    def setConstraintTolerance(self, tol):
        self.tolerance = tol

    def addComputePerDof(self, var, expr):
        self.program.append(["perdof", var, expr])

    def addComputeGlobal(self, var, expr):
        self.program.append(["global", var, expr])

    def addComputeSum(self, var, expr):
        self.program.append(["sum", var, expr])

    def addConstrainPositions(self):
        self.program.append(["constrain_x"])

    def addConstrainVelocities(self):
        self.program.append(["constrain_v"])

    def addUpdateContextState(self):
        self.program.append(["update_context_state"])


class _Mass(float):
    unit = "amu"

    def value_in_unit(self, _u):
        return float(self)


def install_standins():
    simtk = types.ModuleType("simtk")
    mm = types.ModuleType("simtk.openmm")
    mm.CustomIntegrator = Recorder
    mm.version = types.SimpleNamespace(full_version="stand-in (no OpenMM in this image)")
    app = types.ModuleType("simtk.openmm.app")
    mm.app = app
    unit = types.ModuleType("simtk.unit")
    unit.kelvin = 1.0
    unit.picoseconds = 1.0
    unit.picosecond = 1.0
    unit.femtoseconds = 1e-3
    unit.femtosecond = 1e-3
    unit.amu = "amu"
    unit.Quantity = type("Quantity", (), {})
    simtk.openmm = mm
    simtk.unit = unit
    omt = types.ModuleType("openmmtools")
    omtc = types.ModuleType("openmmtools.constants")
    omtc.kB = KB
    omt.constants = omtc
    bench = types.ModuleType("benchmark")
    bench.simulation_parameters = {"temperature": 298.0}
    for name, mod in [("simtk", simtk), ("simtk.openmm", mm), ("simtk.openmm.app", app), ("simtk.unit", unit),
                      ("openmmtools", omt), ("openmmtools.constants", omtc), ("benchmark", bench)]:
        sys.modules[name] = mod


def load_by_path(name, relpath):
    spec = importlib.util.spec_from_file_location(name, os.path.join(REF, relpath))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def extract_functions(relpath, names, namespace):
    """exec only the named top-level function definitions of a reference file."""
    src = open(os.path.join(REF, relpath)).read()
    tree = ast.parse(src)
    found = {}
    for node in tree.body:
        if isinstance(node, ast.FunctionDef) and node.name in names:
            found[node.name] = node  # later definitions shadow earlier ones, as on import
    module = ast.Module(body=[found[n] for n in names if n in found], type_ignores=[])
    exec(compile(module, relpath, "exec"), namespace)
    return namespace


# --------------------------------------------------------------------------- 1. toy integrators
class _Replay:
    def __init__(self, arr):
        self.arr, self.i = arr, 0

    def randn(self):
        v = self.arr[self.i]
        self.i += 1
        return float(v)


class _NpProxy:
    def __init__(self, rnd):
        self.random = rnd

    def __getattr__(self, k):
        return getattr(np, k)


def scalar_stream(seed, replica, tag, n):
    out = []
    for blk in range((n + 3) // 4):
        out.extend(orc.normals4(seed, replica, blk, tag, 0))
    return np.array(out[:n])


def make_toy():
    mod = load_by_path("ref_numba_integrators", "benchmark/integrators/numba_integrators.py")
    systems = {  # low_dimensional_systems.py:54-55 and :133-134, m = 10, beta = 1
        "quartic": (lambda x: x ** 4, lambda x: -4.0 * x ** 3),
        "double_well": (lambda x: x ** 6 + 2 * np.cos(5 * (x + 1)), lambda x: -(6 * x ** 5 - 10 * np.sin(5 * (x + 1)))),
    }
    mass, beta = 10.0, 1.0
    velocity_scale = np.sqrt(1.0 / (beta * mass))
    factories = {"vvvr": mod.vvvr_factory, "orvro": mod.orvro_factory, "baoab": mod.baoab_factory,
                 "aboba": mod.aboba_factory}
    cases_by_system = {"quartic": [(100.0, 0.25, 25), (0.5, 0.9, 40), (10.0, 0.6, 11), (100.0, 1.1, 2)],
                       "double_well": [(100.0, 0.25, 25), (0.5, 0.4, 40), (10.0, 0.6, 11), (100.0, 0.7, 2)]}
    out = {}
    meta = []
    seed = 0x5EED0001
    case_id = 0
    for sysname, (pot, frc) in systems.items():
        for fname, factory in factories.items():
            f = factory(pot, frc, velocity_scale, mass)
            for gamma, dt, n_steps in cases_by_system[sysname]:
                replica = 1000 + case_id
                noise = scalar_stream(seed, replica, 0, 2 * (n_steps - 1))
                x0 = 0.4 * orc.normals4(seed, replica, 0, 3, 0)[1]
                v0 = velocity_scale * orc.normals4(seed, replica, 0, 1, 0)[0]
                saved_np = mod.np
                mod.np = _NpProxy(_Replay(noise))
                try:
                    xs, vs, Q, W = f.py_func(x0, v0, n_steps, gamma, dt)
                    used = mod.np.random.i
                finally:
                    mod.np = saved_np
                assert used == 2 * (n_steps - 1)
                key = "c%03d" % case_id
                out[key + "_xs"], out[key + "_vs"], out[key + "_W"] = xs, vs, W
                out[key + "_Q"] = np.array(Q)
                meta.append(dict(key=key, system=sysname, scheme=fname, gamma=gamma, dt=dt, n_steps=n_steps,
                                 seed=seed, replica=replica, x0=x0, v0=v0, mass=mass, beta=beta))
                case_id += 1
    out["meta"] = np.array(json.dumps(meta))
    np.savez_compressed(os.path.join(HERE, "toy_numba.npz"), **out)
    print("toy_numba.npz:", case_id, "cases")


# --------------------------------------------------------------------------- 2. recorded step programs
SPLITTINGS_OK = [
    "V R O R V", "O V R V O", "O R V R O", "R V O V R", "R O V O R", "V O R O V",
    "R V O", "R O V", "V R R O R R V", "V R R R O R R R V", "v r o r v",
    "O V R R V O", "V R O O R V", "R V O O V R",
    "V0 V1 R O R V1 V0", "V0 V1 R R O R R V1 V1 R R O R R V1 V0",
    "V0 R V1 R V2 R O R V2 R V1 R V0", "V1 V0 R O R V0 V1 V1",
    "V3 V1 V2 V0 R O R V0 V2 V1 V3",
]
SPLITTINGS_BAD = ["V R V", "R O R", "V O V", "V X O R V", "V0 V R O R V V0", "V32 V0 R O", "V0 R O R V0", "", "VRORV"]
CONTINUOUS = [
    [["V", 0.5], ["R", 0.5], ["O", 1.0], ["R", 0.5], ["V", 0.5]],
    [["O", 0.25], ["V", 0.5], ["R", 1.0], ["V", 0.5], ["O", 0.75]],
    [["V", 0.5], ["R", 0.0], ["R", 1.0], ["O", 1.0], ["V", 0.5]],
    [["V", 0.5], ["R", 0.5], ["O", 0.5], ["R", 0.5], ["V", 0.5]],
]


def record(cls, splitting, **kw):
    try:
        integ = cls(splitting=splitting, **kw)
    except Exception as e:  # noqa: BLE001
        return {"error": type(e).__name__}
    return {"program": integ.program, "globals": integ.globals, "perdof": integ.perdof,
            "tolerance": integ.tolerance, "timestep": integ.timestep}


def make_programs():
    lang = load_by_path("ref_langevin", "benchmark/integrators/langevin.py")
    entries = []
    for s in SPLITTINGS_OK + SPLITTINGS_BAD:
        for kw in [dict(temperature=298.0, collision_rate=91.0, timestep=0.002),
                   dict(temperature=298.0, collision_rate=1.0, timestep=0.0045, measure_shadow_work=True),
                   dict(temperature=298.0, collision_rate=1.0, timestep=0.001, override_splitting_checks=True),
                   dict(temperature=300.0, collision_rate=5.0, timestep=0.001, measure_heat=False)]:
            e = record(lang.LangevinSplittingIntegrator, s, **kw)
            e.update(kind="discrete", splitting=s, kwargs=kw)
            entries.append(e)
    for s in CONTINUOUS:
        tuples = [tuple(t) for t in s]
        for kw in [dict(temperature=298.0, collision_rate=91.0, timestep=0.002),
                   dict(temperature=298.0, collision_rate=1.0, timestep=0.004, measure_shadow_work=True)]:
            e = record(lang.ContinuousLangevinSplittingIntegrator, tuples, **kw)
            e.update(kind="continuous", splitting=s, kwargs=kw)
            entries.append(e)
    json.dump({"kB": KB, "entries": entries}, open(os.path.join(HERE, "programs.json"), "w"), indent=0)
    print("programs.json:", len(entries), "entries")
    return lang


# --------------------------------------------------------------------------- 3. interpreted trajectories
def chain_system():
    """5 particles: group 0 = harmonic restraint (K/2)|x|^2 on every atom, group 1 = harmonic bonds."""
    masses = np.array([12.0, 1.0, 16.0, 14.0, 1.0])
    bonds = [(0, 1, 0.11, 2.0e4), (0, 2, 0.14, 3.0e4), (2, 3, 0.13, 2.5e4), (3, 4, 0.10, 2.0e4)]
    x0 = np.array([[0.0, 0.0, 0.0], [0.11, 0.01, 0.0], [-0.05, 0.13, 0.02], [-0.02, 0.22, 0.11], [0.05, 0.21, 0.18]])
    return dict(masses=masses, bonds=bonds, K=50.0, x0=x0)


def chain_forces(sysd, x):
    K = sysd["K"]
    f0 = -K * x
    e0 = 0.5 * K * np.sum(x * x)
    f1 = np.zeros_like(x)
    e1 = 0.0
    for (i, j, r0, k) in sysd["bonds"]:
        d = x[j] - x[i]
        r = np.sqrt(np.sum(d * d))
        e1 += 0.5 * k * (r - r0) ** 2
        g = k * (r - r0) * d / r
        f1[i] += g
        f1[j] -= g
    return {"f0": f0, "f1": f1, "f": f0 + f1, "energy": e0 + e1}


def shake_positions(x_new, x_ref, inv_m, cons, tol=1e-15, max_iter=5000):
    """Solve |x_i - x_j| = d with displacements along the reference bond vectors (SHAKE)."""
    x = x_new.copy()
    for _ in range(max_iter):
        worst = 0.0
        for (i, j, d) in cons:
            rij = x[i] - x[j]
            diff = np.dot(rij, rij) - d * d
            worst = max(worst, abs(diff) / (d * d))
            ref = x_ref[i] - x_ref[j]
            g = diff / (2.0 * np.dot(rij, ref) * (inv_m[i] + inv_m[j]))
            x[i] -= g * inv_m[i] * ref
            x[j] += g * inv_m[j] * ref
        if worst < tol:
            break
    return x


def rattle_velocities(v_in, x, inv_m, cons, tol=1e-15, max_iter=5000):
    v = v_in.copy()
    for _ in range(max_iter):
        worst = 0.0
        for (i, j, d) in cons:
            rij = x[i] - x[j]
            vij = v[i] - v[j]
            rv = np.dot(rij, vij)
            worst = max(worst, abs(rv))
            g = rv / (np.dot(rij, rij) * (inv_m[i] + inv_m[j]))
            v[i] -= g * inv_m[i] * rij
            v[j] += g * inv_m[j] * rij
        if worst < tol:
            break
    return v


def interpret(rec, sysd, x, v, n_steps, noise_fn, cons=()):
    """Minimal CustomIntegrator interpreter (semantics from the OpenMM user guide, [ext]):
    per-DOF expressions see x, v, m, f, f<g>, gaussian, per-DOF and global variables; `energy`
    is the potential energy at the current positions; ConstrainPositions displaces along the
    bond vectors of the last constrained positions; ConstrainVelocities removes the relative
    velocity along each constraint."""
    g = dict(rec["globals"])
    dt = rec["timestep"]
    m = sysd["masses"][:, None] * np.ones((1, 3))
    inv_m = 1.0 / sysd["masses"]
    perdof = {k: np.full_like(x, val) for k, val in rec["perdof"].items()}
    x, v = x.copy(), v.copy()
    traj = []
    draw = 0
    for step in range(n_steps):
        x_ref = x.copy()
        for op in rec["program"]:
            if op[0] in ("perdof", "global", "sum"):
                frc = chain_forces(sysd, x)
                ns = dict(g)
                ns.update(perdof)
                ns.update(x=x, v=v, m=m, dt=dt, sqrt=np.sqrt, energy=frc["energy"],
                          f=frc["f"], f0=frc["f0"], f1=frc["f1"])
                if "gaussian" in op[2]:
                    ns["gaussian"] = noise_fn(draw)
                    draw += 1
                val = eval(op[2], {"__builtins__": {}}, ns)  # noqa: S307  (reference's own expressions)
                if op[0] == "perdof":
                    if op[1] == "x":
                        x = val
                    elif op[1] == "v":
                        v = val
                    else:
                        perdof[op[1]] = val * np.ones_like(x)
                elif op[0] == "global":
                    g[op[1]] = float(val)
                else:
                    g[op[1]] = float(np.sum(val))
            elif op[0] == "constrain_x":
                if cons:
                    x = shake_positions(x, x_ref, inv_m, cons)
                x_ref = x.copy()
            elif op[0] == "constrain_v":
                if cons:
                    v = rattle_velocities(v, x, inv_m, cons)
        frc = chain_forces(sysd, x)
        ke = 0.5 * np.sum(m * v * v)
        traj.append((x.copy(), v.copy(), frc["energy"], ke, g.get("heat", np.nan), g.get("shadow_work", np.nan)))
    return traj


def make_trajectories(lang):
    sysd = chain_system()
    n_atoms = len(sysd["masses"])
    seed = 0x5EED0002
    out = {"masses": sysd["masses"], "bonds": np.array(sysd["bonds"]), "K": np.array(sysd["K"]), "x_init": sysd["x0"]}
    meta = []
    cases = [("V R O R V", False), ("O V R V O", False), ("R V O V R", False), ("O R V R O", False),
             ("V R R O R R V", False), ("V0 V1 R O R V1 V0", False), ("V0 V1 R R O R R V1 V1 R R O R R V1 V0", False),
             ("R V O", False),
             ("V R O R V", True), ("O V R V O", True), ("V0 V1 R O R V1 V0", True), ("R O V O R", True)]
    # constraints (when on): the two light atoms are bound rigidly, as HBonds would do
    cons_all = [(0, 1, 0.11), (3, 4, 0.10)]
    for cid, (splitting, constrained) in enumerate(cases):
        kw = dict(temperature=298.0, collision_rate=5.0, timestep=0.002, measure_shadow_work=True)
        integ = lang.LangevinSplittingIntegrator(splitting=splitting, **kw)
        rec = {"program": integ.program, "globals": integ.globals, "perdof": integ.perdof, "timestep": integ.timestep}
        n_O = sum(t == "O" for t in splitting.upper().split())
        replica = 50 + cid
        n_steps = 12
        kT = KB * 298.0
        sig = np.sqrt(kT / sysd["masses"])[:, None]

        def atom_normals(draw, tag, replica=replica):
            return np.array([orc.normals4(seed, replica, draw, tag, a)[:3] for a in range(n_atoms)])

        sysd_c = dict(sysd)
        cons = cons_all if constrained else []
        if constrained:  # constrained bonds carry no harmonic term (as OpenMM's createSystem does)
            sysd_c["bonds"] = [b for b in sysd["bonds"] if (b[0], b[1]) not in [(c[0], c[1]) for c in cons]]
        x0 = sysd["x0"].copy()
        v0 = sig * atom_normals(0, 1)
        inv_m = 1.0 / sysd["masses"]
        if constrained:
            x0 = shake_positions(x0, x0, inv_m, cons)
            v0 = rattle_velocities(v0, x0, inv_m, cons)
        traj = interpret(rec, sysd_c, x0, v0, n_steps, lambda d: atom_normals(d, 0), cons)
        assert n_O * n_steps == len(traj) * n_O
        key = "t%02d" % cid
        out[key + "_x0"], out[key + "_v0"] = x0, v0
        out[key + "_x"] = np.array([t[0] for t in traj])
        out[key + "_v"] = np.array([t[1] for t in traj])
        out[key + "_pe"] = np.array([t[2] for t in traj])
        out[key + "_ke"] = np.array([t[3] for t in traj])
        out[key + "_heat"] = np.array([t[4] for t in traj])
        out[key + "_shadow_work"] = np.array([t[5] for t in traj])
        meta.append(dict(key=key, splitting=splitting, constrained=constrained, seed=seed, replica=replica,
                         n_steps=n_steps, kT=kT, constraints=cons, **kw))
    out["meta"] = np.array(json.dumps(meta))
    np.savez_compressed(os.path.join(HERE, "lsi_trajectories.npz"), **out)
    print("lsi_trajectories.npz:", len(cases), "cases")


# --------------------------------------------------------------------------- 4. estimators
def make_estimators():
    ns = {"np": np}
    extract_functions("benchmark/evaluation/analysis.py", ["estimate_nonequilibrium_free_energy"], ns)
    unit = types.SimpleNamespace(Quantity=type("Quantity", (), {}))
    ns2 = {"np": np, "unit": unit, "collision_rate": 1.0}
    extract_functions("benchmark/evaluation/compare_near_eq_and_exact.py",
                      ["stdev_log_rho_pi", "stdev_kl_div", "estimate_from_work_samples", "n_steps_",
                       "process_outer_samples"], ns2)
    rng = np.random.RandomState(20261019)
    cases = []
    for n, scale, shift in [(10, 1.0, 0.0), (1000, 0.3, 0.05), (512, 2.0, -0.5), (2, 1.0, 0.0), (2048, 0.01, 1e-4)]:
        wf = shift + scale * rng.randn(n)
        wr = -shift + scale * (0.6 * (wf - shift) / scale + 0.8 * rng.randn(n))
        for n_eff in [None, n / 3.0]:
            dF, var = ns["estimate_nonequilibrium_free_energy"](wf, wr, n_eff)
            cases.append(dict(kind="neq", W_F=wf.tolist(), W_R=wr.tolist(), N_eff=n_eff,
                              DeltaF=float(dF), sq_unc=float(var)))
    for n, scale in [(50, 1.0), (777, 0.2), (1500, 3.0)]:
        w = 0.1 + scale * rng.randn(n)
        cases.append(dict(kind="logmeanexp", w=w.tolist(), estimate=float(ns2["estimate_from_work_samples"](w)),
                          stdev=float(ns2["stdev_log_rho_pi"](w))))
    rows = [0.05 + 0.5 * rng.randn(int(k)) for k in [50, 75, 200, 51, 1000]]
    outer = [{"Ws": r, "estimate": ns2["estimate_from_work_samples"](r)} for r in rows]
    cases.append(dict(kind="outer", rows=[r.tolist() for r in rows], stdev_kl=float(ns2["stdev_kl_div"](outer)),
                      new_estimate=float(ns2["process_outer_samples"](outer))))
    for dt in [0.0001, 0.0005, 0.001, 0.002, 0.004, 0.008, 0.0045]:
        for ncoll in [1, 2]:
            cases.append(dict(kind="n_steps", dt=dt, n_collisions=ncoll, value=int(ns2["n_steps_"](dt, ncoll))))
    json.dump(cases, open(os.path.join(HERE, "estimators.json"), "w"))
    print("estimators.json:", len(cases), "cases")


# --------------------------------------------------------------------------- 5. MTS strings
def make_mts():
    import itertools
    ns = {"np": np, "itertools": itertools}
    names = ["generate_gbaoab_solvent_solute_string", "generate_baoab_mts_string",
             "generate_baoab_mts_string_from_ratios", "generate_gbaoab_string", "generate_sequential_BAOAB_string",
             "generate_all_BAOAB_permutation_strings", "generate_mts_string", "generate_from_ratios"]
    extract_functions("benchmark/utilities/mts_utilities.py", names, ns)
    out = []
    for K_p in [1, 2, 3]:
        for K_r in [1, 2, 3]:
            out.append(dict(fn="generate_gbaoab_solvent_solute_string", kwargs=dict(K_p=K_p, K_r=K_r),
                            value=ns["generate_gbaoab_solvent_solute_string"](K_p=K_p, K_r=K_r)))
    out.append(dict(fn="generate_gbaoab_solvent_solute_string", kwargs=dict(K_p=2, K_r=1, slow_group=[0, 2], fast_group=[1, 3]),
                    value=ns["generate_gbaoab_solvent_solute_string"](K_p=2, K_r=1, slow_group=[0, 2], fast_group=[1, 3])))
    for groups, K_r in [([([0], 1), ([1], 2), ([2], 2)], 1), ([([3], 1), ([1, 2], 2), ([0], 3)], 2), ([([0], 1), ([1], 4)], 1)]:
        out.append(dict(fn="generate_baoab_mts_string", kwargs=dict(groups=[[list(g[0]), g[1]] for g in groups], K_r=K_r),
                        value=ns["generate_baoab_mts_string"](groups, K_r)))
    for b, a in [(5, 7), (1, 1), (2, 3)]:
        out.append(dict(fn="generate_baoab_mts_string_from_ratios", kwargs=dict(bond_steps_per_angle_step=b, angle_steps=a),
                        value=ns["generate_baoab_mts_string_from_ratios"](b, a)))
    for K_r in [1, 2, 5]:
        out.append(dict(fn="generate_gbaoab_string", kwargs=dict(K_r=K_r), value=ns["generate_gbaoab_string"](K_r)))
    for fgl, sym in [((0, 1, 2), True), ((0, 1, 2), False), ((2, 0), True), ((1,), False)]:
        out.append(dict(fn="generate_sequential_BAOAB_string", kwargs=dict(force_group_list=list(fgl), symmetric=sym),
                        value=ns["generate_sequential_BAOAB_string"](fgl, sym)))
    for n, sym in [(2, True), (3, True), (3, False)]:
        val = ns["generate_all_BAOAB_permutation_strings"](n, sym)
        out.append(dict(fn="generate_all_BAOAB_permutation_strings", kwargs=dict(n_force_groups=n, symmetric=sym),
                        value=[[list(p), s] for p, s in val]))
    json.dump(out, open(os.path.join(HERE, "mts_strings.json"), "w"), indent=0)
    print("mts_strings.json:", len(out), "cases")


# --------------------------------------------------------------------------- 6. HMR
class _Atom:
    def __init__(self, symbol):
        self.element = types.SimpleNamespace(symbol=symbol)


class _Topology:
    def __init__(self, symbols, bonds):
        self._atoms = [_Atom(s) for s in symbols]
        self._bonds = bonds

    def atoms(self):
        return iter(self._atoms)

    def bonds(self):
        return iter([(self._atoms[a], self._atoms[b]) for a, b in self._bonds])

    def getNumAtoms(self):
        return len(self._atoms)


class _System:
    def __init__(self, masses):
        self.m = list(masses)

    def getNumParticles(self):
        return len(self.m)

    def getParticleMass(self, i):
        return _Mass(self.m[i])

    def setParticleMass(self, i, val):
        self.m[i] = float(val)


ALANINE_SYMBOLS = ["H", "C", "H", "H", "C", "O", "N", "H", "C", "H", "C", "H", "H", "H", "C", "O", "N", "H", "C", "H", "H", "H"]
ALANINE_BONDS = [(0, 1), (1, 2), (1, 3), (1, 4), (4, 5), (4, 6), (6, 7), (6, 8), (8, 9), (8, 10), (10, 11), (10, 12),
                 (10, 13), (8, 14), (14, 15), (14, 16), (16, 17), (16, 18), (18, 19), (18, 20), (18, 21)]
ELEMENT_MASS = {"H": 1.008, "C": 12.01, "N": 14.01, "O": 16.0}


def make_hmr():
    util = load_by_path("ref_openmm_utilities", "benchmark/utilities/openmm_utilities.py")
    # the stand-in atoms are compared by identity in get_atoms_bonded_to_hydrogen's index lookup
    symbols, bonds = ALANINE_SYMBOLS, ALANINE_BONDS
    masses = [ELEMENT_MASS[s] for s in symbols]
    out = {"symbols": symbols, "bonds": bonds, "masses": masses, "cases": []}
    top = _Topology(symbols, bonds)
    # give atoms the `.index` attribute OpenMM atoms have
    for i, a in enumerate(top._atoms):
        a.index = i
    for scale in [1.0, 1.5, 2.0, 3.0, 4.0]:
        sys_ = _System(masses)
        try:
            hm = util.repartition_hydrogen_mass_amber(top, sys_, scale_factor=scale)
            out["cases"].append(dict(fn="amber", scale_factor=scale, masses=list(hm.m)))
        except Exception as e:  # noqa: BLE001
            out["cases"].append(dict(fn="amber", scale_factor=scale, error=type(e).__name__))
    for mode in ["decrement", "scale"]:
        for atoms in ["connected", "all"]:
            for h_mass in [0.1, 1.0, 4.0]:
                sys_ = _System(masses)
                try:
                    hm = util.repartition_hydrogen_mass(top, sys_, h_mass, mode, atoms)
                    out["cases"].append(dict(fn="generic", mode=mode, atoms=atoms, h_mass=h_mass, masses=list(hm.m)))
                except Exception as e:  # noqa: BLE001
                    out["cases"].append(dict(fn="generic", mode=mode, atoms=atoms, h_mass=h_mass, error=type(e).__name__))
    json.dump(out, open(os.path.join(HERE, "hmr.json"), "w"))
    print("hmr.json:", len(out["cases"]), "cases")


# --------------------------------------------------------------------------- 7. waterbox fixture
def make_waterbox():
    buf = open(os.path.join(REF, "data/tiny_waterbox_constrained_samples.h5"), "rb").read()
    streams = {}
    i = 0
    while True:
        i = buf.find(b"\x78", i)
        if i < 0:
            break
        if buf[i + 1] in (0x9C, 0x5E, 0xDA, 0x01):
            try:
                d = zlib.decompressobj()
                raw = d.decompress(buf[i:])
                streams[i] = raw
            except zlib.error:
                pass
        i += 1
    coords = None
    energies = None
    for off, raw in streams.items():
        if len(raw) == 17 * 306 * 3 * 4:
            coords = np.frombuffer(raw, np.uint8).reshape(4, -1).T.copy().view("<f4").reshape(17, 306, 3)
            coords_off = off
        elif off == 38753:
            energies = np.frombuffer(raw, np.uint8).reshape(4, -1).T.copy().view("<f4").ravel()
    assert coords is not None, sorted((k, len(v)) for k, v in streams.items())
    frames = coords[:4].astype(np.float32)
    e = None if energies is None else energies[:4]
    np.savez_compressed(os.path.join(HERE, "waterbox_frames.npz"), xyz=frames, box=np.array([1.5, 1.5, 1.5]),
                        potential_energy=e if e is not None else np.zeros(0), source_offset=np.array(coords_off))
    print("waterbox_frames.npz:", frames.shape, e)


if __name__ == "__main__":
    install_standins()
    which = sys.argv[1:] or ["toy", "programs", "traj", "estimators", "mts", "hmr", "waterbox"]
    lang = None
    if "toy" in which:
        make_toy()
    if "programs" in which or "traj" in which:
        lang = make_programs()
    if "traj" in which:
        make_trajectories(lang)
    if "estimators" in which:
        make_estimators()
    if "mts" in which:
        make_mts()
    if "hmr" in which:
        make_hmr()
    if "waterbox" in which:
        make_waterbox()

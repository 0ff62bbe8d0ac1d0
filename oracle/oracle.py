"""ctypes front-end of the CPU oracle (oracle/lsi_oracle.c + oracle/splitting.py).

TEST INFRASTRUCTURE ONLY.  Importable from tests/, __graft_entry__.smoke() and
bench.py's cpu_baseline / --impl reference legs; never from the product.
"""
import ctypes
import os
import subprocess

import numpy as np

from . import splitting as _sp

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "liblsi_oracle.so")
_lib = None

POT = {"quartic": 0, "double_well": 1, "harmonic": 2}
MARGINAL = {"configuration": 0, "full": 1}

c_double_p = ctypes.POINTER(ctypes.c_double)
c_int32_p = ctypes.POINTER(ctypes.c_int32)
c_uint32_p = ctypes.POINTER(ctypes.c_uint32)


def build(force=False):
    src = os.path.join(_HERE, "lsi_oracle.c")
    if force or not os.path.exists(_LIB_PATH) or os.path.getmtime(_LIB_PATH) < os.path.getmtime(src):
        subprocess.check_call(["make", "-s", "-C", _HERE, "liblsi_oracle.so"] + (["-B"] if force else []))
    return _LIB_PATH


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_LIB_PATH):
            build()
        _lib = ctypes.CDLL(_LIB_PATH)
        _lib.orc_toy_potential.restype = ctypes.c_double
        _lib.orc_toy_force.restype = ctypes.c_double
        _lib.orc_cache_index.restype = ctypes.c_int64
        _lib.orc_toy_cache_index.restype = ctypes.c_int64
    return _lib


def set_threads(n=None):
    """OpenMP threads of the oracle's replica loops (torchrun exports OMP_NUM_THREADS=1; the CPU baseline
    legs of bench.py ask for every host core explicitly).  Returns the count in effect."""
    lib()
    n = int(n or os.cpu_count() or 1)
    try:
        gomp = ctypes.CDLL("libgomp.so.1")
        gomp.omp_set_num_threads(n)
        gomp.omp_get_max_threads.restype = ctypes.c_int
        return int(gomp.omp_get_max_threads())
    except OSError:
        return int(os.environ.get("OMP_NUM_THREADS", 1))


def _dp(a):
    return None if a is None else a.ctypes.data_as(c_double_p)


def _ops_arrays(program):
    kind = np.array([op[0] for op in program.ops], dtype=np.int32)
    group = np.array([op[1] for op in program.ops], dtype=np.int32)
    frac = np.array([op[2] for op in program.ops], dtype=np.float64)
    a = np.array([op[3] for op in program.ops], dtype=np.float64)
    b = np.array([op[4] for op in program.ops], dtype=np.float64)
    return kind, group, frac, a, b


def philox4x32_10(ctr, key):
    c = (ctypes.c_uint32 * 4)(*[int(x) & 0xFFFFFFFF for x in ctr])
    k = (ctypes.c_uint32 * 2)(*[int(x) & 0xFFFFFFFF for x in key])
    out = (ctypes.c_uint32 * 4)()
    lib().orc_philox4x32_10(c, k, out)
    return [int(x) for x in out]


def normals4(seed, replica, draw, tag, unit):
    out = (ctypes.c_double * 4)()
    lib().orc_normals4(ctypes.c_uint64(seed), ctypes.c_uint64(replica), ctypes.c_uint32(draw),
                       ctypes.c_uint32(tag), ctypes.c_uint32(unit), out)
    return np.array(out[:])


def cache_index(seed, replica, n):
    """particle systems: lane 0 of the replica's TAG_X0 block"""
    return int(lib().orc_cache_index(ctypes.c_uint64(seed), ctypes.c_uint64(replica), ctypes.c_int64(n)))


def toy_cache_index(seed, replica, n):
    """scalar systems: word 2 of the replica's start block (its words 0, 1 are the start / midpoint velocity pair)"""
    return int(lib().orc_toy_cache_index(ctypes.c_uint64(seed), ctypes.c_uint64(replica), ctypes.c_int64(n)))


def toy_potential(kind, x, params=(0.0,)):
    p = (ctypes.c_double * 4)(*(list(params) + [0.0] * 4)[:4])
    return lib().orc_toy_potential(POT[kind], p, ctypes.c_double(x))


def toy_force(kind, x, params=(0.0,)):
    p = (ctypes.c_double * 4)(*(list(params) + [0.0] * 4)[:4])
    return lib().orc_toy_force(POT[kind], p, ctypes.c_double(x))


def toy_protocol(potential, splitting, dt, gamma, n_replicas, n_steps, marginal="configuration",
                 mass=10.0, kT=1.0, seed=0, replica0=0, x_cache=None, x0=None, v0=None, n_legs=2,
                 traces=False, pot_params=(0.0,), program=None, want_direct=False, noise_fp32=False):
    """Returns dict(W (n, n_legs), Q, x_final, v_final[, Wdirect, trace_x, trace_v, trace_w])."""
    prog = program or _sp.compile_splitting(splitting, dt, gamma)
    kind, _group, frac, a, b = _ops_arrays(prog)
    p = (ctypes.c_double * 4)(*(list(pot_params) + [0.0] * 4)[:4])
    n = int(n_replicas)
    W = np.zeros((n, n_legs))
    Q = np.zeros((n, n_legs))
    wd = np.zeros((n, n_legs)) if want_direct else None
    tx = np.zeros((n, n_legs, n_steps + 1)) if traces else None
    tv = np.zeros((n, n_legs, n_steps + 1)) if traces else None
    tw = np.zeros((n, n_legs, n_steps + 1)) if traces else None
    xf, vf = np.zeros(n), np.zeros(n)
    xc = None if x_cache is None else np.ascontiguousarray(x_cache, dtype=np.float64)
    x0a = None if x0 is None else np.ascontiguousarray(x0, dtype=np.float64)
    v0a = None if v0 is None else np.ascontiguousarray(v0, dtype=np.float64)
    assert xc is not None or x0a is not None
    rc = lib().orc_toy_protocol(
        POT[potential], p, ctypes.c_double(mass), ctypes.c_double(kT),
        len(kind), kind.ctypes.data_as(c_int32_p), _dp(frac), _dp(a), _dp(b), ctypes.c_double(dt),
        ctypes.c_uint64(seed), ctypes.c_uint64(replica0), ctypes.c_int64(n),
        _dp(xc), ctypes.c_int64(0 if xc is None else len(xc)), _dp(x0a), _dp(v0a),
        int(n_steps), MARGINAL[marginal], int(n_legs), int(bool(noise_fp32)),
        _dp(W), _dp(Q), _dp(wd), _dp(tx), _dp(tv), _dp(tw), _dp(xf), _dp(vf))
    assert rc == 0
    out = {"W": W, "Q": Q, "x_final": xf, "v_final": vf}
    if want_direct:
        out["Wdirect"] = wd
    if traces:
        out.update(trace_x=tx, trace_v=tv, trace_w=tw)
    return out


def toy_replay(potential, splitting, dt, gamma, x0, v0, n_steps, xi, mass=10.0, kT=1.0, pot_params=(0.0,)):
    """One trajectory driven by an external noise array (one entry per O substep)."""
    prog = _sp.compile_splitting(splitting, dt, gamma)
    kind, _group, frac, a, b = _ops_arrays(prog)
    p = (ctypes.c_double * 4)(*(list(pot_params) + [0.0] * 4)[:4])
    xi = np.ascontiguousarray(xi, dtype=np.float64)
    assert len(xi) >= n_steps * prog.n_O
    tx, tv, tw = np.zeros(n_steps + 1), np.zeros(n_steps + 1), np.zeros(n_steps + 1)
    Q = ctypes.c_double(0.0)
    lib().orc_toy_replay(POT[potential], p, ctypes.c_double(mass), ctypes.c_double(kT),
                         len(kind), kind.ctypes.data_as(c_int32_p), _dp(frac), _dp(a), _dp(b),
                         ctypes.c_double(dt), ctypes.c_double(x0), ctypes.c_double(v0), int(n_steps),
                         _dp(xi), _dp(tx), _dp(tv), _dp(tw), ctypes.byref(Q))
    return tx, tv, Q.value, tw


def sample_equilibrium_scalar(potential, n, n_burn, kT=1.0, seed=0, replica0=0, pot_params=(0.0,)):
    p = (ctypes.c_double * 4)(*(list(pot_params) + [0.0] * 4)[:4])
    out = np.zeros(int(n))
    lib().orc_sample_equilibrium_scalar(POT[potential], p, ctypes.c_double(kT), ctypes.c_int64(n), int(n_burn),
                                        ctypes.c_uint64(seed), ctypes.c_uint64(replica0), _dp(out))
    return out


# ----------------------------------------------------------------------------- particle systems
c_int32_pp = ctypes.POINTER(ctypes.c_int32)


class OrcParticles(ctypes.Structure):
    _fields_ = [
        ("n_atoms", ctypes.c_int32), ("mass", c_double_p), ("charge", c_double_p), ("sigma", c_double_p),
        ("epsilon", c_double_p), ("nonbonded_method", ctypes.c_int32), ("cutoff", ctypes.c_double),
        ("box", ctypes.c_double * 3), ("rf_dielectric", ctypes.c_double),
        ("n_exclusions", ctypes.c_int32), ("exclusions", c_int32_pp),
        ("n_exceptions", ctypes.c_int32), ("exception_atoms", c_int32_pp), ("exception_params", c_double_p),
        ("n_bonds", ctypes.c_int32), ("bond_atoms", c_int32_pp), ("bond_params", c_double_p),
        ("n_angles", ctypes.c_int32), ("angle_atoms", c_int32_pp), ("angle_params", c_double_p),
        ("n_torsions", ctypes.c_int32), ("torsion_atoms", c_int32_pp), ("torsion_params", c_double_p),
        ("restraint_k", ctypes.c_double),
        ("group_nonbonded", ctypes.c_int32), ("group_bonds", ctypes.c_int32), ("group_angles", ctypes.c_int32),
        ("group_torsions", ctypes.c_int32), ("group_restraint", ctypes.c_int32),
        ("n_settle", ctypes.c_int32), ("settle_atoms", c_int32_pp), ("settle_oh", ctypes.c_double),
        ("settle_hh", ctypes.c_double),
        ("n_constraints", ctypes.c_int32), ("constraint_atoms", c_int32_pp), ("constraint_dist", c_double_p),
        ("constraint_tolerance", ctypes.c_double), ("switch_distance", ctypes.c_double),
    ]


def _f64(a, shape=None):
    a = np.ascontiguousarray(a if a is not None else [], dtype=np.float64)
    return a.reshape(shape) if shape is not None else a


def _i32(a, cols):
    a = np.ascontiguousarray(a if a is not None else [], dtype=np.int32)
    return a.reshape(-1, cols)


def make_particles(desc):
    """desc: plain dict of arrays (see integrator_benchmark_b200.testsystems.systems.describe).
    Returns (struct, keepalive)."""
    keep = {}
    S = OrcParticles()
    n = int(desc["n_atoms"])
    S.n_atoms = n
    for name in ("mass", "charge", "sigma", "epsilon"):
        keep[name] = _f64(desc.get(name, np.zeros(n)))
        assert keep[name].shape == (n,)
        setattr(S, name, _dp(keep[name]))
    S.nonbonded_method = {"NoCutoff": 0, "CutoffPeriodic": 1}[desc.get("nonbonded_method", "NoCutoff")]
    S.cutoff = float(desc.get("cutoff", 0.0))
    box = list(desc.get("box", (0.0, 0.0, 0.0)))
    S.box = (ctypes.c_double * 3)(*box)
    S.rf_dielectric = float(desc.get("rf_dielectric", 78.3))
    S.switch_distance = float(desc.get("switch_distance", 0.0) or 0.0)

    def ilist(key, cols):
        a = _i32(desc.get(key), cols)
        keep[key] = a
        return len(a), a.ctypes.data_as(c_int32_pp)

    def dlist(key, cols):
        a = _f64(desc.get(key)).reshape(-1, cols) if cols else _f64(desc.get(key))
        keep[key] = a
        return _dp(a)

    S.n_exclusions, S.exclusions = ilist("exclusions", 2)
    S.n_exceptions, S.exception_atoms = ilist("exception_atoms", 2)
    S.exception_params = dlist("exception_params", 3)
    S.n_bonds, S.bond_atoms = ilist("bond_atoms", 2)
    S.bond_params = dlist("bond_params", 2)
    S.n_angles, S.angle_atoms = ilist("angle_atoms", 3)
    S.angle_params = dlist("angle_params", 2)
    S.n_torsions, S.torsion_atoms = ilist("torsion_atoms", 4)
    S.torsion_params = dlist("torsion_params", 3)
    S.restraint_k = float(desc.get("restraint_k", 0.0))
    groups = desc.get("groups", {})
    for k in ("nonbonded", "bonds", "angles", "torsions", "restraint"):
        setattr(S, "group_" + k, int(groups.get(k, 0)))
    S.n_settle, S.settle_atoms = ilist("settle_atoms", 3)
    S.settle_oh = float(desc.get("settle_oh", 0.0))
    S.settle_hh = float(desc.get("settle_hh", 0.0))
    S.n_constraints, S.constraint_atoms = ilist("constraint_atoms", 2)
    S.constraint_dist = dlist("constraint_dist", 0)
    S.constraint_tolerance = float(desc.get("constraint_tolerance", 1e-8))
    return S, keep


def particles_energy_forces(desc, x, group=-1, force_fp32=False):
    """potential energy and forces; force_fp32: the "mixed" restatement (fp32 pair / bonded arithmetic, energies summed in fp64)"""
    S, _keep = make_particles(desc)
    x = _f64(x).reshape(-1)
    f = np.zeros_like(x)
    fn = lib().orc_particles_energy_forces_f32 if force_fp32 else lib().orc_particles_energy_forces
    fn.restype = ctypes.c_double
    e = fn(ctypes.byref(S), _dp(x), int(group), _dp(f))
    return e, f.reshape(-1, 3)


def constrain_positions(desc, x_ref, x_new):
    S, _keep = make_particles(desc)
    x0 = _f64(x_ref).reshape(-1).copy()
    x1 = _f64(x_new).reshape(-1).copy()
    lib().orc_constrain_positions(ctypes.byref(S), _dp(x0), _dp(x1))
    return x1.reshape(-1, 3)


def constrain_velocities(desc, x, v):
    S, _keep = make_particles(desc)
    x = _f64(x).reshape(-1).copy()
    v = _f64(v).reshape(-1).copy()
    lib().orc_constrain_velocities(ctypes.byref(S), _dp(x), _dp(v))
    return v.reshape(-1, 3)


def particles_protocol(desc, splitting, dt, gamma, kT, n_replicas, n_steps, marginal="configuration", seed=0,
                       replica0=0, x_cache=None, x0=None, v0=None, n_legs=2, traces=False, want_direct=False,
                       noise_fp32=False, round_midpoint=False, program=None, force_fp32=False, mixed=False):
    """mixed=True is the engine's LSI_PRECISION_MIXED: fp32 O noise AND fp32 force / energy arithmetic."""
    if mixed:
        noise_fp32 = force_fp32 = True
    prog = program or _sp.compile_splitting(splitting, dt, gamma)
    kind, group, frac, a, b = _ops_arrays(prog)
    S, _keep = make_particles(desc)
    n, N = int(n_replicas), int(desc["n_atoms"])
    W, Q = np.zeros((n, n_legs)), np.zeros((n, n_legs))
    wd = np.zeros((n, n_legs)) if want_direct else None
    xf, vf = np.zeros((n, N, 3)), np.zeros((n, N, 3))
    T = n_steps + 1
    tx = np.zeros((n, n_legs, T, N, 3)) if traces else None
    tv = np.zeros((n, n_legs, T, N, 3)) if traces else None
    tpe, tke, th, twd = [np.zeros((n, n_legs, T)) if traces else None for _ in range(4)]
    xc = None if x_cache is None else _f64(x_cache).reshape(-1, N, 3)
    x0a = None if x0 is None else _f64(x0).reshape(n, N, 3)
    v0a = None if v0 is None else _f64(v0).reshape(n, N, 3)
    assert xc is not None or x0a is not None
    rc = lib().orc_particles_protocol(
        ctypes.byref(S), ctypes.c_double(kT), len(kind), kind.ctypes.data_as(c_int32_p),
        group.ctypes.data_as(c_int32_p), _dp(frac), _dp(a), _dp(b), ctypes.c_double(dt),
        ctypes.c_uint64(seed), ctypes.c_uint64(replica0), ctypes.c_int64(n),
        _dp(xc), ctypes.c_int64(0 if xc is None else len(xc)), _dp(x0a), _dp(v0a),
        int(n_steps), MARGINAL[marginal], int(n_legs), int(bool(noise_fp32)), int(bool(round_midpoint)) | (2 if force_fp32 else 0),
        _dp(W), _dp(Q), _dp(wd), _dp(xf), _dp(vf), _dp(tx), _dp(tv), _dp(tpe), _dp(tke), _dp(th), _dp(twd))
    assert rc == 0
    out = {"W": W, "Q": Q, "x_final": xf, "v_final": vf}
    if want_direct:
        out["Wdirect"] = wd
    if traces:
        out.update(trace_x=tx, trace_v=tv, trace_pe=tpe, trace_ke=tke, trace_heat=th, trace_wdir=twd)
    return out


def particles_hmc(desc, kT, x_in, n_rounds, steps_per_hmc=10, extra_chances=15, dt=0.001, seed=0, replica0=0):
    """Extra-chance HMC equilibrium sampler; returns (x_out (n, N, 3), v_out, accept (n, 2))."""
    S, _keep = make_particles(desc)
    N = int(desc["n_atoms"])
    x_in = _f64(x_in).reshape(-1, N, 3)
    n = len(x_in)
    x_out, v_out, acc = np.zeros_like(x_in), np.zeros_like(x_in), np.zeros((n, 2))
    rc = lib().orc_particles_hmc(ctypes.byref(S), ctypes.c_double(kT), ctypes.c_int64(n), int(n_rounds), int(steps_per_hmc),
                                 int(extra_chances), ctypes.c_double(dt), ctypes.c_uint64(seed), ctypes.c_uint64(replica0),
                                 _dp(x_in), _dp(x_out), _dp(v_out), _dp(acc))
    assert rc == 0
    return x_out, v_out, acc

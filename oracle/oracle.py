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
    return _lib


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
    return int(lib().orc_cache_index(ctypes.c_uint64(seed), ctypes.c_uint64(replica), ctypes.c_int64(n)))


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

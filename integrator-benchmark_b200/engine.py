"""Host-side handles over the C ABI: compiled programs, systems and the protocol launch.

PyTorch is used here for exactly three things: allocating device buffers, exposing their
raw pointers (`data_ptr()`), and naming the current CUDA stream.  All arithmetic happens
in liblsi.so.
"""
import ctypes

import numpy as np

from . import _cabi
from ._cabi import check, lib

_torch = None


def torch():
    global _torch
    if _torch is None:
        import torch as _t
        _torch = _t
    return _torch


def require_cuda():
    """The engine has no CPU path: fail loudly when no device is usable."""
    if lib.lsi_device_count() <= 0 or not torch().cuda.is_available():
        raise _cabi.LsiCudaError("no CUDA device available: integrator_benchmark_b200 runs on the GPU only "
                                 "(there is no CPU fallback)")


def current_stream_ptr():
    return ctypes.c_void_p(torch().cuda.current_stream().cuda_stream)


class Program:
    """A compiled splitting (lsi_program*)."""

    def __init__(self, splitting, dt, gamma, measure_shadow_work=False, measure_heat=True,
                 override_splitting_checks=False):
        h = ctypes.c_void_p()
        if isinstance(splitting, str):
            check(lib.lsi_compile(splitting.encode(), float(dt), float(gamma), int(bool(measure_shadow_work)),
                                  int(bool(measure_heat)), int(bool(override_splitting_checks)), ctypes.byref(h)))
        else:
            toks = [str(t[0]).encode() for t in splitting]
            fracs = [float(t[1]) for t in splitting]
            arr_t = (ctypes.c_char_p * len(toks))(*toks)
            arr_f = (ctypes.c_double * len(fracs))(*fracs)
            check(lib.lsi_compile_continuous(arr_t, arr_f, len(toks), float(dt), float(gamma),
                                             int(bool(measure_shadow_work)), int(bool(measure_heat)),
                                             int(bool(override_splitting_checks)), ctypes.byref(h)))
        self._h = h

    def __del__(self):
        h, self._h = getattr(self, "_h", None), None
        if h and lib is not None:  # module globals may already be gone at interpreter shutdown
            lib.lsi_program_destroy(h)

    @property
    def handle(self):
        return self._h

    @property
    def dt(self):
        return lib.lsi_program_get_step_size(self._h)

    @dt.setter
    def dt(self, value):
        check(lib.lsi_program_set_step_size(self._h, float(value)))

    @property
    def n_O(self):
        return lib.lsi_program_num_O(self._h)

    @property
    def mts(self):
        return bool(lib.lsi_program_is_mts(self._h))

    def ops(self):
        out = []
        k, g = ctypes.c_int(), ctypes.c_int()
        f, a, b = ctypes.c_double(), ctypes.c_double(), ctypes.c_double()
        for i in range(lib.lsi_program_num_ops(self._h)):
            check(lib.lsi_program_get_op(self._h, i, ctypes.byref(k), ctypes.byref(g), ctypes.byref(f),
                                         ctypes.byref(a), ctypes.byref(b)))
            out.append((k.value, g.value, f.value, a.value, b.value))
        return out


class ScalarSystem:
    """1-DOF toy system (lsi_system*, LSI_SYS_SCALAR)."""
    KINDS = {"quartic": _cabi.POT_QUARTIC, "double_well": _cabi.POT_DOUBLE_WELL, "harmonic": _cabi.POT_HARMONIC}

    def __init__(self, kind, mass=10.0, params=(0.0, 0.0, 0.0, 0.0)):
        p = (ctypes.c_double * 4)(*(list(params) + [0.0] * 4)[:4])
        h = ctypes.c_void_p()
        check(lib.lsi_system_create_scalar(self.KINDS[kind], p, float(mass), ctypes.byref(h)))
        self._h, self.kind, self.mass = h, kind, float(mass)

    def __del__(self):
        h, self._h = getattr(self, "_h", None), None
        if h and lib is not None:
            lib.lsi_system_destroy(h)

    @property
    def handle(self):
        return self._h

    n_dof = 1


class ParticleSystem:
    """Particle system (lsi_system*, LSI_SYS_PARTICLES) built from a plain description dict:

        n_atoms, mass, charge, sigma, epsilon,
        nonbonded_method ("NoCutoff" | "CutoffPeriodic"), cutoff, box, rf_dielectric, switch_distance,
        exclusions (n,2), exception_atoms (n,2), exception_params (n,3: chargeProd, sigma, epsilon),
        bond_atoms (n,2), bond_params (n,2: r0, k), angle_atoms (n,3), angle_params (n,2: theta0, k),
        torsion_atoms (n,4), torsion_params (n,3: periodicity, phase, k), restraint_k,
        groups {nonbonded, bonds, angles, torsions, restraint -> force group id},
        settle_atoms (n,3: O, H1, H2), settle_oh, settle_hh,
        constraint_atoms (n,2), constraint_dist (n,), constraint_tolerance

    (OpenMM units).  The same dict is what testsystems/*.py produce for the reference's systems."""

    def __init__(self, desc):
        self.desc = dict(desc)
        n = int(desc["n_atoms"])
        keep = []
        d = _cabi.ParticlesDesc()
        d.n_atoms = n

        def dbl(key, cols=None, default=None):
            a = desc.get(key, default)
            a = np.ascontiguousarray([] if a is None else a, dtype=np.float64)
            if cols:
                a = a.reshape(-1, cols)
            keep.append(a)
            return a, a.ctypes.data_as(_cabi.c_double_p)

        def i32(key, cols):
            a = np.ascontiguousarray(desc.get(key) if desc.get(key) is not None else [], dtype=np.int32).reshape(-1, cols)
            keep.append(a)
            return a, a.ctypes.data_as(_cabi.c_int32_pp)

        for name, default in (("mass", None), ("charge", np.zeros(n)), ("sigma", np.ones(n)), ("epsilon", np.zeros(n))):
            a, ptr = dbl(name, default=default)
            if a.shape != (n,):
                raise ValueError("%s must have shape (n_atoms,)" % name)
            setattr(d, name, ptr)
        d.nonbonded_method = {"NoCutoff": 0, "CutoffPeriodic": 1}[desc.get("nonbonded_method", "NoCutoff")]
        d.cutoff = float(desc.get("cutoff", 0.0))
        d.box = (ctypes.c_double * 3)(*[float(b) for b in desc.get("box", (0.0, 0.0, 0.0))])
        d.rf_dielectric = float(desc.get("rf_dielectric", 78.3))
        a, d.exclusions = i32("exclusions", 2)
        d.n_exclusions = len(a)
        a, d.exception_atoms = i32("exception_atoms", 2)
        d.n_exceptions = len(a)
        b, d.exception_params = dbl("exception_params", 3)
        assert len(a) == len(b)
        for kind, cols, pcols in (("bond", 2, 2), ("angle", 3, 2), ("torsion", 4, 3)):
            a, ptr = i32(kind + "_atoms", cols)
            b, pptr = dbl(kind + "_params", pcols)
            if len(a) != len(b):
                raise ValueError("%s_atoms and %s_params disagree" % (kind, kind))
            setattr(d, "n_%ss" % kind, len(a))
            setattr(d, kind + "_atoms", ptr)
            setattr(d, kind + "_params", pptr)
        d.restraint_k = float(desc.get("restraint_k", 0.0))
        groups = desc.get("groups", {})
        for k in ("nonbonded", "bonds", "angles", "torsions", "restraint"):
            setattr(d, "group_" + k, int(groups.get(k, 0)))
        a, d.settle_atoms = i32("settle_atoms", 3)
        d.n_settle = len(a)
        d.settle_oh = float(desc.get("settle_oh", 0.0))
        d.settle_hh = float(desc.get("settle_hh", 0.0))
        a, d.constraint_atoms = i32("constraint_atoms", 2)
        d.n_constraints = len(a)
        b, d.constraint_dist = dbl("constraint_dist")
        assert len(a) == len(b)
        d.constraint_tolerance = float(desc.get("constraint_tolerance", 1e-8))
        d.switch_distance = float(desc.get("switch_distance", 0.0) or 0.0)
        h = ctypes.c_void_p()
        check(lib.lsi_system_create_particles(ctypes.byref(d), ctypes.byref(h)))
        self._h = h
        self.n_atoms = n
        self.n_dof = 3 * n
        self.mass = np.array(desc["mass"], dtype=np.float64)

    def __del__(self):
        h, self._h = getattr(self, "_h", None), None
        if h and lib is not None:
            lib.lsi_system_destroy(h)

    @property
    def handle(self):
        return self._h

    def energy_forces(self, x, group=-1, precision="double", want_forces=True):
        """potential energy [n] and forces [n, N, 3] (CUDA tensors) of configurations x [n, N, 3]."""
        require_cuda()
        t = torch()
        x = _to_cuda64(x).reshape(-1, self.n_atoms, 3)
        n = x.shape[0]
        e = t.empty(n, dtype=t.float64, device=x.device)
        f = t.empty_like(x) if want_forces else None
        with t.cuda.device(x.device):
            check(lib.lsi_particles_energy_forces(self._h, _cabi.PRECISION[precision], int(group), n, _ptr(x), _ptr(e),
                                                  _ptr(f), current_stream_ptr()))
        return e, f


def _ptr(t):
    return None if t is None else ctypes.c_void_p(t.data_ptr())


def sample_equilibrium_particles(system, x, kT, n_rounds, steps_per_hmc=10, extra_chances=15, dt=0.001, seed=0, replica0=0,
                                 precision="double", want_velocities=False):
    """Extra-chance HMC chains started from x [n, N, 3] (lsi_sample_equilibrium_particles): returns
    (x_out, v_out or None, accept [n, 2]) as CUDA tensors."""
    require_cuda()
    t = torch()
    x = _to_cuda64(x).reshape(-1, system.n_atoms, 3)
    n = x.shape[0]
    x_out = t.empty_like(x)
    v_out = t.empty_like(x) if want_velocities else None
    acc = t.zeros(n, 2, dtype=t.float64, device=x.device)
    with t.cuda.device(x.device):
        check(lib.lsi_sample_equilibrium_particles(system.handle, _cabi.PRECISION[precision], float(kT), n, int(n_rounds),
                                                   int(steps_per_hmc), int(extra_chances), float(dt), int(seed) & 0xFFFFFFFFFFFFFFFF,
                                                   int(replica0), _ptr(x), _ptr(x_out), _ptr(v_out), _ptr(acc), current_stream_ptr()))
    return x_out, v_out, acc


def _protocol_args(system, n_replicas, n_steps, kT, marginal, n_legs, seed, replica0, x_cache, x0, v0, precision, want_heat,
                   want_direct, want_final, traces, want_moments, dev, out, round_midpoint, moments_out=None):
    """lsi_protocol_args of one launch; output buffers live in (and are reused from) the dict `out`"""
    t = torch()
    n = int(n_replicas)
    dof = system.n_dof

    def buf(name, *shape):
        cur = out.get(name)
        if cur is None or tuple(cur.shape) != tuple(shape) or cur.device != dev:
            cur = t.empty(shape, dtype=t.float64, device=dev)
            out[name] = cur
        return cur

    scalar = isinstance(system, ScalarSystem)
    state_shape = (n,) if scalar else (n, dof // 3, 3)
    a = _cabi.ProtocolArgs()
    a.flags = 1 if (round_midpoint and not scalar) else 0
    a.seed, a.replica0, a.n_replicas = int(seed), int(replica0), n
    a.n_steps, a.n_legs = int(n_steps), int(n_legs)
    if marginal not in _cabi.MARGINAL:
        raise NotImplementedError("`marginal` must be either 'configuration' or 'full'")
    a.marginal, a.precision, a.kT = _cabi.MARGINAL[marginal], _cabi.PRECISION[precision], float(kT)
    a.x_cache = _ptr(x_cache)
    a.n_cache = 0 if x_cache is None else (x_cache.shape[0])
    a.x0, a.v0 = _ptr(x0), _ptr(v0)
    a.W = _ptr(buf("W", n_legs, n))
    if want_heat:
        a.heat = _ptr(buf("heat", n_legs, n))
    if want_direct:
        a.W_direct = _ptr(buf("W_direct", n_legs, n))
    if want_final:
        a.x_final = _ptr(buf("x_final", *state_shape))
        a.v_final = _ptr(buf("v_final", *state_shape))
    if traces == "replica_major":
        if not scalar:
            raise ValueError("replica-major traces are a scalar-system layout")
        a.flags |= 2  # LSI_FLAG_TRACE_REPLICA_MAJOR
        a.trace_x = _ptr(buf("trace_xv_rm", n_legs, n, n_steps + 1, 2))
        a.trace_w = _ptr(buf("trace_w_rm", n_legs, n, n_steps + 1))
    elif traces == "pairs":
        if not scalar:
            raise ValueError("(x, v)-pair traces are a scalar-system layout")
        a.flags |= 4  # LSI_FLAG_TRACE_XV_PAIRS
        a.trace_x = _ptr(buf("trace_xv", n_legs, n_steps + 1, n, 2))
        a.trace_w = _ptr(buf("trace_w", n_legs, n_steps + 1, n))
    elif traces:
        tshape = (n_legs, n_steps + 1, n) if scalar else (n_legs, n_steps + 1, n, dof // 3, 3)
        a.trace_x = _ptr(buf("trace_x", *tshape))
        a.trace_v = _ptr(buf("trace_v", *tshape))
        a.trace_w = _ptr(buf("trace_w", n_legs, n_steps + 1, n))
    if want_moments and n_legs == 2 and n > 0:
        if moments_out is not None:
            if not (moments_out.is_cuda and moments_out.dtype == t.float64 and moments_out.numel() == 8 and moments_out.is_contiguous()):
                raise ValueError("moments_out must be a contiguous float64 CUDA tensor of 8 elements")
            out["moments"] = moments_out
        a.moments = _ptr(moments_out if moments_out is not None else buf("moments", 8))
        nbytes = lib.lsi_protocol_workspace_bytes(n)
        ws = out.get("_workspace")
        if ws is None or ws.numel() * 8 < nbytes or ws.device != dev:
            ws = t.empty((nbytes + 7) // 8, dtype=t.float64, device=dev)
            out["_workspace"] = ws
        a.workspace, a.workspace_bytes = _ptr(ws), ws.numel() * 8
    out["_keepalive"] = (x_cache, x0, v0)
    return a


def _as_dev(a, dev):
    t = torch()
    if a is None:
        return None
    if not t.is_tensor(a):
        a = t.as_tensor(np.ascontiguousarray(a, dtype=np.float64))
    return a.to(device=dev, dtype=t.float64).contiguous()


def run_protocol(system, program, n_replicas, n_steps, kT, marginal="configuration", n_legs=2, seed=0,
                 replica0=0, x_cache=None, x0=None, v0=None, precision="double", want_heat=False,
                 want_direct=False, want_final=False, traces=False, want_moments=True, device=None, out=None,
                 round_midpoint=False, moments_out=None):
    """One launch of the whole protocol for `n_replicas` replicas (lsi_run_protocol).

    Tensors in/out are float64 CUDA tensors; outputs are leg-major: W[leg, replica],
    traces [leg, step, replica(, dof)].  traces="replica_major" (scalar systems) gives the reference's own layout instead:
    trace_w_rm [leg, replica, step] and trace_xv_rm [leg, replica, step, 2]; traces="pairs" keeps the coalesced step-major
    order with (x, v) interleaved, trace_xv [leg, step, replica, 2] (a permuted view of it has the reference's shape).
    `moments_out` (8 float64 on the device) receives the moment vector instead of a buffer of `out`.
    Returns a dict of tensors (no host sync)."""
    require_cuda()
    t = torch()
    dev = t.device("cuda", t.cuda.current_device()) if device is None else t.device(device)
    out = {} if out is None else out
    x_cache, x0, v0 = _as_dev(x_cache, dev), _as_dev(x0, dev), _as_dev(v0, dev)
    a = _protocol_args(system, n_replicas, n_steps, kT, marginal, n_legs, seed, replica0, x_cache, x0, v0, precision, want_heat,
                       want_direct, want_final, traces, want_moments, dev, out, round_midpoint, moments_out)
    with t.cuda.device(dev):
        check(lib.lsi_run_protocol(system.handle, program.handle, ctypes.byref(a), current_stream_ptr()))
    return out


def run_protocols(system, programs, n_replicas, n_steps, kT, marginal="configuration", n_legs=2, seed=0, replica0=0,
                  x_cache=None, x0=None, v0=None, precision="double", want_heat=False, want_final=False, want_moments=True,
                  device=None, outs=None, moments_out=None):
    """Several programs over the SAME ensemble (lsi_run_protocol_batch): the results equal run_protocol(system, p, ...) for
    every p, bit for bit; scalar systems run the six 5-letter palindromes (or the reference's three 6-letter closure
    strings) as ONE kernel launch.  `outs`: one buffer dict per program (reused between calls); `moments_out`: a
    (len(programs), 8) float64 CUDA tensor receiving the moment vectors.  Returns the list of dicts (no host sync)."""
    require_cuda()
    t = torch()
    dev = t.device("cuda", t.cuda.current_device()) if device is None else t.device(device)
    programs = list(programs)
    outs = [{} for _ in programs] if outs is None else outs
    if len(outs) != len(programs):
        raise ValueError("one output dict per program")
    x_cache, x0, v0 = _as_dev(x_cache, dev), _as_dev(x0, dev), _as_dev(v0, dev)
    arr = (_cabi.ProtocolArgs * len(programs))()
    for i, out in enumerate(outs):
        arr[i] = _protocol_args(system, n_replicas, n_steps, kT, marginal, n_legs, seed, replica0, x_cache, x0, v0, precision,
                                want_heat, False, want_final, False, want_moments, dev, out, False,
                                None if moments_out is None else moments_out[i])
    handles = (ctypes.c_void_p * len(programs))(*[p.handle for p in programs])
    with t.cuda.device(dev):
        check(lib.lsi_run_protocol_batch(system.handle, handles, len(programs), arr, current_stream_ptr()))
    return outs


def reduce_moments(W_F, W_R):
    """(count, sum wF, sum wR, sum wF^2, sum wR^2, sum wF wR, n_nonfinite, 0) as a CUDA tensor."""
    require_cuda()
    t = torch()
    W_F = _to_cuda64(W_F)
    W_R = _to_cuda64(W_R).to(W_F.device)
    if W_F.shape != W_R.shape or W_F.dim() != 1:
        raise ValueError("W_F and W_R must be 1-D arrays of equal length")
    n = W_F.numel()
    m = t.empty(8, dtype=t.float64, device=W_F.device)
    ws = t.empty(1024 * 8, dtype=t.float64, device=W_F.device)
    with t.cuda.device(W_F.device):
        check(lib.lsi_reduce_moments(_ptr(W_F), _ptr(W_R), n, _ptr(m), _ptr(ws), ws.numel() * 8, current_stream_ptr()))
    return m


def _ragged(Ws):
    """rows of work samples (list of arrays, 2-D array, or (flat, offsets)) -> (flat CUDA tensor, int64 offsets tensor)"""
    t = torch()
    if isinstance(Ws, tuple) and len(Ws) == 2:
        flat, offs = Ws
        return _to_cuda64(flat), t.as_tensor(np.asarray(offs, dtype=np.int64)).cuda() if not t.is_tensor(offs) else offs.cuda()
    if t.is_tensor(Ws) and Ws.dim() == 2:
        n, m = Ws.shape
        return _to_cuda64(Ws.reshape(-1)), t.arange(0, (n + 1) * m, m, dtype=t.int64, device="cuda")
    rows = [np.asarray(r, dtype=np.float64).reshape(-1) for r in Ws]
    offs = np.concatenate([[0], np.cumsum([len(r) for r in rows])]).astype(np.int64)
    flat = np.concatenate(rows) if rows else np.zeros(0)
    return _to_cuda64(flat), t.as_tensor(offs).cuda()


def reduce_logmeanexp(Ws):
    """per row: log(mean(exp(-w))) and std(exp(-w)) / (mean(exp(-w)) sqrt(n)) (lsi_reduce_logmeanexp;
    reference: compare_near_eq_and_exact.py:62-97).  Returns two CUDA tensors of length n_rows."""
    require_cuda()
    t = torch()
    flat, offs = _ragged(Ws)
    n_rows = offs.numel() - 1
    est = t.empty(n_rows, dtype=t.float64, device=flat.device)
    sd = t.empty(n_rows, dtype=t.float64, device=flat.device)
    if n_rows > 0:
        with t.cuda.device(flat.device):
            check(lib.lsi_reduce_logmeanexp(_ptr(flat), _ptr(offs), n_rows, _ptr(est), _ptr(sd), current_stream_ptr()))
    return est, sd


def bootstrap_nested(Ws, n_boot, seed=0):
    """n_boot two-level bootstrap replicates of mean_rows(log-mean-exp(-w)) (lsi_bootstrap_nested;
    reference: resample_Ws, compare_near_eq_and_exact.py:293-320).  Returns a CUDA tensor (n_boot,)."""
    require_cuda()
    t = torch()
    flat, offs = _ragged(Ws)
    n_rows = offs.numel() - 1
    out = t.empty(int(n_boot), dtype=t.float64, device=flat.device)
    with t.cuda.device(flat.device):
        check(lib.lsi_bootstrap_nested(_ptr(flat), _ptr(offs), n_rows, int(n_boot), int(seed) & 0xFFFFFFFFFFFFFFFF, _ptr(out),
                                       current_stream_ptr()))
    return out


def neq_free_energy_from_moments(moments, N_eff=None):
    m = (ctypes.c_double * 8)(*[float(v) for v in moments])
    dF, var = ctypes.c_double(), ctypes.c_double()
    check(lib.lsi_neq_free_energy(m, -1.0 if N_eff is None else float(N_eff), ctypes.byref(dF), ctypes.byref(var)))
    return dF.value, var.value


def _to_cuda64(a):
    t = torch()
    if t.is_tensor(a):
        return a.to(device="cuda" if not a.is_cuda else a.device, dtype=t.float64).contiguous()
    return t.as_tensor(np.ascontiguousarray(a, dtype=np.float64)).cuda()

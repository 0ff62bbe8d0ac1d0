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
        if h:
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
        if h:
            lib.lsi_system_destroy(h)

    @property
    def handle(self):
        return self._h

    n_dof = 1


def _ptr(t):
    return None if t is None else ctypes.c_void_p(t.data_ptr())


def run_protocol(system, program, n_replicas, n_steps, kT, marginal="configuration", n_legs=2, seed=0,
                 replica0=0, x_cache=None, x0=None, v0=None, precision="double", want_heat=False,
                 want_direct=False, want_final=False, traces=False, want_moments=True, device=None, out=None):
    """One launch of the whole protocol for `n_replicas` replicas (lsi_run_protocol).

    Tensors in/out are float64 CUDA tensors; outputs are leg-major: W[leg, replica],
    traces [leg, step, replica(, dof)].  Returns a dict of tensors (no host sync)."""
    require_cuda()
    t = torch()
    dev = t.device("cuda", t.cuda.current_device()) if device is None else t.device(device)
    n = int(n_replicas)
    dof = system.n_dof
    out = {} if out is None else out

    def buf(name, *shape):
        cur = out.get(name)
        if cur is None or tuple(cur.shape) != tuple(shape) or cur.device != dev:
            cur = t.empty(shape, dtype=t.float64, device=dev)
            out[name] = cur
        return cur

    def as_dev(a):
        if a is None:
            return None
        if not t.is_tensor(a):
            a = t.as_tensor(np.ascontiguousarray(a, dtype=np.float64))
        return a.to(device=dev, dtype=t.float64).contiguous()

    x_cache, x0, v0 = as_dev(x_cache), as_dev(x0), as_dev(v0)
    state_shape = (n,) if dof == 1 else (n, dof)
    a = _cabi.ProtocolArgs()
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
    if traces:
        tshape = (n_legs, n_steps + 1, n) if dof == 1 else (n_legs, n_steps + 1, n, dof)
        a.trace_x = _ptr(buf("trace_x", *tshape))
        a.trace_v = _ptr(buf("trace_v", *tshape))
        a.trace_w = _ptr(buf("trace_w", n_legs, n_steps + 1, n))
    if want_moments and n_legs == 2:
        a.moments = _ptr(buf("moments", 8))
        nbytes = lib.lsi_protocol_workspace_bytes(n)
        ws = out.get("_workspace")
        if ws is None or ws.numel() * 8 < nbytes or ws.device != dev:
            ws = t.empty((nbytes + 7) // 8, dtype=t.float64, device=dev)
            out["_workspace"] = ws
        a.workspace, a.workspace_bytes = _ptr(ws), ws.numel() * 8
    out["_keepalive"] = (x_cache, x0, v0)
    with t.cuda.device(dev):
        check(lib.lsi_run_protocol(system.handle, program.handle, ctypes.byref(a), current_stream_ptr()))
    return out


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

"""Drop-in for the toy half of benchmark/testsystems/low_dimensional_systems.py:52-184.

`quartic` and `double_well` are NumbaBookkeepingSimulator instances with the reference's
parameters (m = 10, beta = 1); `NumbaNonequilibriumSimulator.collect_protocol_samples`
keeps its signature and return convention but runs every protocol of the ensemble in ONE
fused GPU launch instead of a Python loop around numba closures.
"""
import ctypes
import functools

import numpy as np

from .. import _cabi, _rngstate, _staging, engine
from ..integrators.numba_integrators import ToyIntegrator, identify_potential


class NumbaBookkeepingSimulator:
    """Equilibrium side of a 1-D toy system (reference :52-129)."""

    def __init__(self, mass=10.0, beta=1.0, potential=lambda x: x ** 4, force=lambda x: -4.0 * x ** 3,
                 name="quartic"):
        self.mass = mass
        self.beta = beta
        self.velocity_scale = np.sqrt(1.0 / (beta * mass))
        self.potential = potential
        self.force = force
        self.reduced_potential = lambda x: potential(x) * beta
        self.log_q = lambda x: -potential(x) * beta
        self.q = lambda x: np.exp(-potential(x) * beta)
        self.name = name
        self.cached = False
        self.x_samples = None
        self._x_samples_dev = None
        self._kind, self._params = identify_potential(potential, force)

    # --- equilibrium cache: independent Metropolis chains on the GPU (lsi_sample_equilibrium_scalar)
    def collect_equilibrium_samples(self, n_samples=1000000, n_burn=500, seed=None):
        engine.require_cuda()
        t = engine.torch()
        system = engine.ScalarSystem(self._kind, self.mass, self._params)
        out = t.empty(int(n_samples), dtype=t.float64, device="cuda")
        seed = _rngstate.get_seed() if seed is None else seed
        _cabi.check(_cabi.lib.lsi_sample_equilibrium_scalar(system.handle, 1.0 / self.beta, int(n_samples), int(n_burn),
                                                            int(seed), 0, ctypes.c_void_p(out.data_ptr()),
                                                            engine.current_stream_ptr()))
        self._x_samples_dev = out
        return out.cpu().numpy()

    def load_or_simulate_x_samples(self):
        self.x_samples = self.collect_equilibrium_samples()
        self.cached = True

    def set_x_samples(self, x_samples):
        """Install a host-side equilibrium cache (what the reference keeps in `x_samples`).  The copy lives in pinned
        memory so that every upload to the device is an asynchronous DMA from page-locked pages."""
        t = engine.torch()
        host = t.as_tensor(np.ascontiguousarray(x_samples, dtype=np.float64)).clone()
        if t.cuda.is_available():
            host = host.pin_memory()
        self._x_samples_host = host
        self.x_samples = host.numpy()
        self._x_samples_dev = None
        self.cached = True

    def release_device_cache(self):
        """Forget the device copy of the cache: the next protocol call uploads `x_samples` again (host -> device)."""
        self._x_samples_dev = None

    def x_samples_device(self):
        if not self.cached:
            self.load_or_simulate_x_samples()
        if self._x_samples_dev is None:
            host = getattr(self, "_x_samples_host", None)
            if host is None or host.numpy().ctypes.data != np.asarray(self.x_samples).ctypes.data:
                host = engine.torch().as_tensor(np.ascontiguousarray(self.x_samples, dtype=np.float64))
            self._x_samples_dev = host.to("cuda", non_blocking=True)
        return self._x_samples_dev

    def sample_x_from_equilibrium(self):
        if not self.cached:
            self.load_or_simulate_x_samples()
        return self.x_samples[np.random.randint(len(self.x_samples))]

    def sample_v_given_x(self, x):
        return np.random.randn() * self.velocity_scale


def _dw_potential(x):
    return x ** 6 + 2 * np.cos(5 * (x + 1))


def _dw_force(x):
    return -(6 * x ** 5 - 10 * np.sin(5 * (x + 1)))


quartic = NumbaBookkeepingSimulator()
double_well = NumbaBookkeepingSimulator(potential=_dw_potential, force=_dw_force, name="double_well")


class NumbaNonequilibriumSimulator:
    """Reference :139-184.  `integrator` is `partial(scheme, gamma=..., dt=...)` with `scheme`
    one of this package's `*_factory` closures."""

    def __init__(self, equilibrium_simulator, integrator):
        self.equilibrium_simulator, self.integrator = equilibrium_simulator, integrator
        self._buffers = {}

    def sample_x_from_equilibrium(self):
        return self.equilibrium_simulator.sample_x_from_equilibrium()

    def sample_v_given_x(self, x):
        return self.equilibrium_simulator.sample_v_given_x(x)

    def accumulate_shadow_work(self, x_0, v_0, n_steps):
        return self.integrator(x_0, v_0, n_steps)[-1]

    def _unwrap(self):
        integ = self.integrator
        kw = {}
        while isinstance(integ, functools.partial):
            if integ.args:
                raise NotImplementedError("bind gamma and dt by keyword: partial(scheme, gamma=..., dt=...)")
            kw = {**integ.keywords, **kw}
            integ = integ.func
        if not isinstance(integ, ToyIntegrator) or "gamma" not in kw or "dt" not in kw:
            raise NotImplementedError("integrator must be partial(<*_factory closure of this package>, gamma=, dt=); "
                                      "arbitrary Python callables cannot run on the GPU and there is no CPU fallback")
        return integ, kw["gamma"], kw["dt"]

    def collect_protocol_samples(self, n_protocol_samples, protocol_length, marginal="configuration",
                                 traces=True, as_numpy=True, seed=None):
        """Returns (W_shads_F, W_shads_R, xv_F, xv_R) like the reference: cumulative work traces of
        shape (n, protocol_length) and (x, v) traces of shape (n, protocol_length, 2).
        With traces=False only the final works are produced: W arrays have shape (n, 1) (so the
        caller's `W[:, -1]` still works) and xv_* are None -- use this for large ensembles.
        as_numpy=False leaves the results on the GPU as torch tensors; those are VIEWS of this simulator's output
        buffers and are overwritten by its next call (clone them to keep them).  numpy results own their memory."""
        if marginal not in ("configuration", "full"):
            raise NotImplementedError("`marginal` must be either 'configuration' or 'full'")
        integ, gamma, dt = self._unwrap()
        res = integ.ensemble(n_protocol_samples, protocol_length, gamma, dt, marginal=marginal,
                             x_cache=self.equilibrium_simulator.x_samples_device(), seed=seed,
                             traces="pairs" if traces else False, out=self._buffers)
        self.last_moments = res.get("moments")
        if traces:
            # the kernel stores step-major (coalesced); the reference's shapes (n, protocol_length) and
            # (n, protocol_length, 2) are permuted views of those buffers, no extra pass over the traces
            kT = integ.kT
            W_F, W_R = res["trace_w"][0].transpose(0, 1), res["trace_w"][1].transpose(0, 1)
            if kT != 1.0:
                W_F, W_R = W_F * kT, W_R * kT
            xv_F, xv_R = res["trace_xv"][0].permute(1, 0, 2), res["trace_xv"][1].permute(1, 0, 2)
        else:
            W_F, W_R = res["W"][0].unsqueeze(1), res["W"][1].unsqueeze(1)
            if integ.kT != 1.0:
                W_F, W_R = W_F * integ.kT, W_R * integ.kT
            xv_F = xv_R = None
        if as_numpy:
            # one asynchronous copy per array into pinned host memory, one synchronisation (_staging.to_host)
            return tuple(_staging.to_host([W_F, W_R, xv_F, xv_R]))
        return W_F, W_R, xv_F, xv_R

"""Drop-in for the reference's toy integrator factories
(benchmark/integrators/numba_integrators.py:13-252), running on the GPU.

    f = baoab_factory(potential, force, velocity_scale, m)
    xs, vs, Q, W_shads = f(x0, v0, n_steps, gamma, dt)

keeps the reference's contract: `n_steps - 1` integrator steps, arrays of length `n_steps`
whose entry 0 is the start state, `W_shads[i] = (E_i - E_0) - Q_i` cumulative, `Q` the total
heat (numba_integrators.py:141,168-169).  Each closure is the splitting string that performs
the same substeps:

    vvvr  (OVRVO): "O V R R V O"   two half drifts          (:31-61)
    orvro        : "O R V R O"     one full kick            (:85-112)
    baoab (VRORV): "V R O O R V"   two half O draws         (:141-169)
    aboba (RVOVR): "R V O O V R"   two half O draws         (:195-223)

`potential` / `force` must be one of the systems the engine has closed-form kernels for
(quartic, double well, harmonic); they are recognised by probing the callables.  Anything
else raises NotImplementedError -- there is no CPU fallback.

Besides the reference's one-trajectory call, every closure has `.ensemble(...)`, which is
what NumbaNonequilibriumSimulator.collect_protocol_samples launches.
"""
import numpy as np

from .. import _rngstate, engine

SCHEMES = {"vvvr": "O V R R V O", "orvro": "O R V R O", "baoab": "V R O O R V", "aboba": "R V O O V R"}

_PROBE = np.array([-1.7, -0.9, -0.31, 0.0, 0.23, 0.77, 1.4])


def _closed_forms():
    return {
        "quartic": (lambda x: x ** 4, lambda x: -4.0 * x ** 3),
        "double_well": (lambda x: x ** 6 + 2 * np.cos(5 * (x + 1)), lambda x: -(6 * x ** 5 - 10 * np.sin(5 * (x + 1)))),
    }


def identify_potential(potential, force):
    """Map (potential, force) callables onto an engine system kind (+ parameters)."""
    tagged = getattr(potential, "lsi_kind", None)
    if tagged is not None:
        return tagged, getattr(potential, "lsi_params", (0.0, 0.0, 0.0, 0.0))
    try:
        u = np.array([float(potential(x)) for x in _PROBE])
        f = np.array([float(force(x)) for x in _PROBE])
    except Exception as e:  # noqa: BLE001
        raise NotImplementedError("cannot probe the potential/force callables: %s" % e)
    for kind, (pu, pf) in _closed_forms().items():
        if np.allclose(u, pu(_PROBE), rtol=1e-12, atol=1e-12) and np.allclose(f, pf(_PROBE), rtol=1e-12, atol=1e-12):
            return kind, (0.0, 0.0, 0.0, 0.0)
    k = -f[0] / _PROBE[0]
    if np.allclose(u, 0.5 * k * _PROBE ** 2, rtol=1e-12, atol=1e-12) and np.allclose(f, -k * _PROBE, rtol=1e-12, atol=1e-12):
        return "harmonic", (k, 0.0, 0.0, 0.0)
    raise NotImplementedError("the B200 engine has closed-form kernels for the reference's toy systems only "
                              "(quartic x^4, double well x^6 + 2cos(5(x+1)), harmonic); no CPU fallback exists")


class ToyIntegrator:
    """Callable with the signature of the reference's jitted closures."""

    def __init__(self, scheme, potential, force, velocity_scale, m):
        self.scheme = scheme
        self.splitting = SCHEMES[scheme]
        self.kind, self.params = identify_potential(potential, force)
        self.mass = float(m)
        self.velocity_scale = float(velocity_scale)
        self.kT = self.mass * self.velocity_scale ** 2  # velocity_scale = sqrt(1 / (beta m))
        self.system = engine.ScalarSystem(self.kind, self.mass, self.params)
        self._programs = {}

    def program(self, gamma, dt):
        key = (float(gamma), float(dt))
        if key not in self._programs:
            self._programs[key] = engine.Program(self.splitting, dt, gamma)
        return self._programs[key]

    def ensemble(self, n_replicas, n_steps, gamma, dt, marginal="configuration", x_cache=None, x0=None, v0=None,
                 n_legs=2, seed=None, replica0=None, traces=False, want_final=False, want_heat=False, out=None):
        """Whole protocol for an ensemble; `n_steps` counts recorded states as in the reference,
        i.e. `n_steps - 1` integrator steps per leg."""
        if seed is None:
            seed = _rngstate.get_seed()
        if replica0 is None:
            replica0 = _rngstate.reserve_replicas(n_replicas)
        return engine.run_protocol(self.system, self.program(gamma, dt), n_replicas, max(int(n_steps) - 1, 0),
                                   self.kT, marginal=marginal, n_legs=n_legs, seed=seed, replica0=replica0,
                                   x_cache=x_cache, x0=x0, v0=v0, traces=traces, want_final=want_final,
                                   want_heat=want_heat, out=out)

    def __call__(self, x0, v0, n_steps, gamma, dt):
        res = self.ensemble(1, n_steps, gamma, dt, x0=np.array([float(x0)]), v0=np.array([float(v0)]), n_legs=1,
                            traces=True, want_heat=True)
        xs = res["trace_x"][0, :, 0].cpu().numpy()
        vs = res["trace_v"][0, :, 0].cpu().numpy()
        W = res["trace_w"][0, :, 0].cpu().numpy() * self.kT  # reference W_shads are in energy units
        Q = float(res["heat"][0, 0].item())
        return xs, vs, Q, W


def vvvr_factory(potential, force, velocity_scale, m):
    return ToyIntegrator("vvvr", potential, force, velocity_scale, m)


def orvro_factory(potential, force, velocity_scale, m):
    return ToyIntegrator("orvro", potential, force, velocity_scale, m)


def baoab_factory(potential, force, velocity_scale, m):
    return ToyIntegrator("baoab", potential, force, velocity_scale, m)


def aboba_factory(potential, force, velocity_scale, m):
    return ToyIntegrator("aboba", potential, force, velocity_scale, m)


def splitting_factory(splitting, potential, force, velocity_scale, m):
    """Same closure contract for ANY splitting string over {O, R, V} (the reference hard-codes four)."""
    integ = ToyIntegrator("baoab", potential, force, velocity_scale, m)
    integ.scheme, integ.splitting = splitting, splitting
    return integ

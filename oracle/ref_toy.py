"""The reference's own toy path as the CPU arm of bench.py (TEST INFRASTRUCTURE, like everything under oracle/).

oracle/_ref/numba_integrators.py is the reference's file, copied verbatim by oracle/make_ref.sh
(/root/reference/benchmark/integrators/numba_integrators.py:13-227: vvvr / orvro / baoab / aboba closures).  The driver
around it restates NumbaNonequilibriumSimulator.collect_protocol_samples
(/root/reference/benchmark/testsystems/low_dimensional_systems.py:160-184; that module imports simtk and cannot be loaded):
per protocol sample, x0 from the cache, v0 ~ N(0, velocity_scale), one call of the jitted closure for the forward leg, the
midpoint velocity redrawn ("configuration") or kept ("full"), one call for the reverse leg; the same arrays are built and
returned.  `collect_parallel` fans the samples out over processes the way the reference's evaluation scripts do
(multiprocessing.Pool, compare_near_eq_and_exact.py:366-368).
"""
import importlib.util
import os
import time

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
REF_FILE = os.path.join(_HERE, "_ref", "numba_integrators.py")
FACTORIES = {"OVRVO": "vvvr_factory", "ORVRO": "orvro_factory", "VRORV": "baoab_factory", "RVOVR": "aboba_factory"}
_module = None
_closures = {}


def available():
    """(ok, why): the copied reference file exists and numba imports"""
    if not os.path.exists(REF_FILE):
        return False, "oracle/_ref/numba_integrators.py is missing (run oracle/make_ref.sh where /root/reference exists)"
    try:
        import numba  # noqa: F401
    except Exception as e:  # noqa: BLE001
        return False, "numba cannot be imported: %s" % e
    return True, ""


def module():
    global _module
    if _module is None:
        spec = importlib.util.spec_from_file_location("_ref_numba_integrators", REF_FILE)
        _module = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(_module)
    return _module


def system(name):
    """potential, force, velocity_scale, mass of low_dimensional_systems.py:52-76,132-136 (m = 10, beta = 1), as
    numba-jitted scalar functions (the reference passes numba-compilable lambdas to the factories)."""
    from numba import jit
    if name == "quartic":
        potential = jit(nopython=True)(lambda x: x ** 4)
        force = jit(nopython=True)(lambda x: -4.0 * x ** 3)
    elif name == "double_well":
        potential = jit(nopython=True)(lambda x: x ** 6 + 2 * np.cos(5 * (x + 1)))
        force = jit(nopython=True)(lambda x: - (6 * x ** 5 - 10 * np.sin(5 * (x + 1))))
    else:
        raise KeyError(name)
    mass, beta = 10.0, 1.0
    return potential, force, np.sqrt(1.0 / (beta * mass)), mass


def closure(system_name, scheme):
    """the reference's jitted integrator for one scheme: f(x0, v0, n_steps, gamma, dt) -> xs, vs, Q, W_shads"""
    key = (system_name, scheme)
    if key not in _closures:
        potential, force, velocity_scale, mass = system(system_name)
        _closures[key] = (getattr(module(), FACTORIES[scheme])(potential, force, velocity_scale, mass), velocity_scale)
    return _closures[key]


def collect_protocol_samples(system_name, scheme, x_samples, n_protocol_samples, protocol_length, marginal, gamma, dt, seed=None):
    """low_dimensional_systems.py:160-184, restated (tqdm left out)"""
    integrator, velocity_scale = closure(system_name, scheme)
    if seed is not None:
        np.random.seed(seed)
    W_shads_F, W_shads_R = [], []
    xv_F, xv_R = [], []
    for _ in range(n_protocol_samples):
        x_0 = x_samples[np.random.randint(len(x_samples))]
        v_0 = np.random.randn() * velocity_scale
        xs, vs, Q, W_shads = integrator(x_0, v_0, protocol_length, gamma, dt)
        W_shads_F.append(W_shads)
        xv_F.append(np.vstack([xs, vs]).T)
        x_1 = xs[-1]
        if marginal == "configuration":
            v_1 = np.random.randn() * velocity_scale
        elif marginal == "full":
            v_1 = vs[-1]
        else:
            raise NotImplementedError("`marginal` must be either 'configuration' or 'full'")
        xs, vs, Q, W_shads = integrator(x_1, v_1, protocol_length, gamma, dt)
        W_shads_R.append(W_shads)
        xv_R.append(np.vstack([xs, vs]).T)
    return np.array(W_shads_F), np.array(W_shads_R), np.array(xv_F), np.array(xv_R)


def _worker(job):
    system_name, scheme, x_samples, n, protocol_length, marginal, gamma, dt, seed = job
    W_F, W_R, _, _ = collect_protocol_samples(system_name, scheme, x_samples, n, protocol_length, marginal, gamma, dt, seed)
    return W_F[:, -1], W_R[:, -1]


_pool = None


def pool(processes):
    """worker processes, forked AFTER the closures were compiled in the parent (they inherit the machine code)"""
    global _pool
    if _pool is None or _pool._processes != processes:
        import multiprocessing as mp
        if _pool is not None:
            _pool.terminate()
        _pool = mp.get_context("fork").Pool(processes)
        import atexit
        atexit.register(close_pool)
    return _pool


def close_pool():
    global _pool
    if _pool is not None:
        _pool.terminate()
        _pool.join()
        _pool = None


def collect_parallel(system_name, schemes, x_samples, n_per_scheme, protocol_length, marginal, gamma, dt, processes=None, seed=0):
    """All `schemes`, n_per_scheme protocol samples each, spread over `processes` workers.  Returns
    ({scheme: (W_F, W_R)}, seconds): final works only cross the process boundary; seconds is the wall time of the
    collection itself (compilation and pool start-up happen before the clock starts)."""
    processes = int(processes or os.cpu_count() or 1)
    for s in schemes:
        f, scale = closure(system_name, s)
        f(0.1, 0.1, 3, gamma, dt)  # compile before forking
    p = pool(processes) if processes > 1 else None
    jobs = []
    for k, s in enumerate(schemes):
        base, rem = divmod(n_per_scheme, processes)
        for w in range(processes):
            n = base + (1 if w < rem else 0)
            if n:
                jobs.append((s, (system_name, s, x_samples, n, protocol_length, marginal, gamma, dt, seed + 1000 * k + w)))
    t0 = time.perf_counter()
    parts = p.map(_worker, [j for _, j in jobs], chunksize=1) if p is not None else [_worker(j) for _, j in jobs]
    secs = time.perf_counter() - t0
    out = {}
    for (s, _), (wf, wr) in zip(jobs, parts):
        a, b = out.get(s, ([], []))
        a.append(wf)
        b.append(wr)
        out[s] = (a, b)
    return {s: (np.concatenate(a), np.concatenate(b)) for s, (a, b) in out.items()}, secs

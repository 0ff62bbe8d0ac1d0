"""GPU parity of the toy path (configs C1/C2) against the CPU oracle and the golden fixtures.
Everything here goes through the C ABI (integrator_benchmark_b200 -> ctypes -> liblsi.so)."""
import json
import os
from functools import partial

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

SIX = ["O V R V O", "O R V R O", "R V O V R", "R O V O R", "V R O R V", "V O R O V"]
# fp64 on both sides, same Philox words; libm vs CUDA log/sincos differ in the last ulp and the
# double well amplifies that over a trajectory (observed worst case 8e-9 relative) -- still
# three orders inside the 1e-5 parity bar of BASELINE.json
RTOL, ATOL = 1e-7, 1e-9


@pytest.fixture(scope="module")
def ib():
    import integrator_benchmark_b200 as ib
    ib.engine.require_cuda()
    return ib


@pytest.fixture(scope="module")
def orc():
    from oracle import oracle
    return oracle


def _cache(orc, kind, n=4096, seed=11):
    return orc.sample_equilibrium_scalar(kind, n, 60, seed=seed)


@pytest.mark.parametrize("kind", ["quartic", "double_well"])
@pytest.mark.parametrize("marginal", ["configuration", "full"])
def test_protocol_matches_oracle(ib, orc, kind, marginal):
    cache = _cache(orc, kind)
    system = ib.engine.ScalarSystem(kind, 10.0)
    for j, splitting in enumerate(SIX + ["V R R O R R V", "O V R R V O", "V R O O R V"]):
        gamma, dt, n_steps, n = [100.0, 10.0, 0.5][j % 3], [0.1, 0.35, 0.6][j % 3], 10 + j, 3001
        prog = ib.engine.Program(splitting, dt, gamma, measure_shadow_work=True)
        res = ib.engine.run_protocol(system, prog, n, n_steps, 1.0, marginal=marginal, seed=77 + j, replica0=12345,
                                     x_cache=cache, want_heat=True, want_direct=True, want_final=True, traces=True)
        ref = orc.toy_protocol(kind, splitting, dt, gamma, n, n_steps, marginal=marginal, seed=77 + j,
                               replica0=12345, x_cache=cache, traces=True, want_direct=True)
        # replicas that blow up at the large time steps (|x| -> 1e100 and beyond) are chaotic: those are compared
        # through the crash count only
        tame = np.abs(ref["trace_x"]).max(axis=(1, 2)) < 1e3
        assert tame.mean() > 0.9
        R = lambda a: a[tame]  # noqa: E731
        W = res["W"].cpu().numpy().T
        np.testing.assert_allclose(R(W), R(ref["W"]), rtol=RTOL, atol=ATOL)
        np.testing.assert_allclose(R(res["heat"].cpu().numpy().T), R(ref["Q"]), rtol=RTOL, atol=ATOL)
        np.testing.assert_allclose(R(res["W_direct"].cpu().numpy().T), R(ref["Wdirect"]), rtol=1e-7, atol=1e-9)
        np.testing.assert_allclose(R(res["x_final"].cpu().numpy()), R(ref["x_final"]), rtol=RTOL, atol=ATOL)
        np.testing.assert_allclose(R(res["v_final"].cpu().numpy()), R(ref["v_final"]), rtol=RTOL, atol=ATOL)
        np.testing.assert_allclose(R(res["trace_x"].cpu().numpy().transpose(2, 0, 1)), R(ref["trace_x"]), rtol=RTOL, atol=ATOL)
        np.testing.assert_allclose(R(res["trace_w"].cpu().numpy().transpose(2, 0, 1)), R(ref["trace_w"]), rtol=RTOL, atol=ATOL)
        # the two bookkeeping routes agree (reference tests/test_shadow_work.py:30-42)
        np.testing.assert_allclose(R(res["W_direct"].cpu().numpy().T), R(W), rtol=1e-7, atol=1e-9)
        # fused moments == moments of the returned works; crashed replicas are counted, not summed, and the
        # same replicas crash in the oracle
        m = res["moments"].cpu().numpy()
        ok = np.isfinite(W).all(axis=1)
        assert np.array_equal(ok, np.isfinite(ref["W"]).all(axis=1))
        wf, wr = W[ok, 0], W[ok, 1]
        np.testing.assert_allclose(m[:6], [ok.sum(), wf.sum(), wr.sum(), (wf * wf).sum(), (wr * wr).sum(), (wf * wr).sum()],
                                   rtol=1e-6, atol=1e-9)
        assert m[6] == n - ok.sum()
        # the fused (compile-time splitting) kernel and the interpreter agree with the oracle alike:
        # without traces / direct work the whitelisted strings take the fused path
        res2 = ib.engine.run_protocol(system, ib.engine.Program(splitting, dt, gamma), n, n_steps, 1.0, marginal=marginal,
                                      seed=77 + j, replica0=12345, x_cache=cache, want_heat=True, want_final=True)
        np.testing.assert_allclose(R(res2["W"].cpu().numpy().T), R(ref["W"]), rtol=RTOL, atol=ATOL)
        np.testing.assert_allclose(R(res2["heat"].cpu().numpy().T), R(ref["Q"]), rtol=RTOL, atol=ATOL)
        np.testing.assert_allclose(R(res2["x_final"].cpu().numpy()), R(ref["x_final"]), rtol=RTOL, atol=ATOL)
        m2 = res2["moments"].cpu().numpy()
        assert m2[0] == m[0] and m2[6] == m[6]
        W2 = res2["W"].cpu().numpy().T
        ok2 = np.isfinite(W2).all(axis=1)
        np.testing.assert_allclose(m2[1:3], [W2[ok2, 0].sum(), W2[ok2, 1].sum()], rtol=1e-6, atol=1e-9)


@pytest.mark.parametrize("kind", ["quartic", "double_well"])
def test_mixed_precision_noise_matches_oracle(ib, orc, kind):
    """precision="mixed": fp32 Box-Muller noise (as OpenMM's CUDA platform draws `gaussian`), fp64 state.
    Tolerance 1e-5 relative as BASELINE.json states; the works are differences of O(1) energies, so
    they are held to 1e-5 of the energy scale."""
    cache = _cache(orc, kind)
    system = ib.engine.ScalarSystem(kind, 10.0)
    for j, splitting in enumerate(SIX + ["V R R O R R V", "V0 R O R V0".replace("0", "")]):
        gamma, dt, n_steps, n = [100.0, 10.0, 1.0][j % 3], [0.1, 0.3, 0.2][j % 3], 7 + j, 2049
        for traces in (False, True):  # fused kernel / interpreter
            res = ib.engine.run_protocol(system, ib.engine.Program(splitting, dt, gamma), n, n_steps, 1.0, seed=j,
                                         x_cache=cache, precision="mixed", want_final=True, traces=traces)
            ref = orc.toy_protocol(kind, splitting, dt, gamma, n, n_steps, seed=j, x_cache=cache, noise_fp32=True)
            np.testing.assert_allclose(res["x_final"].cpu().numpy(), ref["x_final"], rtol=1e-5, atol=1e-5)
            np.testing.assert_allclose(res["v_final"].cpu().numpy(), ref["v_final"], rtol=1e-5, atol=1e-5)
            np.testing.assert_allclose(res["W"].cpu().numpy().T, ref["W"], rtol=1e-5, atol=1e-5)


def test_replica_ids_key_the_streams(ib, orc):
    """bit-exact replica indexing: a shard [a, b) of a run equals the same slice of the full run."""
    cache = _cache(orc, "double_well")
    system = ib.engine.ScalarSystem("double_well", 10.0)
    prog = ib.engine.Program("V R O R V", 0.4, 10.0)
    full = ib.engine.run_protocol(system, prog, 5000, 9, 1.0, seed=5, replica0=0, x_cache=cache)["W"].cpu().numpy()
    for lo, hi in [(0, 1250), (1250, 2500), (2500, 5000), (4999, 5000)]:
        part = ib.engine.run_protocol(system, prog, hi - lo, 9, 1.0, seed=5, replica0=lo, x_cache=cache)["W"].cpu().numpy()
        assert np.array_equal(part, full[:, lo:hi])
    # cache indices are the oracle's, bit for bit
    res = ib.engine.run_protocol(system, prog, 64, 0, 1.0, seed=5, replica0=1 << 33, x_cache=cache, want_final=True)
    idx = [orc.toy_cache_index(5, (1 << 33) + r, len(cache)) for r in range(64)]
    assert np.array_equal(res["x_final"].cpu().numpy(), cache[idx])


def test_edge_cases(ib, orc):
    system = ib.engine.ScalarSystem("quartic", 10.0)
    prog = ib.engine.Program("O V R V O", 0.25, 100.0)
    cache = _cache(orc, "quartic", 16)
    # zero steps: W = 0 exactly; single replica; n not a multiple of the block size
    for n, steps in [(1, 0), (1, 1), (255, 3), (257, 3)]:
        res = ib.engine.run_protocol(system, prog, n, steps, 1.0, seed=1, x_cache=cache)
        ref = orc.toy_protocol("quartic", "O V R V O", 0.25, 100.0, n, steps, seed=1, x_cache=cache)
        np.testing.assert_allclose(res["W"].cpu().numpy().T, ref["W"], rtol=RTOL, atol=ATOL)
    # empty ensemble is a no-op
    ib.engine.run_protocol(system, prog, 0, 5, 1.0, seed=1, x_cache=cache)
    # unstable step: non-finite works are counted, not summed
    bad = ib.engine.Program("O V R V O", 5.0, 1.0)
    res = ib.engine.run_protocol(system, bad, 512, 200, 1.0, seed=1, x0=np.full(512, 3.0), v0=np.full(512, 1.0))
    m = res["moments"].cpu().numpy()
    assert m[6] == 512 and m[0] == 0
    # a program longer than the in-parameter table takes the staged path
    long_s = " ".join(["V"] + ["R"] * 100 + ["O"] + ["R"] * 100 + ["V"])
    res = ib.engine.run_protocol(system, ib.engine.Program(long_s, 0.3, 5.0), 100, 4, 1.0, seed=3, x_cache=cache)
    ref = orc.toy_protocol("quartic", long_s, 0.3, 5.0, 100, 4, seed=3, x_cache=cache)
    np.testing.assert_allclose(res["W"].cpu().numpy().T, ref["W"], rtol=RTOL, atol=ATOL)
    with pytest.raises(NotImplementedError):
        ib.engine.run_protocol(system, prog, 4, 2, 1.0, marginal="momentum", x_cache=cache)
    with pytest.raises(NotImplementedError):
        ib.engine.run_protocol(system, ib.engine.Program("V0 V1 R O R V1 V0", 0.1, 1.0), 4, 2, 1.0, x_cache=cache)


def test_factories_reproduce_reference_closures(ib, golden_dir):
    """`f(x0, v0, n_steps, gamma, dt)` against the reference's own numba closures: the golden
    noise is the Philox stream of (seed, replica), so pinning both reproduces it on the GPU."""
    from integrator_benchmark_b200.integrators import numba_integrators as ni
    from integrator_benchmark_b200.testsystems import double_well, quartic
    z = np.load(os.path.join(golden_dir, "toy_numba.npz"))
    meta = json.loads(str(z["meta"]))
    systems = {"quartic": quartic, "double_well": double_well}
    factories = {"vvvr": ni.vvvr_factory, "orvro": ni.orvro_factory, "baoab": ni.baoab_factory, "aboba": ni.aboba_factory}
    for c in meta:
        eq = systems[c["system"]]
        f = factories[c["scheme"]](eq.potential, eq.force, eq.velocity_scale, eq.mass)
        ib.set_seed(c["seed"], next_replica=c["replica"])
        xs, vs, Q, W = f(c["x0"], c["v0"], c["n_steps"], c["gamma"], c["dt"])
        k = c["key"]
        assert xs.shape == (c["n_steps"],)
        np.testing.assert_allclose(xs, z[k + "_xs"], rtol=RTOL, atol=ATOL)
        np.testing.assert_allclose(vs, z[k + "_vs"], rtol=RTOL, atol=ATOL)
        np.testing.assert_allclose(W, z[k + "_W"], rtol=1e-8, atol=1e-10)
        np.testing.assert_allclose(Q, float(z[k + "_Q"]), rtol=1e-8, atol=1e-10)


def test_equilibrium_sampler_matches_oracle_and_target(ib, orc):
    from integrator_benchmark_b200.testsystems import double_well
    ib.set_seed(99)
    xs = double_well.collect_equilibrium_samples(n_samples=200000, n_burn=300)
    ref = orc.sample_equilibrium_scalar("double_well", 2000, 300, seed=99)
    np.testing.assert_allclose(xs[:2000], ref, rtol=1e-9, atol=1e-11)
    # histogram against exp(-U)/Z, Z = 4.4848163908 (notebooks/1D figures, updated.ipynb:415)
    edges = np.linspace(-1.6, 1.6, 33)
    hist = np.histogram(xs, bins=edges)[0] / len(xs)
    grid = np.linspace(-1.6, 1.6, 32 * 200 + 1)
    pdf = np.exp(-(grid ** 6 + 2 * np.cos(5 * (grid + 1)))) / 4.4848163908
    expect = np.array([np.trapezoid(pdf[i * 200:(i + 1) * 200 + 1], grid[i * 200:(i + 1) * 200 + 1]) for i in range(32)])
    assert np.max(np.abs(hist - expect)) < 4e-3


def test_collect_protocol_samples_api_and_estimate(ib, orc):
    """reference call pattern (experiments/toy/quartic_kl_validation.py:26-70) end to end."""
    from integrator_benchmark_b200.evaluation import estimate_nonequilibrium_free_energy
    from integrator_benchmark_b200.integrators import baoab_factory, vvvr_factory
    from integrator_benchmark_b200.testsystems import NumbaNonequilibriumSimulator, quartic
    ib.set_seed(2024)
    eq = quartic
    n, length = 20000, 25
    for factory, string in [(baoab_factory, "V R O O R V"), (vvvr_factory, "O V R R V O")]:
        scheme = factory(eq.potential, eq.force, eq.velocity_scale, eq.mass)
        sim = NumbaNonequilibriumSimulator(eq, partial(scheme, gamma=0.5, dt=0.9))
        ib.set_seed(2024, next_replica=1000)
        W_F, W_R, xv_F, xv_R = sim.collect_protocol_samples(n, length, "configuration")
        assert W_F.shape == (n, length) and xv_F.shape == (n, length, 2) and W_F[:, 0].max() == 0.0
        ref = orc.toy_protocol("quartic", string, 0.9, 0.5, n, length - 1, seed=2024, replica0=1000,
                               x_cache=eq.x_samples, traces=True)
        np.testing.assert_allclose(W_F, ref["trace_w"][:, 0], rtol=RTOL, atol=ATOL)
        np.testing.assert_allclose(W_R, ref["trace_w"][:, 1], rtol=RTOL, atol=ATOL)
        np.testing.assert_allclose(xv_R[:, :, 0], ref["trace_x"][:, 1], rtol=RTOL, atol=ATOL)
        dF, var = estimate_nonequilibrium_free_energy(W_F[:, -1], W_R[:, -1])
        dF_ref = 0.5 * (np.mean(ref["W"][:, 0]) - np.mean(ref["W"][:, 1]))
        var_ref = (np.var(ref["W"][:, 0]) + np.var(ref["W"][:, 1]) - 2 * np.cov(ref["W"][:, 0], ref["W"][:, 1])[0, 1]) / (4 * n)
        assert dF == pytest.approx(dF_ref, rel=1e-9, abs=1e-12)
        assert var == pytest.approx(var_ref, rel=1e-7)
        # traces=False gives the same final works
        ib.set_seed(2024, next_replica=1000)
        W_F2, W_R2, _, _ = sim.collect_protocol_samples(n, length, "configuration", traces=False)
        # (fused kernel vs interpreter: same arithmetic up to fma contraction of the cached force)
        np.testing.assert_allclose(W_F2[:, -1], W_F[:, -1], rtol=RTOL, atol=ATOL)
        np.testing.assert_allclose(W_R2[:, -1], W_R[:, -1], rtol=RTOL, atol=ATOL)
        dF2, var2 = ib.evaluation.estimate_from_moments(sim.last_moments.cpu().numpy())
        assert dF2 == pytest.approx(dF, rel=1e-9, abs=1e-12) and var2 == pytest.approx(var, rel=1e-7)


def test_estimators_match_reference_fixtures(ib, golden_dir):
    from integrator_benchmark_b200.evaluation import estimate_nonequilibrium_free_energy
    cases = json.load(open(os.path.join(golden_dir, "estimators.json")))
    n = 0
    for c in cases:
        if c["kind"] != "neq":
            continue
        dF, var = estimate_nonequilibrium_free_energy(np.array(c["W_F"]), np.array(c["W_R"]), c["N_eff"])
        assert dF == pytest.approx(c["DeltaF"], rel=1e-10, abs=1e-14)
        assert var == pytest.approx(c["sq_unc"], rel=1e-7, abs=1e-16)
        n += 1
    assert n >= 8


def test_trace_layouts_and_fused_trace_kernel(ib, orc):
    """Traces from the fused kernel (both layouts) and from the interpreter agree with each other and with the oracle;
    the replica-major layout is the reference's (n, n_steps) / (n, n_steps, 2)."""
    system = ib.engine.ScalarSystem("double_well", 10.0)
    cache = orc.sample_equilibrium_scalar("double_well", 257, 40, seed=5)
    n, steps = 1000, 7
    for splitting in ["O V R V O", "V R O R V", "R V O V R", "V R R R O R R R V", "V O V R O"]:  # the last one is not pre-instantiated
        prog = ib.engine.Program(splitting, 0.3, 10.0)
        a = ib.engine.run_protocol(system, prog, n, steps, 1.0, marginal="configuration", seed=9, replica0=17, x_cache=cache,
                                   traces=True, out={})
        b = ib.engine.run_protocol(system, prog, n, steps, 1.0, marginal="configuration", seed=9, replica0=17, x_cache=cache,
                                   traces="replica_major", out={})
        c = ib.engine.run_protocol(system, prog, n, steps, 1.0, marginal="configuration", seed=9, replica0=17, x_cache=cache,
                                   out={})
        d = ib.engine.run_protocol(system, prog, n, steps, 1.0, marginal="configuration", seed=9, replica0=17, x_cache=cache,
                                   traces="pairs", out={})
        np.testing.assert_array_equal(d["trace_w"].cpu().numpy(), a["trace_w"].cpu().numpy())
        np.testing.assert_array_equal(d["trace_xv"][..., 0].cpu().numpy(), a["trace_x"].cpu().numpy())
        np.testing.assert_array_equal(d["trace_xv"][..., 1].cpu().numpy(), a["trace_v"].cpu().numpy())
        assert b["trace_w_rm"].shape == (2, n, steps + 1) and b["trace_xv_rm"].shape == (2, n, steps + 1, 2)
        np.testing.assert_array_equal(b["trace_w_rm"].cpu().numpy(), a["trace_w"].permute(0, 2, 1).cpu().numpy())
        np.testing.assert_array_equal(b["trace_xv_rm"][..., 0].cpu().numpy(), a["trace_x"].permute(0, 2, 1).cpu().numpy())
        np.testing.assert_array_equal(b["trace_xv_rm"][..., 1].cpu().numpy(), a["trace_v"].permute(0, 2, 1).cpu().numpy())
        # the last trace entry is the work of the leg, up to the rounding of one more energy difference
        np.testing.assert_allclose(a["trace_w"][:, -1].cpu().numpy(), c["W"].cpu().numpy(), rtol=1e-12, atol=1e-13)
        np.testing.assert_allclose(a["W"].cpu().numpy(), c["W"].cpu().numpy(), rtol=1e-13, atol=1e-14)
        ref = orc.toy_protocol("double_well", splitting, 0.3, 10.0, n, steps, marginal="configuration", seed=9, replica0=17,
                               x_cache=cache, traces=True)
        tame = np.isfinite(ref["W"]).all(axis=1) & (np.abs(ref["W"]).max(axis=1) < 50)
        np.testing.assert_allclose(b["trace_xv_rm"][0, :, :, 0].cpu().numpy()[tame], ref["trace_x"][tame, 0], rtol=1e-7, atol=1e-9)
        np.testing.assert_allclose(b["trace_w_rm"][1].cpu().numpy()[tame], ref["trace_w"][tame, 1], rtol=1e-7, atol=1e-8)


def test_full_size_properties(ib):
    """BASELINE config 2 size (1 048 576 replicas x 10): size-independent properties.
    (a) time-reversal/Crooks sanity: for the harmonic well at tiny dt both works vanish;
    (b) the heat + shadow work identity Delta E = W + Q holds per replica (test_shadow_work.py);
    (c) <exp(-W_F)> = 1 (Jarzynski equality for shadow work started in equilibrium)."""
    from integrator_benchmark_b200.testsystems import double_well
    ib.set_seed(7)
    n = 1 << 20
    system = ib.engine.ScalarSystem("double_well", 10.0)
    cache = double_well.x_samples_device()
    prog = ib.engine.Program("V R O R V", 0.3, 100.0, measure_shadow_work=True)
    res = ib.engine.run_protocol(system, prog, n, 9, 1.0, marginal="full", seed=7, x_cache=cache, want_direct=True)
    W = res["W"]
    assert bool(ib.engine.torch().isfinite(W).all())
    assert float((res["W_direct"] - W).abs().max()) < 1e-9
    jarz = float(ib.engine.torch().exp(-W[0]).mean())
    assert abs(jarz - 1.0) < 5e-3
    m = res["moments"].cpu().numpy()
    assert m[0] == n and m[6] == 0
    assert m[1] == pytest.approx(float(W[0].sum()), rel=1e-10)


def test_fastmath_accuracy(ib):
    """csrc/fastmath64.cuh (lean log / sincospi / sincos / sqrt / reciprocal, exact uniforms) against numpy / mpmath:
    about one ulp, on the ranges the kernels use."""
    import ctypes
    import torch
    from integrator_benchmark_b200 import _cabi, engine

    def run(which, x, two=False):
        xd = torch.as_tensor(np.ascontiguousarray(x, dtype=np.float64)).cuda()
        a, b = torch.empty_like(xd), torch.empty_like(xd)
        _cabi.check(_cabi.lib.lsi_test_fastmath(which, xd.numel(), ctypes.c_void_p(xd.data_ptr()), ctypes.c_void_p(a.data_ptr()),
                                                ctypes.c_void_p(b.data_ptr()) if two else None, engine.current_stream_ptr()))
        return (a.cpu().numpy(), b.cpu().numpy()) if two else a.cpu().numpy()

    rng = np.random.RandomState(0)
    words = rng.randint(0, 2 ** 32, size=400000).astype(np.float64)
    words[:4] = [0, 1, 2 ** 32 - 1, 2 ** 31]
    u1, u2 = run(5, words, two=True)
    assert np.array_equal(u1, (words + 0.5) * 2.0 ** -32) and np.array_equal(u2, (words + 0.5) * 2.0 ** -31)  # exact
    ulps = lambda got, ref: np.abs(got - ref) / np.spacing(np.abs(ref))  # noqa: E731
    assert ulps(run(0, u1), np.log(u1)).max() <= 2.0
    pos = np.concatenate([rng.uniform(1e-10, 50.0, 200000), -2.0 * np.log(u1[:200000])])
    assert ulps(run(3, pos), np.sqrt(pos)).max() <= 1.0
    assert ulps(run(4, pos), 1.0 / pos).max() <= 1.0
    s, c = run(1, u2, two=True)
    import mpmath
    mpmath.mp.prec = 160
    idx = rng.randint(0, len(u2), 3000)
    ref_s = np.array([float(mpmath.sin(mpmath.pi * mpmath.mpf(float(u2[i])))) for i in idx])
    ref_c = np.array([float(mpmath.cos(mpmath.pi * mpmath.mpf(float(u2[i])))) for i in idx])
    assert np.abs(s[idx] - ref_s).max() < 3e-16 and np.abs(c[idx] - ref_c).max() < 3e-16
    th = rng.uniform(-40.0, 40.0, 300000)  # 5 (x + 1) for x in [-9, 7]
    s, c = run(2, th, two=True)
    assert np.abs(s - np.sin(th)).max() < 3e-16 and np.abs(c - np.cos(th)).max() < 3e-16
    d0, d1 = run(6, words, two=True)  # lean minus library normals
    assert np.abs(d0).max() < 2e-14 and np.abs(d1).max() < 2e-14
    # table-driven versions (what the toy kernels use): -2 log(u) to 2 ulp, sincos to 3e-16 absolute
    got = run(7, u1)
    sample = np.concatenate([idx, np.argsort(u1)[:50], np.argsort(u1)[-50:], np.argsort(np.abs(u1 - 0.685))[:50],
                             np.argsort(np.abs(u1 - 0.5))[:50]])
    ref_l = np.array([float(-2 * mpmath.log(mpmath.mpf(float(u1[i])))) for i in sample])
    assert ulps(got[sample], ref_l).max() <= 2.0
    assert ulps(got, -2.0 * np.log(u1)).max() <= 3.0
    near_one = 1.0 - np.ldexp(rng.uniform(0.5, 1.0, 2000), rng.randint(-33, -1, 2000))  # u close to 1: result tiny
    ref_n = np.array([float(-2 * mpmath.log(mpmath.mpf(float(v)))) for v in near_one])
    assert ulps(run(7, near_one), ref_n).max() <= 2.0
    s, c = run(8, words, two=True)
    ref_s = np.array([float(mpmath.sin(mpmath.pi * mpmath.mpf(float(u2[i])))) for i in idx])
    ref_c = np.array([float(mpmath.cos(mpmath.pi * mpmath.mpf(float(u2[i])))) for i in idx])
    assert np.abs(s[idx] - ref_s).max() < 3e-16 and np.abs(c[idx] - ref_c).max() < 3e-16
    assert np.abs(s - np.sin(np.pi * u2)).max() < 1.5e-15 and np.abs(s * s + c * c - 1.0).max() < 6e-16
    sub = th[:3000]
    ref_t = np.array([float(mpmath.sin(mpmath.mpf(float(a)))) for a in sub])
    assert np.abs(run(10, sub) - ref_t).max() < 3e-16 and np.abs(run(10, th) - np.sin(th)).max() < 4e-16
    assert ulps(run(11, pos), np.sqrt(pos)).max() <= 2.0
    d0, d1 = run(9, words, two=True)  # table-driven minus library normals
    assert np.abs(d0).max() < 2e-14 and np.abs(d1).max() < 2e-14


def test_one_step_ghmc_acceptance_matches_the_published_value(ib):
    """notebooks/"1D figures, updated.ipynb" cells 65-67: acceptance ratios min(1, exp(-W_shad)) of ONE step of VRORV at
    dt = 0.6, gamma = 10 from equilibrium of the double well: mean 0.84828 +/- 0.00071, std 0.22585 (1e5 samples there,
    4e6 here)."""
    from integrator_benchmark_b200.integrators import baoab_factory
    from integrator_benchmark_b200.testsystems import double_well
    ib.set_seed(314)
    scheme = baoab_factory(double_well.potential, double_well.force, double_well.velocity_scale, double_well.mass)
    res = scheme.ensemble(4000000, 2, 10.0, 0.6, n_legs=1, x_cache=double_well.x_samples_device())
    ratios = ib.engine.torch().exp(-res["W"][0]).clamp(max=1.0)
    mean, std = ratios.mean().item(), ratios.std().item()
    assert abs(mean - 0.84827685) < 3 * 0.00071420 + 3 * std / 2000.0, mean
    assert std == pytest.approx(0.22585, abs=0.003)


def test_quartic_steady_state_histogram_matches_the_published_ground_truth(ib):
    """notes/note_on_dF_neq_approximation.tex:349-380: quartic, m = 10, beta = 1, gamma = 100, dt = 1.1; histogram
    D_KL(rho_x || pi_x) of the steady state: VVVR (OVRVO) ~ 0.01309, BAOAB (VRORV) ~ 0.00013 (scripts/validate_quartic_histogram.py)"""
    import importlib.util
    import os
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    spec = importlib.util.spec_from_file_location("vqh", os.path.join(root, "scripts", "validate_quartic_histogram.py"))
    vqh = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(vqh)
    from integrator_benchmark_b200.integrators import baoab_factory, vvvr_factory
    from integrator_benchmark_b200.testsystems import quartic
    ib.set_seed(99)
    edges = np.linspace(-2.5, 2.5, 101)
    got = {}
    for name, factory in (("VRORV", baoab_factory), ("OVRVO", vvvr_factory)):
        scheme = factory(quartic.potential, quartic.force, quartic.velocity_scale, quartic.mass)
        res = scheme.ensemble(1 << 22, 301, 100.0, 1.1, n_legs=1, x_cache=quartic.x_samples_device(), want_final=True)
        x = res["x_final"].cpu().numpy()
        got[name] = vqh.histogram_kl(x[np.isfinite(x) & (np.abs(x) < 2.5)], edges)
    assert got["OVRVO"] == pytest.approx(0.01309, rel=0.03), got
    assert 0.00007 < got["VRORV"] < 0.00019, got


@pytest.mark.parametrize("kind,precision", [("double_well", "double"), ("quartic", "double"), ("double_well", "mixed"),
                                            ("harmonic", "double")])
def test_batch_launch_equals_individual_launches(ib, orc, kind, precision):
    """lsi_run_protocol_batch == one lsi_run_protocol per program, bit for bit: the one-launch families (six 5-letter
    palindromes, the reference's three 6-letter closure strings), a permuted subset, and batches that fall back to
    separate launches (mixed families, a repeated program, a string outside the families)."""
    import torch
    cache = _cache(orc, "quartic" if kind == "harmonic" else kind)
    system = ib.engine.ScalarSystem(kind, 10.0, (3.0, 0.0, 0.0, 0.0)) if kind == "harmonic" else ib.engine.ScalarSystem(kind, 10.0)
    six_letter = ["O V R R V O", "V R O O R V", "R V O O V R"]
    cases = [SIX, six_letter, SIX[::-1], six_letter[:2], six_letter[1::-1], [SIX[4], SIX[0], SIX[2]], [SIX[0], six_letter[1]], [SIX[1], SIX[1]], [SIX[0], "R V O"],
             [SIX[3]]]
    for n, steps in [(3001, 9), (700, 4), (1, 3), (33, 0), (385, 1), (2049, 2)]:
        for strings in cases:
            progs = [ib.engine.Program(s, 0.35, 10.0) for s in strings]
            singles = [ib.engine.run_protocol(system, p, n, steps, 1.0, seed=3, replica0=77, x_cache=cache, precision=precision,
                                              want_heat=True, want_final=True, out={}) for p in progs]
            mom = torch.full((len(progs), 8), -1.0, dtype=torch.float64, device="cuda")
            batch = ib.engine.run_protocols(system, progs, n, steps, 1.0, seed=3, replica0=77, x_cache=cache, precision=precision,
                                            want_heat=True, want_final=True, moments_out=mom)
            for i, (s, b) in enumerate(zip(singles, batch)):
                for key in ("W", "heat", "x_final", "v_final"):
                    assert torch.equal(s[key], b[key]), (strings, i, key)
                # moment sums: same terms, possibly another summation order (the whole-family kernel maps replicas to
                # threads differently); the count of finite pairs is exact
                assert torch.equal(mom[i], b["moments"]) and mom[i][0] == s["moments"][0] and mom[i][6] == s["moments"][6]
                torch.testing.assert_close(mom[i], s["moments"], rtol=1e-11, atol=1e-9)
    # degenerate batches: no program, no replica
    assert ib.engine.run_protocols(system, [], 10, 3, 1.0, x_cache=cache) == []
    empty = ib.engine.run_protocols(system, [ib.engine.Program(s, 0.35, 10.0) for s in SIX], 0, 3, 1.0, x_cache=cache, want_moments=False)
    assert all(o["W"].shape == (2, 0) for o in empty)
    # "full" marginal and explicit start states travel through the batch as well
    x0 = cache[:500].copy()
    v0 = np.linspace(-0.3, 0.3, 500)
    progs = [ib.engine.Program(s, 0.2, 5.0) for s in SIX]
    singles = [ib.engine.run_protocol(system, p, 500, 7, 1.0, marginal="full", seed=9, x0=x0, v0=v0, precision=precision, out={}) for p in progs]
    batch = ib.engine.run_protocols(system, progs, 500, 7, 1.0, marginal="full", seed=9, x0=x0, v0=v0, precision=precision)
    for s, b in zip(singles, batch):
        assert torch.equal(s["W"], b["W"])
        torch.testing.assert_close(b["moments"], s["moments"], rtol=1e-11, atol=1e-9)
    # one leg and three legs (legs >= 2 draw their start velocities from the TAG_V0 stream that begins at draw 1)
    for n_legs in (1, 3):
        progs = [ib.engine.Program(s, 0.3, 10.0) for s in SIX]
        singles = [ib.engine.run_protocol(system, p, 1500, 6, 1.0, n_legs=n_legs, seed=8, x_cache=cache, precision=precision,
                                          want_final=True, out={}) for p in progs]
        batch = ib.engine.run_protocols(system, progs, 1500, 6, 1.0, n_legs=n_legs, seed=8, x_cache=cache, precision=precision,
                                        want_final=True)
        for s, b in zip(singles, batch):
            assert s["W"].shape == (n_legs, 1500) and torch.equal(s["W"], b["W"]) and torch.equal(s["x_final"], b["x_final"])
    # programs of one family compiled with different time steps share no (a, b sigma): still one launch, still identical
    progs = [ib.engine.Program(s, 0.1 + 0.05 * i, 5.0 + i) for i, s in enumerate(SIX)]
    singles = [ib.engine.run_protocol(system, p, 2000, 9, 1.0, seed=4, x_cache=cache, precision=precision, out={}) for p in progs]
    batch = ib.engine.run_protocols(system, progs, 2000, 9, 1.0, seed=4, x_cache=cache, precision=precision)
    for s, b in zip(singles, batch):
        assert torch.equal(s["W"], b["W"]) and torch.equal(s["moments"], b["moments"])


def test_family_launch_at_baseline_size(ib):
    """BASELINE configs[1] at full size (2^20 replicas, protocol_length 10, the six splittings in one launch) through the
    size-independent properties: a shard of the ensemble reproduces its slice of the full run bit for bit (what multi-GPU
    weak scaling relies on), two programs run separately agree bit for bit with the batch, the moment vector agrees with a
    reduction of the returned works, and the works are finite with the sign structure of the schemes (VRORV ~ 0, OVRVO > 0)."""
    import torch
    from integrator_benchmark_b200.testsystems import double_well
    n, steps = 1 << 20, 9
    system = ib.engine.ScalarSystem("double_well", 10.0)
    cache = torch.from_numpy(double_well.collect_equilibrium_samples(1000000, n_burn=200, seed=12)).cuda()
    progs = [ib.engine.Program(s, 0.35, 10.0) for s in SIX]
    mom = torch.zeros(6, 8, dtype=torch.float64, device="cuda")
    full = ib.engine.run_protocols(system, progs, n, steps, 1.0, seed=21, replica0=0, x_cache=cache, moments_out=mom)
    W = [o["W"].clone() for o in full]
    for lo, hi in [(0, n // 2), (n // 2, n), (n - 1000, n)]:
        part = ib.engine.run_protocols(system, progs, hi - lo, steps, 1.0, seed=21, replica0=lo, x_cache=cache)
        for p in range(6):
            assert torch.equal(part[p]["W"], W[p][:, lo:hi])
    for p in (0, 4):
        single = ib.engine.run_protocol(system, progs[p], n, steps, 1.0, seed=21, replica0=0, x_cache=cache, out={})
        assert torch.equal(single["W"], W[p])
    for p in range(6):
        assert bool(torch.isfinite(W[p]).all())
        m = ib.engine.reduce_moments(W[p][0], W[p][1])
        torch.testing.assert_close(mom[p][:7], m[:7], rtol=1e-11, atol=1e-7)
        assert mom[p][0].item() == n and mom[p][6].item() == 0
    dF = {s: ib.evaluation.estimate_from_moments(mom[i].cpu().numpy()) for i, s in enumerate(SIX)}
    assert abs(dF["V R O R V"][0]) < 5 * dF["V R O R V"][1] ** 0.5 + 2e-4      # BAOAB: configuration-space error ~ 0 at this dt
    assert dF["O V R V O"][0] > 10 * dF["O V R V O"][1] ** 0.5                   # OVRVO: clearly positive

"""Nested ("exact") estimator: log-mean-exp reduction, stopping rules and the two-level bootstrap
(reference: benchmark/evaluation/compare_near_eq_and_exact.py:32-320) on the GPU."""
import json
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ib():
    import integrator_benchmark_b200 as ib
    ib.engine.require_cuda()
    return ib


def test_logmeanexp_matches_reference_fixtures_and_numpy(ib, golden_dir):
    from integrator_benchmark_b200.evaluation import compare_near_eq_and_exact as nested
    cases = json.load(open(os.path.join(golden_dir, "estimators.json")))
    n = 0
    for c in cases:
        if c["kind"] == "logmeanexp":
            w = np.array(c["w"])
            assert nested.estimate_from_work_samples(w) == pytest.approx(c["estimate"], rel=1e-12, abs=1e-14)
            assert nested.stdev_log_rho_pi(w) == pytest.approx(c["stdev"], rel=1e-10)
            n += 1
        elif c["kind"] == "n_steps":
            assert nested.n_steps_(c["dt"], c["n_collisions"]) == c["value"]
            n += 1
    assert n >= 10
    # ragged rows, incl. works large enough to overflow the reference's unshifted exp
    rng = np.random.RandomState(0)
    rows = [rng.randn(k) * s + m for k, s, m in ((1, 1.0, 0.0), (7, 0.1, 3.0), (5000, 2.0, -1.0), (333, 5.0, 40.0), (64, 1.0, -700.0))]
    est, sd = ib.engine.reduce_logmeanexp(rows)
    est, sd = est.cpu().numpy(), sd.cpu().numpy()
    for r, e, s_ in zip(rows, est, sd):
        shift = (-r).max()
        ex = np.exp(-r - shift)
        assert e == pytest.approx(np.log(ex.mean()) + shift, rel=1e-12, abs=1e-12)
        assert s_ == pytest.approx(np.std(ex) / (ex.mean() * np.sqrt(len(r))), rel=1e-9, abs=1e-15)
    # infinite works give what the reference's naive formula gives, not NaN: all +inf -> log(0) = -inf, a -inf -> +inf
    est, _ = ib.engine.reduce_logmeanexp([np.array([np.inf, np.inf]), np.array([0.5, -np.inf, 1.0]), np.array([np.inf, 0.0])])
    est = est.cpu().numpy()
    assert est[0] == -np.inf and est[1] == np.inf and est[2] == pytest.approx(np.log(0.5))
    outer = [{"Ws": r} for r in rows[1:4]]
    per = np.array([np.log(np.mean(np.exp(-r))) for r in rows[1:4]])
    assert nested.stdev_kl_div(outer) == pytest.approx(np.std(per) / np.sqrt(3), rel=1e-10)


def test_bootstrap_nested_statistics(ib):
    from integrator_benchmark_b200.evaluation import compare_near_eq_and_exact as nested
    rng = np.random.RandomState(1)
    rows = [rng.randn(rng.randint(50, 400)) * 0.5 + 0.1 * k for k in range(40)]
    point = np.mean([np.log(np.mean(np.exp(-r))) for r in rows])
    a = nested.bootstrap_estimates(rows, n_bootstrap=400, seed=5)
    b = nested.bootstrap_estimates(rows, n_bootstrap=400, seed=5)
    assert np.array_equal(a, b)                       # counter-based: replicate b is a pure function of (seed, b)
    assert not np.array_equal(a, nested.bootstrap_estimates(rows, n_bootstrap=400, seed=6))
    # same distribution as the reference's host-side resample_Ws
    np.random.seed(0)
    ref = np.array([np.mean([np.log(np.mean(np.exp(-np.asarray(r, dtype=float)))) for r in nested.resample_Ws(rows)])
                    for _ in range(400)])
    assert abs(a.mean() - ref.mean()) < 4 * ref.std() / np.sqrt(400) * 2
    assert a.std() == pytest.approx(ref.std(), rel=0.2)
    assert abs(a.mean() - point) < 0.2
    # degenerate rows: every replicate equals the constant
    const = [np.full(17, 0.25), np.full(3, 0.25)]
    np.testing.assert_allclose(nested.bootstrap_estimates(const, 8, seed=1), -0.25, rtol=1e-14)


def test_adaptive_inner_loop_cuts_where_the_reference_would(ib):
    """collect_inner_samples_until_threshold launches big batches but returns exactly the prefix at which the
    reference's per-batch test (initial + k * batch) first passes."""
    from integrator_benchmark_b200 import unit
    from integrator_benchmark_b200.evaluation import compare_near_eq_and_exact as nested
    from integrator_benchmark_b200.integrators import LangevinSplittingIntegrator
    from integrator_benchmark_b200.testsystems import EquilibriumSimulator, NonequilibriumSimulator, configure_platform, systems
    desc, pos = systems.water_cluster(20, constrained=True, seed=1)
    eq = EquilibriumSimulator(platform=configure_platform("CUDA"), topology=None, system=desc, positions=pos,
                              temperature=298 * unit.kelvin, timestep=1.0 * unit.femtosecond, burn_in_length=300, n_samples=16,
                              thinning_interval=10, name="wc")
    integ = LangevinSplittingIntegrator("O V R V O", collision_rate=1.0 / unit.picosecond, timestep=4.0 * unit.femtosecond)
    sim = NonequilibriumSimulator(eq, integ)
    ib.set_seed(11, next_replica=0)
    rho = nested.sample_from_rho(sim, n_collisions=2)
    x, v = rho["x"], rho["v"]
    assert x.shape == (60, 3)
    for marginal in ("full", "configuration"):
        ib.set_seed(12, next_replica=1000)
        Ws = nested.collect_inner_samples_until_threshold(x, v, sim, marginal, initial_size=50, batch_size=7, n_steps=30,
                                                          threshold=0.02, max_samples=3000)
        ib.set_seed(12, next_replica=1000)
        long = np.concatenate([nested._inner_batch(sim, x, v, 30, marginal, 50), nested._inner_batch(sim, x, v, 30, marginal, 512),
                               nested._inner_batch(sim, x, v, 30, marginal, 512)])
        np.testing.assert_array_equal(Ws, long[:len(Ws)])   # same streams, same order
        # reference rule replayed on the same samples
        ref_sd = lambda w: np.std(np.exp(-w)) / (np.mean(np.exp(-w)) * np.sqrt(len(w)))  # noqa: E731
        k = 50
        while ref_sd(long[:k]) > 0.02 and k <= 3000 - 7:
            k += 7
        assert len(Ws) == k
        assert np.isfinite(Ws).all()
    # a batch of outer samples in one launch, and the estimator end to end on a tiny budget
    batch = nested.sample_from_rho(sim, n=8)
    assert batch["x"].shape == (8, 60, 3) and np.isfinite(batch["v"]).all()
    res = nested.estimate_kl_div_naive_outer_loop(sim, "full", lambda **kw: nested.outer_sample_naive(n_inner_samples=64, **kw),
                                                  n_outer_samples=3, n_steps=20)
    assert np.isfinite(res["new_estimate"]) and len(res["Ws"]) == 3


def test_batched_outer_samples_keep_the_reference_stopping_rule(ib):
    """collect_outer_samples_batched advances the inner loops of all outer samples together; every returned work array
    is cut exactly where the reference's rule (tested every `batch_size` from `initial_size`) first passes"""
    from functools import partial
    from integrator_benchmark_b200 import unit
    from integrator_benchmark_b200.evaluation import compare_near_eq_and_exact as nested
    from integrator_benchmark_b200.integrators import LangevinSplittingIntegrator
    from integrator_benchmark_b200.testsystems import EquilibriumSimulator, NonequilibriumSimulator, configure_platform, systems
    desc, pos = systems.water_cluster(20, constrained=True, seed=1)
    eq = EquilibriumSimulator(platform=configure_platform("CUDA"), topology=None, system=desc, positions=pos,
                              temperature=298 * unit.kelvin, timestep=1.0 * unit.femtosecond, burn_in_length=300, n_samples=16,
                              thinning_interval=10, name="wc")
    eq.metropolis_rounds = 20
    integ = LangevinSplittingIntegrator("O V R V O", collision_rate=1.0 / unit.picosecond, timestep=4.0 * unit.femtosecond)
    sim = NonequilibriumSimulator(eq, integ)
    ib.set_seed(21, next_replica=0)
    ref_sd = lambda w: np.std(np.exp(-w)) / (np.mean(np.exp(-w)) * np.sqrt(len(w)))  # noqa: E731
    for marginal in ("full", "configuration"):
        samples = nested.collect_outer_samples_batched(sim, marginal, 6, 25, initial_size=40, batch_size=3, threshold=0.03,
                                                       max_samples=2000)
        assert len(samples) == 6
        for s in samples:
            w = s["Ws"]
            assert np.isfinite(w).all() and (len(w) - 40) % 3 == 0 and 40 <= len(w) <= 2000
            assert not (ref_sd(w) > 0.03) or not (len(w) <= 2000 - 3)          # the rule passes here ...
            assert all(ref_sd(w[:c]) > 0.03 for c in range(40, len(w), 3))     # ... and nowhere before
            assert s["estimate"] == pytest.approx(np.log(np.mean(np.exp(-w))), rel=1e-10)
            assert s["xv"][0].shape == (60, 3)
    outer = partial(nested.outer_sample_adaptive, initial_size=40, batch_size=3, threshold=0.05, max_samples=500)
    res = nested.estimate_kl_div_adaptive_outer_loop_batched(sim, "full", outer, n_steps=25, initial_size=5, batch_size=4,
                                                             threshold=0.02, max_samples=13)
    assert np.isfinite(res["new_estimate"]) and 5 <= len(res["Ws"]) <= 13 and (len(res["Ws"]) - 5) % 4 == 0
    # the rank-sharded loop (a single rank here) with the batched inner loops
    res = nested.estimate_kl_div_adaptive_outer_loop_sharded(sim, "full", None, n_steps=25, initial_size=5, batch_size=4, threshold=0.02,
                                                             max_samples=13, inner=dict(initial_size=40, batch_size=3, threshold=0.05,
                                                                                        max_samples=500))
    assert np.isfinite(res["new_estimate"]) and res["n_outer_samples"] == len(res["Ws"]) and 5 <= len(res["Ws"]) <= 13

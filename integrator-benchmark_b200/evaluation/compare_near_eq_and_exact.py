"""Drop-in for benchmark/evaluation/compare_near_eq_and_exact.py:32-320 -- the nested ("exact")
estimator of D_KL(rho || pi): outer samples (x, v) ~ rho, inner work samples started from each,
log-mean-exp(-W) per outer sample, adaptive stopping on both loops, two-level bootstrap.

The reference runs every inner trajectory as one Python-level OpenMM call (10^6 - 10^9
replica-steps per point, spread over multiprocessing.Pool(32)).  Here each inner batch is ONE
launch over many replicas and the reductions run on the GPU; the stopping rules are evaluated at
exactly the sample counts the reference would have checked, so the adaptive lengths are the
reference's.
"""
import warnings

import numpy as np

from .. import engine, unit

# experiment constants of the reference (:21-30, :49-59)
splittings = {"OVRVO": "O V R V O", "ORVRO": "O R V R O", "RVOVR": "R V O V R", "VRORV": "V R O R V"}
marginals = ["configuration", "full"]
dt_range = np.array([0.1] + list(np.arange(0.5, 8.001, 0.5))) * unit.femtosecond
collision_rate = 1.0 / unit.picoseconds
inner_loop_initial_size = 50
inner_loop_batch_size = 1
inner_loop_stdev_threshold = 0.01
inner_loop_max_samples = 50000
outer_loop_initial_size = 50
outer_loop_batch_size = 100
outer_loop_stdev_threshold = inner_loop_stdev_threshold
outer_loop_max_samples = 1000
GPU_MIN_BATCH = 512  # trajectories per launch when the caller's batch is smaller


def n_steps_(dt, n_collisions=1, max_steps=1000):
    """min(max_steps, int((n_collisions / collision_rate) / dt))  (:32-46)"""
    dt_ps = dt.value_in_unit(unit.picosecond) if unit.is_quantity(dt) else float(dt)
    return min(max_steps, int((n_collisions / collision_rate.value_in_unit(unit.picosecond ** -1)) / dt_ps))


def estimate_from_work_samples(work_samples):
    """log(mean(exp(-w))) (:95-97), reduced on the GPU in the max-shifted form"""
    est, _ = engine.reduce_logmeanexp([np.asarray(work_samples, dtype=np.float64)])
    return float(est[0])


def stdev_log_rho_pi(w):
    """std(exp(-w)) / (mean(exp(-w)) sqrt(len(w))), ddof = 0 (:62-83)"""
    assert not unit.is_quantity(w)
    assert isinstance(w, np.ndarray)
    _, sd = engine.reduce_logmeanexp([w])
    return float(sd[0])


def stdev_kl_div(outer_samples):
    """np.std of the per-outer-sample estimates / sqrt(n) (:86-92)"""
    est, _ = engine.reduce_logmeanexp([s["Ws"] for s in outer_samples])
    est = est.cpu().numpy()
    return float(np.std(est) / np.sqrt(len(est)))


def inner_sample(noneq_sim, x, v, n_steps, marginal="full"):
    if marginal == "full":
        pass
    elif marginal == "configuration":
        v = noneq_sim.sample_v_given_x(x)
    else:
        raise Exception("marginal must be `full` or `configuration`")
    return noneq_sim.accumulate_shadow_work(x, v, n_steps)["W_shad"]


def _inner_batch(noneq_sim, x, v, n_steps, marginal, n):
    if marginal == "full":
        return noneq_sim.ensemble_shadow_work(x, v, n_steps, n=n)
    if marginal == "configuration":
        return noneq_sim.ensemble_shadow_work(x, None, n_steps, n=n)
    raise Exception("marginal must be `full` or `configuration`")


def collect_inner_samples_naive(x, v, noneq_sim, marginal="full", n_steps=1000, n_inner_samples=100):
    """A fixed number of noneq trajectories starting from x, v: one launch (:111-116)"""
    return _inner_batch(noneq_sim, x, v, n_steps, marginal, n_inner_samples)


def _running_stdev(Ws, counts):
    """stdev_log_rho_pi of the prefixes Ws[:c] for c in counts (one pass of cumulative sums)"""
    e = np.exp(-(Ws - Ws.min()))  # common shift cancels in the ratio
    c1, c2 = np.cumsum(e), np.cumsum(e * e)
    k = np.asarray(counts) - 1
    mean = c1[k] / counts
    var = np.maximum(c2[k] / counts - mean * mean, 0.0)
    return np.sqrt(var) / (mean * np.sqrt(counts))


def collect_inner_samples_until_threshold(x, v, noneq_sim, marginal="full", initial_size=100, batch_size=5, n_steps=1000,
                                          threshold=0.1, max_samples=1000):
    """Up to max_samples trajectories, fewer if the stdev of the estimated log(rho / pi) drops below
    threshold (:119-139).  Trajectories are launched GPU_MIN_BATCH at a time; the stopping rule is applied
    at the counts initial_size + k * batch_size the reference checks, and the returned array is cut there."""
    Ws = _inner_batch(noneq_sim, x, v, n_steps, marginal, initial_size)
    checked = initial_size
    while True:
        # counts the reference would test with the samples in hand
        counts = np.arange(checked, len(Ws) + 1, batch_size)
        sd = _running_stdev(Ws, counts)
        hit = np.nonzero(~(sd > threshold) | ~(counts <= (max_samples - batch_size)))[0]
        if len(hit) > 0:
            Ws = Ws[:counts[hit[0]]]
            break
        checked = counts[-1] + batch_size
        grow = max(batch_size, GPU_MIN_BATCH, checked - len(Ws))
        grow = min(grow, max(batch_size, max_samples - len(Ws)))
        Ws = np.concatenate([Ws, _inner_batch(noneq_sim, x, v, n_steps, marginal, grow)])
    if stdev_log_rho_pi(np.array(Ws)) > threshold:
        warnings.warn("stdev_log_rho_pi(Ws) > threshold\n({:.3f} > {:.3f})".format(stdev_log_rho_pi(np.array(Ws)), threshold),
                      RuntimeWarning)
    return np.array(Ws)


def sample_from_rho(noneq_sim, n_collisions=2, n=None):
    """(x, v) ~ rho: equilibrium draw pushed through n_steps_(dt, n_collisions) steps (:142-158).
    With n given, n such samples from one launch (arrays of shape (n, n_atoms, 3))."""
    n_steps = n_steps_(noneq_sim.integrator.getStepSize(), n_collisions)
    eq = noneq_sim.equilibrium_simulator
    if n is None:
        x0 = noneq_sim.sample_x_from_equilibrium()
        _, x, v = noneq_sim.ensemble_shadow_work(x0, None, n_steps, n=1, want_final=True)
        return {"x": x[0], "v": v[0] * (unit.nanometer / unit.picosecond)}
    eq.load_or_simulate_samples()
    idx = np.random.randint(len(eq.samples), size=int(n))
    _, x, v = noneq_sim.ensemble_shadow_work(eq.samples[idx], None, n_steps, want_final=True)
    return {"x": x, "v": v}


def outer_sample_naive(index=0, noneq_sim=None, marginal="full", n_inner_samples=100, n_steps=1000):
    rho_sample = sample_from_rho(noneq_sim)
    x, v = rho_sample["x"], rho_sample["v"]
    Ws = collect_inner_samples_naive(x, v, noneq_sim, marginal, n_steps, n_inner_samples)
    return {"xv": (x, v), "Ws": Ws, "estimate": estimate_from_work_samples(Ws)}


def outer_sample_adaptive(noneq_sim=None, marginal="full", n_steps=1000, initial_size=100, batch_size=5, threshold=0.1,
                          max_samples=1000):
    rho_sample = sample_from_rho(noneq_sim)
    x, v = rho_sample["x"], rho_sample["v"]
    Ws = collect_inner_samples_until_threshold(x, v, noneq_sim, marginal, initial_size, batch_size, n_steps, threshold, max_samples)
    return {"xv": (x, v), "Ws": Ws, "estimate": estimate_from_work_samples(Ws)}


def noneq_sim_factory(testsystem_name, scheme, dt, collision_rate):
    from .. import testsystems as ts
    from ..integrators import LangevinSplittingIntegrator
    from ..testsystems import NonequilibriumSimulator
    systems = {"alanine_constrained": ts.alanine_constrained, "water_cluster_rigid": ts.water_cluster_rigid}
    assert testsystem_name in systems
    assert scheme in splittings
    assert unit.is_quantity(dt) and dt.unit.is_compatible(unit.femtosecond)
    assert unit.is_quantity(collision_rate) and (1 / collision_rate).unit.is_compatible(unit.picosecond)
    integrator = LangevinSplittingIntegrator(splittings[scheme], temperature=ts.molecular.temperature,
                                             collision_rate=collision_rate, timestep=dt)
    return NonequilibriumSimulator(systems[testsystem_name], integrator)


def process_outer_samples(outer_samples):
    """D_KL estimate: mean over x ~ rho of log(rho(x) / pi(x)) (:227-233)"""
    return np.mean([s["estimate"] for s in outer_samples], 0)


def estimate_kl_div_naive_outer_loop(noneq_sim, marginal, outer_sample_fxn, n_outer_samples=100, n_steps=1000, return_samples=False):
    outer_samples = [outer_sample_fxn(noneq_sim=noneq_sim, marginal=marginal, n_steps=n_steps) for _ in range(n_outer_samples)]
    result = {"new_estimate": process_outer_samples(outer_samples), "Ws": np.array([s["Ws"] for s in outer_samples], dtype=object)}
    if return_samples:
        result["samples"] = outer_samples
    return result


def estimate_kl_div_adaptive_outer_loop(noneq_sim, marginal, outer_sample_fxn, n_steps=1000, initial_size=100, batch_size=1,
                                        threshold=0.1, max_samples=1000, return_samples=False):
    """Up to max_samples outer-loop samples, stopping on stdev_kl_div (:257-290)"""
    collect = lambda: outer_sample_fxn(noneq_sim=noneq_sim, marginal=marginal, n_steps=n_steps)  # noqa: E731
    outer_samples = [collect() for _ in range(initial_size)]
    while (stdev_kl_div(outer_samples) > threshold) and (len(outer_samples) <= (max_samples - batch_size)):
        for _ in range(batch_size):
            outer_samples.append(collect())
    if stdev_kl_div(outer_samples) > threshold:
        warnings.warn("stdev_kl_div(outer_samples) > threshold\n({:.3f} > {:.3f})".format(stdev_kl_div(outer_samples), threshold),
                      RuntimeWarning)
    result = {"new_estimate": process_outer_samples(outer_samples), "Ws": np.array([s["Ws"] for s in outer_samples], dtype=object)}
    if return_samples:
        result["samples"] = outer_samples
    return result


GPU_FILL = 16384  # trajectories per launch that keep a B200 busy for 60-atom systems
GPU_MAX_LAUNCH = 262144  # cap on one launch of the batched outer loop (host staging of the start states)


def collect_outer_samples_batched(noneq_sim, marginal, n_outer, n_steps, initial_size=100, batch_size=5, threshold=0.1,
                                  max_samples=1000, n_collisions=2):
    """`n_outer` outer samples at once: what `[outer_sample_adaptive(...) for _ in range(n_outer)]` computes, with the
    inner loops of ALL outer samples advanced together, one launch per round (the reference spreads outer samples over a
    Pool(32): compare_near_eq_and_exact.py:357-368).  Every outer sample keeps its own stopping rule -- stdev_log_rho_pi
    tested at the counts initial_size + k batch_size, result cut at the first count that passes (:119-139) -- but the
    trajectories of a round are drawn ahead of the rule (each unfinished sample grows by half its size, at least
    GPU_FILL trajectories per launch in total), so some are computed and then cut."""
    if marginal not in ("full", "configuration"):
        raise Exception("marginal must be `full` or `configuration`")
    rho = sample_from_rho(noneq_sim, n_collisions=n_collisions, n=n_outer)
    xs, vs = np.asarray(rho["x"]), np.asarray(rho["v"])
    Ws = [np.zeros(0) for _ in range(n_outer)]
    checked = [initial_size] * n_outer   # next count the stopping rule has to look at
    final = [None] * n_outer
    active = list(range(n_outer))
    while active:
        want = [initial_size if len(Ws[i]) == 0 else max(batch_size, (len(Ws[i]) + 1) // 2) for i in active]
        want = [min(w, max(batch_size, max_samples - len(Ws[i]))) for w, i in zip(want, active)]
        scale = max(1, -(-GPU_FILL // max(1, sum(want)))) if len(Ws[active[0]]) > 0 else 1
        want = [min(w * scale, max(batch_size, max_samples - len(Ws[i]))) for w, i in zip(want, active)]
        offsets = np.concatenate([[0], np.cumsum(want)])
        W = np.empty(int(offsets[-1]))
        lo = 0
        while lo < len(active):  # launches of at most GPU_MAX_LAUNCH trajectories
            hi = lo + 1
            while hi < len(active) and offsets[hi + 1] - offsets[lo] <= GPU_MAX_LAUNCH:
                hi += 1
            x0 = np.repeat(xs[active[lo:hi]], want[lo:hi], axis=0)
            v0 = np.repeat(vs[active[lo:hi]], want[lo:hi], axis=0) if marginal == "full" else None
            W[offsets[lo]:offsets[hi]] = noneq_sim.ensemble_shadow_work(x0, v0, n_steps)
            lo = hi
        still = []
        for k, i in enumerate(active):
            Ws[i] = np.concatenate([Ws[i], W[offsets[k]:offsets[k + 1]]])
            counts = np.arange(checked[i], len(Ws[i]) + 1, batch_size)
            if len(counts) > 0:
                sd = _running_stdev(Ws[i], counts)
                hit = np.nonzero(~(sd > threshold) | ~(counts <= (max_samples - batch_size)))[0]
                if len(hit) > 0:
                    final[i] = Ws[i][:counts[hit[0]]]
                    continue
                checked[i] = counts[-1] + batch_size
            still.append(i)
        active = still
    return [{"xv": (xs[i], vs[i]), "Ws": final[i], "estimate": estimate_from_work_samples(final[i])} for i in range(n_outer)]


def estimate_kl_div_adaptive_outer_loop_batched(noneq_sim, marginal, outer_sample_fxn=None, n_steps=1000, initial_size=100,
                                                batch_size=1, threshold=0.1, max_samples=1000, return_samples=False,
                                                inner=None):
    """estimate_kl_div_adaptive_outer_loop (:257-290) with each batch of outer samples collected by
    collect_outer_samples_batched.  `inner` = dict(initial_size, batch_size, threshold, max_samples) of the inner loop
    (or a functools.partial of outer_sample_adaptive carrying them, as the reference's __main__ builds)."""
    if inner is None:
        inner = dict(getattr(outer_sample_fxn, "keywords", {}) or {})
    collect = lambda n: collect_outer_samples_batched(noneq_sim, marginal, n, n_steps, **inner)  # noqa: E731
    outer_samples = collect(initial_size)
    while (stdev_kl_div(outer_samples) > threshold) and (len(outer_samples) <= (max_samples - batch_size)):
        outer_samples.extend(collect(batch_size))
    if stdev_kl_div(outer_samples) > threshold:
        warnings.warn("stdev_kl_div(outer_samples) > threshold\n({:.3f} > {:.3f})".format(stdev_kl_div(outer_samples), threshold),
                      RuntimeWarning)
    result = {"new_estimate": process_outer_samples(outer_samples), "Ws": np.array([s["Ws"] for s in outer_samples], dtype=object)}
    if return_samples:
        result["samples"] = outer_samples
    return result


def estimate_kl_div_adaptive_outer_loop_sharded(noneq_sim, marginal, outer_sample_fxn, n_steps=1000, initial_size=100,
                                                batch_size=1, threshold=0.1, max_samples=1000, return_samples=False,
                                                inner=None):
    """The adaptive outer loop with the OUTER samples sharded over the ranks of torch.distributed (one process per
    GPU; every rank runs its own inner ensembles on its own device).  Rank r starts with its share of
    `initial_size` (parallel.shard_range) and adds `batch_size` samples per round; the stopping rule of
    estimate_kl_div_adaptive_outer_loop (:257-290) is evaluated on the union through one all-reduce of
    (n, sum, sum of squares) of the per-sample estimates per round -- 24 bytes -- so all ranks stop together.
    `inner` = dict(initial_size, batch_size, threshold, max_samples): collect each rank's share of a round with
    collect_outer_samples_batched instead of calling `outer_sample_fxn` per sample.
    Returns the global estimate and stdev, the global sample count, and this rank's Ws / samples."""
    import torch
    import torch.distributed as dist
    from .. import parallel
    from .. import _rngstate
    rank, world, _ = parallel.world()
    on = dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1
    device = "cuda" if (on and dist.get_backend() == "nccl") else "cpu"
    # Ranks must draw DIFFERENT trajectories, or the union double-counts samples and the stopping rule fires early:
    # (1) replica ids (Philox streams) come from disjoint per-rank ranges; (2) the numpy generator that picks the outer
    # configurations must differ between ranks -- if two ranks were seeded alike, each rank reseeds from (state, rank).
    _rngstate.set_rank(rank)
    if on:
        probe = torch.tensor([float(np.random.get_state()[1][0])], dtype=torch.float64, device=device)
        probes = [torch.zeros_like(probe) for _ in range(world)]
        dist.all_gather(probes, probe)
        if len({float(p.item()) for p in probes}) < world:
            np.random.seed(np.random.SeedSequence([int(probe.item()), rank]).generate_state(4))
    if inner is not None:  # this rank's share of a round in one batched collection (collect_outer_samples_batched)
        collect_n = lambda k: collect_outer_samples_batched(noneq_sim, marginal, k, n_steps, **inner) if k > 0 else []  # noqa: E731
    else:
        collect_n = lambda k: [outer_sample_fxn(noneq_sim=noneq_sim, marginal=marginal, n_steps=n_steps) for _ in range(k)]  # noqa: E731
    lo, hi = parallel.shard_range(initial_size, rank, world)
    outer_samples = collect_n(hi - lo)

    def global_stats():
        e = np.array([s["estimate"] for s in outer_samples], dtype=np.float64)
        m = torch.tensor([float(len(e)), e.sum(), (e * e).sum()], dtype=torch.float64, device=device)
        if on:
            dist.all_reduce(m, op=dist.ReduceOp.SUM)
        n, s1, s2 = (float(v) for v in m.cpu())
        mean = s1 / n
        var = max(s2 / n - (s1 / n) ** 2, 0.0)  # np.std, ddof = 0
        return int(n), mean, float(np.sqrt(var) / np.sqrt(n))
    n, mean, sd = global_stats()
    while sd > threshold and n <= max_samples - batch_size * world:
        outer_samples.extend(collect_n(batch_size))
        n, mean, sd = global_stats()
    if sd > threshold and rank == 0:
        warnings.warn("stdev_kl_div(outer_samples) > threshold\n({:.3f} > {:.3f})".format(sd, threshold), RuntimeWarning)
    result = {"new_estimate": mean, "stdev": sd, "n_outer_samples": n,
              "Ws": np.array([s["Ws"] for s in outer_samples], dtype=object)}
    if return_samples:
        result["samples"] = outer_samples
    return result


def resample_Ws(Ws):
    """One two-level bootstrap sample of the ragged Ws structure on the host, as the reference (:293-320)."""
    Ws = np.asarray(Ws, dtype=object) if not isinstance(Ws, np.ndarray) else Ws
    n_outer = len(Ws)
    Ws_ = Ws[np.random.randint(0, n_outer, n_outer)].copy()
    for i in range(n_outer):
        n_cols = len(Ws_[i])
        Ws_[i] = np.asarray(Ws_[i])[np.random.randint(0, n_cols, n_cols)]
    return Ws_


def bootstrap_estimates(Ws, n_bootstrap=100, seed=0):
    """n_bootstrap replicates of process_outer_samples over resample_Ws(Ws), all on the GPU
    (the notebook loop around resample_Ws: near-eq-vs-exact-5x.ipynb cell 19)."""
    return engine.bootstrap_nested(Ws, n_bootstrap, seed).cpu().numpy()

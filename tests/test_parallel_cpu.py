"""Multi-process host logic on CPU (gloo, world size 2): shard ranges, the moment all-reduce and the
log-mean-exp merge give the single-process answer.  No CUDA calls."""
import os
import socket
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, out_dir):
    sys.path.insert(0, ROOT)
    os.environ.update(RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank), MASTER_ADDR="127.0.0.1",
                      MASTER_PORT=str(port))
    import torch
    import torch.distributed as dist
    from integrator_benchmark_b200 import engine, parallel
    r, w, _ = parallel.init_process_group(backend="gloo")
    assert (r, w) == (rank, world)
    rng = np.random.RandomState(0)
    n_total = 1001
    wF, wR = rng.randn(n_total), rng.randn(n_total) * 0.5 + 0.1
    lo, hi = parallel.shard_range(n_total, rank, world)
    f, b = wF[lo:hi], wR[lo:hi]
    m = torch.tensor([[hi - lo, f.sum(), b.sum(), (f * f).sum(), (b * b).sum(), (f * b).sum(), 0.0, 0.0]] * 3,
                     dtype=torch.float64)
    parallel.all_reduce_moments(m)
    dF, var = engine.neq_free_energy_from_moments(m[1].numpy())
    # nested estimator rows sharded by columns
    rows = rng.randn(4, 600) * 2.0
    clo, chi = parallel.shard_range(600, rank, world)
    part = torch.tensor(-rows[:, clo:chi])
    rmax = part.max(dim=1).values
    rsum = torch.exp(part - rmax[:, None]).sum(dim=1)
    cnt = torch.full((4,), float(chi - clo), dtype=torch.float64)
    lme = parallel.all_reduce_logmeanexp(rmax, rsum, cnt)
    # the bench's grouped exchange: a slice of a ring of moment rows travels in one asynchronous all-reduce, the other rows
    # are untouched; ranks hand out replica ids (Philox streams) from disjoint ranges
    from integrator_benchmark_b200 import _rngstate
    ring = torch.zeros(4, 3, 8, dtype=torch.float64)
    ring[:] = torch.arange(4, dtype=torch.float64)[:, None, None] + 10.0 * (rank + 1)
    h = parallel.all_reduce_moments_async(ring[1:3])
    h.wait()
    ok_ring = bool((ring[0] == 10.0 * (rank + 1)).all() and (ring[3] == 3.0 + 10.0 * (rank + 1)).all()
                   and (ring[1] == 2 * 1.0 + 30.0).all() and (ring[2] == 2 * 2.0 + 30.0).all())
    ok_ids = (_rngstate.reserve_replicas(10) >> _rngstate.RANK_SHIFT) == rank
    np.save(os.path.join(out_dir, "r%d.npy" % rank), np.concatenate([[dF, var], lme.numpy(), [lo, hi], [float(ok_ring), float(ok_ids)]]))
    dist.destroy_process_group()


def test_gloo_world2_matches_single_process(tmp_path):
    import torch.multiprocessing as mp
    port = _free_port()
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    rng = np.random.RandomState(0)
    wF, wR = rng.randn(1001), rng.randn(1001) * 0.5 + 0.1
    dF = 0.5 * (wF.mean() - wR.mean())
    var = (np.var(wF) + np.var(wR) - 2 * np.cov(wF, wR)[0, 1]) / (4 * 1001)
    rows = rng.randn(4, 600) * 2.0
    lme = np.log(np.mean(np.exp(-rows), axis=1))
    got = [np.load(os.path.join(str(tmp_path), "r%d.npy" % r)) for r in range(2)]
    for g in got:
        assert g[0] == pytest.approx(dF, rel=1e-12, abs=1e-14) and g[1] == pytest.approx(var, rel=1e-10)
        np.testing.assert_allclose(g[2:6], lme, rtol=1e-12)
    assert (got[0][6], got[0][7], got[1][6], got[1][7]) == (0, 501, 501, 1001)
    assert all(g[8] == 1.0 and g[9] == 1.0 for g in got)


def _fake_outer_sample(noneq_sim=None, marginal=None, n_steps=None):
    """stands in for outer_sample_adaptive: a deterministic per-rank stream of estimates, no GPU"""
    st = noneq_sim
    est = st["values"][st["next"]]
    st["next"] += 1
    return {"estimate": float(est), "Ws": np.array([est, est])}


def _nested_worker(rank, world, port, out_dir):
    sys.path.insert(0, ROOT)
    os.environ.update(RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank), MASTER_ADDR="127.0.0.1",
                      MASTER_PORT=str(port))
    import torch.distributed as dist
    from integrator_benchmark_b200 import parallel
    from integrator_benchmark_b200.evaluation import compare_near_eq_and_exact as nested
    parallel.init_process_group(backend="gloo")
    values = np.random.RandomState(100 + rank).randn(500) * 0.8 + 0.3
    state = {"values": values, "next": 0}
    res = nested.estimate_kl_div_adaptive_outer_loop_sharded(state, "full", _fake_outer_sample, initial_size=11, batch_size=2,
                                                             threshold=0.05, max_samples=900)
    np.save(os.path.join(out_dir, "n%d.npy" % rank), np.array([res["new_estimate"], res["stdev"], res["n_outer_samples"],
                                                                 state["next"], len(res["Ws"])]))
    dist.destroy_process_group()


def test_gloo_world2_nested_outer_loop_stops_together(tmp_path):
    """outer samples sharded over two ranks: both stop in the same round, on the statistics of the union"""
    import torch.multiprocessing as mp
    port = _free_port()
    mp.spawn(_nested_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    got = [np.load(os.path.join(str(tmp_path), "n%d.npy" % r)) for r in range(2)]
    assert got[0][2] == got[1][2] and got[0][0] == got[1][0] and got[0][1] == got[1][1]
    n0, n1 = int(got[0][3]), int(got[1][3])
    assert n0 + n1 == int(got[0][2]) and (n0 - 6) == (n1 - 5) and (n0 - 6) % 2 == 0  # 6 + 5 initial, then 2 per round each
    vals = np.concatenate([np.random.RandomState(100).randn(500)[:n0] * 0.8 + 0.3,
                           np.random.RandomState(101).randn(500)[:n1] * 0.8 + 0.3])
    assert got[0][0] == pytest.approx(vals.mean(), rel=1e-12)
    assert got[0][1] == pytest.approx(vals.std() / np.sqrt(len(vals)), rel=1e-9)
    assert got[0][1] <= 0.05
    # one round earlier the rule had not been met
    prev = np.concatenate([np.random.RandomState(100).randn(500)[:n0 - 2] * 0.8 + 0.3,
                           np.random.RandomState(101).randn(500)[:n1 - 2] * 0.8 + 0.3])
    assert prev.std() / np.sqrt(len(prev)) > 0.05


def test_shard_ranges_partition():
    sys.path.insert(0, ROOT)
    from integrator_benchmark_b200 import parallel
    for n, w in ((10, 3), (0, 4), (7, 8), (1 << 20, 8)):
        edges = [parallel.shard_range(n, r, w) for r in range(w)]
        assert edges[0][0] == 0 and edges[-1][1] == n
        assert all(edges[i][1] == edges[i + 1][0] for i in range(w - 1))
        sizes = [b - a for a, b in edges]
        assert max(sizes) - min(sizes) <= 1

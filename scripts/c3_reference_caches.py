#!/usr/bin/env python
"""Where does the PUBLISHED DeltaF_neq(dt) curve of water_cluster_rigid / OVRVO (notebooks/near-eq result
plotting.ipynb:1160-1195) fall among curves computed from independent equilibrium caches built the reference's way?

For each of --caches caches: ONE extra-chance-HMC chain from the minimised start, burn-in 100 000 rounds, 1000
configurations every --thin rounds (EquilibriumSimulator.collect_equilibrium_samples_single_chain; reference:
bookkeepers.py:89-113, watercluster.py:92-97 with thinning 10 000), then the reference's protocol
(evaluation/collect_near_eq_samples.py:19-25: N = 100 000 protocols x 1000 steps, gamma = 1/ps, 17 time steps, both marginals).
Every point is appended to --out as a JSON line as soon as it is done; the summary (rank of the published value among the
cache estimates, point by point) goes to --out with suffix .summary.json."""
import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402

from scripts.validate_curves import DTS, PUB  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--caches", type=int, default=20)
    ap.add_argument("--thin", type=int, default=1000)
    ap.add_argument("--burn", type=int, default=100000)
    ap.add_argument("--n", type=int, default=100000)
    ap.add_argument("--protocol-length", type=int, default=1000)
    ap.add_argument("--precision", default="mixed")
    ap.add_argument("--double-dts", default="4.0,5.0,6.0", help="time steps repeated in fp64 (full marginal) on every cache")
    ap.add_argument("--budget-s", type=float, default=1e9, help="stop starting new points after this many seconds")
    ap.add_argument("--seed", type=int, default=31337)
    ap.add_argument("--out", default=os.path.join(ROOT, "gpurun_out", "c3_reference_caches.jsonl"))
    ap.add_argument("--load-caches", default=None)
    args = ap.parse_args()
    import torch
    import integrator_benchmark_b200 as ib
    from integrator_benchmark_b200 import unit
    from integrator_benchmark_b200.evaluation import estimate_from_moments
    from integrator_benchmark_b200.integrators import LangevinSplittingIntegrator
    from integrator_benchmark_b200.testsystems import NonequilibriumSimulator, water_cluster_rigid
    ib.engine.require_cuda()
    t_start = time.time()
    os.makedirs(os.path.dirname(args.out), exist_ok=True)
    log = open(args.out, "a")

    def emit(rec):
        log.write(json.dumps(rec) + "\n")
        log.flush()
        print(json.dumps(rec), flush=True)

    eq = water_cluster_rigid
    eq.n_samples, eq.burn_in_length = 1000, args.burn
    cache_file = os.path.join(os.path.dirname(args.out), "c3_reference_caches_xyz.npz")
    if args.load_caches:
        caches = np.load(args.load_caches)["xyz"].astype(np.float64)
    else:
        t0 = time.time()
        caches = eq.collect_equilibrium_samples_single_chain(n_chains=args.caches, thinning_interval=args.thin, seed=args.seed)
        emit({"event": "caches", "n": int(caches.shape[0]), "frames": int(caches.shape[1]), "thin": args.thin, "burn": args.burn,
              "seconds": time.time() - t0, "acceptance": [float(a) for a in eq.acceptance],
              "burn_in_acceptance": [float(a) for a in eq.burn_in_acceptance],
              "finite": bool(np.isfinite(caches).all())})
        np.savez_compressed(cache_file, xyz=caches.astype(np.float32))
    system = eq.engine_system()
    for c in range(caches.shape[0]):  # potential-energy statistics of every cache (are the chains equilibrated alike?)
        e, _ = system.energy_forces(caches[c], want_forces=False)
        emit({"event": "cache_energy", "cache": c, "mean": float(e.mean()), "std": float(e.std()), "first_100": float(e[:100].mean()),
              "last_100": float(e[-100:].mean())})

    def point(c, marginal, dt_fs, precision, published):
        ib.set_seed(777 + 1000003 * c)
        integ = LangevinSplittingIntegrator("O V R V O", temperature=298 * unit.kelvin, collision_rate=1.0 / unit.picoseconds,
                                            timestep=dt_fs * unit.femtosecond)
        sim = NonequilibriumSimulator(eq, integ, precision=precision)
        sim.collect_protocol_samples(args.n, args.protocol_length, marginal, as_numpy=False)
        m = sim.last_moments.cpu().numpy()
        dF, var = estimate_from_moments(m)
        emit({"event": "point", "cache": c, "marginal": marginal, "dt_fs": dt_fs, "precision": precision, "DeltaF_neq": dF,
              "stderr": float(np.sqrt(max(var, 0.0))), "n_crashed": int(m[6]), "published": published})

    double_dts = [float(v) for v in args.double_dts.split(",") if v]
    stop = False
    for c in range(caches.shape[0]):
        eq.set_samples(caches[c])
        for marginal in ("full", "configuration"):
            for i, dt_fs in enumerate(DTS):
                if time.time() - t_start > args.budget_s:
                    stop = True
                    break
                point(c, marginal, dt_fs, args.precision, PUB[marginal][i])
            if stop:
                break
        if stop:
            break
        for dt_fs in double_dts:
            if time.time() - t_start > args.budget_s:
                break
            point(c, "full", dt_fs, "double", PUB["full"][DTS.index(dt_fs)])
    log.close()
    summarise(args.out)


def summarise(path):
    rows = [json.loads(l) for l in open(path) if l.strip()]
    pts = [r for r in rows if r.get("event") == "point"]
    out = {"points": []}
    inside = stable = 0
    for precision in sorted({r["precision"] for r in pts}):
        for marginal in ("configuration", "full"):
            for i, dt_fs in enumerate(DTS):
                sel = [r for r in pts if r["precision"] == precision and r["marginal"] == marginal and abs(r["dt_fs"] - dt_fs) < 1e-9]
                if len(sel) < 3:
                    continue
                v = np.array([r["DeltaF_neq"] for r in sel])
                pub = PUB[marginal][i]
                crashed = max(r["n_crashed"] for r in sel)
                lo, hi = np.quantile(v, [0.025, 0.975])
                rank = int((v < pub).sum())
                rec = {"precision": precision, "marginal": marginal, "dt_fs": dt_fs, "n_caches": len(v), "published": pub,
                       "mean": float(v.mean()), "std_over_caches": float(v.std(ddof=1)), "mean_stderr_within": float(np.mean([r["stderr"] for r in sel])),
                       "min": float(v.min()), "max": float(v.max()), "q2.5": float(lo), "q97.5": float(hi), "rank_of_published": rank,
                       "inside_central_95": bool(lo <= pub <= hi), "max_crashed": int(crashed),
                       "z_of_published": float((pub - v.mean()) / v.std(ddof=1))}
                out["points"].append(rec)
                if precision != "double" and crashed < 10:
                    stable += 1
                    inside += int(rec["inside_central_95"])
    out["stable_points"], out["published_inside_central_95"] = stable, inside
    json.dump(out, open(path + ".summary.json", "w"), indent=1)
    print(json.dumps({"stable_points": stable, "published_inside_central_95": inside}))


if __name__ == "__main__":
    if len(sys.argv) > 2 and sys.argv[1] == "--summarise":
        summarise(sys.argv[2])
    else:
        main()

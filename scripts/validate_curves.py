#!/usr/bin/env python
"""Parity against the reference's PUBLISHED results:
  * DeltaF_neq(dt) of water_cluster_rigid, OVRVO, both marginals, 17 time steps, N = 1e5 protocols x 1000 steps
    (notebooks/near-eq result plotting.ipynb:1160-1195), and VRORV / configuration at 8 fs;
  * 1-D quartic D_KL(rho_x || pi_x) at dt = 1.1 (notes/note_on_dF_neq_approximation.tex:380).
Writes one JSON document; `z` is (ours - published) / sqrt(stderr_ours^2 + stderr_published^2)."""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

DTS = [0.1] + [0.5 * k for k in range(1, 17)]
PUB = {
    "configuration": [-2.49e-6, -7.84e-5, -2.42e-4, -4.28e-4, -5.26e-4, 9.17e-4, 1.84e-3, 6.93e-3, 1.65e-2, 2.95e-2, 5.39e-2, 8.48e-2,
                      0.1225, 0.1793, 0.2625, 0.3543, 0.5926],
    "full": [-3.76e-6, -5.50e-5, -2.67e-4, -4.76e-4, -2.01e-4, 1.56e-3, 4.89e-3, 9.93e-3, 2.01e-2, 4.30e-2, 6.59e-2, 0.1063, 0.1641,
             0.2342, 0.3483, 0.4648, 0.5778],
}
# the reference quotes its stderr at three time steps only (OVRVO: 4.1e-7 at 0.1 fs, 1.45e-3 at 4 fs, 6.1e-3 at 7 fs);
# its runs used the same N = 1e5 x 1000 steps, so elsewhere our own stderr stands in for the published one
PUB_STDERR_AT = {0.1: 4.1e-7, 4.0: 1.45e-3, 7.0: 6.1e-3}


def pub_stderr(dt, ours):
    return PUB_STDERR_AT.get(round(dt, 3), ours)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--n", type=int, default=100000)
    ap.add_argument("--protocol-length", type=int, default=1000)
    ap.add_argument("--n-cache", type=int, default=1000)
    ap.add_argument("--burn", type=int, default=100000, help="reference burn_in_length (x 10 steps)")
    ap.add_argument("--precision", default="mixed", choices=["mixed", "double"])
    ap.add_argument("--max-dt", type=float, default=100.0, help="only time steps up to this many fs")
    ap.add_argument("--skip-quartic", action="store_true")
    ap.add_argument("--cache-dt-fs", type=float, default=1.0, help="time step of the Langevin equilibration of the cache")
    ap.add_argument("--cache-seed", type=int, default=None, help="seed of the equilibrium cache (default: the global seed)")
    ap.add_argument("--out", default=os.path.join(ROOT, "gpurun_out", "validate_curves.json"))
    args = ap.parse_args()
    import integrator_benchmark_b200 as ib
    from integrator_benchmark_b200 import unit
    from integrator_benchmark_b200.evaluation import estimate_from_moments
    from integrator_benchmark_b200.integrators import LangevinSplittingIntegrator, baoab_factory, vvvr_factory
    from integrator_benchmark_b200.testsystems import NonequilibriumSimulator, NumbaNonequilibriumSimulator, quartic, water_cluster_rigid
    from functools import partial
    ib.engine.require_cuda()
    ib.set_seed(20261019)
    out = {"water_cluster_rigid": [], "quartic": [], "n": args.n, "protocol_length": args.protocol_length}

    eq = water_cluster_rigid
    eq.n_samples, eq.burn_in_length = args.n_cache, args.burn
    eq.timestep = args.cache_dt_fs * unit.femtosecond
    out["precision"], out["cache_dt_fs"] = args.precision, args.cache_dt_fs
    t0 = time.time()
    eq.collect_equilibrium_samples(seed=args.cache_seed)
    out["cache_seed"] = args.cache_seed
    out["equilibration_s"] = time.time() - t0
    out["cache"] = {"n": len(eq.samples), "steps_each": int(min(eq.burn_in_length * eq.steps_per_hmc, eq.max_equilibration_steps)) + 1000}
    t0 = time.time()
    runs = [("O V R V O", m, dt, PUB[m][i]) for m in ("configuration", "full") for i, dt in enumerate(DTS)]
    runs.append(("V R O R V", "configuration", 8.0, 0.038225))
    runs = [r for r in runs if r[2] <= args.max_dt]
    for splitting, marginal, dt, published in runs:
        integ = LangevinSplittingIntegrator(splitting, temperature=298 * unit.kelvin, collision_rate=1.0 / unit.picoseconds,
                                            timestep=dt * unit.femtosecond)
        sim = NonequilibriumSimulator(eq, integ, precision=args.precision)
        sim.collect_protocol_samples(args.n, args.protocol_length, marginal, as_numpy=False)
        m = sim.last_moments.cpu().numpy()
        dF, var = estimate_from_moments(m)
        se = float(np.sqrt(max(var, 0.0)))
        pse = 0.0058 if splitting == "V R O R V" else pub_stderr(dt, se)
        out["water_cluster_rigid"].append({"splitting": splitting, "marginal": marginal, "dt_fs": dt, "DeltaF_neq": dF, "stderr": se,
                                           "published": published, "published_stderr": pse, "n_finite": int(m[0]),
                                           "n_crashed": int(m[6]), "z": (dF - published) / float(np.hypot(se, pse))})
        print(out["water_cluster_rigid"][-1], flush=True)
    out["water_cluster_s"] = time.time() - t0

    # 1-D quartic, m = 10, beta = 1, gamma = 100, dt = 1.1: D_KL(rho_x || pi_x) from histograms in the reference's note
    for name, factory, published in (() if args.skip_quartic else (("VRORV", baoab_factory, 0.00013), ("OVRVO", vvvr_factory, 0.01309))):
        scheme = factory(quartic.potential, quartic.force, quartic.velocity_scale, quartic.mass)
        sim = NumbaNonequilibriumSimulator(quartic, partial(scheme, gamma=100.0, dt=1.1))
        W_F, W_R, _, _ = sim.collect_protocol_samples(2000000, 100, "configuration", traces=False)
        dF, var = estimate_from_moments(sim.last_moments.cpu().numpy())
        out["quartic"].append({"scheme": name, "dt": 1.1, "gamma": 100.0, "DeltaF_neq": dF, "stderr": float(np.sqrt(var)),
                               "published_D_KL": published})
        print(out["quartic"][-1], flush=True)
    stable = [r for r in out["water_cluster_rigid"] if r["n_crashed"] < 10]  # 8 fs is past the stability limit (crashes)
    zs = np.array([r["z"] for r in stable])
    out["summary"] = {"n_points": len(out["water_cluster_rigid"]), "n_stable": len(stable), "max_abs_z_stable": float(np.abs(zs).max()),
                      "n_within_2_sigma": int((np.abs(zs) < 2).sum()), "n_within_3_sigma": int((np.abs(zs) < 3).sum()),
                      "rms_z": float(np.sqrt(np.mean(zs ** 2)))}
    os.makedirs(os.path.dirname(args.out), exist_ok=True)
    json.dump(out, open(args.out, "w"), indent=1)
    print(json.dumps(out["summary"]))


if __name__ == "__main__":
    main()

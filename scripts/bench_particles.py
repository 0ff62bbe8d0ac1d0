#!/usr/bin/env python
"""Throughput of the particle-system protocol kernels (configs C3-C5) on one GPU.
Prints one JSON line per (system, precision)."""
import argparse
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
KT = 8.3144621e-3 * 298.0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--systems", default="water_cluster,alanine,waterbox")
    ap.add_argument("--precisions", default="mixed,double")
    ap.add_argument("--splitting", default="V R O R V")
    ap.add_argument("--steps", type=int, default=50, help="integrator steps per leg")
    ap.add_argument("--replicas", type=int, default=0)
    ap.add_argument("--reps", type=int, default=3)
    ap.add_argument("--dt", type=float, default=0.002)
    args = ap.parse_args()
    import torch
    import integrator_benchmark_b200 as ib
    from integrator_benchmark_b200.testsystems import systems as ts
    ib.engine.require_cuda()
    frames = np.load(os.path.join(ROOT, "tests", "golden", "waterbox_frames.npz"))["xyz"].astype(np.float64)
    table = {
        "water_cluster": (ts.water_cluster(20, constrained=True, seed=1), 148 * 64),
        "water_cluster_flexible": (ts.water_cluster(20, constrained=False, seed=1), 148 * 64),
        "alanine": (ts.alanine_dipeptide(constrained=True), 148 * 128),
        "alanine_unconstrained": (ts.alanine_dipeptide(constrained=False), 148 * 128),
        "waterbox": ((ts.tiny_waterbox()[0], frames), 148 * 16),
    }
    for name in args.systems.split(","):
        (desc, x), n_default = table[name]
        n = args.replicas or n_default
        cache = x if x.ndim == 3 else x[None]
        system = ib.engine.ParticleSystem(desc)
        prog = ib.engine.Program(args.splitting, args.dt if "flex" not in name and "uncon" not in name else args.dt / 4, 1.0)
        cache_dev = torch.as_tensor(cache).cuda()
        for prec in args.precisions.split(","):
            out = {}
            best = 1e30
            for rep in range(args.reps + 1):
                a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a.record()
                res = ib.engine.run_protocol(system, prog, n, args.steps, KT, seed=1, x_cache=cache_dev, precision=prec, out=out)
                b.record()
                b.synchronize()
                if rep > 0:
                    best = min(best, a.elapsed_time(b))
            W = res["W"].cpu().numpy()
            rs = n * 2 * args.steps / (best * 1e-3)
            print(json.dumps({"system": name, "precision": prec, "splitting": args.splitting, "replicas": n,
                              "steps_per_leg": args.steps, "ms": best, "replica_steps_per_s": rs,
                              "W_mean": [float(np.nanmean(W[0])), float(np.nanmean(W[1]))],
                              "nonfinite": int((~np.isfinite(W)).sum())}), flush=True)


if __name__ == "__main__":
    main()

#!/usr/bin/env python
"""Is the PUBLISHED DeltaF_neq(dt) curve of water_cluster_rigid / OVRVO one plausible draw from the distribution of curves that
independent reference-style equilibrium caches produce?  Reads the JSON lines written by scripts/c3_reference_caches.py and
tests the published vector at dt = 3, 4, 5, 6, 7 fs against the cache-to-cache distribution with Hotelling's T^2 (a new
observation against a sample of n curves: T^2 n (n - p) / (p (n - 1) (n + 1)) ~ F(p, n - p)), which accounts for the strong
correlation of a cache's effect across time steps (one cache moves a whole sweep together).
    python scripts/c3_published_vs_caches.py profiles/r2_c3_reference_caches.jsonl [more.jsonl ...]"""
import json
import os
import sys

import numpy as np
from scipy import stats

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from scripts.validate_curves import DTS, PUB  # noqa: E402

TEST_DTS = (3.0, 4.0, 5.0, 6.0, 7.0)


def load(path):
    rows = [json.loads(l) for l in open(path) if l.strip()]
    pts = [r for r in rows if r.get("event") == "point" and r["precision"] == "mixed"]
    out = {}
    for marg in ("full", "configuration"):
        n = 1 + max(r["cache"] for r in pts)
        M = np.full((n, len(DTS)), np.nan)
        for r in pts:
            if r["marginal"] == marg:
                M[r["cache"], DTS.index(r["dt_fs"])] = r["DeltaF_neq"]
        out[marg] = M
    return out


def main(paths):
    sets = [load(p) for p in paths]
    report = {"files": [os.path.basename(p) for p in paths], "test_dts_fs": TEST_DTS, "marginals": {}}
    for marg in ("full", "configuration"):
        sel = [DTS.index(d) for d in TEST_DTS]
        X = np.vstack([s[marg] for s in sets])[:, sel]
        X = X[~np.isnan(X).any(axis=1)]
        n, p = X.shape
        m, S = X.mean(axis=0), np.cov(X.T)
        pub = np.array(PUB[marg])[sel]
        T2 = float((pub - m) @ np.linalg.solve(S, pub - m))
        F = T2 * n * (n - p) / (p * (n - 1) * (n + 1))
        report["marginals"][marg] = {
            "n_caches": n, "published": pub.tolist(), "cache_mean": m.tolist(), "cache_std": np.sqrt(np.diag(S)).tolist(),
            "z_pointwise": ((pub - m) / np.sqrt(np.diag(S))).tolist(), "correlation_4fs_6fs": float(np.corrcoef(X.T)[1, 3]),
            "hotelling_T2": T2, "F": float(F), "p_value": float(1.0 - stats.f.cdf(F, p, n - p))}
    print(json.dumps(report, indent=1))


if __name__ == "__main__":
    main(sys.argv[1:])

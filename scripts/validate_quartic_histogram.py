"""Ground-truth check of the toy integrators against the reference's note (notes/note_on_dF_neq_approximation.tex:349-380):
quartic potential, m = 10, beta = 1, gamma = 100, dt = 1.1; D_KL(rho_x || pi_x) between the histogram of the Langevin steady
state and the exact equilibrium histogram on the same bins: VVVR (OVRVO) ~ 0.01309, BAOAB (VRORV) ~ 0.00013.  The steady state
is sampled as the final x of n independent replicas after `steps` steps from equilibrium (one launch)."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402

import integrator_benchmark_b200 as ib  # noqa: E402
from integrator_benchmark_b200.integrators import baoab_factory, vvvr_factory  # noqa: E402
from integrator_benchmark_b200.testsystems import quartic  # noqa: E402


def histogram_kl(x, edges):
    from math import fsum
    counts, _ = np.histogram(x, bins=edges)
    rho = counts / counts.sum()
    grid = np.linspace(edges[0], edges[-1], 200 * (len(edges) - 1) + 1)   # exact pi per bin by Simpson on a fine grid
    p = np.exp(-grid ** 4)
    per = p.reshape(-1)[:-1].reshape(len(edges) - 1, 200)
    h = grid[1] - grid[0]
    mass = np.array([(h / 3) * (p[200 * i] + p[200 * (i + 1)] + 4 * p[200 * i + 1:200 * (i + 1):2].sum()
                                + 2 * p[200 * i + 2:200 * (i + 1) - 1:2].sum()) for i in range(len(edges) - 1)])
    pi = mass / mass.sum()
    ok = rho > 0
    return fsum(rho[ok] * np.log(rho[ok] / pi[ok]))


def main():
    n, steps = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 24, 300
    ib.set_seed(99)
    edges = np.linspace(-2.5, 2.5, 101)
    out = {"n": n, "steps": steps, "bins": 100}
    for name, factory, published in (("VRORV", baoab_factory, 0.00013), ("OVRVO", vvvr_factory, 0.01309)):
        scheme = factory(quartic.potential, quartic.force, quartic.velocity_scale, quartic.mass)
        res = scheme.ensemble(n, steps + 1, 100.0, 1.1, n_legs=1, x_cache=quartic.x_samples_device(), want_final=True)
        x = res["x_final"].cpu().numpy()
        x = x[np.isfinite(x) & (np.abs(x) < 2.5)]
        out[name] = {"D_KL_histogram": histogram_kl(x, edges), "published": published, "samples": int(len(x)),
                     "histogram_bias_estimate": 99 / (2.0 * len(x))}
    cache = quartic.x_samples_device().cpu().numpy()
    out["equilibrium_cache"] = {"D_KL_histogram": histogram_kl(cache[np.abs(cache) < 2.5], edges), "samples": int(len(cache))}
    print(json.dumps(out))


if __name__ == "__main__":
    main()

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


def histogram_kl(x, edges, potential=lambda g: g ** 4):
    from math import fsum
    counts, _ = np.histogram(x, bins=edges)
    rho = counts / counts.sum()
    grid = np.linspace(edges[0], edges[-1], 200 * (len(edges) - 1) + 1)   # exact pi per bin by Simpson on a fine grid
    p = np.exp(-potential(grid))
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
    if len(sys.argv) > 2 and sys.argv[2] == "double_well":
        double_well_check(int(sys.argv[3]) if len(sys.argv) > 3 else 1 << 22)


def double_well_check(n, steps=20000, dt=0.7, gamma=10.0):
    """1-D double well, gamma = 10, configuration-space D_KL at dt ~ 0.7 read off the reference's figure
    (notebooks/conf_summary.jpg, code in "1D figures, updated.ipynb" cells 12-14, 100 bins on [-2.25, 2.25]):
    OVRVO ~ 0.23, ORVRO ~ 0.11, RVOVR ~ 0.08, VRORV ~ 1.5e-3."""
    from integrator_benchmark_b200.integrators import aboba_factory, orvro_factory
    from integrator_benchmark_b200.testsystems import double_well
    U = lambda g: g ** 6 + 2.0 * np.cos(5.0 * (g + 1.0))  # noqa: E731
    edges = np.linspace(-2.25, 2.25, 101)
    out = {"system": "double_well", "n": n, "steps": steps, "dt": dt, "gamma": gamma}
    for name, factory, published in (("OVRVO", vvvr_factory, 0.23), ("ORVRO", orvro_factory, 0.11), ("RVOVR", aboba_factory, 0.08),
                                     ("VRORV", baoab_factory, 1.5e-3)):
        scheme = factory(double_well.potential, double_well.force, double_well.velocity_scale, double_well.mass)
        res = scheme.ensemble(n, steps + 1, gamma, dt, n_legs=1, x_cache=double_well.x_samples_device(), want_final=True)
        x = res["x_final"].cpu().numpy()
        keep = np.isfinite(x) & (np.abs(x) < 2.25)
        out[name] = {"D_KL_histogram": histogram_kl(x[keep], edges, U), "figure_read_off": published, "kept": int(keep.sum())}
    cache = double_well.x_samples_device().cpu().numpy()
    out["equilibrium_cache"] = {"D_KL_histogram": histogram_kl(cache[np.abs(cache) < 2.25], edges, U), "samples": int(len(cache))}
    print(json.dumps(out))


if __name__ == "__main__":
    main()

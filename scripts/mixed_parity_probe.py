#!/usr/bin/env python
"""Observed differences between the engine's mixed precision and the oracle's mixed restatement (same Philox words) on
equilibrated configurations: sets the tolerances of tests/test_particles_gpu.py::test_mixed_precision_matches_mixed_oracle."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402

import integrator_benchmark_b200 as ib  # noqa: E402
from integrator_benchmark_b200.testsystems import systems as ts  # noqa: E402
from oracle import oracle as orc  # noqa: E402

KT = 8.3144621e-3 * 298.0
frames = np.load(os.path.join(ROOT, "tests", "golden", "waterbox_frames.npz"))["xyz"].astype(np.float64)
S = {"cluster_rigid": ts.water_cluster(20, constrained=True, seed=1), "cluster_flexible": ts.water_cluster(20, constrained=False, seed=2),
     "alanine_constrained": ts.alanine_dipeptide(constrained=True), "alanine_hmr": ts.alanine_dipeptide(constrained=True, mass_scale=3.0),
     "waterbox": (ts.tiny_waterbox()[0], frames)}
CASES = [("cluster_rigid", "O V R V O", 0.002), ("cluster_rigid", "V R O R V", 0.004), ("cluster_flexible", "V R O R V", 0.0005),
         ("alanine_constrained", "V R O R V", 0.002), ("alanine_constrained", "V0 V1 R O R V1 V0", 0.002),
         ("alanine_hmr", "O V R V O", 0.003), ("waterbox", "O V R V O", 0.002), ("waterbox", "V R O R V", 0.002)]


def equilibrated(desc, x, n, dt, seed):
    system = ib.engine.ParticleSystem(desc)
    x0 = x[np.arange(n) % len(x)] if x.ndim == 3 else np.repeat(x[None], n, axis=0)
    for k, (h, g, ns) in enumerate([(0.25 * dt, 50.0, 400), (dt, 5.0, 3000)]):
        res = ib.engine.run_protocol(system, ib.engine.Program("V R O R V", h, g), n, ns, KT, n_legs=1, seed=seed + k, x0=x0,
                                     want_final=True, want_moments=False, out={})
        x0 = res["x_final"].cpu().numpy()
    return x0


for name, splitting, dt in CASES:
    desc, x = S[name]
    n = 4
    x0 = equilibrated(desc, x, n, dt, 21)
    system = ib.engine.ParticleSystem(desc)
    for n_steps in (8, 20):
        prog = ib.engine.Program(splitting, dt, 1.0)
        res = ib.engine.run_protocol(system, prog, n, n_steps, KT, seed=11, replica0=300, x0=x0, precision="mixed", want_heat=True,
                                     want_final=True, out={})
        ref = orc.particles_protocol(desc, splitting, dt, 1.0, KT, n, n_steps, seed=11, replica0=300, x0=x0, mixed=True)
        ref64 = orc.particles_protocol(desc, splitting, dt, 1.0, KT, n, n_steps, seed=11, replica0=300, x0=x0)
        pe = orc.particles_energy_forces(desc, x0[0])[0]
        em, fm = system.energy_forces(x0, precision="mixed")
        e32, f32 = orc.particles_energy_forces(desc, x0[0], force_fp32=True)
        print(json.dumps({"case": [name, splitting, dt, n_steps], "PE": pe, "KE_thermal": 1.5 * desc["n_atoms"] * KT,
                          "dW_kJ": float(np.abs(res["W"].cpu().numpy().T - ref["W"]).max() * KT),
                          "dW_vs_f64_oracle_kJ": float(np.abs(res["W"].cpu().numpy().T - ref64["W"]).max() * KT),
                          "dQ_kJ": float(np.abs(res["heat"].cpu().numpy().T - ref["Q"]).max()),
                          "dx_nm": float(np.abs(res["x_final"].cpu().numpy() - ref["x_final"]).max()),
                          "dv": float(np.abs(res["v_final"].cpu().numpy() - ref["v_final"]).max()),
                          "dE_static_kJ": float(abs(em[0].item() - e32)), "dF_static_rel": float(np.abs(fm[0].cpu().numpy() - f32).max() / np.abs(f32).max())}), flush=True)

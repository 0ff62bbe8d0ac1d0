#!/usr/bin/env python
"""Times the six bench.py splittings (C2: double well, 2^20 replicas, protocol_length 10) for every liblsi variant in
integrator-benchmark_b200/lib/variants (built by scripts/build_variant.sh), one subprocess per variant.
    python scripts/bench_toy_variants.py [--child]     one JSON line per variant
"""
import glob
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
SPLITTINGS = ["O V R V O", "O R V R O", "R V O V R", "R O V O R", "V R O R V", "V O R O V"]


def child():
    import torch
    import integrator_benchmark_b200 as ib
    from integrator_benchmark_b200.testsystems import double_well, quartic
    if os.environ.get("TOY_CASE", "c2") == "c1":  # quartic, the reference's vvvr + baoab closure strings, protocol_length 100
        global SPLITTINGS
        SPLITTINGS = ["O V R R V O", "V R O O R V"]
        n, steps = 1 << 20, int(os.environ.get("TOY_STEPS", 99))
        system = ib.engine.ScalarSystem("quartic", 10.0)
        cache = torch.from_numpy(quartic.collect_equilibrium_samples(1000000, n_burn=100, seed=2)).cuda()
        progs = [ib.engine.Program(s, 1.1, 100.0) for s in SPLITTINGS]
    else:
        n, steps = 1 << 20, int(os.environ.get("TOY_STEPS", 9))
        system = ib.engine.ScalarSystem("double_well", 10.0)
        cache = torch.from_numpy(double_well.collect_equilibrium_samples(1000000, n_burn=100, seed=2)).cuda()
        progs = [ib.engine.Program(s, 0.35, 10.0) for s in SPLITTINGS]
    bufs = [{} for _ in SPLITTINGS]
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    for _ in range(20):
        for p, b in zip(progs, bufs):
            ib.engine.run_protocol(system, p, n, steps, 1.0, seed=1, x_cache=cache, out=b)
    torch.cuda.synchronize()
    reps = 30
    ms = [0.0] * len(progs)
    evs = []
    for _ in range(reps):
        for j, (p, b) in enumerate(zip(progs, bufs)):
            flush.zero_()
            a, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            ib.engine.run_protocol(system, p, n, steps, 1.0, seed=1, x_cache=cache, out=b)
            e.record()
            evs.append((j, a, e))
    torch.cuda.synchronize()
    for j, a, e in evs:
        ms[j] += a.elapsed_time(e) / reps
    W = bufs[0]["W"].double().sum().item()
    # the six programs as one batch launch (lsi_run_protocol_batch)
    mom = torch.zeros(len(progs), 8, dtype=torch.float64, device="cuda")
    for _ in range(10):
        ib.engine.run_protocols(system, progs, n, steps, 1.0, seed=1, x_cache=cache, outs=bufs, moments_out=mom)
    torch.cuda.synchronize()
    bev = []
    for _ in range(reps):
        flush.zero_()
        a, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        ib.engine.run_protocols(system, progs, n, steps, 1.0, seed=1, x_cache=cache, outs=bufs, moments_out=mom)
        e.record()
        bev.append((a, e))
    torch.cuda.synchronize()
    batch_ms = sum(a.elapsed_time(e) for a, e in bev) / reps
    Wb = bufs[0]["W"].double().sum().item()
    print(json.dumps({"variant": os.environ.get("LSI_VARIANT", "default"), "ms": [round(m, 4) for m in ms], "sum_ms": round(sum(ms), 4),
                      "replica_steps_per_s": len(progs) * n * 2 * steps / sum(ms) * 1e3, "checksum_W_OVRVO": W, "batch_ms": round(batch_ms, 4),
                      "batch_replica_steps_per_s": len(progs) * n * 2 * steps / batch_ms * 1e3, "batch_checksum_equal": W == Wb}))


if __name__ == "__main__":
    if "--child" in sys.argv:
        child()
    else:
        libs = [("default", None)] + [(os.path.basename(p)[7:-3], p) for p in sorted(glob.glob(os.path.join(ROOT, "integrator-benchmark_b200", "lib", "variants", "liblsi_*.so")))]
        for name, path in libs:
            env = dict(os.environ, LSI_VARIANT=name)
            if path:
                env["LSI_LIB_PATH"] = path
            r = subprocess.run([sys.executable, os.path.abspath(__file__), "--child"], env=env, capture_output=True, text=True)
            print(r.stdout.strip().split("\n")[-1] if r.returncode == 0 else json.dumps({"variant": name, "error": r.stderr[-400:]}), flush=True)

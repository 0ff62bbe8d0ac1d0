"""Throughput of the reference-facing toy call WITH traces (the reference's default output:
low_dimensional_systems.py:160-184 returns work traces (n, n_steps) and (x, v) traces (n, n_steps, 2)).

    python scripts/bench_toy_traces.py [n_replicas] [protocol_length]
"""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import integrator_benchmark_b200 as ib  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 20
T = int(sys.argv[2]) if len(sys.argv) > 2 else 10
system = ib.engine.ScalarSystem("double_well", 10.0)
import ctypes  # noqa: E402
cache = torch.empty(100000, dtype=torch.float64, device="cuda")
ib._cabi.check(ib._cabi.lib.lsi_sample_equilibrium_scalar(system.handle, 1.0, cache.numel(), 200, 1, 0, ctypes.c_void_p(cache.data_ptr()),
                                                          ib.engine.current_stream_ptr()))
out = {}
rows = {}
for label, kw in [("works only (fused)", dict(traces=False)),
                  ("traces, step-major (fused)", dict(traces=True)),
                  ("traces, step-major with (x, v) pairs (fused; what the shim returns views of)", dict(traces="pairs")),
                  ("traces, replica-major = reference layout (fused)", dict(traces="replica_major")),
                  ("traces, step-major, interpreter (mixed-precision noise)", dict(traces=True, precision="mixed"))]:
    prog = ib.engine.Program("O V R V O", 0.35, 10.0)
    bufs = {}
    for _ in range(3):
        ib.engine.run_protocol(system, prog, n, T - 1, 1.0, seed=1, x_cache=cache, out=bufs, **kw)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    reps = 20
    a.record()
    for _ in range(reps):
        ib.engine.run_protocol(system, prog, n, T - 1, 1.0, seed=1, x_cache=cache, out=bufs, **kw)
    b.record()
    b.synchronize()
    ms = a.elapsed_time(b) / reps
    nbytes = 0 if not kw.get("traces") else 2 * n * T * 24
    rows[label] = {"ms": ms, "replica_steps_per_s": n * 2 * (T - 1) / ms * 1e3, "trace_GB_per_s": nbytes / ms / 1e6}
print(json.dumps({"n": n, "protocol_length": T, "rows": rows}, indent=1))

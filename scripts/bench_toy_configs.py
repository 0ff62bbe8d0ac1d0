"""Device-resident throughput of the toy configurations of SURVEY 8d that bench.py does not time: C1 (quartic, 2^20 and 2^24
replicas x 100 steps, OVRVO / VRORV, gamma 100) next to C2's double well at the same length.  One JSON line per case."""
import ctypes
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import integrator_benchmark_b200 as ib  # noqa: E402

for kind, n, T, dt, gamma in (("quartic", 1 << 20, 100, 1.1, 100.0), ("quartic", 1 << 24, 100, 1.1, 100.0),
                              ("double_well", 1 << 20, 100, 0.35, 10.0), ("double_well", 1 << 20, 10, 0.35, 10.0)):
    system = ib.engine.ScalarSystem(kind, 10.0)
    cache = torch.empty(1000000, dtype=torch.float64, device="cuda")
    ib._cabi.check(ib._cabi.lib.lsi_sample_equilibrium_scalar(system.handle, 1.0, cache.numel(), 200, 1, 0,
                                                              ctypes.c_void_p(cache.data_ptr()), ib.engine.current_stream_ptr()))
    for splitting in ("O V R V O", "V R O R V"):
        prog = ib.engine.Program(splitting, dt, gamma)
        bufs = {}
        for _ in range(3):
            ib.engine.run_protocol(system, prog, n, T - 1, 1.0, seed=1, x_cache=cache, out=bufs)
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        reps = 10
        a.record()
        for _ in range(reps):
            res = ib.engine.run_protocol(system, prog, n, T - 1, 1.0, seed=1, x_cache=cache, out=bufs)
        b.record()
        b.synchronize()
        ms = a.elapsed_time(b) / reps
        dF, var = ib.evaluation.estimate_from_moments(res["moments"].cpu().numpy())
        print(json.dumps({"system": kind, "replicas": n, "protocol_length": T, "splitting": splitting, "dt": dt, "gamma": gamma,
                          "ms": ms, "replica_steps_per_s": n * 2 * (T - 1) / ms * 1e3, "DeltaF_neq": dF, "stderr": var ** 0.5}))

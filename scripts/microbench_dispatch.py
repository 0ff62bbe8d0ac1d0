#!/usr/bin/env python
"""Does integer work hide under fp64 instructions on a B200 sub-partition?  (lsi_microbench_fma, precision 10 + k)"""
import ctypes
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from integrator_benchmark_b200 import _cabi, engine  # noqa: E402

engine.require_cuda()
sink = torch.zeros(8, dtype=torch.float64, device="cuda")
blocks, iters = 148 * 8, 20000
base = None
for k in range(5):
    best = 1e30
    for _ in range(5):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        _cabi.check(_cabi.lib.lsi_microbench_fma(10 + k, iters, blocks, ctypes.c_void_p(sink.data_ptr()), engine.current_stream_ptr()))
        b.record()
        b.synchronize()
        best = min(best, a.elapsed_time(b))
    base = base or best
    print(json.dumps({"int_per_fp64": k, "ms": best, "relative_to_fp64_only": best / base,
                      "fp64_tflops": blocks * 256 * iters * 16 / (best * 1e-3) / 1e12,
                      "model_if_dispatch_blocked": 1 + k / 2.0, "model_if_hidden": max(1.0, (1 + k) / 2.0)}), flush=True)

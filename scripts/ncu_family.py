#!/usr/bin/env python
"""One batch launch of the six 5-letter splittings (config C2) for profiling: python scripts/ncu_family.py [n_replicas] [steps_per_leg] [reps]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import integrator_benchmark_b200 as ib  # noqa: E402
from integrator_benchmark_b200.testsystems import double_well  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 20
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 9
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 3
system = ib.engine.ScalarSystem("double_well", 10.0)
cache = torch.from_numpy(double_well.collect_equilibrium_samples(1000000, n_burn=100, seed=2)).cuda()
progs = [ib.engine.Program(s, 0.35, 10.0) for s in ["O V R V O", "O R V R O", "R V O V R", "R O V O R", "V R O R V", "V O R O V"]]
bufs = [{} for _ in progs]
mom = torch.zeros(6, 8, dtype=torch.float64, device="cuda")
for _ in range(reps):
    ib.engine.run_protocols(system, progs, n, steps, 1.0, seed=1, x_cache=cache, outs=bufs, moments_out=mom)
torch.cuda.synchronize()
print(mom[:, :3].cpu().numpy())

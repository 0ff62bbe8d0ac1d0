#!/usr/bin/env python
"""Headline benchmark: replica-steps/s of the fused Langevin-splitting protocol.

Workload (BASELINE.json configs[1]): 1-D double well
(m = 10, beta = 1, gamma = 10 as in the reference's six-scheme comparison, dt = 0.35), 1 048 576 replicas PER GPU, protocol_length 10
(the reference's convention: 9 integrator steps per leg, forward + reverse leg), all six
5-letter symmetric splittings over {O, R, V}; one bench "step" = one pass over the batch =
6 fused protocol launches + the DeltaF_neq moment reduction (all-reduced over ranks).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]

Prints ONE JSON line (rank 0).  `value` is device-resident throughput (CUDA events, max over
ranks); `e2e` goes through the public API with pinned HOST buffers (H2D of the equilibrium
cache, D2H of the works, every step); `roofline` holds the dominant kernel against the FP64
FMA peak measured in the same run; `cpu_baseline` is the CPU oracle port on the host cores.
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

SPLITTINGS = {"OVRVO": "O V R V O", "ORVRO": "O R V R O", "RVOVR": "R V O V R",
              "ROVOR": "R O V O R", "VRORV": "V R O R V", "VOROV": "V O R O V"}
SYSTEM, MASS, BETA, GAMMA, DT = "double_well", 10.0, 1.0, 10.0, 0.35
N_CACHE = 1_000_000
METRIC, UNIT = "replica-steps/sec (double well, six 5-letter splittings, F+R protocol)", "replica-steps/s"

# Algorithmic flops per replica-step (DESIGN.md "flop convention"): FMA = 2, add/mul = 1,
# transcendental = 1; Philox/Box-Muller work is NOT counted.  O: 9/DOF, V: 2/DOF, R: 2/DOF,
# one force evaluation (10 flops: x^5, sin, 2 FMA) per distinct position visited.
F_FORCE_DW = 10


def flops_per_step(splitting):
    toks = splitting.split()
    n_evals = 0
    valid = False  # force known at current position? steady state: last op of previous step
    seq = toks + toks  # second pass sees the steady-state validity
    for i, t in enumerate(seq):
        if t == "R":
            valid = False
        elif t == "V":
            if not valid and i >= len(toks):
                n_evals += 1
            valid = True
    return 9 * toks.count("O") + 2 * toks.count("V") + 2 * toks.count("R") + F_FORCE_DW * n_evals


class ClockSampler:
    """SM clock + clock-event reasons sampled DURING the timed region (NVML from a thread, every
    ~2 ms; the nvidia-smi -lms loop of the profiling recipe cannot resolve a region this short)."""

    def __init__(self, gpu_index=0, period_s=0.002):
        import threading
        self.rows, self.marks, self._stop = [], [], threading.Event()
        self.ok = False
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(gpu_index)
            self.max_mhz = float(pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM))
            self.ok = True
        except Exception:  # noqa: BLE001
            return
        self.period = period_s
        self.t = threading.Thread(target=self._run, daemon=True)
        self.t.start()

    def _run(self):
        nv = self.nv
        reasons_fn = getattr(nv, "nvmlDeviceGetCurrentClocksEventReasons", None) or nv.nvmlDeviceGetCurrentClocksThrottleReasons
        while not self._stop.is_set():
            try:
                self.rows.append((time.perf_counter(), float(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM)),
                                  int(reasons_fn(self.h))))
            except Exception:  # noqa: BLE001
                pass
            time.sleep(self.period)

    def mark(self):
        self.marks.append(time.perf_counter())

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0, "source": "nvml"}
        if not self.ok:
            return out
        self._stop.set()
        self.t.join(timeout=2)
        lo, hi = (self.marks[0], self.marks[-1]) if len(self.marks) >= 2 else (0.0, 1e30)
        rows = [r for r in self.rows if lo <= r[0] <= hi]
        out["sm_max_mhz"] = self.max_mhz
        if rows:
            out["sm_mhz"] = float(np.median([r[1] for r in rows]))
            out["samples"] = len(rows)
            bits = 0
            for r in rows:
                bits |= r[2]
            names = {0x8: "hw_slowdown", 0x40: "hw_thermal_slowdown", 0x20: "sw_thermal_slowdown", 0x4: "sw_power_cap"}
            out["reasons"] = [n for b, n in names.items() if bits & b]
        return out


def cpu_port_throughput(n_replicas, protocol_length, repeats=1):
    """CPU oracle port (OpenMP over replicas, all host cores) on the same workload."""
    from oracle import oracle
    oracle.set_threads()
    cache = oracle.sample_equilibrium_scalar(SYSTEM, 20000, 100, seed=1)
    steps = protocol_length - 1
    t0 = time.perf_counter()
    total = 0
    for _ in range(repeats):
        for name, s in SPLITTINGS.items():
            oracle.toy_protocol(SYSTEM, s, DT, GAMMA, n_replicas, steps, mass=MASS, kT=1.0 / BETA, seed=1, x_cache=cache)
            total += n_replicas * 2 * steps
    dt = time.perf_counter() - t0
    return total / dt, dt


def run_reference(args):
    """--impl reference: the reference's CPU implementation of the path.  The reference is pure
    Python (numba closures / OpenMM) and cannot travel to the GPU box, so this arm times the CPU
    oracle port (oracle/lsi_oracle.c, pinned to the reference's closures by tests/golden) with all
    host threads, each step a bounded sample of the workload."""
    rank = int(os.environ.get("RANK", 0))
    if rank != 0:
        return
    n_sample = args.ref_replicas
    cores = os.cpu_count()
    for _ in range(args.warmup):
        cpu_port_throughput(max(n_sample // 8, 1024), args.protocol_length)
    t0 = time.perf_counter()
    total = 0.0
    for _ in range(args.steps):
        v, dt = cpu_port_throughput(n_sample, args.protocol_length)
        total += v * dt
    wall = time.perf_counter() - t0
    value = total / wall
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * wall / args.steps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic", "impl": "reference",
        "config": workload_config(args, args.replicas),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port",
                         "sample": "%d replicas x 6 splittings x 2 legs x %d steps per bench step" %
                                   (n_sample, args.protocol_length - 1)},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


def workload_config(args, n_per_gpu):
    return {"workload": "C2 double well (BASELINE configs[1]): %d replicas/GPU x protocol_length %d "
                        "(%d steps/leg, 2 legs) x 6 splittings" % (n_per_gpu, args.protocol_length,
                                                                    args.protocol_length - 1),
            "system": SYSTEM, "mass": MASS, "beta": BETA, "gamma": GAMMA, "dt": DT, "marginal": "configuration",
            "splittings": list(SPLITTINGS), "replicas_per_gpu": n_per_gpu, "protocol_length": args.protocol_length,
            "equilibrium_cache": N_CACHE, "l2": "flushed between timed steps (256 MiB memset, untimed)",
            "parallelism": "replica shards, no data-path collective; 48-double moment all-reduce per step"}


# ------------------------------------------------------------------------------------------------
# configs[2..4]: water cluster, alanine dipeptide, tiny water box
KT_298 = 8.31446261815324e-3 * 298.0
PARTICLE_WORKLOADS = {
    # name: (config label, replicas/GPU, protocol_length, dt ps, splittings, algorithmic flops per replica-step)
    "water_cluster": ("C3 rigid TIP3P water cluster, 20 waters, no cutoff, restraint K=1, SETTLE (BASELINE configs[2])",
                      47360, 200, 0.002, {"OVRVO": "O V R V O", "VRORV": "V R O R V"}, {"OVRVO": 56200, "VRORV": 57400}),
    "alanine": ("C4 alanine dipeptide vacuum, H-bonds constrained, MTS V0/V1 (BASELINE configs[3])",
                59200, 200, 0.002, {"VRORV": "V R O R V", "MTS": "V0 V1 R O R V1 V1 R O R V1 V0"}, {"VRORV": 29000, "MTS": 59000}),
    "waterbox": ("C5 tiny water box, 102 TIP3P waters, 1.5 nm, cutoff 0.6 nm + reaction field, SETTLE (BASELINE configs[4])",
                 8192, 200, 0.002, {"OVRVO": "O V R V O", "VRORV": "V R O R V"}, {"OVRVO": 1.20e6, "VRORV": 1.20e6}),
}
HBM_BYTES_PER_STEP_STREAMING = {"waterbox": 19584.0}  # 2 x 306 x 32 B if the state were streamed every step


def particle_system(name):
    from integrator_benchmark_b200.testsystems import systems as ts
    if name == "water_cluster":
        return ts.water_cluster(20, constrained=True, seed=1)
    if name == "alanine":
        return ts.alanine_dipeptide(constrained=True)
    desc, _ = ts.tiny_waterbox(constrained=True)
    frames = np.load(os.path.join(ROOT, "tests", "golden", "waterbox_frames.npz"))["xyz"].astype(np.float64)
    return desc, frames


def particles_cpu_port(name, cache, n_replicas, steps_per_leg, dt, splittings, seed=1):
    """CPU oracle port (fp64, OpenMP over replicas, all host cores) on a bounded sample of the workload."""
    from oracle import oracle
    oracle.set_threads()
    desc, _ = particle_system(name)
    t0 = time.perf_counter()
    total = 0
    for s in splittings.values():
        oracle.particles_protocol(desc, s, dt, 1.0, KT_298, n_replicas, steps_per_leg, seed=seed, x_cache=cache)
        total += n_replicas * 2 * steps_per_leg
    secs = time.perf_counter() - t0
    return total / secs, secs


def run_particles(args):
    name = args.workload
    label, n_default, plen_default, dt, splittings, flops = PARTICLE_WORKLOADS[name]
    n = args.replicas or n_default
    steps_per_leg = args.protocol_length or plen_default
    metric = "replica-steps/sec (%s, %s, F+R protocol)" % (name, "+".join(splittings))
    precision = args.precision or "mixed"
    cfg = {"workload": "%s: %d replicas/GPU x protocol_length %d (2 legs) x %d splittings" % (label, n, steps_per_leg, len(splittings)),
           "system": name, "temperature_K": 298.0, "collision_rate_per_ps": 1.0, "dt_ps": dt, "marginal": "configuration",
           "splittings": list(splittings), "replicas_per_gpu": n, "protocol_length": steps_per_leg, "precision": precision,
           "l2": "flushed between timed steps (256 MiB memset, untimed); state is shared-memory resident",
           "parallelism": "replica shards, no data-path collective; moment all-reduce per step"}
    rank = int(os.environ.get("RANK", 0))
    if args.impl == "reference":
        if rank != 0:
            return
        desc, x = particle_system(name)
        cache = x if x.ndim == 3 else x[None]
        n_ref = args.ref_replicas or {"water_cluster": 256, "alanine": 1024, "waterbox": 32}[name]
        steps_ref = min(steps_per_leg, 50)
        for _ in range(args.warmup):
            particles_cpu_port(name, cache, max(n_ref // 8, 8), 5, dt, splittings)
        t0 = time.perf_counter()
        tot = 0.0
        for _ in range(args.steps):
            v, secs = particles_cpu_port(name, cache, n_ref, steps_ref, dt, splittings)
            tot += v * secs
        wall = time.perf_counter() - t0
        value = tot / wall
        print(json.dumps({"metric": metric, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
                          "warmup": args.warmup, "ms_per_step": 1e3 * wall / args.steps, "higher_is_better": True, "scaling": "weak",
                          "vs_baseline": None, "dtype": "f64", "data": "synthetic", "impl": "reference", "config": cfg,
                          "cpu_baseline": {"value": value, "unit": UNIT, "cores": os.cpu_count(), "kind": "port",
                                           "sample": "%d replicas x %d splittings x 2 legs x %d steps per bench step (restated CPU "
                                                     "oracle, OpenMP; not OpenMM)" % (n_ref, len(splittings), steps_ref)},
                          "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}, "gpu_launches": 0}))
        return

    import ctypes
    import torch
    import torch.distributed as dist
    import integrator_benchmark_b200 as ib
    from integrator_benchmark_b200 import _cabi, engine, parallel
    rank, world, local_rank = parallel.init_process_group()
    engine.require_cuda()
    dev = torch.device("cuda", local_rank)
    torch.cuda.set_device(dev)
    seed = 0x5EED0003
    replica0 = rank * n
    desc, x = particle_system(name)
    system = engine.ParticleSystem(desc)
    # equilibrium cache (input of the path): the reference's own frames for the box; otherwise 1024
    # configurations equilibrated on the GPU from the start geometry (same on every rank)
    # (the box: the reference's 4 fixture frames tiled, then equilibrated for the reaction-field potential)
    n_eq = 1024 if x.ndim == 2 else 256
    if True:
        x0 = torch.as_tensor(np.ascontiguousarray(x if x.ndim == 3 else x[None]))
        x0 = x0[torch.arange(n_eq) % x0.shape[0]].contiguous().to(dev)
        for k, (h, gamma, ns) in enumerate([(0.0002, 100.0, 500), (0.001, 20.0, 500), (0.002, 5.0, 2000)]):
            res = engine.run_protocol(system, engine.Program("V R O R V", h, gamma), n_eq, ns, KT_298, n_legs=1, seed=seed + k,
                                      x0=x0, want_final=True, want_moments=False, out={})
            x0 = res["x_final"]
        cache_host = x0.cpu()
    cache_host = cache_host.pin_memory()
    cache = cache_host.to(dev)
    programs = {k: engine.Program(s, dt, 1.0) for k, s in splittings.items()}
    bufs = {k: {} for k in splittings}
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    moments_all = torch.zeros(len(splittings), 8, dtype=torch.float64, device=dev)

    def one_pass(cache_dev):
        for j, (k, prog) in enumerate(programs.items()):
            res = engine.run_protocol(system, prog, n, steps_per_leg, KT_298, marginal="configuration", seed=seed, replica0=replica0,
                                      x_cache=cache_dev, precision=precision, out=bufs[k])
            moments_all[j].copy_(res["moments"], non_blocking=True)
        parallel.all_reduce_moments(moments_all)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(args.warmup):
        one_pass(cache)
    barrier()
    sampler = ClockSampler(local_rank) if rank == 0 else None
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    barrier()
    if sampler:
        sampler.mark()
    for s_ in range(args.steps):
        flush.zero_()
        ev[s_][0].record()
        one_pass(cache)
        ev[s_][1].record()
    barrier()
    if sampler:
        sampler.mark()
    dev_ms = sum(a.elapsed_time(b) for a, b in ev)
    clocks = sampler.stop() if sampler else None
    tmax = torch.tensor([dev_ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
    dev_ms = float(tmax.item())
    steps_per_pass = len(splittings) * n * 2 * steps_per_leg
    value = world * steps_per_pass * args.steps / (dev_ms * 1e-3)

    # dominant kernel alone
    k_ms = {k: 0.0 for k in splittings}
    for _ in range(max(1, args.steps // 2)):
        for k, prog in programs.items():
            flush.zero_()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            engine.run_protocol(system, prog, n, steps_per_leg, KT_298, seed=seed, replica0=replica0, x_cache=cache,
                                precision=precision, out=bufs[k], want_moments=False)
            b.record()
            b.synchronize()
            k_ms[k] += a.elapsed_time(b)
    kernel_ms = {k: v / max(1, args.steps // 2) for k, v in k_ms.items()}

    sink = torch.zeros(8, dtype=torch.float64, device=dev)
    peaks = {}
    for prec, pname in [(0, "fp64"), (1, "fp32")]:
        blocks, iters = 148 * 8, 20000 if prec == 0 else 40000
        best = 1e30
        for _ in range(6):
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            _cabi.check(_cabi.lib.lsi_microbench_fma(prec, iters, blocks, ctypes.c_void_p(sink.data_ptr()), engine.current_stream_ptr()))
            b.record()
            b.synchronize()
            best = min(best, a.elapsed_time(b))
        peaks[pname] = blocks * 256 * iters * 16 / (best * 1e-3) / 1e12

    # end to end with host buffers: H2D of the cache, D2H of the works and moments, every step
    w_host = torch.empty(len(splittings), 2, n, dtype=torch.float64).pin_memory()
    m_host = torch.empty(len(splittings), 8, dtype=torch.float64).pin_memory()
    cache2 = torch.empty_like(cache)

    def e2e_pass(copy_works=True):
        cache2.copy_(cache_host, non_blocking=True)
        for j, (k, prog) in enumerate(programs.items()):
            res = engine.run_protocol(system, prog, n, steps_per_leg, KT_298, seed=seed, replica0=replica0, x_cache=cache2,
                                      precision=precision, out=bufs[k])
            if copy_works:
                w_host[j].copy_(res["W"], non_blocking=True)
            moments_all[j].copy_(res["moments"], non_blocking=True)
        parallel.all_reduce_moments(moments_all)
        m_host.copy_(moments_all, non_blocking=True)
        torch.cuda.current_stream().synchronize()
        return [ib.evaluation.estimate_from_moments(m_host[j].numpy()) for j in range(len(splittings))]

    e2e_pass()
    barrier()
    t0 = time.perf_counter()
    e2e_steps = max(2, args.steps // 2)
    for _ in range(e2e_steps):
        estimates = e2e_pass()
    barrier()
    e2e_s = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(e2e_s, op=dist.ReduceOp.MAX)
    e2e_value = world * steps_per_pass * e2e_steps / float(e2e_s.item())
    # the same call for a caller that keeps the works on the device and reads back only the estimate's moments
    for _ in range(2):
        e2e_pass(copy_works=False)
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        e2e_pass(copy_works=False)
    barrier()
    e2e_m = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(e2e_m, op=dist.ReduceOp.MAX)
    e2e_estimate_only = world * steps_per_pass * e2e_steps / float(e2e_m.item())
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    kernel_s = sum(kernel_ms.values()) * 1e-3
    alg = sum(flops[k] * n * 2 * steps_per_leg for k in splittings)
    achieved = alg / kernel_s / 1e12
    peak = peaks["fp32"] if precision == "mixed" else peaks["fp64"]
    roofline = {"bound": "fp32_fma" if precision == "mixed" else "fp64_fma", "kernel": "particles_protocol_kernel<%s>" % name,
                "achieved": achieved, "peak": peak, "unit": "TFLOP/s", "frac": achieved / peak, "traffic": None,
                "peak_source": "lsi_microbench_fma, best of 6, measured in this run (fp32 %.1f, fp64 %.1f TFLOP/s)" % (peaks["fp32"], peaks["fp64"]),
                "note": "algorithmic flops per replica-step from SURVEY 8d (half-shell pair count; FMA = 2); the kernel runs a "
                        "full-shell loop (twice the pair evaluations, no atomics) with fp64 constraints and bookkeeping",
                "replica_steps_per_s_kernel_only": len(splittings) * n * 2 * steps_per_leg / kernel_s,
                "kernel_ms_per_splitting": kernel_ms, "flops_per_replica_step": flops}
    if name in HBM_BYTES_PER_STEP_STREAMING:
        peaks_file = {}
        try:
            peaks_file = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:  # noqa: BLE001
            pass
        hbm_peak = peaks_file.get("hbm_gbs", 6650.0)
        eq = HBM_BYTES_PER_STEP_STREAMING[name] * roofline["replica_steps_per_s_kernel_only"] / 1e9
        roofline["hbm_equivalent"] = {"achieved": eq, "peak": hbm_peak, "unit": "GB/s", "frac": eq / hbm_peak,
                                      "note": "bytes a per-step-streaming design would move (19.6 KB per replica-step); this "
                                              "kernel keeps the state in shared memory and moves 7.4 KB per replica-PROTOCOL",
                                      "peak_source": "MEASURED_PEAKS.json" if peaks_file else "fallback"}
    cpu = None
    if world == 1 and not args.no_cpu_baseline:
        n_ref = args.ref_replicas or {"water_cluster": 512, "alanine": 2048, "waterbox": 48}[name]
        v, secs = particles_cpu_port(name, cache_host.numpy(), n_ref, min(steps_per_leg, 50), dt, splittings)
        cpu = {"value": v, "unit": UNIT, "cores": os.cpu_count(), "kind": "port",
               "sample": "%d replicas x %d splittings x 2 legs x %d steps (%.1f s; restated fp64 CPU oracle, OpenMP over replicas; not "
                         "OpenMM -- the reference's own figure is 1.28e4 replica-steps/s on the OpenMM Reference platform for alanine "
                         "dipeptide, BASELINE.md)" % (n_ref, len(splittings), min(steps_per_leg, 50), secs)}
    print(json.dumps({
        "metric": metric, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": dev_ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32 forces / f64 state" if precision == "mixed" else "f64", "data": "synthetic", "config": cfg, "clocks": clocks,
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": int(cache_host.numel() * 8),
                "d2h_bytes_per_step": int(w_host.numel() * 8 + m_host.numel() * 8), "steps": e2e_steps,
                "estimate_only": {"value": e2e_estimate_only, "unit": UNIT, "d2h_bytes_per_step": int(m_host.numel() * 8),
                                  "note": "works stay on the device (as_numpy=False); only the moments of the estimate are read back"}},
        "gpu_launches": args.steps * len(splittings) * 3, "roofline": roofline, "cpu_baseline": cpu,
        "DeltaF_neq": {k: {"estimate": e[0], "stderr": float(np.sqrt(max(e[1], 0.0)))} for k, e in zip(splittings, estimates)}}))
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=0, help="0 = 200 (double_well) / 4 (particle workloads)")
    ap.add_argument("--warmup", type=int, default=0, help="0 = 10 (double_well) / 3")
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="double_well", choices=["double_well"] + sorted(PARTICLE_WORKLOADS),
                    help="double_well = BASELINE configs[1] (the headline); the others are configs[2..4]")
    ap.add_argument("--replicas", type=int, default=0, help="replicas per GPU (0 = the workload's default)")
    ap.add_argument("--protocol-length", type=int, default=0, help="0 = the workload's default")
    ap.add_argument("--ref-replicas", type=int, default=0)
    ap.add_argument("--precision", default=None, choices=[None, "double", "mixed"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    if args.impl == "ours" and args.warmup:
        args.warmup = max(args.warmup, 3)  # timing rule: at least three warm-up steps
    if args.workload != "double_well":
        args.steps, args.warmup = args.steps or 4, args.warmup or 3
        return run_particles(args)
    args.steps, args.warmup = args.steps or 200, args.warmup or 10
    args.replicas = args.replicas or (1 << 20)
    args.protocol_length = args.protocol_length or 10
    args.ref_replicas = args.ref_replicas or (1 << 17)

    if args.impl == "reference":
        return run_reference(args)

    import torch
    import torch.distributed as dist

    import integrator_benchmark_b200 as ib
    from integrator_benchmark_b200 import _cabi, engine, parallel
    import ctypes

    rank, world, local_rank = parallel.init_process_group()
    engine.require_cuda()
    dev = torch.device("cuda", local_rank)
    torch.cuda.set_device(dev)
    n = args.replicas
    steps_per_leg = args.protocol_length - 1
    replica0 = rank * n  # global replica ids: rank r owns [r n, (r + 1) n)
    seed = 0x5EED0002

    from integrator_benchmark_b200.testsystems import double_well
    system = engine.ScalarSystem(SYSTEM, MASS)
    programs = {k: engine.Program(s, DT, GAMMA) for k, s in SPLITTINGS.items()}
    # equilibrium cache (input of the path, generated once on the GPU, same on every rank)
    x_cache_host = torch.from_numpy(double_well.collect_equilibrium_samples(N_CACHE, n_burn=300, seed=seed)).pin_memory()
    x_cache = x_cache_host.to(dev)
    bufs = {k: {} for k in SPLITTINGS}
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    moments_all = torch.zeros(len(SPLITTINGS), 8, dtype=torch.float64, device=dev)

    toy_precision = args.precision or "double"

    def one_pass(x_cache_dev):
        for j, (k, prog) in enumerate(programs.items()):
            res = engine.run_protocol(system, prog, n, steps_per_leg, 1.0 / BETA, marginal="configuration", seed=seed,
                                      replica0=replica0, x_cache=x_cache_dev, out=bufs[k], precision=toy_precision)
            moments_all[j].copy_(res["moments"], non_blocking=True)
        parallel.all_reduce_moments(moments_all)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(args.warmup):
        one_pass(x_cache)
    barrier()
    # a step is ~1.5 ms: keep warming (untimed) for at least half a second so that clocks and the
    # allocator have settled on a fresh box before anything is timed
    t_warm = time.perf_counter()
    while time.perf_counter() - t_warm < 0.5:
        for _ in range(20):
            one_pass(x_cache)
        torch.cuda.synchronize()
    barrier()

    # ---- device-resident timing: K steps, per-step CUDA events, L2 flushed (untimed) in between
    sampler = ClockSampler(local_rank) if rank == 0 else None
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    barrier()
    if sampler:
        sampler.mark()
    t_wall0 = time.perf_counter()
    for s in range(args.steps):
        flush.zero_()
        ev[s][0].record()
        one_pass(x_cache)
        ev[s][1].record()
    barrier()
    t_wall = time.perf_counter() - t_wall0
    if sampler:
        sampler.mark()
    dev_ms = sum(a.elapsed_time(b) for a, b in ev)
    clocks = sampler.stop() if sampler else None
    tmax = torch.tensor([dev_ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
    dev_ms = float(tmax.item())
    steps_per_pass = len(SPLITTINGS) * n * 2 * steps_per_leg  # replica-steps per rank per bench step
    value = world * steps_per_pass * args.steps / (dev_ms * 1e-3)

    # ---- dominant kernel alone: events around each protocol launch (kernel + 1-block finalize), L2 flushed before
    # each one; everything is enqueued first and synchronised once, so the GPU does not idle between launches
    kev = {k: [] for k in SPLITTINGS}
    for s in range(args.steps):
        for j, (k, prog) in enumerate(programs.items()):
            flush.zero_()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            engine.run_protocol(system, prog, n, steps_per_leg, 1.0 / BETA, marginal="configuration", seed=seed,
                                replica0=replica0, x_cache=x_cache, out=bufs[k], precision=toy_precision)
            b.record()
            kev[k].append((a, b))
    torch.cuda.synchronize()
    kernel_ms = {k: sum(a.elapsed_time(b) for a, b in v) / args.steps for k, v in kev.items()}

    # ---- measured FMA peaks (same run, burst = best of 5)
    sink = torch.zeros(8, dtype=torch.float64, device=dev)
    peaks = {}
    for prec, name in [(0, "fp64"), (1, "fp32")]:
        blocks, iters = 148 * 8, 20000 if prec == 0 else 40000
        best = 1e30
        for _ in range(6):
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            _cabi.check(_cabi.lib.lsi_microbench_fma(prec, iters, blocks, ctypes.c_void_p(sink.data_ptr()),
                                                     engine.current_stream_ptr()))
            b.record()
            b.synchronize()
            best = min(best, a.elapsed_time(b))
        peaks[name] = blocks * 256 * iters * 16 / (best * 1e-3) / 1e12  # TFLOP/s

    # ---- end to end through the public API with HOST buffers (pinned), every step:
    #      H2D of the equilibrium cache, 6 launches, D2H of both work arrays per splitting
    w_host = torch.empty(len(SPLITTINGS), 2, n, dtype=torch.float64).pin_memory()
    m_host = torch.empty(len(SPLITTINGS), 8, dtype=torch.float64).pin_memory()
    x_dev = torch.empty_like(x_cache)

    # D2H of the works rides a copy stream so that it overlaps the next splitting's kernel (PCIe, not the SMs,
    # bounds this loop: 100 MB per step); events order kernel -> copy and copy -> next reuse of the buffer
    copy_stream = torch.cuda.Stream(device=dev)
    done_k = [torch.cuda.Event() for _ in SPLITTINGS]
    done_c = [torch.cuda.Event() for _ in SPLITTINGS]
    for e in done_c:
        e.record(copy_stream)

    def e2e_pass(copy_works=True):
        main = torch.cuda.current_stream()
        x_dev.copy_(x_cache_host, non_blocking=True)
        for j, (k, prog) in enumerate(programs.items()):
            main.wait_event(done_c[j])  # the previous pass's copy of this buffer has left
            res = engine.run_protocol(system, prog, n, steps_per_leg, 1.0 / BETA, marginal="configuration", seed=seed,
                                      replica0=replica0, x_cache=x_dev, out=bufs[k], precision=toy_precision)
            moments_all[j].copy_(res["moments"], non_blocking=True)
            if copy_works:
                done_k[j].record(main)
                with torch.cuda.stream(copy_stream):
                    copy_stream.wait_event(done_k[j])
                    w_host[j].copy_(res["W"], non_blocking=True)
                    done_c[j].record(copy_stream)
        parallel.all_reduce_moments(moments_all)
        m_host.copy_(moments_all, non_blocking=True)
        main.synchronize()
        copy_stream.synchronize()
        return [ib.evaluation.estimate_from_moments(m_host[j].numpy()) for j in range(len(SPLITTINGS))]

    for _ in range(3):
        e2e_pass()
    barrier()
    t0 = time.perf_counter()
    e2e_steps = max(3, args.steps // 2)
    for _ in range(e2e_steps):
        estimates = e2e_pass()
    barrier()
    e2e_s = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(e2e_s, op=dist.ReduceOp.MAX)
    e2e_value = world * steps_per_pass * e2e_steps / float(e2e_s.item())
    # the same call for a caller that keeps the works on the device and reads back only the estimate's moments
    for _ in range(2):
        e2e_pass(copy_works=False)
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        e2e_pass(copy_works=False)
    barrier()
    e2e_m = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(e2e_m, op=dist.ReduceOp.MAX)
    e2e_estimate_only = world * steps_per_pass * e2e_steps / float(e2e_m.item())

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    alg_flops = sum(flops_per_step(s) * n * 2 * steps_per_leg for s in SPLITTINGS.values())
    kernel_s = sum(kernel_ms.values()) * 1e-3
    achieved = alg_flops / kernel_s / 1e12
    hbm_bytes = len(SPLITTINGS) * n * (8 + 16)  # per pass: gather x0 (8 B) + write W_F, W_R (16 B) per replica
    peaks_file = {}
    try:
        peaks_file = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:  # noqa: BLE001
        pass
    hbm_peak = peaks_file.get("hbm_gbs", 6650.0)
    roofline = {
        "bound": "fp64_fma", "kernel": "toy_fused_kernel<double_well, splitting>", "achieved": achieved, "peak": peaks["fp64"],
        "unit": "TFLOP/s", "frac": achieved / peaks["fp64"],
        # dram__bytes_read.sum + dram__bytes_write.sum per launch of toy_fused_kernel, ncu --set full
        # (profiles/r1_toy_fused_ncu_details.txt): the 8 MB equilibrium cache is read once, the works stay in L2
        "traffic": 8.04e6 if n == (1 << 20) else None,
        "peak_source": "lsi_microbench_fma fp64, best of 6, measured in this run (fp32: %.1f TFLOP/s)" % peaks["fp32"],
        "note": "algorithmic flops exclude the Philox + Box-Muller noise generation that dominates the instruction count "
                "(n_O Gaussians per replica-step); ncu (profiles/r1_ncu_summary.json): issue slots busy + FP64 pipe busy / 2 "
                "= 86-91 % of the dispatch limit of the 16-lane fp64 sub-partitions",
        "replica_steps_per_s_kernel_only": len(SPLITTINGS) * n * 2 * steps_per_leg / kernel_s,
        "kernel_ms_per_splitting": kernel_ms,
        "flops_per_replica_step": {k: flops_per_step(s) for k, s in SPLITTINGS.items()},
        "hbm": {"achieved": hbm_bytes / kernel_s / 1e9, "peak": hbm_peak, "unit": "GB/s",
                "frac": hbm_bytes / kernel_s / 1e9 / hbm_peak,
                "peak_source": "MEASURED_PEAKS.json" if peaks_file else "fallback"},
    }
    cpu = None
    if world == 1 and not args.no_cpu_baseline:
        _, warm = cpu_port_throughput(args.ref_replicas * 2, args.protocol_length)        # also warms the thread pool
        repeats = int(min(64, max(1, round(10.0 / max(warm, 1e-3)))))                     # about 10 s of CPU work
        v, secs = cpu_port_throughput(args.ref_replicas * 2, args.protocol_length, repeats=repeats)
        cpu = {"value": v, "unit": UNIT, "cores": os.cpu_count(), "kind": "port",
               "sample": "%d replicas x 6 splittings x %d passes (%.1f s, OpenMP over replicas)"
                         % (args.ref_replicas * 2, repeats, secs)}
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": dev_ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f64" if toy_precision == "double" else "f64 state, f32-generated Gaussians", "data": "synthetic",
        "config": workload_config(args, n),
        "wall_ms_per_step_incl_flush": 1e3 * t_wall / args.steps,
        "clocks": clocks,
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": int(x_cache_host.numel() * 8),
                "d2h_bytes_per_step": int(w_host.numel() * 8 + m_host.numel() * 8), "steps": e2e_steps,
                "estimate_only": {"value": e2e_estimate_only, "unit": UNIT, "d2h_bytes_per_step": int(m_host.numel() * 8),
                                  "note": "works stay on the device (as_numpy=False); only the moments of the estimate are read back"}},
        "gpu_launches": args.steps * len(SPLITTINGS) * 2,
        "roofline": roofline, "cpu_baseline": cpu,
        "DeltaF_neq": {k: {"estimate": e[0], "stderr": float(np.sqrt(max(e[1], 0.0)))} for k, e in zip(SPLITTINGS, estimates)},
    }
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()

#!/usr/bin/env python
"""Benchmark of the replica-ensemble / shadow-work path: replica-steps/s per (system, splitting).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--workload NAME]

Default (`--workload all`): ONE JSON line whose top level is the headline, BASELINE.json configs[1] (C2: 1-D double well,
1 048 576 replicas per GPU, protocol_length 10, the six 5-letter symmetric splittings over {O, R, V}), and whose
`per_workload` key holds the same record (value / e2e / roofline / cpu_baseline) for the other four configs, each on a
short fixed number of steps: C1 quartic (the reference's vvvr / baoab closures), C3 rigid water cluster, C4 alanine
dipeptide (VRORV + an MTS string), C5 tiny water box.  `--workload NAME` measures one of them alone as the top level.

One bench "step" = one pass over the batch = the fused protocol (forward leg, midpoint velocity redraw, reverse leg) of
every splitting -- scalar systems: all splittings in ONE launch (lsi_run_protocol_batch), particle systems: one launch per
splitting -- + the DeltaF_neq moment reduction, all-reduced over ranks.
  value     device-resident throughput: inputs in HBM, CUDA events around every step, L2 flushed (untimed) between steps,
            max over ranks.  The moment vectors of 8 consecutive steps travel in one all-reduce that overlaps the following
            steps (moments are additive; nothing downstream needs them earlier); the last ones are drained inside the
            timed region.
  e2e       the same work through the reference-facing drop-in call with DEFAULT arguments, as a reference caller makes it
            (NumbaNonequilibriumSimulator.collect_protocol_samples / NonequilibriumSimulator.collect_protocol_samples):
            host-side equilibrium cache uploaded every step, numpy results on the host (the toy call returns the full
            work and (x, v) traces, as the reference does).  The timed region ends when the call's results are on the host;
            estimates computed from them afterwards are reported, not timed (the reference arm does not time them either).
            Sub-keys give the same call with `traces=False` (final works only) and `as_numpy=False` (estimate only: the
            moment vector is read back).
  roofline  dominant kernel alone (events around each launch) against the FMA peak measured in the same run.
  cpu_baseline / --impl reference: the reference's own numba closures (oracle/_ref, installed by oracle/make_ref.sh)
            under the restated collect_protocol_samples loop on all host cores for C1 / C2; the restated C oracle
            (OpenMP) for C3-C5, where the reference needs OpenMM.
"""
import argparse
import ctypes
import json
import os
import sys
import time
from functools import partial

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

UNIT = "replica-steps/s"
KT_298 = 8.31446261815324e-3 * 298.0
FIVE_LETTER = {"OVRVO": "O V R V O", "ORVRO": "O R V R O", "RVOVR": "R V O V R",
               "ROVOR": "R O V O R", "VRORV": "V R O R V", "VOROV": "V O R O V"}
# the four closures of the reference's numba file and the substeps they perform (integrators/numba_integrators.py)
REF_SCHEMES = {"OVRVO": "vvvr", "ORVRO": "orvro", "VRORV": "baoab", "RVOVR": "aboba"}

WORKLOADS = {
    # toy workloads: protocol_length counts recorded states (reference convention): protocol_length - 1 steps per leg
    "double_well": dict(
        config="C2", kind="toy", system="double_well", mass=10.0, kT=1.0, gamma=10.0, dt=0.35, replicas=1 << 20, protocol_length=10,
        splittings=FIVE_LETTER, f_force=10, n_cache=1_000_000, e2e_replicas=1 << 20, ref_replicas=1 << 16, precision="double",
        label="C2 double well (BASELINE configs[1]), six 5-letter symmetric splittings"),
    "quartic": dict(
        config="C1", kind="toy", system="quartic", mass=10.0, kT=1.0, gamma=100.0, dt=1.1, replicas=1 << 20, protocol_length=100,
        splittings={"OVRVO": "O V R R V O", "VRORV": "V R O O R V"}, closures={"OVRVO": "vvvr", "VRORV": "baoab"}, f_force=3,
        n_cache=1_000_000, e2e_replicas=1 << 17, ref_replicas=1 << 13, precision="double",
        label="C1 1-D quartic (BASELINE configs[0]), OVRVO vs VRORV as the reference's vvvr / baoab closures perform them"),
    # particle workloads: protocol_length = steps per leg (NonequilibriumSimulator.collect_protocol_samples)
    "water_cluster": dict(
        config="C3", kind="particles", replicas=47360, protocol_length=200, dt=0.002, gamma=1.0, kT=KT_298,
        splittings={"OVRVO": "O V R V O", "VRORV": "V R O R V"}, flops={"OVRVO": 56200, "VRORV": 57400}, precision="mixed",
        ref_replicas=256, cpu_replicas=512, e2e_replicas=47360,
        label="C3 rigid TIP3P water cluster, 20 waters, no cutoff, restraint K=1, SETTLE (BASELINE configs[2])"),
    "alanine": dict(
        config="C4", kind="particles", replicas=59200, protocol_length=200, dt=0.002, gamma=1.0, kT=KT_298,
        splittings={"VRORV": "V R O R V", "MTS": "V0 V1 R O R V1 V1 R O R V1 V0"}, flops={"VRORV": 29000, "MTS": 59000},
        precision="mixed", ref_replicas=1024, cpu_replicas=2048, e2e_replicas=59200,
        label="C4 alanine dipeptide vacuum, H-bonds constrained, MTS V0/V1 (BASELINE configs[3])"),
    "waterbox": dict(
        config="C5", kind="particles", replicas=8192, protocol_length=200, dt=0.002, gamma=1.0, kT=KT_298,
        splittings={"OVRVO": "O V R V O", "VRORV": "V R O R V"}, flops={"OVRVO": 1.20e6, "VRORV": 1.20e6}, precision="mixed",
        ref_replicas=32, cpu_replicas=48, e2e_replicas=8192, hbm_bytes_per_step_streaming=19584.0,
        label="C5 tiny water box, 102 TIP3P waters, 1.5 nm, cutoff 0.6 nm + reaction field, switched LJ, SETTLE (BASELINE configs[4])"),
}
ORDER = ["double_well", "quartic", "water_cluster", "alanine", "waterbox"]


def flops_per_step(splitting, f_force):
    """Algorithmic flops per replica-step of a 1-DOF system (DESIGN.md flop convention): FMA = 2, add/mul = 1,
    transcendental = 1; Philox / Box-Muller work is NOT counted.  O: 9, V: 2, R: 2, one force evaluation per distinct
    position visited (steady state)."""
    toks = splitting.split()
    n_evals, valid = 0, False
    for i, t in enumerate(toks + toks):  # the second pass sees the steady-state validity
        if t == "R":
            valid = False
        elif t == "V":
            if not valid and i >= len(toks):
                n_evals += 1
            valid = True
    return 9 * toks.count("O") + 2 * toks.count("V") + 2 * toks.count("R") + f_force * n_evals


def metric_name(name, W):
    return "replica-steps/sec (%s, %s, F+R protocol)" % (name, "+".join(W["splittings"]))


def steps_per_leg(W, plen=None):
    plen = plen or W["protocol_length"]
    return plen - 1 if W["kind"] == "toy" else plen


def workload_config(name, W, n, plen):
    cfg = {"workload": "%s: %d replicas/GPU x protocol_length %d (%d steps/leg, 2 legs) x %d splittings"
                       % (W["label"], n, plen, steps_per_leg(W, plen), len(W["splittings"])),
           "system": name, "gamma": W["gamma"], "dt": W["dt"], "kT": W["kT"], "marginal": "configuration",
           "splittings": list(W["splittings"].values()), "replicas_per_gpu": n, "protocol_length": plen, "precision": W["precision"],
           "l2": "flushed between timed steps (256 MiB memset, untimed)",
           "parallelism": "replica shards, no data-path collective; one moment all-reduce per 8 steps, overlapped with the following steps"}
    if W["kind"] == "toy":
        cfg["mass"], cfg["equilibrium_cache"] = W["mass"], W["n_cache"]
    return cfg


class ClockSampler:
    """SM clock + clock-event reasons sampled DURING the timed region (NVML from a thread, back to back -- a query takes a
    fraction of a millisecond and releases the GIL --; the nvidia-smi -lms loop of the profiling recipe cannot resolve a region of
    a few milliseconds).  The interpreter's thread switch interval is shortened while the sampler lives, so that the main
    thread's launch loop cannot starve it."""

    def __init__(self, gpu_index=0, period_s=0.0):
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
        self._switch = sys.getswitchinterval()
        sys.setswitchinterval(2e-4)
        self.t = threading.Thread(target=self._run, daemon=True)
        self.t.start()
        t_wait = time.perf_counter()  # the first queries after nvmlInit can take tens of milliseconds: let them pass
        while len(self.rows) < 5 and time.perf_counter() - t_wait < 1.0:
            time.sleep(0.002)

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
        sys.setswitchinterval(self._switch)
        lo, hi = (self.marks[0], self.marks[-1]) if len(self.marks) >= 2 else (0.0, 1e30)
        rows = [r for r in self.rows if lo <= r[0] <= hi]
        if len(rows) < 2 and self.rows:  # a region shorter than two queries: the samples that bracket it
            before = [r for r in self.rows if r[0] < lo][-1:]
            after = [r for r in self.rows if r[0] > hi][:1]
            rows = before + rows + after
            out["bracketing_samples"] = True
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


# ------------------------------------------------------------------------------------------------
# inputs
def particle_description(name):
    """(system description, start geometry or frames) of a particle workload"""
    from integrator_benchmark_b200.testsystems import systems as ts
    if name == "water_cluster":
        return ts.water_cluster(20, constrained=True, seed=1)
    if name == "alanine":
        return ts.alanine_dipeptide(constrained=True)
    desc, _ = ts.tiny_waterbox(constrained=True)
    frames = np.load(os.path.join(ROOT, "tests", "golden", "waterbox_frames.npz"))["xyz"].astype(np.float64)
    return desc, frames


def particle_simulator(name):
    """the reference-facing EquilibriumSimulator of a particle workload (testsystems/molecular.py)"""
    from integrator_benchmark_b200 import testsystems
    return {"water_cluster": testsystems.water_cluster_rigid, "alanine": testsystems.alanine_constrained,
            "waterbox": testsystems.tiny_waterbox}[name]


# ------------------------------------------------------------------------------------------------
# CPU legs (test infrastructure: the only places bench.py touches oracle/)
def cpu_reference_toy(name, W, n_per_scheme, x_cache, repeats=1):
    """The reference's own numba closures (oracle/_ref) under the restated collect_protocol_samples loop, all host cores.
    Only the schemes the reference's file implements are run.  Returns (replica-steps/s, seconds, description) or None."""
    from oracle import ref_toy
    ok, why = ref_toy.available()
    if not ok:
        return None, why
    schemes = [k for k in W["splittings"] if k in REF_SCHEMES]
    cores = os.cpu_count() or 1
    plen = W["protocol_length"]
    ref_toy.collect_parallel(name, schemes, x_cache, max(cores * 16, 64), plen, "configuration", W["gamma"], W["dt"], processes=cores)
    secs = 0.0
    for r in range(repeats):
        _, s = ref_toy.collect_parallel(name, schemes, x_cache, n_per_scheme, plen, "configuration", W["gamma"], W["dt"],
                                        processes=cores, seed=17 * r)
        secs += s
    total = repeats * len(schemes) * n_per_scheme * 2 * (plen - 1)
    return (total / secs, secs, "%d protocols x %d schemes (%s) x 2 legs x %d steps x %d passes, %d worker processes: the reference's "
            "numba_integrators.py closures (oracle/_ref) under the restated low_dimensional_systems.py:160-184 loop, traces "
            "collected as the reference does" % (n_per_scheme, len(schemes), "/".join(schemes), plen - 1, repeats, cores)), ""


def cpu_port_toy(name, W, n, x_cache, repeats=1):
    """restated C oracle (OpenMP over replicas) on the same workload: (replica-steps/s, seconds)"""
    from oracle import oracle
    oracle.set_threads()
    steps = W["protocol_length"] - 1
    oracle.toy_protocol(name, "V R O R V", W["dt"], W["gamma"], 4096, steps, mass=W["mass"], kT=W["kT"], seed=1, x_cache=x_cache)
    t0 = time.perf_counter()
    for _ in range(repeats):
        for s in W["splittings"].values():
            oracle.toy_protocol(name, s, W["dt"], W["gamma"], n, steps, mass=W["mass"], kT=W["kT"], seed=1, x_cache=x_cache)
    secs = time.perf_counter() - t0
    return repeats * len(W["splittings"]) * n * 2 * steps / secs, secs


def cpu_port_particles(name, W, n, steps, cache):
    from oracle import oracle
    oracle.set_threads()
    desc, _ = particle_description(name)
    oracle.particles_protocol(desc, "V R O R V", W["dt"], W["gamma"], W["kT"], max(os.cpu_count() or 1, 8), 2, seed=1, x_cache=cache)
    t0 = time.perf_counter()
    for s in W["splittings"].values():
        oracle.particles_protocol(desc, s, W["dt"], W["gamma"], W["kT"], n, steps, seed=1, x_cache=cache)
    secs = time.perf_counter() - t0
    return len(W["splittings"]) * n * 2 * steps / secs, secs


def toy_cache_cpu(name, n=20000):
    from oracle import oracle
    oracle.set_threads()
    return oracle.sample_equilibrium_scalar(name, n, 100, seed=1)


def particle_cache_cpu(name, n=16):
    """a few start configurations for the CPU leg: the committed water-box frames, or the start geometry relaxed by a
    short oracle run (the CPU leg measures throughput; it does not need an equilibrium ensemble)"""
    desc, x = particle_description(name)
    if x.ndim == 3:
        return x
    return x[None]


def cpu_baseline(name, W, target_s=10.0):
    """cpu_baseline object of one workload: about `target_s` seconds of CPU work on all host cores"""
    cores = os.cpu_count() or 1
    if W["kind"] == "toy":
        x_cache = toy_cache_cpu(name)
        n = W["ref_replicas"]
        out, why = cpu_reference_toy(name, W, n, x_cache)
        port_v, _ = cpu_port_toy(name, W, 1 << 17, x_cache)
        if out is not None:
            v, secs, _ = out
            repeats = int(min(32, max(1, round(target_s / max(secs, 1e-3)))))
            (v, secs, what), _ = cpu_reference_toy(name, W, n, x_cache, repeats=repeats)
            return {"value": v, "unit": UNIT, "cores": cores, "kind": "reference", "sample": what + " (%.1f s)" % secs,
                    "port_value": port_v, "port_note": "restated C oracle (oracle/lsi_oracle.c, OpenMP over replicas, final works only)"}
        return {"value": port_v, "unit": UNIT, "cores": cores, "kind": "port",
                "sample": "131072 replicas x %d splittings (restated C oracle; reference closures unavailable: %s)" % (len(W["splittings"]), why)}
    steps = min(W["protocol_length"], 50)
    v, secs = cpu_port_particles(name, W, W["cpu_replicas"], steps, particle_cache_cpu(name))
    return {"value": v, "unit": UNIT, "cores": cores, "kind": "port",
            "sample": "%d replicas x %d splittings x 2 legs x %d steps (%.1f s; restated fp64 C oracle, OpenMP over replicas; not OpenMM -- "
                      "the reference's own figure is 1.28e4 replica-steps/s on the OpenMM Reference platform for alanine dipeptide, "
                      "BASELINE.md)" % (W["cpu_replicas"], len(W["splittings"]), steps, secs)}


def reference_record(name, W, steps, warmup, n_gpus):
    """--impl reference: the reference's CPU implementation of the path on the host cores; each step a bounded sample"""
    cores = os.cpu_count() or 1
    plen = W["protocol_length"]
    if W["kind"] == "toy":
        x_cache = toy_cache_cpu(name)
        n = W["ref_replicas"]
        out, why = cpu_reference_toy(name, W, max(n // 8, 256), x_cache, repeats=max(1, warmup))  # compiles, forks, warms
        if out is not None:
            total_s, units = 0.0, 0.0
            for _ in range(steps):
                (v, secs, what), _ = cpu_reference_toy(name, W, n, x_cache)
                total_s += secs
                units += v * secs
            kind = "reference"
        else:
            total_s, units, what, kind = 0.0, 0.0, "restated C oracle (reference closures unavailable: %s)" % why, "port"
            for _ in range(steps):
                v, secs = cpu_port_toy(name, W, 1 << 17, x_cache)
                total_s += secs
                units += v * secs
    else:
        cache = particle_cache_cpu(name)
        nsteps = min(plen, 50)
        for _ in range(max(1, warmup)):
            cpu_port_particles(name, W, max(W["ref_replicas"] // 8, 8), 5, cache)
        total_s, units, kind = 0.0, 0.0, "port"
        for _ in range(steps):
            v, secs = cpu_port_particles(name, W, W["ref_replicas"], nsteps, cache)
            total_s += secs
            units += v * secs
        what = "%d replicas x %d splittings x 2 legs x %d steps per bench step (restated fp64 C oracle, OpenMP; not OpenMM)" % (
            W["ref_replicas"], len(W["splittings"]), nsteps)
    value = units / total_s
    return {"metric": metric_name(name, W), "value": value, "unit": UNIT, "n_gpus": n_gpus, "steps": steps, "warmup": warmup,
            "ms_per_step": 1e3 * total_s / steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
            "data": "synthetic", "impl": "reference", "config": workload_config(name, W, W["replicas"], plen),
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": kind, "sample": what},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}, "gpu_launches": 0}


def run_reference(args):
    if int(os.environ.get("RANK", 0)) != 0:
        return
    names = ORDER if args.workload == "all" else [args.workload]
    head = reference_record(names[0], WORKLOADS[names[0]], args.steps, args.warmup, args.gpus)
    if len(names) > 1:
        head["per_workload"] = {}
        for nm in names[1:]:
            rec = reference_record(nm, WORKLOADS[nm], 2, 1, args.gpus)
            head["per_workload"]["%s %s" % (WORKLOADS[nm]["config"], nm)] = {k: rec[k] for k in ("metric", "value", "unit", "ms_per_step", "steps",
                                                                                                "config", "cpu_baseline", "e2e")}
    print(json.dumps(head))


# ------------------------------------------------------------------------------------------------
# GPU arm
class Context:
    """process-wide state of the GPU arm: ranks, device, L2-flush buffer, FMA peaks measured once"""

    def __init__(self):
        import torch
        import torch.distributed as dist
        from integrator_benchmark_b200 import engine, parallel
        self.torch, self.dist = torch, dist
        self.rank, self.world, self.local_rank = parallel.init_process_group()
        engine.require_cuda()
        self.dev = torch.device("cuda", self.local_rank)
        torch.cuda.set_device(self.dev)
        self.flush = torch.empty(256 << 20, dtype=torch.uint8, device=self.dev)
        self._peaks = None

    def barrier(self):
        if self.world > 1:
            self.dist.barrier()
        self.torch.cuda.synchronize()

    def max_over_ranks(self, x):
        t = self.torch.tensor([float(x)], dtype=self.torch.float64, device=self.dev)
        if self.world > 1:
            self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return float(t.item())

    def peaks(self):
        """FP64 / FP32 FMA peaks of this GPU, TFLOP/s (lsi_microbench_fma, burst = best of 6)"""
        if self._peaks is None:
            torch = self.torch
            from integrator_benchmark_b200 import _cabi, engine
            sink = torch.zeros(8, dtype=torch.float64, device=self.dev)
            self._peaks = {}
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
                self._peaks[name] = blocks * 256 * iters * 16 / (best * 1e-3) / 1e12
        return self._peaks


def hbm_peak():
    try:
        return json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"], "MEASURED_PEAKS.json"
    except Exception:  # noqa: BLE001
        return 6650.0, "fallback (B200_PROFILING.md)"


def gpu_record(ctx, name, W, steps, warmup, n, plen, with_cpu, min_warm_s=0.0):
    """Measures one workload on the GPU arm; returns the JSON record (rank 0) or None (other ranks)."""
    torch = ctx.torch
    import integrator_benchmark_b200 as ib
    from integrator_benchmark_b200 import engine, parallel
    from integrator_benchmark_b200.integrators import LangevinSplittingIntegrator, numba_integrators
    from integrator_benchmark_b200.testsystems import (NonequilibriumSimulator, NumbaBookkeepingSimulator,
                                                       NumbaNonequilibriumSimulator)
    from integrator_benchmark_b200 import unit

    dev, rank, world = ctx.dev, ctx.rank, ctx.world
    toy = W["kind"] == "toy"
    nsl = steps_per_leg(W, plen)
    seed = 0x5EED0000 + ORDER.index(name) + 1
    replica0 = rank * n  # global replica ids: rank r owns [r n, (r + 1) n)
    S = len(W["splittings"])
    precision = W["precision"]

    # ---- system, programs, equilibrium cache (the path's input; generated once on the GPU, same on every rank)
    if toy:
        system = engine.ScalarSystem(W["system"], W["mass"])
        eq = NumbaBookkeepingSimulator(mass=W["mass"], beta=1.0 / W["kT"], name=name,
                                       **({} if name == "quartic" else {"potential": ib.testsystems.double_well.potential,
                                                                        "force": ib.testsystems.double_well.force}))
        eq.set_x_samples(eq.collect_equilibrium_samples(W["n_cache"], n_burn=300, seed=seed))
        cache = eq.x_samples_device()
    else:
        eq = particle_simulator(name)
        system = eq.engine_system()
        desc, x = particle_description(name)
        n_eq = 1024 if x.ndim == 2 else 256
        x0 = torch.as_tensor(np.ascontiguousarray(x if x.ndim == 3 else x[None]))
        x0 = x0[torch.arange(n_eq) % x0.shape[0]].contiguous().to(dev)
        for k, (h, gamma, ns) in enumerate([(0.0002, 100.0, 500), (0.001, 20.0, 500), (0.002, 5.0, 2000)]):
            res = engine.run_protocol(system, engine.Program("V R O R V", h, gamma), n_eq, ns, W["kT"], n_legs=1, seed=seed + 100 + k,
                                      x0=x0, want_final=True, want_moments=False, out={})
            x0 = res["x_final"]
        eq.set_samples(x0.cpu().numpy())
        cache = eq.samples_device()
    programs = {k: engine.Program(s, W["dt"], W["gamma"]) for k, s in W["splittings"].items()}
    bufs = {k: {} for k in programs}
    mom = torch.zeros(16, S, 8, dtype=torch.float64, device=dev)  # ring of 2 x GROUP rows (GROUP = 8 below)
    pending = [None, None]

    def launch(k, prog, x_cache_dev, moments_out=None):
        return engine.run_protocol(system, prog, n, nsl, W["kT"], marginal="configuration", seed=seed, replica0=replica0,
                                   x_cache=x_cache_dev, out=bufs[k], precision=precision, moments_out=moments_out)

    def launch_all(moments_out):
        """all splittings of the step: scalar systems in ONE launch (lsi_run_protocol_batch), particle systems one by one"""
        if toy:
            engine.run_protocols(system, list(programs.values()), n, nsl, W["kT"], marginal="configuration", seed=seed,
                                 replica0=replica0, x_cache=cache, precision=precision, outs=list(bufs.values()), moments_out=moments_out)
        else:
            for j, (k, prog) in enumerate(programs.items()):
                launch(k, prog, cache, moments_out[j])

    GROUP = 8  # bench steps whose moment vectors travel in one all-reduce

    def one_pass(i):
        """bench step i writes its S x 8 moment vectors into row i % (2 GROUP) of a ring; every GROUP steps the finished
        group of rows is all-reduced in ONE collective, started on NCCL's stream after this step's launch so that it overlaps
        the following steps (moments are additive and nothing downstream needs them earlier); a group's rows are reused only
        after its all-reduce has finished (stream-ordered wait, no host block)"""
        row = i % (2 * GROUP)
        half = row // GROUP
        if row % GROUP == 0 and pending[half] is not None:
            pending[half].wait()
            pending[half] = None
        launch_all(mom[row])
        if row % GROUP == GROUP - 1:
            pending[half] = parallel.all_reduce_moments_async(mom[half * GROUP:(half + 1) * GROUP])

    def drain(i_last):
        """all-reduce the rows of the unfinished group, wait for everything in flight"""
        row = i_last % (2 * GROUP)
        half = row // GROUP
        if row % GROUP != GROUP - 1:
            pending[half] = parallel.all_reduce_moments_async(mom[half * GROUP:row + 1])
        for q in (0, 1):
            if pending[q] is not None:
                pending[q].wait()
                pending[q] = None

    for i in range(warmup):
        one_pass(i)
    drain(warmup - 1)
    ctx.barrier()
    t_warm = time.perf_counter()
    while time.perf_counter() - t_warm < min_warm_s:  # short steps: keep warming until clocks and allocator have settled
        for i in range(20):
            one_pass(i)
        drain(19)
        torch.cuda.synchronize()
    ctx.barrier()

    # ---- device-resident timing
    sampler = ClockSampler(ctx.local_rank) if rank == 0 else None
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
    ctx.barrier()
    if sampler:
        sampler.mark()
    t_wall0 = time.perf_counter()
    for s in range(steps):
        ctx.flush.zero_()
        ev[s][0].record()
        one_pass(s)
        if s == steps - 1:
            drain(s)
        ev[s][1].record()
    ctx.barrier()
    t_wall = time.perf_counter() - t_wall0
    if sampler:
        sampler.mark()
    clocks = sampler.stop() if sampler else None
    dev_ms = ctx.max_over_ranks(sum(a.elapsed_time(b) for a, b in ev))
    units_per_pass = S * n * 2 * nsl  # replica-steps per rank per bench step
    value = world * units_per_pass * steps / (dev_ms * 1e-3)

    # ---- dominant kernel alone: events around each protocol launch, L2 flushed before each; enqueue all, sync once
    k_reps = max(2, min(steps, 50))
    kev = {k: [] for k in programs}
    for _ in range(k_reps):
        for k, prog in programs.items():
            ctx.flush.zero_()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            launch(k, prog, cache)
            b.record()
            kev[k].append((a, b))
    bev = []
    scratch_mom = torch.zeros(S, 8, dtype=torch.float64, device=dev)
    for _ in range(k_reps if toy else 0):  # the step's launch as the bench issues it (one batch launch + one finalize)
        ctx.flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        launch_all(scratch_mom)
        b.record()
        bev.append((a, b))
    torch.cuda.synchronize()
    kernel_ms = {k: sum(a.elapsed_time(b) for a, b in v) / k_reps for k, v in kev.items()}
    batch_ms = sum(a.elapsed_time(b) for a, b in bev) / k_reps if bev else None
    estimates_dev = [ib.evaluation.estimate_from_moments(mom[(steps - 1) % 16][j].cpu().numpy()) for j in range(S)]

    # ---- end to end through the reference-facing call (host cache in, numpy out)
    n_e2e = min(n, W["e2e_replicas"])
    if toy:
        kind_of = W.get("closures")
        sims = {}
        for k, s in W["splittings"].items():
            if kind_of:
                f = getattr(numba_integrators, kind_of[k] + "_factory")(eq.potential, eq.force, eq.velocity_scale, eq.mass)
            else:
                f = numba_integrators.splitting_factory(s, eq.potential, eq.force, eq.velocity_scale, eq.mass)
            sims[k] = NumbaNonequilibriumSimulator(eq, partial(f, gamma=W["gamma"], dt=W["dt"]))
        h2d = int(eq.x_samples.nbytes)

        def call(sim, mode):
            if mode == "default":  # the reference's call: full work and (x, v) traces as numpy arrays
                W_F, W_R, xv_F, xv_R = sim.collect_protocol_samples(n_e2e, plen, "configuration")
                return (W_F[:, -1], W_R[:, -1]), W_F.nbytes + W_R.nbytes + xv_F.nbytes + xv_R.nbytes
            if mode == "final_works":
                W_F, W_R, _, _ = sim.collect_protocol_samples(n_e2e, plen, "configuration", traces=False)
                return (W_F[:, -1], W_R[:, -1]), W_F.nbytes + W_R.nbytes
            sim.collect_protocol_samples(n_e2e, plen, "configuration", traces=False, as_numpy=False)
            return sim.last_moments.cpu().numpy(), 64
        modes = ["default", "final_works", "estimate_only"]
    else:
        sims = {}
        for k, s in W["splittings"].items():
            integ = LangevinSplittingIntegrator(s, temperature=298.0 * unit.kelvin, collision_rate=W["gamma"] / unit.picoseconds,
                                                timestep=W["dt"] * unit.picoseconds)
            sims[k] = NonequilibriumSimulator(eq, integ, precision=precision)
        h2d = int(eq.samples.nbytes)

        def call(sim, mode):
            if mode == "default":  # the reference's call: {"W_shads_F", "W_shads_R"} as numpy arrays
                r = sim.collect_protocol_samples(n_e2e, plen, "configuration")
                return (r["W_shads_F"], r["W_shads_R"]), r["W_shads_F"].nbytes + r["W_shads_R"].nbytes
            sim.collect_protocol_samples(n_e2e, plen, "configuration", as_numpy=False)
            return sim.last_moments.cpu().numpy(), 64
        modes = ["default", "estimate_only"]

    e2e_steps = max(2, min(steps // 2, 10))
    e2e = {}
    for mode in modes:
        d2h = 0
        for _ in range(2):  # first calls page-lock their staging buffers
            eq.release_device_cache()
            for sim in sims.values():
                call(sim, mode)
        ctx.barrier()
        t0 = time.perf_counter()
        for _ in range(e2e_steps):
            d2h = 0
            # the step's input, the host-side equilibrium cache, is uploaded again every step (the simulators of a step share
            # their equilibrium simulator, as the reference's callers do, hence its device copy)
            eq.release_device_cache()
            for sim in sims.values():  # the timed region ends with the call's results on the host (works, or the moment vector);
                got, nbytes = call(sim, mode)  # they are dropped at once, so their pinned staging blocks are recycled
                d2h += nbytes
                del got
        ctx.barrier()
        secs = ctx.max_over_ranks(time.perf_counter() - t0)
        e2e[mode] = {"value": world * S * n_e2e * 2 * nsl * e2e_steps / secs, "unit": UNIT, "h2d_bytes_per_step": h2d,
                     "d2h_bytes_per_step": int(d2h), "steps": e2e_steps, "replicas_per_call": n_e2e}
        if mode == "default":  # what a caller computes from the returned works: one more pass, outside the timed region
            estimates_e2e = []
            for sim in sims.values():
                (wf, wr), _ = call(sim, mode)
                estimates_e2e.append(ib.evaluation.estimate_nonequilibrium_free_energy(wf, wr))
                del wf, wr

    if rank != 0:
        return None

    # ---- roofline of the dominant kernel
    peaks = ctx.peaks()
    kernel_s = (batch_ms if batch_ms is not None else sum(kernel_ms.values())) * 1e-3
    if toy:
        fl = {k: flops_per_step(s, W["f_force"]) for k, s in W["splittings"].items()}
        bound, peak, kname = "fp64_fma", peaks["fp64"], "toy_fused2_batch_kernel<%s, %d splittings in one launch>" % (name, S)
        note = ("algorithmic flops exclude the Philox + Box-Muller noise generation that dominates the instruction count "
                "(n_O Gaussians per replica-step)")
    else:
        fl = W["flops"]
        fp32 = precision == "mixed"
        bound, peak, kname = ("fp32_fma" if fp32 else "fp64_fma"), (peaks["fp32"] if fp32 else peaks["fp64"]), "particles_protocol_kernel<%s>" % name
        note = ("algorithmic flops per replica-step from SURVEY 8d (half-shell pair count; FMA = 2); the kernel runs a full-shell "
                "loop (twice the pair evaluations, no atomics) with fp64 constraints and bookkeeping")
    achieved = sum(fl[k] * n * 2 * nsl for k in programs) / kernel_s / 1e12
    hpk, hsrc = hbm_peak()
    # per pass: start state in, two works out per splitting (the toy batch launch gathers the start state once for all splittings)
    hbm_bytes = n * (8 + 16 * S) if toy else S * n * (24 * system.n_atoms + 16)
    roofline = {"bound": bound, "kernel": kname, "achieved": achieved, "peak": peak, "unit": "TFLOP/s", "frac": achieved / peak,
                # dram__bytes_read.sum + dram__bytes_write.sum of one toy_family_kernel launch, ncu --set full
                # (profiles/r2f_toy_family_ncu_details.txt: 8.2 MB read, 43-45 MB written; the rest of the 100 MB of works is still
                # in L2 when the kernel ends)
                "traffic": 5.2e7 if (name == "double_well" and n == (1 << 20) and S == 6) else None,
                "peak_source": "lsi_microbench_fma, best of 6, measured in this run (fp64 %.1f, fp32 %.1f TFLOP/s)" % (peaks["fp64"], peaks["fp32"]),
                "note": note, "replica_steps_per_s_kernel_only": units_per_pass / kernel_s, "kernel_ms_per_splitting": kernel_ms,
                "kernel_ms_batch_launch": batch_ms,
                "flops_per_replica_step": fl,
                "hbm": {"achieved": hbm_bytes / kernel_s / 1e9, "peak": hpk, "unit": "GB/s", "frac": hbm_bytes / kernel_s / 1e9 / hpk,
                        "peak_source": hsrc, "note": "algorithmic bytes: start state gathered once, two works written per replica"}}
    if "hbm_bytes_per_step_streaming" in W:
        eqv = W["hbm_bytes_per_step_streaming"] * units_per_pass / kernel_s / 1e9
        roofline["hbm_equivalent"] = {"achieved": eqv, "peak": hpk, "unit": "GB/s", "frac": eqv / hpk, "peak_source": hsrc,
                                      "note": "bytes a per-step-streaming design would move (19.6 KB per replica-step); this kernel "
                                              "keeps the state in shared memory and moves 7.4 KB per replica-PROTOCOL"}
    main = dict(e2e["default"])
    main["call"] = ("NumbaNonequilibriumSimulator.collect_protocol_samples(n, protocol_length, marginal)" if toy else
                    "NonequilibriumSimulator.collect_protocol_samples(n, protocol_length, marginal)") + ", default arguments, numpy results"
    if "final_works" in e2e:
        main["final_works_only"] = dict(e2e["final_works"], note="same call with traces=False: the two final work arrays only")
    main["estimate_only"] = dict(e2e["estimate_only"], note="same call with as_numpy=False: works stay on the device, the moments of the estimate are read back")
    launches_per_step = 2 if toy else 3 * S  # toy: one batch kernel + one finalize; particles: protocol + partials + finalize each
    rec = {"metric": metric_name(name, W), "value": value, "unit": UNIT, "n_gpus": world, "steps": steps, "warmup": warmup,
           "ms_per_step": dev_ms / steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
           "dtype": "f64" if precision == "double" else "f32 forces and O noise / f64 state, constraints and accumulators",
           "data": "synthetic", "config": workload_config(name, W, n, plen), "wall_ms_per_step_incl_flush": 1e3 * t_wall / steps,
           "clocks": clocks, "e2e": main, "gpu_launches": steps * launches_per_step, "roofline": roofline,
           "cpu_baseline": cpu_baseline(name, W) if with_cpu else None,
           "DeltaF_neq": {k: {"estimate": e[0], "stderr": float(np.sqrt(max(e[1], 0.0)))} for k, e in zip(programs, estimates_dev)},
           "DeltaF_neq_e2e": {k: {"estimate": float(e[0]), "stderr": float(np.sqrt(max(e[1], 0.0)))} for k, e in zip(programs, estimates_e2e)}}
    if name == "water_cluster":
        # the second half of BASELINE's metric, "wall time to DeltaF_neq at fixed stderr": the reference's published point
        # (water_cluster_rigid, OVRVO, 4 fs: stderr 1.45e-3 kT from N = 100 000 protocols x 1000 steps) costs
        # N x 2 legs x 1000 replica-steps at the measured throughput
        rec["time_to_fixed_stderr"] = {"stderr_kT": 1.45e-3, "protocols": 100000, "protocol_length": 1000,
                                       "seconds": 100000 * 2 * 1000 / value,
                                       "note": "one point of the reference's near-equilibrium sweep at its published standard error"}
    return rec


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=0, help="0 = 200 (toy headline) / 4 (particle workloads)")
    ap.add_argument("--warmup", type=int, default=0, help="0 = 10 (toy headline) / 3")
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="all", choices=["all"] + ORDER,
                    help="all = C2 headline + per_workload records of the other configs; or one config alone")
    ap.add_argument("--replicas", type=int, default=0, help="replicas per GPU (0 = the workload's default)")
    ap.add_argument("--protocol-length", type=int, default=0, help="0 = the workload's default")
    ap.add_argument("--ref-replicas", type=int, default=0, help="protocols per scheme (toy) / replicas (particles) of one reference-arm step")
    ap.add_argument("--precision", default=None, choices=[None, "double", "mixed"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    names = ORDER if args.workload == "all" else [args.workload]
    head = names[0]
    toy_head = WORKLOADS[head]["kind"] == "toy"
    if args.ref_replicas:
        WORKLOADS[head] = dict(WORKLOADS[head], ref_replicas=args.ref_replicas)
    if args.impl == "reference":
        args.steps, args.warmup = args.steps or 5, args.warmup or 1
        return run_reference(args)
    args.steps = args.steps or (200 if toy_head else 4)
    args.warmup = max(args.warmup or (10 if toy_head else 3), 3)  # timing rule: at least three warm-up steps
    if args.precision:
        for nm in names:
            WORKLOADS[nm] = dict(WORKLOADS[nm], precision=args.precision)

    ctx = Context()
    W = WORKLOADS[head]
    with_cpu = ctx.world == 1 and not args.no_cpu_baseline
    line = gpu_record(ctx, head, W, args.steps, args.warmup, args.replicas or W["replicas"], args.protocol_length or W["protocol_length"],
                      with_cpu, min_warm_s=0.5 if toy_head else 0.0)
    per = {}
    for nm in names[1:]:
        Wn = WORKLOADS[nm]
        k = 20 if Wn["kind"] == "toy" else 4
        rec = gpu_record(ctx, nm, Wn, k, 3, Wn["replicas"], Wn["protocol_length"], with_cpu, min_warm_s=0.2 if Wn["kind"] == "toy" else 0.0)
        if rec is not None:
            per["%s %s" % (Wn["config"], nm)] = rec
    if ctx.rank == 0:
        if per:
            line["per_workload"] = per
        print(json.dumps(line))
    if ctx.world > 1:
        ctx.dist.destroy_process_group()


if __name__ == "__main__":
    main()

"""Drop-in for benchmark/testsystems/bookkeepers.py:45-320 (molecular systems).

`EquilibriumSimulator` owns a system description and its equilibrium-configuration cache;
`NonequilibriumSimulator` runs the forward / reverse protocol and returns shadow works.  The
reference drives one OpenMM Context from a Python loop over protocol samples; here the whole
ensemble of `n_protocol_samples` replicas is ONE launch of the fused protocol kernel
(lsi_run_protocol), each replica keyed by its global id.  There is no OpenMM, no mdtraj and no
CPU path: frames are plain (n_atoms, 3) arrays in nm (an object with an `.xyz` attribute, as an
mdtraj frame has, is accepted too).
"""
import os

import numpy as np

import warnings

from .. import _rngstate, _staging, engine, unit
from ..constants import kB


def _frame_xyz(frame):
    xyz = getattr(frame, "xyz", frame)
    xyz = np.asarray(xyz.value_in_unit(unit.nanometer) if unit.is_quantity(xyz) else xyz, dtype=np.float64)
    if xyz.ndim == 3:
        if xyz.shape[0] != 1:
            raise TypeError("Expected length-1 frame, got length {}".format(xyz.shape[0]))
        xyz = xyz[0]
    return xyz


class EquilibriumSimulator:
    """Equilibrium side of a molecular test system (reference :45-156).

    `system` is the plain description dict of testsystems/systems.py (the counterpart of the OpenMM
    System the reference passes); `platform` and `topology` are carried for callers that read them."""

    def __init__(self, platform, topology, system, positions, temperature, timestep, burn_in_length, n_samples,
                 thinning_interval, name):
        self.platform = platform
        self.topology = topology
        self._engine_systems = {}  # constraint tolerance -> engine.ParticleSystem
        self.system = system
        self.positions = np.asarray(positions, dtype=np.float64)
        self.temperature = temperature
        self.timestep = timestep
        self.burn_in_length = burn_in_length
        self.n_samples = n_samples
        self.thinning_interval = thinning_interval
        self.name = name
        self.cached = False
        self.constraint_tolerance = 1e-8
        self.steps_per_hmc = 10
        self.kT = self.temperature * kB
        self.samples = None
        self._samples_dev = None
        # sampler settings (see collect_equilibrium_samples)
        self.extra_chances = 15
        self.equilibration_collision_rate = 5.0  # 1/ps, Langevin mixing stage
        self.metropolis_rounds = 2000            # final extra-chance HMC rounds (x steps_per_hmc steps)
        self.sampler_precision = "double"
        self.acceptance = None
        self.max_equilibration_steps = 1000000  # 1 ns at 1 fs, the length of the reference's burn-in
        self.use_disk_cache = False              # the reference always caches to DATA_PATH; opt in here

    # ---- engine handles
    @property
    def system(self):
        return self._system_desc

    @system.setter
    def system(self, desc):
        # a new description (e.g. repartitioned hydrogen masses) invalidates the compiled system tables
        self._system_desc = desc
        self._engine_systems = {}

    @property
    def kT_value(self):
        return self.kT.value_in_unit(unit.kilojoule_per_mole) if unit.is_quantity(self.kT) else float(self.kT)

    def engine_system(self, constraint_tolerance=None):
        """The engine's handle of this system at a constraint tolerance (default: this simulator's own, 1e-8).  Handles
        are cached per tolerance, so simulators and GHMC probes that share this object never change each other's."""
        tol = float(self.constraint_tolerance if constraint_tolerance is None else constraint_tolerance)
        if tol not in self._engine_systems:
            desc = dict(self.system)
            desc["constraint_tolerance"] = tol
            self._engine_systems[tol] = engine.ParticleSystem(desc)
        return self._engine_systems[tol]

    @property
    def precision(self):
        """arithmetic the platform this simulator was built on stands for: "Reference" -> "double", CUDA / OpenCL ->
        "mixed" (benchmark/testsystems/configuration.py:36-45)"""
        return getattr(self.platform, "precision", "mixed")

    # ---- equilibrium cache
    def collect_equilibrium_samples(self, seed=None):
        """x ~ pi for the cache (reference :66-113: ONE chain of XC-GHMC with steps_per_hmc = 10, 15 extra chances
        and no collision rate, i.e. extra-chance HMC; `burn_in_length` rounds of burn-in, then a sample every
        `thinning_interval` rounds).  Here `n_samples` INDEPENDENT chains start from `positions`:
          1. a short small-step, high-friction Langevin relaxation (the reference minimises the energy instead);
          2. VRORV Langevin dynamics at `timestep`, low friction, for min(burn_in_length * steps_per_hmc,
             max_equilibration_steps) steps -- the mixing stage (an HMC chain that redraws every velocity each
             10 fs diffuses far too slowly to equilibrate 1000 chains from a common start);
          3. `metropolis_rounds` rounds of the reference's extra-chance HMC (lsi_sample_equilibrium_particles, fp64),
             which removes the O(dt^2) bias of stage 2: the cache is distributed as exp(-U/kT) on the constraint
             manifold.
        Returns (n_samples, n_atoms, 3) in nm; HMC acceptance statistics in `self.acceptance`."""
        engine.require_cuda()
        t = engine.torch()
        system = self.engine_system()
        n = int(self.n_samples)
        seed = _rngstate.get_seed() if seed is None else int(seed)
        dt = self.timestep.value_in_unit(unit.picosecond) if unit.is_quantity(self.timestep) else float(self.timestep)
        x = t.as_tensor(np.ascontiguousarray(self.positions))[None].expand(n, -1, -1).contiguous().cuda()
        burn = int(min(self.burn_in_length * self.steps_per_hmc, self.max_equilibration_steps))
        for k, (h, gamma, n_steps) in enumerate([(0.1 * dt, 100.0, 500), (0.5 * dt, 20.0, 500),
                                                 (dt, self.equilibration_collision_rate, burn)]):
            prog = engine.Program("V R O R V", h, gamma)
            res = engine.run_protocol(system, prog, n, n_steps, self.kT_value, n_legs=1, seed=seed + 7919 * (k + 1),
                                      replica0=0, x0=x, want_final=True, want_moments=False,
                                      precision="double" if k < 2 else "mixed", out={})
            x = res["x_final"]
        rounds = int(min(self.metropolis_rounds, max(1, burn // self.steps_per_hmc)))
        x, _, acc = engine.sample_equilibrium_particles(system, x, self.kT_value, rounds, self.steps_per_hmc, self.extra_chances, dt,
                                                        seed=seed, replica0=0, precision=self.sampler_precision)
        self.acceptance = acc.mean(dim=0).cpu().numpy()  # (rounds accepted, accepted at the first chance)
        ok = t.isfinite(x).all(dim=(1, 2))
        self.n_dropped = int((~ok).sum().item())
        if self.n_dropped:
            warnings.warn("%s: %d of %d equilibration chains ended in a non-finite state and were dropped from the cache"
                          % (self.name, self.n_dropped, n))
            x = x[ok]
        self._samples_dev = x.contiguous()
        self.samples = self._samples_dev.cpu().numpy()
        self.cached = True
        return self.samples

    def collect_equilibrium_samples_single_chain(self, n_chains=1, thinning_interval=None, seed=None, quench_steps=20000,
                                                 progress=None):
        """Equilibrium caches built THE REFERENCE'S WAY (reference :89-113): one extra-chance-HMC chain per cache
        (steps_per_hmc = 10, 15 extra chances, `timestep`), started from the energy-minimised `positions`, burnt in for
        `burn_in_length` rounds, then `n_samples` configurations taken every `thinning_interval` rounds (default: this
        simulator's, the reference's 10 000).  `n_chains` independent chains run side by side (one warp or CTA each) and
        cost the same wall time as one; returns (n_chains, n_samples, n_atoms, 3) in nm, float64, on the host.  The
        minimisation is a quench: Langevin dynamics at 1 K (the reference calls OpenMM's L-BFGS minimiser).
        A lone chain is latency-bound (about 0.2 ms per round for the 20-water cluster), which is why
        collect_equilibrium_samples() -- many independent chains -- is the default."""
        engine.require_cuda()
        t = engine.torch()
        system = self.engine_system()
        n = int(n_chains)
        thin = int(self.thinning_interval if thinning_interval is None else thinning_interval)
        seed = _rngstate.get_seed() if seed is None else int(seed)
        dt = self.timestep.value_in_unit(unit.picosecond) if unit.is_quantity(self.timestep) else float(self.timestep)
        x = t.as_tensor(np.ascontiguousarray(self.positions))[None].expand(n, -1, -1).contiguous().cuda()
        # energy "minimisation": relax at small step, then quench at 1 K (each chain gets its own noise)
        for k, (h, gamma, kT, n_steps) in enumerate([(0.1 * dt, 100.0, self.kT_value, 500), (0.5 * dt, 20.0, self.kT_value, 2000),
                                                     (dt, 10.0, self.kT_value / 298.0, int(quench_steps))]):
            res = engine.run_protocol(system, engine.Program("V R O R V", h, gamma), n, n_steps, kT, n_legs=1,
                                      seed=seed + 104729 * (k + 1), replica0=0, x0=x, want_final=True, want_moments=False,
                                      precision="double", out={})
            x = res["x_final"]
        call = [0]

        def rounds(x_in, n_rounds):
            call[0] += 1  # a fresh seed per launch: every launch counts its rounds from zero
            x_out, _, acc = engine.sample_equilibrium_particles(system, x_in, self.kT_value, int(n_rounds), self.steps_per_hmc,
                                                                self.extra_chances, dt, seed=seed + 15485863 * call[0], replica0=0,
                                                                precision=self.sampler_precision)
            return x_out, acc
        x, acc = rounds(x, self.burn_in_length)
        burn_acc = acc.mean(dim=0).cpu().numpy()
        out = np.empty((n, int(self.n_samples)) + tuple(self.positions.shape), dtype=np.float64)
        acc_sum = t.zeros(2, dtype=t.float64, device=x.device)
        for i in range(int(self.n_samples)):
            x, acc = rounds(x, thin)
            acc_sum += acc.mean(dim=0)
            out[:, i] = x.cpu().numpy()
            if progress is not None and (i + 1) % 100 == 0:
                progress(i + 1)
        self.acceptance = (acc_sum / max(1, int(self.n_samples))).cpu().numpy()
        self.burn_in_acceptance = burn_acc
        if not np.isfinite(out).all():
            warnings.warn("%s: a single-chain cache contains non-finite configurations" % self.name)
        return out

    def get_acceptance_rate(self):
        """fraction of HMC rounds accepted (reference :80-83)"""
        return float(self.acceptance[0])

    def load_or_simulate_samples(self):
        """If equilibrium samples were collected and stored before, load them, else collect and store them
        (reference :66-76).  The cache is `{name}_samples.h5` in DATA_PATH, an mdtraj HDF5 trajectory as the reference
        writes it (its own files, e.g. data/tiny_waterbox_constrained_samples.h5, load directly).
        Off by default (collecting a cache takes seconds here, not hours): set `use_disk_cache = True`."""
        if self.cached:
            return
        if self.use_disk_cache and self.check_for_cached_samples():
            self.set_samples(self._read_cache())
            return
        self.collect_equilibrium_samples()
        if self.use_disk_cache:
            self.save_samples()

    def get_path_to_samples(self):
        """Samples are {name}_samples.h5 in DATA_PATH (reference :115-118): an mdtraj HDF5 trajectory"""
        from .. import DATA_PATH
        return os.path.join(DATA_PATH, "{}_samples.h5".format(self.name))

    def _candidate_paths(self):
        h5 = self.get_path_to_samples()
        return [h5, h5[:-3] + ".npz"]  # .npz: caches written by earlier versions of this package

    def check_for_cached_samples(self):
        """Check if there's a file where we expect to find cached samples (reference :120-125)"""
        return any(os.path.exists(p) for p in self._candidate_paths())

    def _read_cache(self):
        n_atoms = self.positions.shape[0]
        for path in self._candidate_paths():
            if not os.path.exists(path):
                continue
            if path.endswith(".npz"):
                with np.load(path) as d:
                    xyz = np.asarray(d["xyz"], dtype=np.float64)
            else:
                from ..serialization.mdtraj_hdf5 import load_mdtraj_hdf5
                xyz = load_mdtraj_hdf5(path, n_atoms)["xyz"].astype(np.float64)
            if xyz.ndim != 3 or xyz.shape[1:] != (n_atoms, 3):
                raise ValueError("{}: cached samples have shape {}, expected (*, {}, 3)".format(path, xyz.shape, n_atoms))
            return xyz
        raise FileNotFoundError(self.get_path_to_samples())

    def save_samples(self, path=None, potential_energy=None):
        """Store the cache the way the reference's HDF5Reporter does (:106-113: coordinates, cell, potentialEnergy; float32 nm):
        an mdtraj HDF5 trajectory that `md.load` opens (serialization/mdtraj_hdf5.py).  Potential energies (kJ/mol) are
        evaluated on the GPU when one is available and none are given."""
        from ..serialization.mdtraj_hdf5 import save_mdtraj_hdf5, topology_from_description
        path = self.get_path_to_samples() if path is None else path
        os.makedirs(os.path.dirname(os.path.abspath(path)), exist_ok=True)
        xyz = np.asarray(self.samples, dtype=np.float32)
        desc = self.system if isinstance(self.system, dict) else {}
        if potential_energy is None and desc.get("n_atoms") and engine.lib.lsi_device_count() > 0:
            pe, _ = self.engine_system().energy_forces(np.asarray(self.samples, dtype=np.float64), want_forces=False)
            potential_energy = pe.cpu().numpy()
        periodic = desc.get("nonbonded_method") == "CutoffPeriodic"
        topology = topology_from_description(desc) if desc.get("n_atoms") else None
        tmp = path + ".tmp"
        save_mdtraj_hdf5(tmp, xyz, cell_lengths=desc.get("box") if periodic else None, potential_energy=potential_energy,
                         topology=topology, title=str(self.name))
        os.replace(tmp, path)
        return path

    def set_samples(self, samples):
        """Install an externally produced cache, e.g. the frames of the reference's
        data/tiny_waterbox_constrained_samples.h5 (tests/golden/waterbox_frames.npz)."""
        xyz = np.asarray(getattr(samples, "xyz", samples), dtype=np.float64)
        t = engine.torch()
        host = t.as_tensor(np.ascontiguousarray(xyz[None] if xyz.ndim == 2 else xyz)).clone()
        if t.cuda.is_available():
            host = host.pin_memory()  # uploads of the cache are asynchronous DMAs from page-locked memory
        self._samples_host = host
        self.samples = host.numpy()
        self._samples_dev = None
        self.cached = True

    def release_device_cache(self):
        """Forget the device copy of the cache: the next protocol call uploads `samples` again (host -> device)."""
        self._samples_dev = None

    def samples_device(self):
        self.load_or_simulate_samples()
        if self._samples_dev is None:
            host = getattr(self, "_samples_host", None)
            if host is None or host.numpy().ctypes.data != np.asarray(self.samples).ctypes.data:
                host = engine.torch().as_tensor(np.ascontiguousarray(self.samples, dtype=np.float64))
            self._samples_dev = host.to("cuda", non_blocking=True)
        return self._samples_dev

    def sample_x_from_equilibrium(self):
        """Draw sample (uniformly, with replacement) from cache of configuration samples (:127-131)"""
        self.load_or_simulate_samples()
        return self.samples[np.random.randint(len(self.samples))]

    def sample_v_given_x(self, x, seed=None):
        """Maxwell-Boltzmann velocities projected onto the velocity constraints at x (:133-140)."""
        return _sample_v(self, x, self.constraint_tolerance, seed)


def _sample_v(eq, x, tol, seed=None):
    system = eq.engine_system(tol)
    prog = engine.Program("V R O R V", 0.001, 1.0)
    seed = _rngstate.get_seed() if seed is None else int(seed)
    rid = _rngstate.reserve_replicas(1)
    res = engine.run_protocol(system, prog, 1, 0, eq.kT_value, n_legs=1, seed=seed, replica0=rid,
                              x0=_frame_xyz(x)[None], want_final=True, want_moments=False)
    return res["v_final"][0].cpu().numpy() * (unit.nanometer / unit.picosecond)


class NonequilibriumSimulator:
    """Nonequilibrium simulator, supporting shadow_work accumulation, and drawing x, v, from
    equilibrium (reference :159-295)."""

    def __init__(self, equilibrium_simulator, integrator, precision=None, round_midpoint=True):
        self.equilibrium_simulator, self.integrator = equilibrium_simulator, integrator
        self.constraint_tolerance = self.integrator.getConstraintTolerance()
        # the arithmetic follows the platform the equilibrium simulator was built on, as the reference's Simulation does
        # (bookkeepers.py:142-156): "Reference" -> "double", CUDA -> "mixed" (how the reference configures OpenMM's CUDA
        # platform, testsystems/configuration.py:42-45); an explicit `precision` overrides it
        self.precision = equilibrium_simulator.precision if precision is None else precision
        if self.precision not in ("double", "mixed"):
            raise ValueError("precision must be 'double' or 'mixed'")
        # positions pass through a float32 frame between the legs, as in the reference (:266,298-305)
        self.round_midpoint = round_midpoint
        self._buffers = {}
        self._last = None  # final (x, v) of the most recent single-trajectory call

    def _system(self):
        return self.equilibrium_simulator.engine_system(self.constraint_tolerance)

    # ---- sampling
    def sample_x_from_equilibrium(self):
        return self.equilibrium_simulator.sample_x_from_equilibrium()

    def sample_v_given_x(self, x):
        return _sample_v(self.equilibrium_simulator, x, self.constraint_tolerance)

    def _kT(self):
        return self.equilibrium_simulator.kT_value

    # ---- one trajectory (the reference's unit of work; here a 1-replica launch)
    def accumulate_shadow_work(self, x_0, v_0, n_steps, store_potential_energy=False, store_W_shad_trace=False):
        """Run the integrator for n_steps and return a dictionary containing "W_shad" (kT) and optionally
        "potential_energies" (length n_steps + 1, kJ/mol) and "W_shad_trace" (per-step increments, length
        n_steps).  Non-finite start states give an empty dictionary (:199-204)."""
        result = {}
        x0 = _frame_xyz(x_0)
        v0 = np.asarray(v_0.value_in_unit(unit.nanometer / unit.picosecond) if unit.is_quantity(v_0) else v_0,
                        dtype=np.float64)
        if not (np.isfinite(x0).all() and np.isfinite(v0).all()):
            print("Error setting positions or velocities!")
            return result
        eq = self.equilibrium_simulator
        traces = bool(store_potential_energy or store_W_shad_trace)
        rid = _rngstate.reserve_replicas(1)
        direct = self.integrator._measure_shadow_work
        res = engine.run_protocol(self._system(), self.integrator.program, 1, int(n_steps), self._kT(), n_legs=1,
                                  seed=_rngstate.get_seed(), replica0=rid, x0=x0[None], v0=v0[None],
                                  precision=self.precision, want_heat=True, want_direct=direct, want_final=True,
                                  traces=traces, want_moments=False, out={})
        kT = self._kT()
        self.integrator._accumulate(heat=res["heat"][0, 0].item(),
                                    shadow_work=res["W_direct"][0, 0].item() * kT if direct else None)
        self._last = (res["x_final"][0].cpu().numpy(), res["v_final"][0].cpu().numpy())
        if store_potential_energy:
            frames = res["trace_x"][0, :, 0]
            pe, _ = self._system().energy_forces(frames, precision=self.precision, want_forces=False)
            result["potential_energies"] = pe.cpu().numpy() * unit.kilojoule_per_mole
        if store_W_shad_trace:
            result["W_shad_trace"] = np.diff(res["trace_w"][0, :, 0].cpu().numpy())
        result["W_shad"] = res["W"][0, 0].item()
        return result

    # ---- many trajectories at once (what the nested estimator's inner and outer loops batch into)
    def ensemble_shadow_work(self, x_0, v_0, n_steps, n=None, seed=None, want_final=False):
        """`n` independent trajectories of `n_steps` in ONE launch.  x_0: (n_atoms, 3) shared by all or
        (n, n_atoms, 3); v_0: same shapes, or None for fresh Maxwell-Boltzmann velocities per trajectory.
        Returns the works (kT) as a numpy array, plus final (x, v) arrays when want_final."""
        eq = self.equilibrium_simulator
        t = engine.torch()
        x0 = np.asarray(getattr(x_0, "xyz", x_0), dtype=np.float64)
        if x0.ndim == 2:
            if n is None:
                raise ValueError("give n when x_0 is a single configuration")
            x0 = np.broadcast_to(x0, (int(n),) + x0.shape)
        n = x0.shape[0]
        v0 = None
        if v_0 is not None:
            v0 = np.asarray(v_0.value_in_unit(unit.nanometer / unit.picosecond) if unit.is_quantity(v_0) else v_0, dtype=np.float64)
            v0 = np.broadcast_to(v0, x0.shape)
        rid = _rngstate.reserve_replicas(n)
        res = engine.run_protocol(self._system(), self.integrator.program, n, int(n_steps), self._kT(), n_legs=1,
                                  seed=_rngstate.get_seed() if seed is None else int(seed), replica0=rid,
                                  x0=np.array(x0, dtype=np.float64, order="C"), v0=None if v0 is None else np.array(v0, dtype=np.float64, order="C"),
                                  precision=self.precision, want_heat=True, want_final=want_final, want_moments=False, out={})
        self.integrator._accumulate(heat=res["heat"].nan_to_num(0.0).sum().item())
        W = res["W"][0].cpu().numpy()
        if want_final:
            return W, res["x_final"].cpu().numpy(), res["v_final"].cpu().numpy()
        return W

    # ---- the ensemble
    def collect_protocol_samples(self, n_protocol_samples, protocol_length, marginal="configuration",
                                 store_potential_energy_traces=False, store_W_shad_traces=False, seed=None,
                                 as_numpy=True):
        """Perform nonequilibrium measurements, aimed at measuring the free energy difference for the
        chosen marginal (:247-295): per sample, x_0 ~ cache, v_0 ~ MB, forward leg of `protocol_length`
        steps -> W_F; v redrawn ("configuration") or kept ("full"); reverse leg -> W_R.
        Returns {"W_shads_F", "W_shads_R"[, "W_shad_traces", "potential_energies"]} in kT."""
        if marginal not in ["configuration", "full"]:
            raise NotImplementedError("`marginal` must be either 'configuration' or 'full'")
        eq = self.equilibrium_simulator
        n = int(n_protocol_samples)
        traces = bool(store_potential_energy_traces or store_W_shad_traces)
        seed = _rngstate.get_seed() if seed is None else int(seed)
        rid = _rngstate.reserve_replicas(n)
        res = engine.run_protocol(self._system(), self.integrator.program, n, int(protocol_length), self._kT(),
                                  marginal=marginal, n_legs=2, seed=seed, replica0=rid, x_cache=eq.samples_device(),
                                  precision=self.precision, want_heat=True, traces=traces, out=self._buffers,
                                  round_midpoint=self.round_midpoint)
        self.last_moments = res.get("moments")
        W = res["W"]
        self.integrator._accumulate(heat=res["heat"].nan_to_num(0.0).sum().item())
        inc = pe = None
        if store_W_shad_traces:
            tw = res["trace_w"]  # (leg, step, replica), cumulative
            inc = (tw[:, 1:] - tw[:, :-1]).permute(2, 0, 1)  # (replica, leg, step)
        if store_potential_energy_traces:
            frames = res["trace_x"][1].reshape(-1, self._system().n_atoms, 3)
            pe, _ = self._system().energy_forces(frames, precision=self.precision, want_forces=False)
            pe = pe.reshape(int(protocol_length) + 1, n).transpose(0, 1)
        if as_numpy:
            # asynchronous copies into pinned host memory, one synchronisation (_staging.to_host); the arrays own their memory
            wF, wR, inc, pe = _staging.to_host([W[0], W[1], inc, pe])
        else:
            wF, wR = W[0], W[1]  # views of this simulator's output buffer: overwritten by its next call
        result = {"W_shads_F": wF, "W_shads_R": wR}
        if inc is not None:
            result["W_shad_traces"] = [(inc[i, 0], inc[i, 1]) for i in range(n)]
        if pe is not None:
            result["potential_energies"] = [pe[i] for i in range(n)]
        return result

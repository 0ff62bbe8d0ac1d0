"""One (system, splitting, dt, marginal, collision rate, H-mass factor) experiment as a single ensemble launch.

API-compatible with the reference's canonical caller of the hot path (benchmark/experiments/driver.py:12-82):
the same descriptor fields, `Experiment(descriptor, filename).run() / .save() / .run_and_save()`, and the same
pickle layout {"result": {"W_shads_F", "W_shads_R", ...}, "descriptor": {...without the simulator...}}, so that
benchmark/plotting/plotting_utilities.py:125-176 reads the output unchanged.
"""
import pickle
from collections import namedtuple

from .. import unit
from ..evaluation import estimate_nonequilibrium_free_energy
from ..integrators import LangevinSplittingIntegrator
from ..testsystems import NonequilibriumSimulator
from ..utilities.openmm_utilities import repartition_hydrogen_mass_amber_masses

_FIELDS = ("experiment_name", "system_name", "equilibrium_simulator", "splitting_name", "splitting_string",
           "timestep_in_fs", "marginal", "collision_rate_name", "collision_rate", "n_protocol_samples",
           "protocol_length", "h_mass_factor")
ExperimentDescriptor = namedtuple("ExperimentDescriptor", _FIELDS)


def _apply_hydrogen_mass_factor(eq_sim, factor):
    """Hydrogen-mass repartitioning on the system description (utilities/openmm_utilities.py:223-269); a system
    can be modified once, as in the reference."""
    if getattr(eq_sim, "MODIFIED_H_MASS", False):
        raise Exception("H mass of this system was already modified! Can't guarantee correct behavior")
    desc = dict(eq_sim.system)
    missing = [k for k in ("elements", "bonds_topology") if k not in desc]
    if missing:
        raise Exception("hydrogen-mass repartitioning needs %s in the system description" % " and ".join(missing))
    desc["mass"] = repartition_hydrogen_mass_amber_masses(desc["elements"], desc["bonds_topology"], desc["mass"], factor)
    eq_sim.system = desc  # the setter drops the compiled system tables (masses are part of them)
    eq_sim.MODIFIED_H_MASS = True


class Experiment:
    def __init__(self, experiment_descriptor, filename, store_potential_energy_traces=False):
        self.experiment_descriptor = experiment_descriptor
        self.filename = filename
        self.store_potential_energy_traces = store_potential_energy_traces
        self.result = None
        if experiment_descriptor.h_mass_factor != 1:
            _apply_hydrogen_mass_factor(experiment_descriptor.equilibrium_simulator, experiment_descriptor.h_mass_factor)

    def run(self):
        d = self.experiment_descriptor
        integrator = LangevinSplittingIntegrator(splitting=d.splitting_string, timestep=d.timestep_in_fs * unit.femtosecond,
                                                 collision_rate=d.collision_rate)
        want_pe = bool(self.store_potential_energy_traces) and d.marginal == "full"
        self.result = NonequilibriumSimulator(d.equilibrium_simulator, integrator).collect_protocol_samples(
            d.n_protocol_samples, d.protocol_length, d.marginal, store_potential_energy_traces=want_pe)
        self.DeltaF_neq, self.squared_uncertainty = estimate_nonequilibrium_free_energy(self.result["W_shads_F"],
                                                                                        self.result["W_shads_R"])
        print("%s\n\t%.3f +/- %.3f" % (self, self.DeltaF_neq, self.squared_uncertainty ** 0.5))

    def save(self):
        descriptor = {k: v for k, v in self.experiment_descriptor._asdict().items() if k != "equilibrium_simulator"}
        with open(self.filename, "wb") as f:
            pickle.dump({"result": self.result, "descriptor": descriptor}, f)

    def run_and_save(self):
        self.run()
        self.save()

    def __str__(self):
        d = self.experiment_descriptor
        return "\n\t".join([str(d.system_name), str(d.splitting_name), "dt=%sfs" % d.timestep_in_fs, "marginal: %s" % d.marginal,
                            "collision rate: %s" % d.collision_rate_name, "H-mass scale: %s" % d.h_mass_factor])

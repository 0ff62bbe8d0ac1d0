"""Drop-in for benchmark/experiments/driver.py:12-82 -- the canonical caller of the hot path: one
(system, splitting, dt, marginal, collision rate, H-mass factor) experiment, run as ONE ensemble launch,
saved as the reference's pickle {"result": {"W_shads_F", "W_shads_R"[, ...]}, "descriptor": {...}} so that
benchmark/plotting/plotting_utilities.py:125-176 consumes the output unchanged."""
from collections import namedtuple
from pickle import dump

import numpy as np

from .. import unit
from ..evaluation import estimate_nonequilibrium_free_energy
from ..integrators import LangevinSplittingIntegrator
from ..testsystems import NonequilibriumSimulator
from ..utilities.openmm_utilities import repartition_hydrogen_mass_amber_masses

ExperimentDescriptor = namedtuple("ExperimentDescriptor", ["experiment_name",
                                                           "system_name", "equilibrium_simulator",
                                                           "splitting_name", "splitting_string",
                                                           "timestep_in_fs",
                                                           "marginal",
                                                           "collision_rate_name", "collision_rate",
                                                           "n_protocol_samples", "protocol_length", "h_mass_factor"])


class Experiment():
    def __init__(self, experiment_descriptor, filename, store_potential_energy_traces=False):
        self.experiment_descriptor = experiment_descriptor
        self.filename = filename
        exp = self.experiment_descriptor
        self.store_potential_energy_traces = store_potential_energy_traces

        if exp.h_mass_factor != 1:
            eq = exp.equilibrium_simulator
            if hasattr(eq, "MODIFIED_H_MASS"):
                raise (Exception("H mass of this system was already modified! Can't guarantee correct behavior"))
            desc = dict(eq.system)
            if "elements" not in desc or "bonds_topology" not in desc:
                raise Exception("hydrogen-mass repartitioning needs `elements` and `bonds_topology` in the system description")
            desc["mass"] = repartition_hydrogen_mass_amber_masses(desc["elements"], desc["bonds_topology"], desc["mass"],
                                                                  exp.h_mass_factor)
            eq.system = desc
            eq._engine_system = None  # masses are part of the compiled system tables
            eq.MODIFIED_H_MASS = True

    def run(self):
        exp = self.experiment_descriptor
        simulator = NonequilibriumSimulator(exp.equilibrium_simulator,
                                            LangevinSplittingIntegrator(
                                                splitting=exp.splitting_string,
                                                timestep=exp.timestep_in_fs * unit.femtosecond,
                                                collision_rate=exp.collision_rate))

        self.result = simulator.collect_protocol_samples(
            exp.n_protocol_samples, exp.protocol_length, exp.marginal,
            store_potential_energy_traces=(exp.marginal == "full" and self.store_potential_energy_traces))

        W_shads_F, W_shads_R = self.result["W_shads_F"], self.result["W_shads_R"]
        DeltaF_neq, squared_uncertainty = estimate_nonequilibrium_free_energy(W_shads_F, W_shads_R)
        self.DeltaF_neq, self.squared_uncertainty = DeltaF_neq, squared_uncertainty
        print(self)
        print("\t{:.3f} +/- {:.3f}".format(DeltaF_neq, np.sqrt(squared_uncertainty)))

    def save(self):
        with open(self.filename, "wb") as f:
            everything_but_the_simulator = self.experiment_descriptor._asdict()
            everything_but_the_simulator.pop("equilibrium_simulator")
            dump({"result": self.result, "descriptor": everything_but_the_simulator}, f)

    def run_and_save(self):
        self.run()
        self.save()

    def __str__(self):
        exp = self.experiment_descriptor
        properties = [exp.system_name,
                      exp.splitting_name,
                      "dt={}fs".format(exp.timestep_in_fs),
                      "marginal: {}".format(exp.marginal),
                      "collision rate: {}".format(exp.collision_rate_name),
                      "H-mass scale: {}".format(exp.h_mass_factor)
                      ]
        return "\n\t".join(["{}"] * len(properties)).format(*properties)

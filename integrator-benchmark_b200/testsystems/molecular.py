"""The reference's molecular EquilibriumSimulator instances, rebuilt on plain system descriptions:
water_cluster_rigid / water_cluster_flexible (watercluster.py:87-108), tiny_waterbox
(waterbox.py:36-48), alanine_constrained / alanine_unconstrained (alanine_dipeptide.py:41-56).
Same names, temperature (298 K), timesteps, burn-in lengths and cache sizes.  Constructing them is
free (no GPU work until samples are requested)."""
from .. import unit
from . import systems
from .bookkeepers import EquilibriumSimulator
from .configuration import configure_platform

temperature = 298 * unit.kelvin
simulation_parameters = {"temperature": temperature, "pressure": 1 * unit.atmosphere, "tolerance": 1.0 / (10 ** 8)}


def _make(builder, name, platform, timestep_fs, burn_in_length, n_samples=1000, thinning_interval=10000, **kw):
    desc, positions = builder(**kw)
    return EquilibriumSimulator(platform=configure_platform(platform), topology=desc.get("name"), system=desc,
                                positions=positions, temperature=temperature, timestep=timestep_fs * unit.femtosecond,
                                burn_in_length=burn_in_length, n_samples=n_samples,
                                thinning_interval=thinning_interval, name=name)


def load_alanine(constrained=True):
    desc, positions = systems.alanine_dipeptide(constrained=constrained)
    return desc.get("name"), desc, positions


water_cluster_rigid = _make(systems.water_cluster, "water_cluster_rigid", "Reference", 1.0, 100000, constrained=True)
water_cluster_flexible = _make(systems.water_cluster, "water_cluster_flexible", "Reference", 1.0, 100000, constrained=False)
tiny_waterbox = _make(systems.tiny_waterbox, "tiny_waterbox_constrained", "CUDA", 1.0, 100000, constrained=True)
alanine_constrained = _make(systems.alanine_dipeptide, "alanine_constrained", "Reference", 1.0, 50000, constrained=True)
alanine_unconstrained = _make(systems.alanine_dipeptide, "alanine_unconstrained", "Reference", 1.0, 50000,
                              constrained=False)
system_loaders = {"alanine": load_alanine}

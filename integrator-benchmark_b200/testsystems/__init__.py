from .low_dimensional_systems import (NumbaBookkeepingSimulator, NumbaNonequilibriumSimulator, double_well,  # noqa: F401
                                      quartic)
from .bookkeepers import EquilibriumSimulator, NonequilibriumSimulator  # noqa: F401
from .configuration import configure_platform  # noqa: F401
from .molecular import (alanine_constrained, alanine_unconstrained, load_alanine, system_loaders,  # noqa: F401
                        tiny_waterbox, water_cluster_flexible, water_cluster_rigid)

from .low_dimensional_systems import (NumbaBookkeepingSimulator, NumbaNonequilibriumSimulator, double_well,  # noqa: F401
                                      quartic)

from .numba_integrators import (aboba_factory, baoab_factory, orvro_factory, splitting_factory,  # noqa: F401
                                vvvr_factory)
from .langevin import ContinuousLangevinSplittingIntegrator, LangevinSplittingIntegrator  # noqa: F401
from .ghmc import CustomizableGHMC  # noqa: F401

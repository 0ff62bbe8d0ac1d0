from .analysis import estimate_from_moments, estimate_nonequilibrium_free_energy  # noqa: F401
from . import compare_near_eq_and_exact  # noqa: F401
from . import thresholds  # noqa: F401

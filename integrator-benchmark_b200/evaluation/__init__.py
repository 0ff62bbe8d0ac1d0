from .analysis import estimate_from_moments, estimate_nonequilibrium_free_energy  # noqa: F401

from .driver import Experiment, ExperimentDescriptor  # noqa: F401

"""B200-native engine for the replica-ensemble / shadow-work path of
choderalab/integrator-benchmark, behind the reference's Python API.

Everything numeric runs in liblsi.so (hand-written sm_100a CUDA behind the C ABI of
include/lsi.h); importing the package loads that library and fails if it is missing.
"""
from . import _cabi, constants, engine, unit  # noqa: F401
from ._rngstate import get_seed, set_seed  # noqa: F401
from . import evaluation, integrators, testsystems, parallel, utilities, experiments  # noqa: F401,E402

__version__ = "0.1.0"

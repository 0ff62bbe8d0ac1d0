"""B200-native engine for the replica-ensemble / shadow-work path of
choderalab/integrator-benchmark, behind the reference's Python API.

Everything numeric runs in liblsi.so (hand-written sm_100a CUDA behind the C ABI of
include/lsi.h); importing the package loads that library and fails if it is missing.
"""
import os

# Where equilibrium-sample caches and experiment outputs go (benchmark/__init__.py:3-11); override with LSI_DATA_PATH.
DATA_PATH = os.environ.get("LSI_DATA_PATH", os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "data"))
FIGURE_PATH = os.environ.get("LSI_FIGURE_PATH", os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "figures"))

from . import _cabi, constants, engine, unit  # noqa: F401,E402
from ._rngstate import get_seed, set_seed  # noqa: F401,E402
from . import evaluation, integrators, testsystems, parallel, utilities, experiments  # noqa: F401,E402

__version__ = "0.1.0"

"""`openmmtools.constants` stand-in (reference: benchmark/integrators/langevin.py:6)."""
from . import unit

kB = unit.MOLAR_GAS_CONSTANT_R

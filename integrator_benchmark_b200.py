"""Import alias: the package directory is named `integrator-benchmark_b200` (not a valid
Python identifier); this module makes it importable as `integrator_benchmark_b200`."""
import os as _os

__path__ = [_os.path.join(_os.path.dirname(_os.path.abspath(__file__)), "integrator-benchmark_b200")]
__package__ = __name__
if __spec__ is not None:
    __spec__.submodule_search_locations = __path__
__file__ = _os.path.join(__path__[0], "__init__.py")
with open(__file__) as _f:
    exec(compile(_f.read(), __file__, "exec"))

"""ctypes binding of liblsi.so (include/lsi.h).  This is the whole FFI surface: the
Python host code above it only passes plain numbers, strings and raw device pointers.

The library is loaded from integrator-benchmark_b200/lib/liblsi.so (built in-tree by
build.py).  There is no fallback: if it cannot be loaded, importing this module raises.
"""
import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
# LSI_LIB_PATH: load another build of the same library (kernel experiments, scripts/build_variant.sh)
LIB_PATH = os.environ.get("LSI_LIB_PATH") or os.path.join(_HERE, "lib", "liblsi.so")

OK, ERR_ASSERT, ERR_VALUE, ERR_ARG, ERR_CUDA, ERR_UNSUPPORTED, ERR_KEY = 0, -1, -2, -3, -4, -5, -6
OP_R, OP_V, OP_O = 0, 1, 2
POT_QUARTIC, POT_DOUBLE_WELL, POT_HARMONIC = 0, 1, 2
MARGINAL = {"configuration": 0, "full": 1}
PRECISION = {"double": 0, "mixed": 1}
TAG_O, TAG_V0, TAG_VMID, TAG_X0, TAG_BOOT, TAG_EQ = range(6)

c_double_p = ctypes.POINTER(ctypes.c_double)
c_int64_p = ctypes.POINTER(ctypes.c_int64)


class LsiCudaError(RuntimeError):
    pass


class ProtocolArgs(ctypes.Structure):
    """struct lsi_protocol_args (include/lsi.h)"""
    _fields_ = [
        ("seed", ctypes.c_uint64), ("replica0", ctypes.c_uint64), ("n_replicas", ctypes.c_int64),
        ("n_steps", ctypes.c_int32), ("n_legs", ctypes.c_int32), ("marginal", ctypes.c_int32),
        ("precision", ctypes.c_int32), ("flags", ctypes.c_int32), ("reserved", ctypes.c_int32), ("kT", ctypes.c_double),
        ("x_cache", ctypes.c_void_p), ("n_cache", ctypes.c_int64), ("x0", ctypes.c_void_p), ("v0", ctypes.c_void_p),
        ("W", ctypes.c_void_p), ("heat", ctypes.c_void_p), ("W_direct", ctypes.c_void_p),
        ("x_final", ctypes.c_void_p), ("v_final", ctypes.c_void_p),
        ("trace_x", ctypes.c_void_p), ("trace_v", ctypes.c_void_p), ("trace_w", ctypes.c_void_p),
        ("moments", ctypes.c_void_p), ("workspace", ctypes.c_void_p), ("workspace_bytes", ctypes.c_int64),
    ]


c_int32_pp = ctypes.POINTER(ctypes.c_int32)


class ParticlesDesc(ctypes.Structure):
    """struct lsi_particles_desc (include/lsi.h)"""
    _fields_ = [
        ("n_atoms", ctypes.c_int32), ("mass", c_double_p), ("charge", c_double_p), ("sigma", c_double_p),
        ("epsilon", c_double_p), ("nonbonded_method", ctypes.c_int32), ("cutoff", ctypes.c_double),
        ("box", ctypes.c_double * 3), ("rf_dielectric", ctypes.c_double),
        ("n_exclusions", ctypes.c_int32), ("exclusions", c_int32_pp),
        ("n_exceptions", ctypes.c_int32), ("exception_atoms", c_int32_pp), ("exception_params", c_double_p),
        ("n_bonds", ctypes.c_int32), ("bond_atoms", c_int32_pp), ("bond_params", c_double_p),
        ("n_angles", ctypes.c_int32), ("angle_atoms", c_int32_pp), ("angle_params", c_double_p),
        ("n_torsions", ctypes.c_int32), ("torsion_atoms", c_int32_pp), ("torsion_params", c_double_p),
        ("restraint_k", ctypes.c_double),
        ("group_nonbonded", ctypes.c_int32), ("group_bonds", ctypes.c_int32), ("group_angles", ctypes.c_int32),
        ("group_torsions", ctypes.c_int32), ("group_restraint", ctypes.c_int32),
        ("n_settle", ctypes.c_int32), ("settle_atoms", c_int32_pp), ("settle_oh", ctypes.c_double),
        ("settle_hh", ctypes.c_double),
        ("n_constraints", ctypes.c_int32), ("constraint_atoms", c_int32_pp), ("constraint_dist", c_double_p),
        ("constraint_tolerance", ctypes.c_double), ("switch_distance", ctypes.c_double),
    ]


def _load():
    if not os.path.exists(LIB_PATH):
        # build in-tree when a toolchain is present (the build container); never fall back to CPU code
        import importlib.util
        spec = importlib.util.spec_from_file_location("_lsi_build", os.path.join(_HERE, "build.py"))
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
        mod.build()
    lib = ctypes.CDLL(LIB_PATH)
    vp, i32, i64, u64, dbl = ctypes.c_void_p, ctypes.c_int, ctypes.c_int64, ctypes.c_uint64, ctypes.c_double
    sig = {
        "lsi_last_error": (ctypes.c_char_p, []),
        "lsi_abi_version": (i32, []),
        "lsi_device_count": (i32, []),
        "lsi_compile": (i32, [ctypes.c_char_p, dbl, dbl, i32, i32, i32, ctypes.POINTER(vp)]),
        "lsi_compile_continuous": (i32, [ctypes.POINTER(ctypes.c_char_p), c_double_p, i32, dbl, dbl, i32, i32, i32,
                                         ctypes.POINTER(vp)]),
        "lsi_program_num_ops": (i32, [vp]),
        "lsi_program_is_mts": (i32, [vp]),
        "lsi_program_num_O": (i32, [vp]),
        "lsi_program_get_op": (i32, [vp, i32, ctypes.POINTER(i32), ctypes.POINTER(i32), c_double_p, c_double_p, c_double_p]),
        "lsi_program_get_step_size": (dbl, [vp]),
        "lsi_program_set_step_size": (i32, [vp, dbl]),
        "lsi_program_flags": (i32, [vp, ctypes.POINTER(i32), ctypes.POINTER(i32)]),
        "lsi_program_destroy": (None, [vp]),
        "lsi_system_create_scalar": (i32, [i32, c_double_p, dbl, ctypes.POINTER(vp)]),
        "lsi_system_create_particles": (i32, [ctypes.POINTER(ParticlesDesc), ctypes.POINTER(vp)]),
        "lsi_particles_energy_forces": (i32, [vp, ctypes.c_int32, ctypes.c_int32, i64, vp, vp, vp, vp]),
        "lsi_system_destroy": (None, [vp]),
        "lsi_system_num_dof": (i32, [vp]),
        "lsi_protocol_workspace_bytes": (i64, [i64]),
        "lsi_run_protocol": (i32, [vp, vp, ctypes.POINTER(ProtocolArgs), vp]),
        "lsi_run_protocol_batch": (i32, [vp, ctypes.POINTER(vp), i32, ctypes.POINTER(ProtocolArgs), vp]),
        "lsi_sample_equilibrium_scalar": (i32, [vp, dbl, i64, ctypes.c_int32, u64, u64, vp, vp]),
        "lsi_sample_equilibrium_particles": (i32, [vp, ctypes.c_int32, dbl, i64, ctypes.c_int32, ctypes.c_int32, ctypes.c_int32, dbl,
                                                   u64, u64, vp, vp, vp, vp, vp]),
        "lsi_reduce_moments": (i32, [vp, vp, i64, vp, vp, i64, vp]),
        "lsi_neq_free_energy": (i32, [c_double_p, dbl, c_double_p, c_double_p]),
        "lsi_reduce_logmeanexp": (i32, [vp, vp, i64, vp, vp, vp]),
        "lsi_microbench_fma": (i32, [i32, i64, i32, vp, vp]),
        "lsi_test_fastmath": (i32, [i32, i64, vp, vp, vp, vp]),
        "lsi_bootstrap_nested": (i32, [vp, vp, i64, ctypes.c_int32, u64, vp, vp]),
    }
    for name, (res, args) in sig.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    return lib, sorted(sig)


lib, EXPORTED = _load()

_EXC = {ERR_ASSERT: AssertionError, ERR_VALUE: ValueError, ERR_ARG: ValueError, ERR_CUDA: LsiCudaError,
        ERR_UNSUPPORTED: NotImplementedError, ERR_KEY: KeyError}


def check(rc):
    """Raise the exception type the reference raises at the same place."""
    if rc != OK:
        msg = (lib.lsi_last_error() or b"").decode()
        raise _EXC.get(rc, RuntimeError)(msg)
    return rc

"""Host-side checks that need no GPU: the C-ABI library loads, exports what include/lsi.h
declares, and its splitting compiler agrees with the reference's own constructor."""
import ctypes
import json
import os
import re
import sys

import pytest

import integrator_benchmark_b200 as ib
from integrator_benchmark_b200 import _cabi, engine
from tests.test_oracle_golden import _expected_ops_from_recorded

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_functions():
    text = open(os.path.join(ROOT, "include", "lsi.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(lsi_[A-Za-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    names = declared_functions()
    assert len(names) >= 20
    lib = ctypes.CDLL(_cabi.LIB_PATH)
    for n in names:
        assert hasattr(lib, n), "liblsi.so does not export %s" % n
    # and the ctypes binding covers each of them
    assert set(names) == set(_cabi.EXPORTED)
    assert _cabi.lib.lsi_abi_version() == 3


def test_compiler_matches_reference_constructor(golden_dir):
    data = json.load(open(os.path.join(golden_dir, "programs.json")))
    n_ok = n_err = 0
    for e in data["entries"]:
        kw = e["kwargs"]
        args = dict(dt=kw["timestep"], gamma=kw["collision_rate"],
                    override_splitting_checks=kw.get("override_splitting_checks", False),
                    measure_shadow_work=kw.get("measure_shadow_work", False),
                    measure_heat=kw.get("measure_heat", True))
        splitting = e["splitting"] if e["kind"] == "discrete" else [tuple(t) for t in e["splitting"]]
        if "error" in e:
            with pytest.raises(Exception) as ei:
                engine.Program(splitting, **args)
            assert type(ei.value).__name__ == e["error"], (splitting, kw)
            n_err += 1
            continue
        expected = _expected_ops_from_recorded(e)
        if any(op[0] == _cabi.OP_V and op[2] == float("inf") for op in expected):
            with pytest.raises(ValueError):  # documented deviation (lsi.h)
                engine.Program(splitting, **args)
            continue
        ops = engine.Program(splitting, **args).ops()
        assert len(ops) == len(expected), (splitting, ops, expected)
        for got, exp in zip(ops, expected):
            assert got[0] == exp[0] and got[1] == exp[1], (splitting, got, exp)
            if exp[2] is not None:
                assert got[2] == pytest.approx(exp[2], rel=1e-15)
            if exp[0] == _cabi.OP_O:
                assert got[3] == pytest.approx(exp[3], rel=1e-14) and got[4] == pytest.approx(exp[4], rel=1e-14)
        n_ok += 1
    assert n_ok > 60 and n_err > 20


def test_set_step_size_keeps_O_coefficients():
    """setStepSize rescales R and V only (langevin.py:192-196 bakes a, b)."""
    p = engine.Program("V R O R V", 0.001, 10.0)
    before = p.ops()
    p.dt = 0.004
    assert p.dt == 0.004
    assert p.ops() == before


def test_no_gpu_means_loud_failure():
    if _cabi.lib.lsi_device_count() > 0:
        pytest.skip("a GPU is present")
    with pytest.raises(_cabi.LsiCudaError):
        engine.require_cuda()
    from integrator_benchmark_b200.evaluation import estimate_nonequilibrium_free_energy
    with pytest.raises(_cabi.LsiCudaError):
        estimate_nonequilibrium_free_energy([0.0, 1.0], [1.0, 0.0])


def test_product_does_not_import_oracle():
    pkg = os.path.join(ROOT, "integrator-benchmark_b200")
    for dirpath, _dirs, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h")):
                text = open(os.path.join(dirpath, f)).read()
                for pat in (r"^\s*(from|import)\s+oracle", r"liblsi_oracle", r"oracle[/.]", r"#include.*oracle"):
                    assert not re.search(pat, text, flags=re.M), (pat, os.path.join(dirpath, f))
    assert ib.__version__


def test_particle_system_validation_is_host_only_and_loud():
    """lsi_system_create_particles copies and checks its tables on the host (no CUDA call): bad descriptions are
    rejected with the boundary's error codes, and compute entry points fail loudly without a device."""
    import numpy as np
    import integrator_benchmark_b200 as ib
    from integrator_benchmark_b200.testsystems import systems as ts
    desc, x = ts.water_cluster(4, constrained=True)
    system = ib.engine.ParticleSystem(desc)            # fine without a GPU
    assert system.n_dof == 36
    with pytest.raises(ValueError):                    # atom index out of range
        ib.engine.ParticleSystem(dict(desc, exclusions=[(0, 99)]))
    with pytest.raises(ValueError):                    # a pair constraint on an atom of a rigid water
        ib.engine.ParticleSystem(dict(desc, constraint_atoms=[(0, 3)], constraint_dist=[0.3]))
    with pytest.raises(ValueError):                    # impossible triangle
        ib.engine.ParticleSystem(dict(desc, settle_hh=0.3))
    with pytest.raises(ValueError):                    # massless site
        ib.engine.ParticleSystem(dict(desc, mass=np.zeros(12)))
    with pytest.raises(ValueError):                    # box smaller than twice the cutoff
        ib.engine.ParticleSystem(dict(desc, nonbonded_method="CutoffPeriodic", cutoff=0.6, box=(1.0, 1.0, 1.0)))
    with pytest.raises(NotImplementedError):           # waters with different masses cannot share the SETTLE constants
        m = np.array(desc["mass"], dtype=float)
        m[3] = 18.0
        ib.engine.ParticleSystem(dict(desc, mass=m))
    big, _ = ts.tiny_waterbox(n_waters=300, box_edge=2.2)
    with pytest.raises(ValueError):                    # more atoms than one CTA holds
        ib.engine.ParticleSystem(big)
    if ib._cabi.lib.lsi_device_count() == 0:
        with pytest.raises(ib._cabi.LsiCudaError):
            system.energy_forces(x[None])
        with pytest.raises(ib._cabi.LsiCudaError):
            ib.engine.run_protocol(system, ib.engine.Program("V R O R V", 0.001, 1.0), 2, 3, 2.5, x0=np.repeat(x[None], 2, 0))


def test_bench_reference_arm_runs_without_a_gpu():
    """`bench.py --impl reference` times the reference's own numba closures (oracle/_ref, when installed; otherwise the CPU
    oracle port) on the host cores and prints the contract's JSON line (no CUDA needed)"""
    import json
    import subprocess
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "2", "--warmup", "1",
                          "--workload", "double_well", "--ref-replicas", "2048"], capture_output=True, text=True, timeout=300, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    line = json.loads(out.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["value"] > 0 and line["unit"] == "replica-steps/s"
    assert line["e2e"] == {"value": line["value"], "unit": line["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    have_ref = os.path.exists(os.path.join(ROOT, "oracle", "_ref", "numba_integrators.py"))
    assert line["cpu_baseline"]["kind"] == ("reference" if have_ref else "port") and line["cpu_baseline"]["value"] == line["value"]
    assert line["higher_is_better"] is True and line["steps"] == 2 and line["gpu_launches"] == 0
    assert line["config"]["replicas_per_gpu"] == 1 << 20 and "C2 double well" in line["config"]["workload"]


def test_bench_workload_table_and_flop_convention():
    """bench.py's workload table covers BASELINE.json's five configs and its algorithmic-flop convention is the one of SURVEY 8d
    (O 9, V 2, R 2 per substep, one force evaluation per distinct position visited)"""
    import importlib.util
    spec = importlib.util.spec_from_file_location("bench_module", os.path.join(ROOT, "bench.py"))
    bench = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(bench)
    assert bench.ORDER == ["double_well", "quartic", "water_cluster", "alanine", "waterbox"]
    assert [bench.WORKLOADS[k]["config"] for k in bench.ORDER] == ["C2", "C1", "C3", "C4", "C5"]
    W = bench.WORKLOADS["double_well"]
    assert W["replicas"] == 1 << 20 and W["protocol_length"] == 10 and len(W["splittings"]) == 6
    assert bench.steps_per_leg(W) == 9 and bench.steps_per_leg(bench.WORKLOADS["waterbox"]) == 200
    fl = {k: bench.flops_per_step(s, W["f_force"]) for k, s in W["splittings"].items()}
    assert fl == {"OVRVO": 34, "ORVRO": 34, "RVOVR": 27, "ROVOR": 34, "VRORV": 27, "VOROV": 34} and sum(fl.values()) == 190
    assert bench.flops_per_step("O V R R V O", 3) == 18 + 4 + 4 + 3 and bench.flops_per_step("V R O O R V", 3) == 4 + 4 + 18 + 3
    cfg = bench.workload_config("double_well", W, W["replicas"], W["protocol_length"])
    assert "C2 double well" in cfg["workload"] and "model" not in cfg and cfg["replicas_per_gpu"] == 1 << 20

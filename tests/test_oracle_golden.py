"""The CPU oracle against fixtures produced by the reference's own code
(tests/golden/make_golden.py) and against published known answers."""
import json
import os

import numpy as np
import pytest

from oracle import oracle as orc
from oracle import splitting as sp

# numba closures (numba_integrators.py) expressed as splitting strings:
# two half-O draws per step in baoab/aboba, two half-R in vvvr, one full V in orvro
SCHEME_STRING = {"vvvr": "O V R R V O", "orvro": "O R V R O", "baoab": "V R O O R V", "aboba": "R V O O V R"}


def test_philox_known_answers():
    # Random123 kat_vectors, philox4x32 10 rounds
    assert orc.philox4x32_10([0, 0, 0, 0], [0, 0]) == [0x6627E8D5, 0xE169C58D, 0xBC57AC4C, 0x9B00DBD8]
    assert orc.philox4x32_10([0xFFFFFFFF] * 4, [0xFFFFFFFF] * 2) == [0x408F276D, 0x41C83B0E, 0xA20BC7C6, 0x6D5451FD]
    assert orc.philox4x32_10([0x243F6A88, 0x85A308D3, 0x13198A2E, 0x03707344],
                             [0xA4093822, 0x299F31D0]) == [0xD16CFE09, 0x94FDCCEB, 0x5001E420, 0x24126EA1]


def test_normals_are_standard():
    z = np.concatenate([orc.normals4(7, r, 0, 0, 0) for r in range(5000)])
    assert abs(z.mean()) < 0.03 and abs(z.std() - 1.0) < 0.03
    assert abs(np.mean(z ** 4) - 3.0) < 0.2


def test_toy_potentials_known_values():
    # partition function of the double well on [-2.25, 2.25]: notebooks/1D figures, updated.ipynb:415
    xs = np.linspace(-2.25, 2.25, 200001)
    u = np.array([orc.toy_potential("double_well", x) for x in xs[::20]])
    Z = np.trapezoid(np.exp(-u), xs[::20])
    assert abs(Z - 4.4848163908) < 1e-6
    for x in [-1.3, -0.2, 0.0, 0.7, 1.9]:
        for kind in ["quartic", "double_well"]:
            h = 1e-6
            fd = -(orc.toy_potential(kind, x + h) - orc.toy_potential(kind, x - h)) / (2 * h)
            assert abs(fd - orc.toy_force(kind, x)) < 1e-6 * max(1.0, abs(fd))


def _toy_cases(golden_dir):
    z = np.load(os.path.join(golden_dir, "toy_numba.npz"))
    return z, json.loads(str(z["meta"]))


def test_oracle_matches_reference_numba_closures(golden_dir):
    """oracle (replayed noise) == reference numba_integrators.py bodies, all 4 schemes x 2 systems."""
    z, meta = _toy_cases(golden_dir)
    assert len(meta) == 32
    for c in meta:
        k = c["key"]
        n = c["n_steps"]
        noise = []
        for blk in range((2 * (n - 1) + 3) // 4):
            noise.extend(orc.normals4(c["seed"], c["replica"], blk, 0, 0))
        tx, tv, Q, tw = orc.toy_replay(c["system"], SCHEME_STRING[c["scheme"]], c["dt"], c["gamma"], c["x0"], c["v0"],
                                       n - 1, noise, mass=c["mass"], kT=1.0 / c["beta"])
        np.testing.assert_allclose(tx, z[k + "_xs"], rtol=1e-12, atol=1e-13)
        np.testing.assert_allclose(tv, z[k + "_vs"], rtol=1e-12, atol=1e-13)
        np.testing.assert_allclose(tw, z[k + "_W"], rtol=1e-10, atol=1e-11)
        np.testing.assert_allclose(Q, float(z[k + "_Q"]), rtol=1e-10, atol=1e-11)


def test_oracle_philox_protocol_equals_replay(golden_dir):
    """the Philox-driven protocol consumes the scalar stream exactly as the replay does."""
    z, meta = _toy_cases(golden_dir)
    for c in meta[::3]:
        n = c["n_steps"]
        out = orc.toy_protocol(c["system"], SCHEME_STRING[c["scheme"]], c["dt"], c["gamma"], 1, n - 1,
                               mass=c["mass"], kT=1.0 / c["beta"], seed=c["seed"], replica0=c["replica"],
                               x0=[c["x0"]], v0=[c["v0"]], n_legs=1, traces=True)
        np.testing.assert_allclose(out["trace_x"][0, 0], z[c["key"] + "_xs"], rtol=1e-12, atol=1e-13)
        np.testing.assert_allclose(out["trace_w"][0, 0], z[c["key"] + "_W"], rtol=1e-10, atol=1e-11)
        np.testing.assert_allclose(out["W"][0, 0], z[c["key"] + "_W"][-1], rtol=1e-10, atol=1e-11)


# ----------------------------------------------------------------------------- recorded step programs
def _expected_ops_from_recorded(entry):
    """Decode the recorded CustomIntegrator program back into (kind, group, frac, a, b)."""
    import re
    ops = []
    g = entry["globals"]
    for op in entry["program"]:
        if op[0] != "perdof":
            continue
        var, expr = op[1], op[2]
        if var == "x" and "v" in expr:
            m = re.match(r"x \+ \(\(dt ([/*]) ([0-9.e+-]+)\) \* v\)", expr)
            val = float(m.group(2))
            ops.append((sp.R, -1, 1.0 / val if m.group(1) == "/" else val, 0.0, 0.0))
        elif var == "v" and "gaussian" in expr:
            m = re.match(r"\(([0-9.e+-]+|a) \* v\) \+ \(([0-9.e+-]+|b) \* sqrt", expr)
            a = g["a"] if m.group(1) == "a" else float(m.group(1))
            b = g["b"] if m.group(2) == "b" else float(m.group(2))
            ops.append((sp.O, -1, None, a, b))
        elif var == "v" and re.search(r"\* f\d* / m", expr):
            m = re.match(r"v \+ \(?\(dt ([/*]) ([0-9.e+-]+)\) \* f(\d*) / m\)?", expr)
            val = float(m.group(2))
            grp = int(m.group(3)) if m.group(3) else -1
            ops.append((sp.V, grp, (1.0 / val if val != 0 else float("inf")) if m.group(1) == "/" else val, 0.0, 0.0))
    return ops


def test_splitting_compiler_matches_reference_constructor(golden_dir):
    data = json.load(open(os.path.join(golden_dir, "programs.json")))
    n_checked = 0
    for e in data["entries"]:
        kw = e["kwargs"]
        args = dict(dt=kw["timestep"], gamma=kw["collision_rate"],
                    override_splitting_checks=kw.get("override_splitting_checks", False),
                    measure_shadow_work=kw.get("measure_shadow_work", False),
                    measure_heat=kw.get("measure_heat", True))
        fn = sp.compile_splitting if e["kind"] == "discrete" else sp.compile_continuous
        splitting = e["splitting"] if e["kind"] == "discrete" else [tuple(t) for t in e["splitting"]]
        if "error" in e:
            with pytest.raises(Exception) as ei:
                fn(splitting, **args)
            assert type(ei.value).__name__ == e["error"], (splitting, kw)
            continue
        expected = _expected_ops_from_recorded(e)
        if any(op[0] == sp.V and op[2] == float("inf") for op in expected):
            # reference quirk Q2 ("dt / 0" for a lone V<g> token): the oracle refuses instead
            with pytest.raises(ZeroDivisionError):
                fn(splitting, **args)
            continue
        prog = fn(splitting, **args)
        assert len(prog.ops) == len(expected), (splitting, prog.ops, expected)
        for got, exp in zip(prog.ops, expected):
            assert got[0] == exp[0] and got[1] == exp[1], (splitting, got, exp)
            if exp[2] is not None:
                assert got[2] == pytest.approx(exp[2], rel=1e-15)
            if exp[0] == sp.O:
                assert got[3] == pytest.approx(exp[3], rel=1e-15) and got[4] == pytest.approx(exp[4], rel=1e-15)
        n_checked += 1
    assert n_checked > 60

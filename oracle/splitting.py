"""CPU restatement of the reference's splitting-string "compiler".

TEST INFRASTRUCTURE ONLY (see oracle/lsi_oracle.c header): the product has its
own compiler in integrator-benchmark_b200/csrc/program.cpp and never imports
this file.

Follows /root/reference/benchmark/integrators/langevin.py:
  * tokenisation and substep counts           :89-94
  * multi-timestep detection, per-group n_Vs  :98-105
  * sanity checks and their exception types   :108-121
  * substep lengths dt/n_R, dt/n_V, dt/n_Vs   :130,155,157
  * O coefficients from h = dt/max(1, n_O)    :192-196
and, for the explicit-fraction variant, langevin.py:277-302,361-364,406-408.

Pinned by tests/golden/programs.json, recorded from the reference's own
constructor (tests/golden/make_golden.py).
"""
import math

R, V, O = 0, 1, 2  # op kinds shared with lsi_oracle.c


class Program:
    """ops: list of (kind, group, frac, a, b); group == -1 means "all forces"."""

    def __init__(self, ops, mts, measure_shadow_work, measure_heat):
        self.ops = ops
        self.mts = mts
        self.measure_shadow_work = measure_shadow_work
        self.measure_heat = measure_heat

    @property
    def n_O(self):
        return sum(1 for op in self.ops if op[0] == O)


def compile_splitting(splitting, dt, gamma, override_splitting_checks=False,
                      measure_shadow_work=False, measure_heat=True):
    tokens = splitting.upper().split()
    n_R = sum(t == "R" for t in tokens)
    n_V = sum(t == "V" for t in tokens)  # bare "V" only (langevin.py:93)
    n_O = sum(t == "O" for t in tokens)

    v_tokens = [t for t in tokens if t[0] == "V"]
    mts = len(set(v_tokens)) > 1
    n_Vs = {}
    if mts:
        for t in set(v_tokens):
            # counts every token whose tail equals the group label (langevin.py:103);
            # for the bare "V" (label "") that also matches "R" and "O"
            n_Vs[t[1:]] = sum(s[1:] == t[1:] for s in tokens)

    if not override_splitting_checks:
        assert "R" in tokens
        assert "V" in [t[0] for t in tokens]
        assert "O" in tokens
        allowed = set("RVO") | {"V%d" % i for i in range(32)}
        assert set(tokens).issubset(allowed)
        if mts and n_V > 0:
            raise ValueError("Splitting string includes an evaluation of all forces and "
                             "evaluation of subsets of forces.")

    h = dt / max(1, n_O)
    a = math.exp(-gamma * h)
    b = math.sqrt(1.0 - math.exp(-2.0 * gamma * h))

    ops = []
    for t in tokens:
        if t == "O":
            ops.append((O, -1, 1.0 / max(1, n_O), a, b))
        elif t == "R":
            ops.append((R, -1, 1.0 / n_R, 0.0, 0.0))
        elif t[0] == "V":
            label = t[1:]
            if mts:
                frac = 1.0 / n_Vs[label]
            else:
                if n_V == 0:
                    raise ZeroDivisionError("dt / n_V with n_V == 0 (single force-group token, langevin.py:157)")
                frac = 1.0 / n_V
            ops.append((V, int(label) if label else -1, frac, 0.0, 0.0))
        # any other token is silently skipped by substep_function (langevin.py:177-183)
    return Program(ops, mts, measure_shadow_work, measure_heat)


def compile_continuous(splitting, dt, gamma, override_splitting_checks=False,
                       measure_shadow_work=False, measure_heat=True):
    """langevin.py:223-408: list of (token, fraction) pairs."""
    mts = len(set(s for s in splitting if s[0] == "V")) > 1  # quirk: compares whole tuples' first item
    if not override_splitting_checks:
        assert "R" in [s[0] for s in splitting]
        assert "V" in [s[0][0] for s in splitting]
        assert "O" in [s[0] for s in splitting]
        lengths = {}
        for step, length in splitting:
            lengths[step] = lengths.get(step, 0.0) + length
        assert lengths["R"] == 1.0
        assert lengths["O"] == 1.0
        assert lengths["V"] == 1.0
    ops = []
    for step, length in splitting:
        if not length > 0:
            continue
        if step == "O":
            h = dt * length
            ops.append((O, -1, length, math.exp(-gamma * h), math.sqrt(1.0 - math.exp(-2.0 * gamma * h))))
        elif step == "R":
            ops.append((R, -1, length, 0.0, 0.0))
        elif step[0] == "V":
            label = step[1:]
            ops.append((V, int(label) if label else -1, length, 0.0, 0.0))
    return Program(ops, mts, measure_shadow_work, measure_heat)

"""Multi-timestep splitting-string generators
(reference: benchmark/utilities/mts_utilities.py:88-108,184-224,283-307; of the functions the
reference defines twice, the later definition is the live one).

The strings are inputs of the splitting compiler and are taken literally: merging or
reordering tokens changes the substep lengths, so nothing here canonicalises.
"""
import itertools


def _v(groups):
    return ["V%d" % g for g in groups]


def generate_gbaoab_string(K_r=1):
    """'V R..R O R..R V' with K_r drifts on each side."""
    return " ".join(["V"] + ["R"] * K_r + ["O"] + ["R"] * K_r + ["V"])


def generate_gbaoab_solvent_solute_string(K_p=2, K_r=1, slow_group=[0], fast_group=[1]):
    """slow (fast R^K_r O R^K_r fast)^K_p slow"""
    inner = _v(fast_group) + ["R"] * K_r + ["O"] + ["R"] * K_r + _v(fast_group)
    return " ".join(_v(slow_group) + inner * K_p + _v(slow_group))


def generate_baoab_mts_string(groups, K_r=1):
    """groups = [(force groups, repeat), ...] from slowest to fastest: the innermost block
    'fastest R^K_r O R^K_r fastest' is repeated, wrapped by the next slower groups, repeated ..."""
    tokens = None
    for fgs, repeat in reversed(list(groups)):
        vs = _v(fgs)
        core = ["R"] * K_r + ["O"] + ["R"] * K_r if tokens is None else tokens
        tokens = (vs + core + vs) * repeat
    return " ".join(tokens)


def generate_baoab_mts_string_from_ratios(bond_steps_per_angle_step=5, angle_steps=7):
    """four groups: 0 bonds, 1-2 angles/torsions, 3 nonbonded"""
    return generate_baoab_mts_string([([3], 1), ([1, 2], angle_steps), ([0], bond_steps_per_angle_step)])


def generate_sequential_BAOAB_string(force_group_list, symmetric=True):
    """'V0 R V1 R ... O ...' : the V R pair split into sequential per-group updates."""
    vr = []
    for g in force_group_list:
        vr += ["V%d" % g, "R"]
    return " ".join(vr + ["O"] + (vr[::-1] if symmetric else vr))


def generate_all_BAOAB_permutation_strings(n_force_groups, symmetric=True):
    return [(perm, generate_sequential_BAOAB_string(perm, symmetric))
            for perm in itertools.permutations(range(n_force_groups))]

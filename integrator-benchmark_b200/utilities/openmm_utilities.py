"""Hydrogen-mass repartitioning on plain mass vectors
(reference: benchmark/utilities/openmm_utilities.py:134-269).

The reference edits an OpenMM System in place through a Topology; the engine's systems are
tables, so the same transforms act on (elements, bonds, masses).  The functions that take
(topology, system) accept any objects with the OpenMM methods the reference calls
(`topology.atoms()`, `topology.bonds()`, `system.getParticleMass/setParticleMass`) and also the
engine's description dicts.
"""
from copy import deepcopy

import numpy as np


def _hydrogens(elements):
    return [i for i, e in enumerate(elements) if e == "H"]


def _heavy_partners(elements, bonds):
    """one entry per heavy-atom--hydrogen bond (repeats when an atom carries several hydrogens)"""
    out = []
    for a, b in bonds:
        ha, hb = elements[a] == "H", elements[b] == "H"
        if ha and not hb:
            out.append(b)
        elif hb and not ha:
            out.append(a)
    return out


def repartition_hydrogen_mass_amber_masses(elements, bonds, masses, scale_factor=3):
    """m_H <- s m_H; every heavy atom loses (s - 1) m_H per bonded hydrogen
    (Hopkins et al. 2015; reference :223-269).  Hydrogens must start with equal masses."""
    m = np.array(masses, dtype=np.float64)
    hs = _hydrogens(elements)
    if len(set(m[hs].tolist())) > 1:
        raise NotImplementedError("Initial hydrogen masses aren't all equal. "
                                  "Implementation currently assumes all hydrogen masses are initially equal.")
    out = m.copy()
    for h in hs:
        out[h] = m[h] * scale_factor
    added = out[hs[0]] - m[hs[0]]
    for heavy in _heavy_partners(elements, bonds):
        target = out[heavy] - added
        if target <= 0:
            raise RuntimeError("Trying to remove too much mass!")
        out[heavy] = target
    return out


def repartition_hydrogen_mass_masses(elements, bonds, masses, h_mass=4.0, mode="decrement", atoms="connected"):
    """set every hydrogen to h_mass and take the added mass from the atoms bonded to hydrogen
    ("connected") or from all heavy atoms ("all"), by a constant ("decrement") or proportionally
    ("scale") -- reference :134-220."""
    m = np.array(masses, dtype=np.float64)
    hs = _hydrogens(elements)
    if atoms == "connected":
        others = sorted(set(_heavy_partners(elements, bonds)))
        to_remove = (h_mass - m[hs].sum() / len(hs)) * len(hs)
    elif atoms == "all":
        others = sorted(set(range(len(m))) - set(hs))
        to_remove = h_mass * len(hs) - m[hs].sum()
    else:
        raise NotImplementedError("`atoms` must be either `all` or `connected`!")
    initial_others = m[others].sum()
    out = m.copy()
    out[hs] = h_mass
    if mode == "decrement":
        dec = to_remove / len(others)
        for o in others:
            if out[o] - dec <= 0:
                raise RuntimeError("Trying to remove too much mass!")
            out[o] -= dec
    elif mode == "scale":
        factor = (initial_others - to_remove) / initial_others
        for o in others:
            out[o] *= factor
    else:
        raise NotImplementedError("`mode` must be either `decrement` or `scale`!")
    return out


def _tables(topology, system):
    """(elements, bonds, masses) from an engine description or OpenMM-like objects"""
    if isinstance(system, dict):
        return list(system["elements"]), [tuple(b) for b in system["bonds_topology"]], np.array(system["mass"], float)
    atoms = list(topology.atoms())
    index = {id(a): i for i, a in enumerate(atoms)}
    elements = [a.element.symbol for a in atoms]
    bonds = [(index[id(a)], index[id(b)]) for a, b in topology.bonds()]
    masses = []
    for i in range(system.getNumParticles()):
        mi = system.getParticleMass(i)
        masses.append(mi.value_in_unit(None) if hasattr(mi, "value_in_unit") and not isinstance(mi, float) else float(mi))
    return elements, bonds, np.array(masses, float)


def _with_masses(system, masses):
    out = deepcopy(system)
    if isinstance(out, dict):
        out["mass"] = np.array(masses)
        return out
    for i, mi in enumerate(masses):
        out.setParticleMass(i, float(mi))
    return out


def repartition_hydrogen_mass_amber(topology, system, scale_factor=3):
    elements, bonds, masses = _tables(topology, system)
    return _with_masses(system, repartition_hydrogen_mass_amber_masses(elements, bonds, masses, scale_factor))


def repartition_hydrogen_mass(topology, system, h_mass=4.0, mode="decrement", atoms="connected"):
    elements, bonds, masses = _tables(topology, system)
    return _with_masses(system, repartition_hydrogen_mass_masses(elements, bonds, masses, h_mass, mode, atoms))


def get_masses(system):
    if isinstance(system, dict):
        return np.array(system["mass"], float)
    return np.array([float(system.getParticleMass(i)) for i in range(system.getNumParticles())])


def get_sum_of_masses(system, atom_indices=None):
    m = get_masses(system)
    return float(m.sum() if atom_indices is None else m[list(atom_indices)].sum())

"""OpenMM System XML <-> the engine's plain system description.

The reference writes its test systems with XmlSerializer.serialize(system)
(benchmark/serialization/serialize_testsystems.py:5-13).  `load_system_xml` turns such a file into the
description dict that engine.ParticleSystem / EquilibriumSimulator take, so a system built by the reference
(with the real force-field files) runs on the engine unchanged; `system_to_xml` writes the same layout back
(round trip, and a way to hand the engine's systems to OpenMM for cross-checks where OpenMM exists).

What is read: Particles (masses), Constraints, PeriodicBoxVectors (orthorhombic), HarmonicBondForce,
HarmonicAngleForce, PeriodicTorsionForce, NonbondedForce (particles, exceptions; NoCutoff or CutoffPeriodic),
the reference's restraint CustomExternalForce "(K/2.0) * (x^2 + y^2 + z^2); K = ..." (watercluster.py:70-80),
forceGroup of every force.  CMMotionRemover and MonteCarloBarostat are dropped, as NonequilibriumSimulator
does (bookkeepers.py:167-171).  Triangles of three constraints over three atoms are rigid waters (SETTLE).
Ewald / PME electrostatics are not built: with pme_as_reaction_field=True such a force is read as
CutoffPeriodic with the same cutoff (the defined variant of DESIGN.md), otherwise NotImplementedError.
"""
import re
import xml.etree.ElementTree as ET
from collections import defaultdict

import numpy as np

_METHODS = {0: "NoCutoff", 1: "CutoffNonPeriodic", 2: "CutoffPeriodic", 3: "Ewald", 4: "PME", 5: "LJPME"}


def _f(el, key, default=None):
    v = el.get(key)
    if v is None:
        if default is None:
            raise ValueError("missing attribute %r in <%s>" % (key, el.tag))
        return default
    return float(v)


def load_system_xml(source, pme_as_reaction_field=False, ignore_unknown_forces=False):
    """`source`: path, file object or XML string.  Returns the description dict."""
    if isinstance(source, str) and source.lstrip().startswith("<"):
        root = ET.fromstring(source)
    else:
        root = ET.parse(source).getroot()
    if root.tag != "System":
        raise ValueError("not an OpenMM System XML (root element <%s>)" % root.tag)
    masses = [_f(p, "mass") for p in root.find("Particles")]
    n = len(masses)
    desc = dict(n_atoms=n, mass=np.array(masses), charge=np.zeros(n), sigma=np.ones(n), epsilon=np.zeros(n),
                nonbonded_method="NoCutoff", groups={}, constraint_tolerance=1e-8)
    box = root.find("PeriodicBoxVectors")
    if box is not None:
        a, b, c = box.find("A"), box.find("B"), box.find("C")
        off = [_f(a, "y", 0.0), _f(a, "z", 0.0), _f(b, "x", 0.0), _f(b, "z", 0.0), _f(c, "x", 0.0), _f(c, "y", 0.0)]
        if any(abs(o) > 1e-12 for o in off):
            raise NotImplementedError("triclinic periodic boxes are not built")
        desc["box"] = (_f(a, "x"), _f(b, "y"), _f(c, "z"))

    forces = root.find("Forces")
    for force in (forces if forces is not None else []):
        kind = force.get("type")
        group = int(force.get("forceGroup", 0))
        if kind == "HarmonicBondForce":
            bonds = list(force.find("Bonds"))
            desc["bond_atoms"] = [(int(b.get("p1")), int(b.get("p2"))) for b in bonds]
            desc["bond_params"] = [(_f(b, "d"), _f(b, "k")) for b in bonds]
            desc["groups"]["bonds"] = group
        elif kind == "HarmonicAngleForce":
            angles = list(force.find("Angles"))
            desc["angle_atoms"] = [(int(a.get("p1")), int(a.get("p2")), int(a.get("p3"))) for a in angles]
            desc["angle_params"] = [(_f(a, "a"), _f(a, "k")) for a in angles]
            desc["groups"]["angles"] = group
        elif kind == "PeriodicTorsionForce":
            tors = list(force.find("Torsions"))
            desc["torsion_atoms"] = [tuple(int(t.get("p%d" % i)) for i in (1, 2, 3, 4)) for t in tors]
            desc["torsion_params"] = [(_f(t, "periodicity"), _f(t, "phase"), _f(t, "k")) for t in tors]
            desc["groups"]["torsions"] = group
        elif kind == "NonbondedForce":
            method = _METHODS.get(int(float(force.get("method", 0))), "?")
            if method in ("Ewald", "PME", "LJPME"):
                if not pme_as_reaction_field:
                    raise NotImplementedError("NonbondedForce method %s is not built (pass pme_as_reaction_field=True to read it "
                                              "as CutoffPeriodic with the same cutoff)" % method)
                method = "CutoffPeriodic"
            if method not in ("NoCutoff", "CutoffPeriodic"):
                raise NotImplementedError("NonbondedForce method %s is not built" % method)
            desc["nonbonded_method"] = method
            desc["cutoff"] = _f(force, "cutoff", 0.0)
            if force.get("useSwitchingFunction", "0") not in ("0", "false"):
                if method != "CutoffPeriodic":
                    raise NotImplementedError("a switching function needs a cutoff method")
                desc["switch_distance"] = _f(force, "switchingDistance", 0.0)
            desc["rf_dielectric"] = _f(force, "rfDielectric", 78.3)
            parts = list(force.find("Particles"))
            if len(parts) != n:
                raise ValueError("NonbondedForce has %d particles, the System %d" % (len(parts), n))
            desc["charge"] = np.array([_f(p, "q") for p in parts])
            desc["sigma"] = np.array([_f(p, "sig") for p in parts])
            desc["epsilon"] = np.array([_f(p, "eps") for p in parts])
            excl, exc_atoms, exc_params = [], [], []
            ex = force.find("Exceptions")
            for e in (ex if ex is not None else []):
                pair = (int(e.get("p1")), int(e.get("p2")))
                q, sig, eps = _f(e, "q"), _f(e, "sig"), _f(e, "eps")
                if q == 0.0 and eps == 0.0:
                    excl.append(pair)
                else:
                    exc_atoms.append(pair)
                    exc_params.append((q, sig, eps))
            desc["exclusions"], desc["exception_atoms"], desc["exception_params"] = excl, exc_atoms, exc_params
            desc["groups"]["nonbonded"] = group
        elif kind == "CustomExternalForce":
            energy = force.get("energy", "").replace(" ", "")
            m = re.search(r"\(K/2(?:\.0*)?\)\*\(x\^2\+y\^2\+z\^2\);K=([-+0-9.eE]+)", energy)
            parts = force.find("Particles")
            if m is None or parts is None or len(parts) != n:
                raise NotImplementedError("only the restraint (K/2) (x^2 + y^2 + z^2) on every particle is built; got %r" % force.get("energy"))
            desc["restraint_k"] = float(m.group(1))
            desc["groups"]["restraint"] = group
        elif kind in ("CMMotionRemover", "MonteCarloBarostat", "MonteCarloAnisotropicBarostat"):
            continue  # removed by the reference before the nonequilibrium runs (bookkeepers.py:167-171)
        elif not ignore_unknown_forces:
            raise NotImplementedError("force type %s is not built" % kind)

    # constraints: triangles over three atoms are rigid waters
    cons_el = root.find("Constraints")
    cons = [(int(c.get("p1")), int(c.get("p2")), _f(c, "d")) for c in (cons_el if cons_el is not None else [])]
    adj = defaultdict(dict)
    for i, j, d in cons:
        adj[i][j] = d
        adj[j][i] = d
    used, settle, oh_hh = set(), [], {}
    for i in sorted(adj):
        if i in used or len(adj[i]) != 2:
            continue
        j, k = sorted(adj[i])
        if len(adj[j]) == 2 and len(adj[k]) == 2 and k in adj[j] and not ({j, k} & used):
            tri = [i, j, k]
            # the oxygen is the atom whose two constraints have equal length
            for o in tri:
                h1, h2 = [a for a in tri if a != o]
                if abs(adj[o][h1] - adj[o][h2]) < 1e-9 and masses[h1] == masses[h2]:
                    settle.append((o, h1, h2))
                    oh_hh[(round(adj[o][h1], 9), round(adj[h1][h2], 9))] = (adj[o][h1], adj[h1][h2])
                    used.update(tri)
                    break
    if len(oh_hh) > 1:
        raise NotImplementedError("rigid waters with different geometries in one system")
    if settle:
        (oh, hh), = oh_hh.values()
        desc.update(settle_atoms=settle, settle_oh=oh, settle_hh=hh)
    rest = [(i, j, d) for i, j, d in cons if i not in used and j not in used]
    desc["constraint_atoms"] = [(i, j) for i, j, _ in rest]
    desc["constraint_dist"] = [d for _, _, d in rest]
    for key in ("bonds", "angles", "torsions", "nonbonded", "restraint"):
        desc["groups"].setdefault(key, 0)
    return desc


def system_to_xml(desc, openmm_version="7.2"):
    """The description as OpenMM System XML (the layout XmlSerializer writes)."""
    n = int(desc["n_atoms"])
    g = desc.get("groups", {})
    root = ET.Element("System", openmmVersion=openmm_version, type="System", version="1")
    box = desc.get("box", (2.0, 2.0, 2.0))
    pbv = ET.SubElement(root, "PeriodicBoxVectors")
    for tag, vec in (("A", (box[0], 0, 0)), ("B", (0, box[1], 0)), ("C", (0, 0, box[2]))):
        ET.SubElement(pbv, tag, x=repr(float(vec[0])), y=repr(float(vec[1])), z=repr(float(vec[2])))
    parts = ET.SubElement(root, "Particles")
    for m in np.asarray(desc["mass"], dtype=float):
        ET.SubElement(parts, "Particle", mass=repr(float(m)))
    cons = ET.SubElement(root, "Constraints")
    for (o, h1, h2) in desc.get("settle_atoms", []) or []:
        for a, b, d in ((o, h1, desc["settle_oh"]), (o, h2, desc["settle_oh"]), (h1, h2, desc["settle_hh"])):
            ET.SubElement(cons, "Constraint", d=repr(float(d)), p1=str(int(a)), p2=str(int(b)))
    for (a, b), d in zip(desc.get("constraint_atoms", []) or [], desc.get("constraint_dist", []) or []):
        ET.SubElement(cons, "Constraint", d=repr(float(d)), p1=str(int(a)), p2=str(int(b)))
    forces = ET.SubElement(root, "Forces")
    if len(desc.get("bond_atoms", []) or []):
        f = ET.SubElement(forces, "Force", forceGroup=str(g.get("bonds", 0)), type="HarmonicBondForce", usesPeriodic="0", version="2")
        bonds = ET.SubElement(f, "Bonds")
        for (a, b), (d, k) in zip(desc["bond_atoms"], desc["bond_params"]):
            ET.SubElement(bonds, "Bond", d=repr(float(d)), k=repr(float(k)), p1=str(int(a)), p2=str(int(b)))
    if len(desc.get("angle_atoms", []) or []):
        f = ET.SubElement(forces, "Force", forceGroup=str(g.get("angles", 0)), type="HarmonicAngleForce", usesPeriodic="0", version="2")
        angles = ET.SubElement(f, "Angles")
        for (a, b, c), (th, k) in zip(desc["angle_atoms"], desc["angle_params"]):
            ET.SubElement(angles, "Angle", a=repr(float(th)), k=repr(float(k)), p1=str(int(a)), p2=str(int(b)), p3=str(int(c)))
    if len(desc.get("torsion_atoms", []) or []):
        f = ET.SubElement(forces, "Force", forceGroup=str(g.get("torsions", 0)), type="PeriodicTorsionForce", usesPeriodic="0", version="2")
        tors = ET.SubElement(f, "Torsions")
        for (a, b, c, d_), (per, ph, k) in zip(desc["torsion_atoms"], desc["torsion_params"]):
            ET.SubElement(tors, "Torsion", k=repr(float(k)), p1=str(int(a)), p2=str(int(b)), p3=str(int(c)), p4=str(int(d_)),
                          periodicity=str(int(round(per))), phase=repr(float(ph)))
    method = {"NoCutoff": 0, "CutoffPeriodic": 2}[desc.get("nonbonded_method", "NoCutoff")]
    sw = float(desc.get("switch_distance", 0.0) or 0.0)
    f = ET.SubElement(forces, "Force", alpha="0", cutoff=repr(float(desc.get("cutoff", 1.0) or 1.0)), dispersionCorrection="0",
                      ewaldTolerance=".0005", forceGroup=str(g.get("nonbonded", 0)), method=str(method), nx="0", ny="0", nz="0",
                      recipForceGroup="-1", rfDielectric=repr(float(desc.get("rf_dielectric", 78.3))),
                      switchingDistance=repr(sw) if sw > 0 else "-1",
                      type="NonbondedForce", useSwitchingFunction="1" if sw > 0 else "0", version="2")
    nbp = ET.SubElement(f, "Particles")
    q = np.asarray(desc.get("charge", np.zeros(n)), dtype=float)
    sig = np.asarray(desc.get("sigma", np.ones(n)), dtype=float)
    eps = np.asarray(desc.get("epsilon", np.zeros(n)), dtype=float)
    for i in range(n):
        ET.SubElement(nbp, "Particle", eps=repr(float(eps[i])), q=repr(float(q[i])), sig=repr(float(sig[i])))
    exc = ET.SubElement(f, "Exceptions")
    for (a, b) in desc.get("exclusions", []) or []:
        ET.SubElement(exc, "Exception", eps="0", p1=str(int(a)), p2=str(int(b)), q="0", sig="1")
    for (a, b), (qq, s_, e_) in zip(desc.get("exception_atoms", []) or [], desc.get("exception_params", []) or []):
        ET.SubElement(exc, "Exception", eps=repr(float(e_)), p1=str(int(a)), p2=str(int(b)), q=repr(float(qq)), sig=repr(float(s_)))
    if float(desc.get("restraint_k", 0.0) or 0.0) != 0.0:
        f = ET.SubElement(forces, "Force", energy="(K/2.0) * (x^2 + y^2 + z^2);K = %r;" % float(desc["restraint_k"]),
                          forceGroup=str(g.get("restraint", 0)), type="CustomExternalForce", version="1")
        ET.SubElement(f, "PerParticleParameters")
        ET.SubElement(f, "GlobalParameters")
        ps = ET.SubElement(f, "Particles")
        for i in range(n):
            ET.SubElement(ps, "Particle", index=str(i))
    return '<?xml version="1.0" ?>\n' + ET.tostring(root, encoding="unicode")

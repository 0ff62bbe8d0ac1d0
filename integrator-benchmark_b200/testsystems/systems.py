"""Descriptions (plain data) of the reference's molecular test systems.

The reference builds these with openmmtools / OpenMM force-field files
(benchmark/testsystems/watercluster.py:10-84, waterbox.py:36-48, alanine_dipeptide.py:12-21),
none of which exist in this environment.  The parameter values below are therefore restated
from the published force fields (TIP3P: Jorgensen et al. 1983 as shipped in OpenMM's tip3p.xml;
alanine dipeptide: AMBER parm94/ff96 atom types, charges and valence terms) and CANNOT be
cross-checked against OpenMM here: parity of force-field VALUES is unpinned.  The
engine itself is table-driven (engine.ParticleSystem), so a description exported from a real
OpenMM System replaces these without touching any kernel.

Every builder returns (desc, positions): `desc` feeds engine.ParticleSystem, `positions` is a
reasonable (not equilibrated) start configuration in nm.
"""
import itertools

import numpy as np

# TIP3P (OpenMM tip3p.xml)
TIP3P = dict(qO=-0.834, qH=0.417, sigmaO=0.3150752406575124, epsO=0.635968, rOH=0.09572,
             theta=1.82421813418, mO=15.99943, mH=1.007947, kb=462750.4, ktheta=836.8)
TIP3P["rHH"] = 2.0 * TIP3P["rOH"] * np.sin(0.5 * TIP3P["theta"])  # 0.15139 nm

KCAL = 4.184


def _water_geometry():
    s, c = 0.5 * TIP3P["rHH"], np.sqrt(TIP3P["rOH"] ** 2 - (0.5 * TIP3P["rHH"]) ** 2)
    return np.array([[0.0, 0.0, 0.0], [s, c, 0.0], [-s, c, 0.0]])


def _random_rotation(rng):
    q = rng.randn(4)
    q /= np.linalg.norm(q)
    a, b, c, d = q
    return np.array([[a * a + b * b - c * c - d * d, 2 * (b * c - a * d), 2 * (b * d + a * c)],
                     [2 * (b * c + a * d), a * a - b * b + c * c - d * d, 2 * (c * d - a * b)],
                     [2 * (b * d - a * c), 2 * (c * d + a * b), a * a - b * b - c * c + d * d]])


def _water_tables(n_waters, constrained):
    n = 3 * n_waters
    desc = dict(
        n_atoms=n,
        mass=np.tile([TIP3P["mO"], TIP3P["mH"], TIP3P["mH"]], n_waters),
        charge=np.tile([TIP3P["qO"], TIP3P["qH"], TIP3P["qH"]], n_waters),
        sigma=np.tile([TIP3P["sigmaO"], 1.0, 1.0], n_waters),
        epsilon=np.tile([TIP3P["epsO"], 0.0, 0.0], n_waters),
        exclusions=[(3 * w + a, 3 * w + b) for w in range(n_waters) for a, b in ((0, 1), (0, 2), (1, 2))],
        groups=dict(nonbonded=0, bonds=1, angles=1, torsions=1, restraint=0),
        constraint_tolerance=1e-8,
    )
    if constrained:
        desc.update(settle_atoms=[(3 * w, 3 * w + 1, 3 * w + 2) for w in range(n_waters)],
                    settle_oh=TIP3P["rOH"], settle_hh=TIP3P["rHH"])
    else:
        desc.update(bond_atoms=[(3 * w, 3 * w + h) for w in range(n_waters) for h in (1, 2)],
                    bond_params=[(TIP3P["rOH"], TIP3P["kb"])] * (2 * n_waters),
                    angle_atoms=[(3 * w + 1, 3 * w, 3 * w + 2) for w in range(n_waters)],
                    angle_params=[(TIP3P["theta"], TIP3P["ktheta"])] * n_waters)
    return desc


def _lattice_waters(n_waters, spacing, centre, rng, box=None):
    m = int(np.ceil(n_waters ** (1.0 / 3.0)))
    sites = np.array(list(itertools.product(range(m), repeat=3)), dtype=float)
    sites = sites[np.argsort(np.linalg.norm(sites - 0.5 * (m - 1), axis=1), kind="stable")][:n_waters]
    sites = (sites - 0.5 * (m - 1)) * spacing + centre
    tri = _water_geometry()
    tri = tri - tri.mean(axis=0)
    return np.concatenate([tri @ _random_rotation(rng).T + s for s in sites])


def water_cluster(n_waters=20, K=1.0, constrained=True, seed=0):
    """WaterCluster of watercluster.py:10-84: TIP3P, NoCutoff, no CMM remover, harmonic restraint
    (K/2)|x|^2 with K = 1 kJ/mol/nm^2 on every atom; rigid (SETTLE) or flexible."""
    desc = _water_tables(n_waters, constrained)
    desc.update(nonbonded_method="NoCutoff", restraint_k=float(K), name="water_cluster_%s" % ("rigid" if constrained else "flexible"))
    pos = _lattice_waters(n_waters, 0.31, np.zeros(3), np.random.RandomState(seed))
    return desc, pos


def tiny_waterbox(n_side=None, box_edge=1.5, cutoff=0.6, n_waters=102, constrained=True, seed=0, switch_width=0.15):
    """tiny_waterbox of waterbox.py:36-48: 102 TIP3P waters in a 1.5 nm cube, cutoff 0.6 nm.
    The reference's WaterBox (openmmtools defaults) uses PME electrostatics, a Lennard-Jones switching function over the
    last `switch_width` = 1.5 A before the cutoff, and a dispersion correction (a constant at fixed volume: it changes no
    force and no work).  This engine's variant keeps the Lennard-Jones part exactly (switching function included) and
    replaces PME by the cutoff + reaction-field electrostatics that BASELINE.json asks for ("cutoff nonbonded",
    SURVEY 8a note on R20).  switch_width = 0 gives the plain truncation."""
    desc = _water_tables(n_waters, constrained)
    desc.update(nonbonded_method="CutoffPeriodic", cutoff=float(cutoff), box=(box_edge,) * 3, rf_dielectric=78.3,
                switch_distance=float(cutoff - switch_width) if switch_width and switch_width > 0 else 0.0,
                name="tiny_waterbox_%s" % ("constrained" if constrained else "flexible"))
    m = int(np.ceil(n_waters ** (1.0 / 3.0)))
    pos = _lattice_waters(n_waters, box_edge / m, 0.5 * box_edge * np.ones(3), np.random.RandomState(seed))
    return desc, pos


# ------------------------------------------------------------------------------ alanine dipeptide
ALA_NAMES = ["HH31", "CH3", "HH32", "HH33", "C", "O", "N", "H", "CA", "HA", "CB", "HB1", "HB2", "HB3", "C", "O",
             "N", "H", "CH3", "HH31", "HH32", "HH33"]
ALA_ELEMENTS = ["H", "C", "H", "H", "C", "O", "N", "H", "C", "H", "C", "H", "H", "H", "C", "O", "N", "H", "C", "H", "H", "H"]
ALA_TYPES = ["HC", "CT", "HC", "HC", "C", "O", "N", "H", "CT", "H1", "CT", "HC", "HC", "HC", "C", "O", "N", "H", "CT",
             "H1", "H1", "H1"]
ALA_CHARGES = [0.1123, -0.3662, 0.1123, 0.1123, 0.5972, -0.5679, -0.4157, 0.2719, 0.0337, 0.0823, -0.1825, 0.0603,
               0.0603, 0.0603, 0.5973, -0.5679, -0.4157, 0.2719, -0.1490, 0.0976, 0.0976, 0.0976]
ALA_BONDS = [(0, 1), (1, 2), (1, 3), (1, 4), (4, 5), (4, 6), (6, 7), (6, 8), (8, 9), (8, 10), (10, 11), (10, 12),
             (10, 13), (8, 14), (14, 15), (14, 16), (16, 17), (16, 18), (18, 19), (18, 20), (18, 21)]
ELEMENT_MASS = {"H": 1.008, "C": 12.01, "N": 14.01, "O": 16.0}

# parm94 van der Waals: Rmin/2 (Angstrom), epsilon (kcal/mol)
_LJ = {"HC": (1.4870, 0.0157), "CT": (1.9080, 0.1094), "C": (1.9080, 0.0860), "O": (1.6612, 0.2100),
       "N": (1.8240, 0.1700), "H": (0.6000, 0.0157), "H1": (1.3870, 0.0157)}
# bonds: k (kcal/mol/A^2, AMBER convention k (r - r0)^2), r0 (A)
_BOND = {("CT", "HC"): (340.0, 1.090), ("CT", "H1"): (340.0, 1.090), ("CT", "CT"): (310.0, 1.526),
         ("C", "CT"): (317.0, 1.522), ("C", "O"): (570.0, 1.229), ("C", "N"): (490.0, 1.335),
         ("H", "N"): (434.0, 1.010), ("CT", "N"): (337.0, 1.449)}
# angles: k (kcal/mol/rad^2, AMBER convention), theta0 (deg)
_ANGLE = {("HC", "CT", "HC"): (35.0, 109.5), ("H1", "CT", "H1"): (35.0, 109.5), ("C", "CT", "HC"): (50.0, 109.5),
          ("CT", "C", "O"): (80.0, 120.4), ("CT", "C", "N"): (70.0, 116.6), ("N", "C", "O"): (80.0, 122.9),
          ("C", "N", "H"): (30.0, 119.8), ("C", "N", "CT"): (50.0, 121.9), ("CT", "N", "H"): (30.0, 118.4),
          ("C", "CT", "N"): (63.0, 110.1), ("CT", "CT", "N"): (80.0, 109.7), ("H1", "CT", "N"): (50.0, 109.5),
          ("C", "CT", "H1"): (50.0, 109.5), ("C", "CT", "CT"): (63.0, 111.1), ("CT", "CT", "H1"): (50.0, 109.5),
          ("CT", "CT", "HC"): (50.0, 109.5)}
# proper torsions: specific (ff96 phi/psi) then generic by central bond; (divider, barrier kcal/mol, phase deg, n)
_TORSION_SPECIFIC = {
    ("C", "N", "CT", "C"): [(1, 0.85, 180.0, 2), (1, 0.80, 0.0, 1)],
    ("N", "CT", "C", "N"): [(1, 0.85, 180.0, 2), (1, 0.80, 0.0, 1)],
    ("CT", "CT", "N", "C"): [(1, 0.50, 180.0, 4), (1, 0.15, 180.0, 3), (1, 0.53, 0.0, 1)],
    ("CT", "CT", "C", "N"): [(1, 0.10, 0.0, 4), (1, 0.07, 0.0, 2)],
    ("H", "N", "C", "O"): [(1, 2.5, 180.0, 2), (1, 2.0, 0.0, 1)],
}
_TORSION_GENERIC = {("C", "CT"): [], ("CT", "CT"): [(9, 1.40, 0.0, 3)], ("CT", "N"): [], ("C", "N"): [(4, 10.0, 180.0, 2)]}
_IMPROPERS = [((1, 6, 4, 5), 10.5), ((8, 16, 14, 15), 10.5), ((4, 8, 6, 7), 1.0), ((14, 18, 16, 17), 1.0)]


def _key(*t):
    return t if t <= t[::-1] else t[::-1]


def alanine_dipeptide(constrained=True, mass_scale=None):
    """AlanineDipeptideVacuum(constraints=HBonds or None) keeping bond/angle/torsion/nonbonded
    forces (alanine_dipeptide.py:12-21, utilities/openmm_utilities.py:345-357); force groups as
    experiments/submission_scripts/4_mts_integrator_splittings.py:36-42: nonbonded 0, valence 1."""
    n = len(ALA_TYPES)
    nbrs = {i: set() for i in range(n)}
    for a, b in ALA_BONDS:
        nbrs[a].add(b)
        nbrs[b].add(a)
    mass = np.array([ELEMENT_MASS[e] for e in ALA_ELEMENTS])
    sigma = np.array([_LJ[t][0] * 2.0 / 2.0 ** (1.0 / 6.0) * 0.1 for t in ALA_TYPES])
    eps = np.array([_LJ[t][1] * KCAL for t in ALA_TYPES])
    is_h = [e == "H" for e in ALA_ELEMENTS]

    bonds, bparams, cons, cdist = [], [], [], []
    for a, b in ALA_BONDS:
        k, r0 = _BOND[_key(ALA_TYPES[a], ALA_TYPES[b])]
        if constrained and (is_h[a] or is_h[b]):
            cons.append((a, b))
            cdist.append(0.1 * r0)
        else:
            bonds.append((a, b))
            bparams.append((0.1 * r0, 2.0 * k * KCAL * 100.0))
    angles, aparams = [], []
    for j in range(n):
        for i, k_ in itertools.combinations(sorted(nbrs[j]), 2):
            k, th = _ANGLE[_key(ALA_TYPES[i], ALA_TYPES[j], ALA_TYPES[k_])]
            angles.append((i, j, k_))
            aparams.append((np.deg2rad(th), 2.0 * k * KCAL))
    torsions, tparams, pairs14 = [], [], set()
    for j, k_ in ALA_BONDS:
        for i in sorted(nbrs[j] - {k_}):
            for l in sorted(nbrs[k_] - {j}):
                tt = (ALA_TYPES[i], ALA_TYPES[j], ALA_TYPES[k_], ALA_TYPES[l])
                terms = _TORSION_SPECIFIC.get(tt) or _TORSION_SPECIFIC.get(tt[::-1])
                if terms is None:
                    terms = _TORSION_GENERIC[_key(ALA_TYPES[j], ALA_TYPES[k_])]
                for div, pk, phase, per in terms:
                    torsions.append((i, j, k_, l))
                    tparams.append((float(per), np.deg2rad(phase), pk / div * KCAL))
                pairs14.add((min(i, l), max(i, l)))
    for atoms, pk in _IMPROPERS:
        torsions.append(atoms)
        tparams.append((2.0, np.pi, pk * KCAL))
    excl = set()
    for a, b in ALA_BONDS:
        excl.add((min(a, b), max(a, b)))
    for i, j, k_ in angles:
        excl.add((min(i, k_), max(i, k_)))
    pairs14 = sorted(p for p in pairs14 if p not in excl)
    q = np.array(ALA_CHARGES)
    exc_params = [(q[i] * q[j] / 1.2, 0.5 * (sigma[i] + sigma[j]), np.sqrt(eps[i] * eps[j]) / 2.0) for i, j in pairs14]
    desc = dict(n_atoms=n, mass=mass, charge=q, sigma=sigma, epsilon=eps, nonbonded_method="NoCutoff",
                exclusions=sorted(excl), exception_atoms=pairs14, exception_params=exc_params,
                bond_atoms=bonds, bond_params=bparams, angle_atoms=angles, angle_params=aparams,
                torsion_atoms=torsions, torsion_params=tparams,
                groups=dict(nonbonded=0, bonds=1, angles=1, torsions=1, restraint=1),
                constraint_atoms=cons, constraint_dist=cdist, constraint_tolerance=1e-8,
                name="alanine_%s" % ("constrained" if constrained else "unconstrained"),
                elements=list(ALA_ELEMENTS), bonds_topology=list(ALA_BONDS))
    if mass_scale is not None:
        from ..utilities.openmm_utilities import repartition_hydrogen_mass_amber_masses
        desc["mass"] = repartition_hydrogen_mass_amber_masses(ALA_ELEMENTS, ALA_BONDS, mass, mass_scale)
    return desc, _alanine_positions()


def _place(a, b, c, r, theta, phi):
    """NeRF: new atom d bonded to c with |cd| = r, angle b-c-d = theta, dihedral a-b-c-d = phi (degrees)."""
    theta, phi = np.deg2rad(theta), np.deg2rad(phi)
    bc = c - b
    bc /= np.linalg.norm(bc)
    nrm = np.cross(b - a, bc)
    nrm /= np.linalg.norm(nrm)
    m = np.cross(nrm, bc)
    d2 = np.array([-r * np.cos(theta), r * np.sin(theta) * np.cos(phi), r * np.sin(theta) * np.sin(phi)])
    return c + d2[0] * bc + d2[1] * m + d2[2] * nrm


def _alanine_positions(phi=-75.0, psi=145.0):
    x = {}
    x[1] = np.array([0.0, 0.0, 0.0])
    x[4] = np.array([0.1522, 0.0, 0.0])
    x[5] = x[4] + 0.1229 * np.array([np.cos(np.deg2rad(180 - 120.4)), np.sin(np.deg2rad(180 - 120.4)), 0.0])
    x[6] = _place(x[5], x[1], x[4], 0.1335, 116.6, 180.0)
    for idx, dih in ((0, 180.0), (2, 60.0), (3, -60.0)):
        x[idx] = _place(x[6], x[4], x[1], 0.1090, 109.5, dih)
    x[7] = _place(x[5], x[4], x[6], 0.1010, 119.8, 180.0)
    x[8] = _place(x[5], x[4], x[6], 0.1449, 121.9, 0.0)
    x[14] = _place(x[4], x[6], x[8], 0.1522, 110.1, phi)
    x[10] = _place(x[6], x[14], x[8], 0.1526, 111.1, -122.6)
    x[9] = _place(x[6], x[14], x[8], 0.1090, 109.5, 119.0)
    for idx, dih in ((11, 180.0), (12, 60.0), (13, -60.0)):
        x[idx] = _place(x[6], x[8], x[10], 0.1090, 109.5, dih)
    x[16] = _place(x[6], x[8], x[14], 0.1335, 116.6, psi)
    x[15] = _place(x[6], x[8], x[14], 0.1229, 120.4, psi + 180.0)
    x[17] = _place(x[15], x[14], x[16], 0.1010, 119.8, 180.0)
    x[18] = _place(x[15], x[14], x[16], 0.1449, 121.9, 0.0)
    for idx, dih in ((19, 180.0), (20, 60.0), (21, -60.0)):
        x[idx] = _place(x[14], x[16], x[18], 0.1090, 109.5, dih)
    return np.array([x[i] for i in range(22)])

"""The equilibrium-sample caches of the reference are mdtraj HDF5 trajectories: EquilibriumSimulator writes
`{name}_samples.h5` with mdtraj's HDF5Reporter (coordinates, cell, potentialEnergy; benchmark/testsystems/bookkeepers.py:106-113)
and loads it back with `md.load` (:75); the one file the reference ships is data/tiny_waterbox_constrained_samples.h5.

mdtraj's layout (HDF5TrajectoryFile, "Pande" conventions 1.1), every array an extendible chunked dataset filtered with
byte-shuffle + zlib, float32 little-endian:

    coordinates      (frames, n_atoms, 3)  units = "nanometers"
    cell_lengths     (frames, 3)           units = "nanometers"
    cell_angles      (frames, 3)           units = "degrees"
    potentialEnergy  (frames,)             units = "kilojoules_per_mole"
    topology         (1,) fixed-length string: JSON {"chains": [{"index", "residues": [{"index", "name", "resSeq",
                     "atoms": [{"index", "name", "element"}]}]}], "bonds": [[i, j], ...]}

Both directions go through serialization/hdf5_min.py (no h5py / PyTables / mdtraj in this image): the reader walks the file
structurally (superblock -> root group -> dataset headers -> chunk B-trees), the writer emits the same structures, so a cache
written here is a file the reference's `md.load` opens, and vice versa.
"""
import json

import numpy as np

from . import hdf5_min

_ELEMENTS = [(1.008, "H"), (12.011, "C"), (14.007, "N"), (15.999, "O"), (32.06, "S"), (30.974, "P"), (35.45, "Cl"), (22.99, "Na")]


def load_mdtraj_hdf5(path, n_atoms=None):
    """Frames of an mdtraj HDF5 trajectory: dict(xyz (frames, n_atoms, 3) float32 nm, cell_lengths (frames, 3) or None,
    cell_angles (frames, 3) or None, potential_energy (frames,) or None, topology (dict) or None).
    Raises ValueError when the file has no coordinates, when `n_atoms` is given and differs, when the per-frame arrays
    disagree about the number of frames, or when a coordinate is not finite."""
    d = hdf5_min.read(path)
    if "coordinates" not in d:
        raise ValueError("{}: no `coordinates` dataset".format(path))
    xyz = np.ascontiguousarray(d["coordinates"].data, dtype=np.float32)
    if xyz.ndim != 3 or xyz.shape[2] != 3:
        raise ValueError("{}: coordinates have shape {}".format(path, xyz.shape))
    if n_atoms is not None and xyz.shape[1] != int(n_atoms):
        raise ValueError("{}: {} atoms per frame, expected {}".format(path, xyz.shape[1], n_atoms))
    units = d["coordinates"].attrs.get("units", "nanometers")
    if units not in ("nanometers", b"nanometers"):
        raise ValueError("{}: coordinates in {!r}, expected nanometers".format(path, units))
    if not np.all(np.isfinite(xyz)):
        raise ValueError("{}: non-finite coordinates".format(path))
    out = dict(xyz=xyz, cell_lengths=None, cell_angles=None, potential_energy=None, topology=None)
    for key, name, tail in (("cell_lengths", "cell_lengths", (3,)), ("cell_angles", "cell_angles", (3,)), ("potential_energy", "potentialEnergy", ())):
        if name in d:
            a = np.asarray(d[name].data)
            if a.shape != (xyz.shape[0],) + tail:
                raise ValueError("{}: {} has shape {}, expected {} frames".format(path, name, a.shape, xyz.shape[0]))
            out[key] = a
    if "topology" in d:
        raw = d["topology"].data.reshape(-1)[0]
        try:
            out["topology"] = json.loads(bytes(raw).split(b"\0")[0].decode())
        except ValueError:
            out["topology"] = None
    return out


def topology_from_description(desc):
    """mdtraj topology JSON (as a dict) for one of this package's system descriptions: rigid / flexible waters become HOH
    residues (O, H1, H2), everything else one residue per connected molecule; elements from `desc["elements"]` when the
    description carries them, else from the masses; bonds from the bond, constraint, SETTLE (and `bonds_topology`) tables."""
    n = int(desc["n_atoms"])
    mass = np.asarray(desc["mass"], dtype=np.float64)
    bonds = set()
    for key in ("bond_atoms", "constraint_atoms", "bonds_topology"):
        for i, j in np.asarray(desc.get(key) if desc.get(key) is not None else [], dtype=np.int64).reshape(-1, 2):
            bonds.add((int(min(i, j)), int(max(i, j))))
    waters = {}
    for o, h1, h2 in np.asarray(desc.get("settle_atoms") if desc.get("settle_atoms") is not None else [], dtype=np.int64).reshape(-1, 3):
        bonds.add((int(min(o, h1)), int(max(o, h1))))
        bonds.add((int(min(o, h2)), int(max(o, h2))))
        waters[int(o)] = (int(o), int(h1), int(h2))
    parent = list(range(n))

    def find(a):
        while parent[a] != a:
            parent[a] = parent[parent[a]]
            a = parent[a]
        return a
    for i, j in bonds:
        parent[find(i)] = find(j)
    groups = {}
    for a in range(n):
        groups.setdefault(find(a), []).append(a)

    given = desc.get("elements")

    def element(a):
        if given is not None:
            return str(given[a])
        m = mass[a]
        if m < 4.5 and any(mass[b] > 4.5 for b in range(n)):  # hydrogens (possibly repartitioned up to ~4 u)
            return "H"
        return min(_ELEMENTS, key=lambda e: abs(e[0] - m))[1]
    residues = []
    for k, atoms in enumerate(sorted(groups.values(), key=lambda g: g[0])):
        els = [element(a) for a in atoms]
        water = len(atoms) == 3 and sorted(els) == ["H", "H", "O"]
        names, count = [], {}
        for e in els:
            count[e] = count.get(e, 0) + 1
            names.append(e if water and e == "O" else "%s%d" % (e, count[e]))
        residues.append({"index": k, "name": "HOH" if water else "MOL", "resSeq": k + 1,
                         "atoms": [{"index": a, "name": nm, "element": e} for a, nm, e in zip(atoms, names, els)]})
    return {"chains": [{"index": 0, "residues": residues}], "bonds": [[i, j] for i, j in sorted(bonds)]}


def save_mdtraj_hdf5(path, xyz, cell_lengths=None, cell_angles=None, potential_energy=None, topology=None, title="title",
                     application="integrator_benchmark_b200"):
    """Write frames as an mdtraj HDF5 trajectory (the reference's cache format, bookkeepers.py:106-118): float32 coordinates
    in nm, optional cell (nm, degrees) and potential energies (kJ/mol), the topology as JSON.  Returns `path`."""
    xyz = np.ascontiguousarray(xyz, dtype=np.float32)
    if xyz.ndim != 3 or xyz.shape[2] != 3:
        raise ValueError("xyz must have shape (frames, n_atoms, 3)")
    f, n = xyz.shape[0], xyz.shape[1]
    per = max(1, min(f if f else 1, (1 << 16) // max(1, n * 12)))  # ~64 KB chunks like mdtraj, never more than 64 chunks (hdf5_min)
    D = hdf5_min.Dataset
    ds = {"coordinates": D(xyz, {"units": "nanometers"}, chunks=(per, n, 3))}

    def per_frame(a, tail):
        a = np.ascontiguousarray(np.broadcast_to(np.asarray(a, dtype=np.float32), (f,) + tail))
        return a
    if cell_lengths is not None:
        ds["cell_lengths"] = D(per_frame(cell_lengths, (3,)), {"units": "nanometers"}, chunks=(5461, 3))
        ds["cell_angles"] = D(per_frame(90.0 if cell_angles is None else cell_angles, (3,)), {"units": "degrees"}, chunks=(5461, 3))
    if potential_energy is not None:
        ds["potentialEnergy"] = D(per_frame(potential_energy, ()), {"units": "kilojoules_per_mole"}, chunks=(16384,))
    if topology is not None:
        text = json.dumps(topology).encode()
        ds["topology"] = D(np.array([text], dtype="S%d" % len(text)))
    root = {"conventions": "Pande", "conventionVersion": "1.1", "program": "MDTraj", "programVersion": "1.8.0", "title": title,
            "application": application}
    return hdf5_min.write(path, ds, root)

"""Reader for the equilibrium-sample caches the reference stores as mdtraj HDF5 trajectories
(benchmark/testsystems/bookkeepers.py:66-76 loads `{name}_samples.h5` with md.load; the one file the
reference ships is data/tiny_waterbox_constrained_samples.h5).

There is no h5py / PyTables / mdtraj in this image, and the cache is the only HDF5 the path touches,
so this is NOT an HDF5 implementation: it recognises the layout mdtraj's HDF5TrajectoryFile writes --
every array is a chunked dataset filtered with byte-shuffle + zlib, float32 little-endian:

    coordinates      (frames, n_atoms, 3)  nm        chunk = (k, n_atoms, 3)
    cell_lengths     (frames, 3)           nm        chunk = (5461, 3)
    cell_angles      (frames, 3)           degrees   chunk = (5461, 3)
    time / potentialEnergy / ...           (frames,)

and recovers the chunks by locating the zlib streams in the file, undoing the shuffle and matching
chunk byte counts against `n_atoms`.  Chunks are pre-allocated to full size, so the tail of the last
chunk is zero padding; frames that are entirely zero are dropped.
"""
import zlib

import numpy as np

_ZLIB_SECOND_BYTES = (0x01, 0x5E, 0x9C, 0xDA)


def _zlib_streams(buf):
    """yield (offset, decompressed bytes) for every complete zlib stream in `buf`, in file order"""
    i = 0
    end = 0
    while True:
        i = buf.find(b"\x78", i)
        if i < 0:
            return
        if i >= end and i + 1 < len(buf) and buf[i + 1] in _ZLIB_SECOND_BYTES:
            d = zlib.decompressobj()
            try:
                raw = d.decompress(buf[i:])
            except zlib.error:
                raw = None
            if raw is not None and d.eof and len(raw) >= 12:
                yield i, raw
                end = len(buf) - len(d.unused_data)  # do not rescan inside an accepted stream
        i += 1


def _unshuffle_f4(raw):
    """undo HDF5's byte-shuffle filter for 4-byte elements"""
    n = len(raw) // 4
    return np.frombuffer(raw[:4 * n], np.uint8).reshape(4, n).T.copy().view("<f4").ravel()


def load_mdtraj_hdf5(path, n_atoms):
    """Frames of an mdtraj HDF5 trajectory: dict(xyz (frames, n_atoms, 3) float32 nm, cell_lengths (frames, 3) or
    None, cell_angles (frames, 3) or None).  Raises ValueError when no coordinate chunk for `n_atoms` is found."""
    with open(path, "rb") as f:
        buf = f.read()
    if buf[:8] != b"\x89HDF\r\n\x1a\n":
        raise ValueError("{}: not an HDF5 file".format(path))
    frame_bytes = int(n_atoms) * 3 * 4
    coords, triples = [], []
    for _, raw in _zlib_streams(buf):
        if len(raw) % frame_bytes == 0:
            coords.append(_unshuffle_f4(raw).reshape(-1, int(n_atoms), 3))
        elif len(raw) % 12 == 0:
            triples.append(_unshuffle_f4(raw).reshape(-1, 3))
    if not coords:
        raise ValueError("{}: no coordinate chunk of {} atoms found".format(path, n_atoms))
    xyz = np.concatenate(coords, axis=0)
    valid = np.any(xyz != 0.0, axis=(1, 2))
    if not np.all(np.isfinite(xyz[valid])):
        raise ValueError("{}: non-finite coordinates".format(path))
    n_frames = int(valid.sum())
    xyz = np.ascontiguousarray(xyz[valid])
    lengths = angles = None
    for arr in triples:  # lengths are written before angles; angles of a rectangular box are 90 degrees
        head = arr[:n_frames]
        if len(head) < n_frames or not np.all(head > 0.0):
            continue
        if lengths is None:
            lengths = head.copy()
        elif angles is None:
            angles = head.copy()
    return dict(xyz=xyz, cell_lengths=lengths, cell_angles=angles)

"""A minimal HDF5 reader and writer for the one HDF5 layout this path touches: the mdtraj trajectory files that hold the
reference's equilibrium-sample caches (benchmark/testsystems/bookkeepers.py:75 `md.load(...)`, :106-113 `HDF5Reporter`).

There is no h5py / PyTables / mdtraj in this image, so both directions are written against the HDF5 file-format
specification (version 0 superblock, version 1 object headers, symbol-table groups = B-tree v1 + local heap + symbol
node, chunked datasets indexed by a B-tree v1, filters byte-shuffle + deflate) -- exactly the structures PyTables 3.x /
libhdf5 1.8-1.10 produce for mdtraj's HDF5TrajectoryFile, as found in the reference's
data/tiny_waterbox_constrained_samples.h5 (the parser below walks that file structurally; tests/test_serialization.py).

Reader: `read(path)` -> {dataset name: Dataset(array, attrs)}, root attributes under "/".
Writer: `write(path, datasets, root_attrs)` emits the same structures (one leaf B-tree node per dataset, at most 64 chunks).
Not supported (raises ValueError): superblock >= 2, object header v2, fractal-heap groups, datatypes other than
fixed-point / IEEE float / fixed-length string, filters other than shuffle and deflate.
"""
import struct
import time
import zlib

import numpy as np

UNDEF = 0xFFFFFFFFFFFFFFFF
SIGNATURE = b"\x89HDF\r\n\x1a\n"
GROUP_LEAF_K, GROUP_INTERNAL_K, CHUNK_K = 4, 16, 32  # libhdf5 defaults (the v0 superblock stores the first two)


class Dataset:
    def __init__(self, data, attrs=None, chunks=None, maxshape=None):
        self.data, self.attrs, self.chunks, self.maxshape = data, dict(attrs or {}), chunks, maxshape


# ------------------------------------------------------------------------------------------------ reader
class _Reader:
    def __init__(self, buf):
        self.b = buf
        if buf[:8] != SIGNATURE:
            raise ValueError("not an HDF5 file")
        if buf[8] != 0:
            raise ValueError("HDF5 superblock version %d is not supported (only 0)" % buf[8])
        if buf[13] != 8 or buf[14] != 8:
            raise ValueError("only 8-byte offsets and lengths are supported")
        self.root_ohdr, = struct.unpack_from("<Q", buf, 56 + 8)

    def u(self, fmt, off):
        return struct.unpack_from("<" + fmt, self.b, off)

    # ---- object headers
    def messages(self, addr):
        ver, _, nmsg, _, hsize = self.u("BBHII", addr)
        if ver != 1:
            raise ValueError("object header version %d is not supported (only 1)" % ver)
        out, blocks = [], [(addr + 16, addr + 16 + hsize)]
        while blocks and len(out) < nmsg:
            p, end = blocks.pop(0)
            while p + 8 <= end and len(out) < nmsg:
                mtype, msize, flags = self.u("HHB", p)
                body = self.b[p + 8:p + 8 + msize]
                if mtype == 0x10:
                    off, ln = struct.unpack("<QQ", body[:16])
                    blocks.append((off, off + ln))
                out.append((mtype, body))
                p += 8 + msize
        return out

    # ---- groups
    def group_entries(self, btree, heap):
        if self.b[heap:heap + 4] != b"HEAP":
            raise ValueError("bad local heap")
        heap_data, = self.u("Q", heap + 24)
        out = []

        def walk(addr):
            if self.b[addr:addr + 4] != b"TREE":
                raise ValueError("bad group B-tree node")
            ntype, level, nent = self.u("BBH", addr + 4)
            if ntype != 0:
                raise ValueError("expected a group B-tree node")
            for i in range(nent):
                child, = self.u("Q", addr + 24 + 16 * i + 8)
                if level > 0:
                    walk(child)
                else:
                    if self.b[child:child + 4] != b"SNOD":
                        raise ValueError("bad symbol node")
                    nsym, = self.u("H", child + 6)
                    for j in range(nsym):
                        name_off, ohdr = self.u("QQ", child + 8 + 40 * j)
                        s = heap_data + name_off
                        out.append((self.b[s:self.b.index(b"\0", s)].decode(), ohdr))
        walk(btree)
        return out

    # ---- message bodies
    @staticmethod
    def dataspace(body):
        ver, rank, flags = body[0], body[1], body[2]
        if ver == 1:
            p = 8
        elif ver == 2:
            if body[3] == 2:  # null dataspace
                return None, None
            p = 4
        else:
            raise ValueError("dataspace version %d" % ver)
        dims = struct.unpack_from("<%dQ" % rank, body, p)
        maxd = struct.unpack_from("<%dQ" % rank, body, p + 8 * rank) if flags & 1 else dims
        return tuple(dims), tuple(None if m == UNDEF else m for m in maxd)

    @staticmethod
    def datatype(body):
        cls, size = body[0] & 0x0F, struct.unpack_from("<I", body, 4)[0]
        if cls == 0:
            if body[1] & 1:
                raise ValueError("big-endian integers are not supported")
            return np.dtype("<%s%d" % ("i" if body[1] & 0x08 else "u", size))
        if cls == 1:
            if body[1] & 1:
                raise ValueError("big-endian floats are not supported")
            return np.dtype("<f%d" % size)
        if cls == 3:
            return np.dtype("S%d" % size)
        raise ValueError("datatype class %d is not supported" % cls)

    def attribute(self, body):
        ver, _, nsz, tsz, ssz = struct.unpack_from("<BBHHH", body, 0)
        if ver != 1:
            raise ValueError("attribute message version %d" % ver)
        pad = lambda n: (n + 7) & ~7  # noqa: E731
        p = 8
        name = body[p:p + nsz].split(b"\0")[0].decode()
        p += pad(nsz)
        dt = self.datatype(body[p:p + tsz])
        p += pad(tsz)
        dims, _ = self.dataspace(body[p:p + ssz])
        p += pad(ssz)
        if dims is None:
            return name, (b"" if dt.kind == "S" else None)
        count = int(np.prod(dims)) if dims else 1
        val = np.frombuffer(body[p:p + count * dt.itemsize], dt, count)
        if dt.kind == "S":
            val = [v.split(b"\0")[0] for v in val.tolist()]
        val = val[0] if not dims else val
        return name, (val.decode() if isinstance(val, bytes) else val)

    def chunks_of(self, btree, rank):
        out = []

        def walk(addr):
            if self.b[addr:addr + 4] != b"TREE":
                raise ValueError("bad chunk B-tree node")
            ntype, level, nent = self.u("BBH", addr + 4)
            if ntype != 1:
                raise ValueError("expected a chunk B-tree node")
            key = 8 + 8 * (rank + 1)
            for i in range(nent):
                p = addr + 24 + i * (key + 8)
                nbytes, mask = self.u("II", p)
                offs = self.u("%dQ" % (rank + 1), p + 8)
                child, = self.u("Q", p + key)
                if level > 0:
                    walk(child)
                else:
                    out.append((offs[:rank], nbytes, mask, child))
        if btree != UNDEF:
            walk(btree)
        return out

    def dataset(self, ohdr):
        msgs = self.messages(ohdr)
        dims = maxd = dt = layout = None
        filters, attrs = [], {}
        for mtype, body in msgs:
            if mtype == 0x01:
                dims, maxd = self.dataspace(body)
            elif mtype == 0x03:
                dt = self.datatype(body)
            elif mtype == 0x08:
                layout = body
            elif mtype == 0x0B:
                if body[0] != 1:
                    raise ValueError("filter pipeline version %d" % body[0])
                p = 8
                for _ in range(body[1]):
                    fid, nlen, _, ncd = struct.unpack_from("<HHHH", body, p)
                    p += 8 + ((nlen + 7) & ~7)
                    cd = struct.unpack_from("<%dI" % ncd, body, p)
                    p += 4 * (ncd + (ncd & 1))
                    filters.append((fid, cd))
            elif mtype == 0x0C:
                k, v = self.attribute(body)
                attrs[k] = v
        if dims is None or dt is None or layout is None:
            return None
        if layout[0] != 3:
            raise ValueError("data layout message version %d" % layout[0])
        count = int(np.prod(dims)) if dims else 1
        cls = layout[1]
        chunk_shape = None
        if cls == 0:  # compact
            size, = struct.unpack_from("<H", layout, 2)
            arr = np.frombuffer(layout[4:4 + size], dt, count).reshape(dims)
        elif cls == 1:  # contiguous
            addr, size = struct.unpack_from("<QQ", layout, 2)
            arr = (np.zeros(dims, dt) if addr == UNDEF else np.frombuffer(self.b[addr:addr + count * dt.itemsize], dt, count).reshape(dims))
        elif cls == 2:  # chunked
            rank = layout[2] - 1
            btree, = struct.unpack_from("<Q", layout, 3)
            chunk_shape = struct.unpack_from("<%dI" % rank, layout, 11)
            arr = np.zeros(dims, dt)
            for offs, nbytes, mask, addr in self.chunks_of(btree, rank):
                raw = self.b[addr:addr + nbytes]
                for k, (fid, cd) in reversed(list(enumerate(filters))):
                    if (mask >> k) & 1:
                        continue
                    if fid == 1:
                        raw = zlib.decompress(raw)
                    elif fid == 2:
                        es = cd[0] if cd else dt.itemsize
                        n = len(raw) // es
                        raw = np.frombuffer(raw[:n * es], np.uint8).reshape(es, n).T.tobytes() + raw[n * es:]
                    else:
                        raise ValueError("filter %d is not supported" % fid)
                block = np.frombuffer(raw, dt, int(np.prod(chunk_shape))).reshape(chunk_shape)
                sel = tuple(slice(o, min(o + c, d)) for o, c, d in zip(offs, chunk_shape, dims))
                if all(s.start < s.stop for s in sel):
                    arr[sel] = block[tuple(slice(0, s.stop - s.start) for s in sel)]
        else:
            raise ValueError("data layout class %d" % cls)
        return Dataset(arr, attrs, chunks=chunk_shape, maxshape=maxd)

    def read(self):
        out = {}
        root_attrs, sym = {}, None
        for mtype, body in self.messages(self.root_ohdr):
            if mtype == 0x0C:
                k, v = self.attribute(body)
                root_attrs[k] = v
            elif mtype == 0x11:
                sym = struct.unpack_from("<QQ", body, 0)
        out["/"] = Dataset(None, root_attrs)
        if sym is None:
            raise ValueError("root group has no symbol table (new-style groups are not supported)")
        for name, ohdr in self.group_entries(*sym):
            ds = self.dataset(ohdr)
            if ds is not None:
                out[name] = ds
        return out


def read(path):
    """{name: Dataset} of the root group's datasets; root attributes under "/".  Anything that is not a well-formed file of the
    supported kind -- wrong signature, unsupported structure, truncated or corrupt offsets -- raises ValueError."""
    with open(path, "rb") as f:
        buf = f.read()
    try:
        return _Reader(buf).read()
    except (struct.error, IndexError, zlib.error, OverflowError, MemoryError) as e:
        raise ValueError("{}: truncated or corrupt HDF5 file ({}: {})".format(path, type(e).__name__, e))


# ------------------------------------------------------------------------------------------------ writer
def _pad8(b):
    return b + b"\0" * (-len(b) % 8)


def _msg(mtype, body, flags=0):
    body = _pad8(body)
    return struct.pack("<HHB3x", mtype, len(body), flags) + body


def _dtype_msg(dt):
    dt = np.dtype(dt)
    if dt.kind == "f" and dt.itemsize == 4:
        return bytes.fromhex("11201f00") + struct.pack("<IHHBBBBI", 4, 0, 32, 23, 8, 0, 23, 127)
    if dt.kind == "f" and dt.itemsize == 8:
        return bytes.fromhex("11203f00") + struct.pack("<IHHBBBBI", 8, 0, 64, 52, 11, 0, 52, 1023)
    if dt.kind in "iu":
        return bytes([0x10, 0x08 if dt.kind == "i" else 0x00, 0, 0]) + struct.pack("<IHH", dt.itemsize, 0, 8 * dt.itemsize)
    if dt.kind == "S":
        return bytes.fromhex("13000000") + struct.pack("<I", dt.itemsize)
    raise ValueError("unsupported dtype %s" % dt)


def _string_attr(name, value, utf8=True):
    """scalar fixed-length string attribute, the way PyTables writes them (empty strings get a null dataspace)"""
    nm = name.encode() + b"\0"
    v = value.encode() if isinstance(value, str) else bytes(value)
    dtype = bytes([0x13, 0x10 if utf8 else 0x00, 0, 0]) + struct.pack("<I", max(len(v), 1))  # numpy-style: no terminator counted
    if not v:
        space, data, ssz = bytes.fromhex("02000002"), b"", 4
    else:
        space, data, ssz = struct.pack("<BBBx4x", 1, 0, 0), v, 8
    body = struct.pack("<BBHHH", 1, 0, len(nm), len(dtype), ssz) + _pad8(nm) + _pad8(dtype) + _pad8(space) + _pad8(data)
    return _msg(0x0C, body)


def _int_attr(name, value):
    nm = name.encode() + b"\0"
    dtype = _dtype_msg(np.int32)
    body = struct.pack("<BBHHH", 1, 0, len(nm), len(dtype), 8) + _pad8(nm) + _pad8(dtype) + struct.pack("<BBBx4x", 1, 0, 0) + _pad8(struct.pack("<i", int(value)))
    return _msg(0x0C, body)


def _object_header(messages):
    body = b"".join(messages)
    return struct.pack("<BBHII4x", 1, 0, len(messages), 1, len(body)) + body


def _shuffle(raw, es):
    n = len(raw) // es
    return np.frombuffer(raw, np.uint8).reshape(n, es).T.tobytes()


class _Writer:
    def __init__(self):
        self.buf = bytearray()

    def tell(self):
        return len(self.buf)

    def put(self, b):
        self.buf += b"\0" * (-len(self.buf) % 8)
        addr = len(self.buf)
        self.buf += b
        return addr

    def reserve(self, n):
        return self.put(b"\0" * n)

    def patch(self, addr, b):
        self.buf[addr:addr + len(b)] = b


def write(path, datasets, root_attrs=None, mtime=None):
    """datasets: {name: Dataset}; chunked (extendible along axis 0, shuffle + deflate) when `chunks` is given, else contiguous.
    A chunked dataset is written with at most 64 chunks (one leaf node of the chunk B-tree): `chunks[0]` is raised if needed."""
    mtime = int(time.time()) if mtime is None else int(mtime)
    w = _Writer()
    w.reserve(96)  # superblock, patched at the end
    names = sorted(datasets)
    if len(names) > 2 * GROUP_LEAF_K:
        raise ValueError("at most %d datasets in the root group" % (2 * GROUP_LEAF_K))
    # local heap data: names, 8-byte aligned, after the empty string at offset 0 (file order = insertion order is free)
    heap_data, name_off = bytearray(8), {}
    for nm in names:
        name_off[nm] = len(heap_data)
        heap_data += _pad8(nm.encode() + b"\0")
    # ---- root object header
    attrs = [("TITLE", ""), ("CLASS", "GROUP"), ("VERSION", "1.0"), ("PYTABLES_FORMAT_VERSION", "2.1")] + list((root_attrs or {}).items())
    root_msgs = [None] + [_string_attr(k, v) for k, v in attrs]  # slot 0: symbol table message, filled once the addresses exist
    root_len = 16 + 24 + sum(len(m) for m in root_msgs[1:])
    root_addr = w.reserve(root_len)
    btree_addr = w.reserve(24 + (2 * GROUP_INTERNAL_K + 1) * 8 + 2 * GROUP_INTERNAL_K * 8)
    heap_addr = w.reserve(32)
    heap_data_addr = w.put(bytes(heap_data))
    snod_addr = w.reserve(8 + 2 * GROUP_LEAF_K * 40)
    root_msgs[0] = _msg(0x11, struct.pack("<QQ", btree_addr, heap_addr))
    w.patch(root_addr, _object_header(root_msgs))
    w.patch(btree_addr, b"TREE" + struct.pack("<BBHQQ", 0, 0, 1, UNDEF, UNDEF) + struct.pack("<QQQ", 0, snod_addr, name_off[names[-1]] if names else 0))
    w.patch(heap_addr, b"HEAP" + struct.pack("<B3xQQQ", 0, len(heap_data), 1, heap_data_addr))

    ohdr_addr = {}
    for nm in names:
        ds = datasets[nm]
        arr = np.ascontiguousarray(ds.data)
        dt = arr.dtype.newbyteorder("<") if arr.dtype.byteorder == ">" else arr.dtype
        arr = arr.astype(dt, copy=False)
        rank = arr.ndim
        msgs = []
        attr_msgs = [(_int_attr(k, v) if isinstance(v, (int, np.integer)) else _string_attr(k, v, utf8=(k == "units"))) for k, v in ds.attrs.items()]
        if ds.chunks is not None:
            chunks = list(ds.chunks)
            chunks[0] = max(int(chunks[0]), -(-arr.shape[0] // (2 * CHUNK_K)), 1)
            maxd = [UNDEF] + list(arr.shape[1:])
            msgs.append(_msg(0x01, struct.pack("<BBB5x", 1, rank, 1) + struct.pack("<%dQ" % rank, *arr.shape) + struct.pack("<%dQ" % rank, *maxd)))
            msgs.append(_msg(0x03, _dtype_msg(dt), flags=1))
            msgs.append(_msg(0x05, bytes.fromhex("0203000100000000"), flags=1))
            msgs.append(_msg(0x0B, struct.pack("<BB6x", 1, 2) + struct.pack("<HHHH", 2, 8, 1, 1) + b"shuffle\0" + struct.pack("<II", dt.itemsize, 0) +
                             struct.pack("<HHHH", 1, 8, 1, 1) + b"deflate\0" + struct.pack("<II", 1, 0), flags=1))
            # chunk data, then the leaf node that indexes it
            key = 8 + 8 * (rank + 1)
            entries = []
            cshape = tuple(chunks)
            for c0 in range(0, arr.shape[0], cshape[0]):
                block = np.zeros(cshape, dt)
                part = arr[c0:c0 + cshape[0]]
                block[:part.shape[0]] = part
                raw = zlib.compress(_shuffle(block.tobytes(), dt.itemsize), 1)
                entries.append((len(raw), (c0,) + (0,) * rank, w.put(raw)))
            node = bytearray(24 + (2 * CHUNK_K + 1) * key + 2 * CHUNK_K * 8)
            node[:24] = b"TREE" + struct.pack("<BBHQQ", 1, 0, len(entries), UNDEF, UNDEF)
            p = 24
            for nbytes, offs, addr in entries:
                node[p:p + key + 8] = struct.pack("<II%dQQ" % (rank + 1), nbytes, 0, *offs, addr)
                p += key + 8
            last = (entries[-1][1][0] + cshape[0],) + tuple(cshape[1:]) + (dt.itemsize,) if entries else (0,) * (rank + 1)
            node[p:p + key] = struct.pack("<II%dQ" % (rank + 1), 0, 0, *last)
            tree_addr = w.put(bytes(node)) if entries else UNDEF
            msgs.append(_msg(0x08, struct.pack("<BBB", 3, 2, rank + 1) + struct.pack("<Q", tree_addr) + struct.pack("<%dI" % (rank + 1), *cshape, dt.itemsize), flags=1))
            msgs.append(_msg(0x12, struct.pack("<B3xI", 1, mtime)))
            attr_msgs = [_string_attr("CLASS", "EARRAY", utf8=False), _string_attr("VERSION", "1.1", utf8=False), _string_attr("TITLE", "", utf8=False),
                         _int_attr("EXTDIM", 0)] + attr_msgs
        else:
            data_addr = w.put(arr.tobytes())
            msgs.append(_msg(0x01, struct.pack("<BBB5x", 1, rank, 0) + struct.pack("<%dQ" % rank, *arr.shape)))
            msgs.append(_msg(0x03, _dtype_msg(dt), flags=1))
            msgs.append(_msg(0x05, bytes.fromhex("0202020100000000"), flags=1))
            msgs.append(_msg(0x08, struct.pack("<BB", 3, 1) + struct.pack("<QQ", data_addr, arr.nbytes), flags=1))
            msgs.append(_msg(0x12, struct.pack("<B3xI", 1, mtime)))
            attr_msgs = [_string_attr("CLASS", "ARRAY"), _string_attr("VERSION", "2.4"), _string_attr("TITLE", ""),
                         _string_attr("FLAVOR", "python")] + attr_msgs
        ohdr_addr[nm] = w.put(_object_header(msgs + attr_msgs))

    snod = bytearray(8 + 2 * GROUP_LEAF_K * 40)
    snod[:8] = b"SNOD" + struct.pack("<BBH", 1, 0, len(names))
    for j, nm in enumerate(names):
        snod[8 + 40 * j:8 + 40 * j + 24] = struct.pack("<QQII", name_off[nm], ohdr_addr[nm], 0, 0)
    w.patch(snod_addr, bytes(snod))
    eof = len(w.buf)
    sb = SIGNATURE + struct.pack("<BBBBBBBBHHI", 0, 0, 0, 0, 0, 8, 8, 0, GROUP_LEAF_K, GROUP_INTERNAL_K, 0) + \
        struct.pack("<QQQQ", 0, UNDEF, eof, UNDEF) + struct.pack("<QQII", 0, root_addr, 1, 0) + struct.pack("<QQ", btree_addr, heap_addr)
    assert len(sb) == 96
    w.patch(0, sb)
    with open(path, "wb") as f:
        f.write(bytes(w.buf))
    return path

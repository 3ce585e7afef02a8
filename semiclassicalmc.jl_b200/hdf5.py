"""Minimal HDF5 writer / reader in pure NumPy for the result files of /root/reference/src/IO.jl and `saveMeans!`
(ObservablesGeneric.jl:49-72, ObservablesTgHbn.jl:101-186, MonteCarlo.jl:232-234): nested groups of dense datasets.

Neither libhdf5 nor h5py exists in this image, so the file is laid out by hand after the HDF5 File Format Specification 1.8,
using only its oldest structures -- the ones every libhdf5 / HDF5.jl reads:

* superblock version 0 (8-byte offsets and lengths), root group as a symbol-table entry;
* groups = object header v1 with one Symbol Table message -> v1 B-tree (one leaf level) + local heap + one symbol-table node
  ("SNOD"), entries sorted by name; the file-wide "group leaf node K" is raised so that the largest group fits one node;
* datasets = object header v1 with Dataspace (v1), Datatype (v1), Fill Value (v2, undefined) and Data Layout (v3, contiguous);
* datatypes: little-endian float64 / float32 / int64 / int32 / uint64 / uint8, and complex128 as the compound {r, i} HDF5.jl uses.

Dimension order follows HDF5.jl: a Julia array of size (a, b) (column-major) is stored with HDF5 dims (b, a); the NumPy array
handed to `write` is C-ordered and written as is, so NumPy shape (b, a) reads back in Julia as an a x b matrix.

STATUS: written to the specification and checked by the independent parser below (`read`, which walks superblock -> B-tree ->
heap -> object headers by file address) and by byte-level known answers of the superblock; it has NOT been opened with libhdf5
(none available offline).  The Julia-`serialize`d `cfg` / `obs` blobs of checkpoint! (IO.jl:9-13) are not written: they are
Julia-internal; the state itself is the `state` dataset (IO.jl:146-150 layout).
"""
import struct

import numpy as np

SIG = b"\x89HDF\r\n\x1a\n"
UNDEF = 0xFFFFFFFFFFFFFFFF
K_INTERNAL = 16                      # group internal node K (superblock default)
FREE_NULL = 1                        # end marker of a local heap's free list


def _pad8(n):
    return (n + 7) & ~7


# ---- datatype messages ---------------------------------------------------------------------------------------------
def _dt_fixed(size, signed):
    return struct.pack("<BBBBI", 0x10, 0x08 if signed else 0x00, 0, 0, size) + struct.pack("<HH", 0, 8 * size)


def _dt_float(size):
    if size == 8:
        prop = struct.pack("<HHBBBBI", 0, 64, 52, 11, 0, 52, 1023)
        return struct.pack("<BBBBI", 0x11, 0x20, 63, 0, 8) + prop
    prop = struct.pack("<HHBBBBI", 0, 32, 23, 8, 0, 23, 127)
    return struct.pack("<BBBBI", 0x11, 0x20, 31, 0, 4) + prop


def _dt_complex():
    """compound {r: f64 @ 0, i: f64 @ 8} -- the on-disk form HDF5.jl gives ComplexF64 (version-1 member encoding)."""
    out = struct.pack("<BBBBI", 0x16, 2, 0, 0, 16)
    for name, off in ((b"r", 0), (b"i", 8)):
        out += name.ljust((len(name) + 8) // 8 * 8, b"\0") + struct.pack("<IB3xI4x4I", off, 0, 0, 0, 0, 0, 0) + _dt_float(8)
    return out


def _datatype(a):
    k, s = a.dtype.kind, a.dtype.itemsize
    if k == "f" and s in (4, 8):
        return _dt_float(s)
    if k in "iu" and s in (1, 2, 4, 8):
        return _dt_fixed(s, k == "i")
    if k == "c" and s == 16:
        return _dt_complex()
    raise TypeError(f"hdf5.write: unsupported dtype {a.dtype}")


def _canon(v):
    a = np.asarray(v)
    if a.dtype.kind == "b":
        a = a.astype(np.uint8)
    if a.dtype.kind == "c":
        a = a.astype(np.complex128)
    if a.dtype.kind == "f" and a.dtype.itemsize not in (4, 8):
        a = a.astype(np.float64)
    return np.asarray(a.astype(a.dtype.newbyteorder("<"), copy=False), order="C")       # (ascontiguousarray would make scalars 1-d)


# ---- writer --------------------------------------------------------------------------------------------------------
class _Writer:
    def __init__(self, leaf_k):
        self.buf = bytearray(96)
        self.leaf_k = leaf_k

    def alloc(self, data):
        addr = _pad8(len(self.buf))
        self.buf.extend(b"\0" * (addr - len(self.buf)))
        self.buf.extend(data)
        return addr

    @staticmethod
    def header(msgs):
        body = b""
        for mtype, data, flags in msgs:
            data = data.ljust(_pad8(len(data)), b"\0")
            body += struct.pack("<HHB3x", mtype, len(data), flags) + data
        return struct.pack("<BBHII4x", 1, 0, len(msgs), 1, len(body)) + body

    def dataset(self, value):
        a = _canon(value)
        raw = a.tobytes()
        daddr = self.alloc(raw) if raw else UNDEF
        space = struct.pack("<BBB5x", 1, a.ndim, 0) + b"".join(struct.pack("<Q", n) for n in a.shape)
        fill = struct.pack("<BBBB", 2, 2, 2, 0)
        layout = struct.pack("<BBQQ", 3, 1, daddr, len(raw))
        return self.alloc(self.header([(0x0001, space, 0), (0x0003, _datatype(a), 1), (0x0005, fill, 1), (0x0008, layout, 0)]))

    def group(self, tree):
        names = sorted(tree, key=lambda s: s.encode("utf-8"))
        entries = []
        for n in names:
            v = tree[n]
            entries.append((n, *(self.group(v) if isinstance(v, dict) else (self.dataset(v), None, None))))
        heap = bytearray(8)                                        # offset 0: the empty name
        offs = []
        for n in names:
            offs.append(len(heap))
            b = n.encode("utf-8") + b"\0"
            heap += b.ljust(_pad8(len(b)), b"\0")
        free_off = len(heap)
        heap += struct.pack("<QQ", FREE_NULL, 16)                  # one free block closes the segment (head must be a valid offset)
        snod = struct.pack("<4sBBH", b"SNOD", 1, 0, len(names))
        for (n, oh, bt, hp), off in zip(entries, offs):
            snod += struct.pack("<QQII", off, oh, 1 if bt is not None else 0, 0) + (struct.pack("<QQ", bt, hp) if bt is not None else b"\0" * 16)
        snod = snod.ljust(8 + 2 * self.leaf_k * 40, b"\0")
        snod_addr = self.alloc(snod) if names else None
        tree_node = struct.pack("<4sBBHQQ", b"TREE", 0, 0, 1 if names else 0, UNDEF, UNDEF)
        if names:
            tree_node += struct.pack("<QQQ", 0, snod_addr, offs[-1])
        tree_node = tree_node.ljust(24 + (2 * K_INTERNAL + 1) * 8 + 2 * K_INTERNAL * 8, b"\0")
        bt_addr = self.alloc(tree_node)
        hp_addr = _pad8(len(self.buf))
        self.alloc(struct.pack("<4sB3xQQQ", b"HEAP", 0, len(heap), free_off, hp_addr + 32) + bytes(heap))
        oh_addr = self.alloc(self.header([(0x0011, struct.pack("<QQ", bt_addr, hp_addr), 0)]))
        return oh_addr, bt_addr, hp_addr


def _max_group(tree):
    return max([len(tree)] + [_max_group(v) for v in tree.values() if isinstance(v, dict)])


def _insert(tree, path, value):
    parts = [p for p in path.split("/") if p]
    for p in parts[:-1]:
        tree = tree.setdefault(p, {})
        if not isinstance(tree, dict):
            raise ValueError(f"hdf5.write: {path!r} passes through a dataset")
    tree[parts[-1]] = value


def nest(flat):
    """{'means/energy/mean': x, 'rs': y} -> nested dict of groups (the `file["a/b"] = x` spelling of HDF5.jl)."""
    tree = {}
    for k, v in flat.items():
        if isinstance(v, dict):
            _insert(tree, k, nest(v))
        else:
            _insert(tree, k, v)
    return tree


def write(filename, tree):
    """Write a nested dict (str -> dict | array | scalar; '/' in a key opens groups) as an HDF5 file (h5open(filename, "w"))."""
    tree = nest(tree)
    w = _Writer(max(4, (_max_group(tree) + 1) // 2))
    oh, bt, hp = w.group(tree)
    sb = SIG + struct.pack("<BBBBBBBBHHI", 0, 0, 0, 0, 0, 8, 8, 0, w.leaf_k, K_INTERNAL, 0)
    sb += struct.pack("<QQQQ", 0, UNDEF, len(w.buf), UNDEF) + struct.pack("<QQII", 0, oh, 1, 0) + struct.pack("<QQ", bt, hp)
    assert len(sb) == 96
    w.buf[:96] = sb
    with open(filename, "wb") as f:
        f.write(w.buf)


def update(filename, tree):
    """h5open(filename, "cw"): add / replace entries of an existing file (rewritten as a whole), or create it."""
    import os
    old = read(filename) if os.path.exists(filename) else {}

    def merge(dst, src):
        for k, v in src.items():
            if isinstance(v, dict) and isinstance(dst.get(k), dict):
                merge(dst[k], v)
            else:
                dst[k] = v
    merge(old, nest(tree))
    write(filename, old)


# ---- reader (independent walk of the same structures; also reads v2 dataspaces and v2/v3 compound members) ----------
class _Reader:
    def __init__(self, data):
        self.d = data
        if data[:8] != SIG:
            raise ValueError("not an HDF5 file")
        ver, _, _, _, _, so, sl = struct.unpack_from("<7B", data, 8)
        if ver != 0 or so != 8 or sl != 8:
            raise ValueError("hdf5.read: only superblock version 0 with 8-byte offsets is supported")
        self.leaf_k, self.int_k = struct.unpack_from("<HH", data, 16)
        self.base, _, self.eof, _ = struct.unpack_from("<4Q", data, 24)
        if self.eof > len(data):
            raise ValueError("hdf5.read: truncated file")
        _, oh, cache, _, bt, hp = struct.unpack_from("<QQIIQQ", data, 56)
        self.root = (oh, bt, hp) if cache == 1 else (oh, None, None)

    def messages(self, addr):
        ver, _, nmsg, _, size = struct.unpack_from("<BBHII", self.d, addr)
        if ver != 1:
            raise ValueError("hdf5.read: only version-1 object headers are supported")
        out, blocks = [], [(addr + 16, size)]
        while blocks and len(out) < nmsg:
            p, n = blocks.pop(0)
            end = p + n
            while p + 8 <= end and len(out) < nmsg:
                mtype, msize, _ = struct.unpack_from("<HHB", self.d, p)
                body = self.d[p + 8:p + 8 + msize]
                if mtype == 0x0010:
                    blocks.append(struct.unpack_from("<QQ", body, 0))
                out.append((mtype, body))
                p += 8 + msize
        return out

    def heap_name(self, hp, off):
        sig, = struct.unpack_from("<4s", self.d, hp)
        assert sig == b"HEAP"
        size, _, seg = struct.unpack_from("<QQQ", self.d, hp + 8)
        assert off < size
        end = self.d.index(b"\0", seg + off)
        return bytes(self.d[seg + off:end]).decode("utf-8")

    def tree_nodes(self, addr):
        sig, ntype, level, used, _, _ = struct.unpack_from("<4sBBHQQ", self.d, addr)
        assert sig == b"TREE" and ntype == 0
        for i in range(used):
            child, = struct.unpack_from("<Q", self.d, addr + 24 + 8 + 16 * i)
            if level:
                yield from self.tree_nodes(child)
            else:
                yield child

    def group(self, bt, hp):
        out = {}
        for snod in self.tree_nodes(bt):
            sig, _, _, n = struct.unpack_from("<4sBBH", self.d, snod)
            assert sig == b"SNOD"
            for i in range(n):
                off, oh = struct.unpack_from("<QQ", self.d, snod + 8 + 40 * i)
                out[self.heap_name(hp, off)] = self.object(oh)
        return out

    def datatype(self, b, p=0):
        """-> (numpy dtype, bytes consumed)"""
        cv, b0, b1, b2, size = struct.unpack_from("<BBBBI", b, p)
        cls, ver = cv & 15, cv >> 4
        if cls == 0:
            return np.dtype(("<" if not b0 & 1 else ">") + ("i" if b0 & 8 else "u") + str(size)), 12
        if cls == 1:
            return np.dtype(("<" if not b0 & 1 else ">") + "f" + str(size)), 20
        if cls == 6:
            n, q, fields = b0 | (b1 << 8), p + 8, []
            for _ in range(n):
                e = b.index(b"\0", q)
                name = bytes(b[q:e]).decode()
                q = q + ((e - q + 8) // 8 * 8 if ver < 3 else e - q + 1)
                if ver < 3:
                    off, = struct.unpack_from("<I", b, q); q += 4
                    if ver == 1:
                        q += 28
                else:
                    nb = 1 if size < 256 else 2 if size < 65536 else 4
                    off = int.from_bytes(b[q:q + nb], "little"); q += nb
                dt, used = self.datatype(b, q)
                q += used
                fields.append((name, dt, off))
            dt = np.dtype({"names": [f[0] for f in fields], "formats": [f[1] for f in fields], "offsets": [f[2] for f in fields], "itemsize": size})
            return dt, q - p
        raise ValueError(f"hdf5.read: datatype class {cls} is not supported")

    def object(self, addr):
        shape = dt = layout = None
        for mtype, b in self.messages(addr):
            if mtype == 0x0011:
                return self.group(*struct.unpack_from("<QQ", b, 0))
            if mtype == 0x0001:
                ver, rank = b[0], b[1]
                shape = struct.unpack_from(f"<{rank}Q", b, 8 if ver == 1 else 4)
            elif mtype == 0x0003:
                dt, _ = self.datatype(b)
            elif mtype == 0x0008:
                if b[0] != 3 or b[1] != 1:
                    raise ValueError("hdf5.read: only contiguous version-3 layouts are supported")
                layout = struct.unpack_from("<QQ", b, 2)
        if shape is None or dt is None or layout is None:
            raise ValueError("hdf5.read: object is neither a group nor a dense dataset")
        n = int(np.prod(shape, dtype=np.int64)) if shape else 1
        a = np.frombuffer(self.d, dtype=dt, count=n, offset=layout[0]).reshape(shape) if n else np.zeros(shape, dtype=dt)
        if dt.names == ("r", "i"):
            a = a["r"] + 1j * a["i"]
        return a.copy() if a.ndim else a[()]


def read(filename):
    """Read an HDF5 file of the subset above into a nested dict (groups -> dicts, datasets -> arrays / scalars)."""
    with open(filename, "rb") as f:
        r = _Reader(f.read())
    if r.root[1] is None:
        return r.object(r.root[0])
    return r.group(r.root[1], r.root[2])

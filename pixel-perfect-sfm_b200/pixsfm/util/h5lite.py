"""h5lite — the subset of HDF5 the reference's files use, in pure Python (h5py / libhdf5 are not installable offline).

What pixsfm reads and writes through h5py / HighFive (reference features/store_features.py:1-88,
features/src/featuremap.cc:138-267, util/hloc.py:11-70):  groups nested by image name, numeric datasets (contiguous, or
chunked with chunk = one patch), numeric / fixed-string attributes.  Files h5py writes with default settings use the
"classic" layout — superblock version 0, version-1 object headers, groups as symbol tables (version-1 B-tree + local
heap + SNOD nodes), layout message version 3 — and that is what this module parses and emits (HDF5 File Format
Specification 1.8/3.0, sections III.A-E, IV.A.2).  Validated on the read side against real HDF5 files (the ten
`datasets/sacre_coeur/ground_truth/calibration_*.h5` of the reference tree, tests/test_h5lite.py); the write side is
validated by round trip only — say so wherever it matters.  Not supported (raises): new-style groups (link messages /
fractal heaps, `libver='latest'`), filters other than deflate / shuffle, variable-length and compound types.

API (the slice of h5py the callers need):
    f = File(path, "r" | "w");  g = f["a/b"];  name in g;  g.keys();  g.create_group(name);  g.attrs[name]
    d = g.create_dataset(name, data=array, chunks=None | tuple);  d[...] / d[i] / np.asarray(d);  d.shape, d.dtype, d.attrs
    f.visititems(fn);  f.close() (writing happens at close)."""
import mmap
import os
import struct
import zlib

import numpy as np

UNDEF = 0xFFFFFFFFFFFFFFFF
_SIG = b"\x89HDF\r\n\x1a\n"


# ------------------------------------------------------------------------------------------------------------ datatypes
def _dtype_from_message(buf):
    """datatype message -> numpy dtype"""
    cv, b0, b1, b2, size = struct.unpack_from("<BBBBI", buf, 0)
    cls, ver = cv & 0x0F, cv >> 4
    if ver not in (1, 2, 3):
        raise NotImplementedError("datatype message version %d" % ver)
    order = ">" if (b0 & 1) else "<"
    if cls == 0:                                   # fixed point
        signed = bool(b0 & 0x08)
        return np.dtype("%s%s%d" % (order, "i" if signed else "u", size))
    if cls == 1:                                   # IEEE float
        return np.dtype("%sf%d" % (order, size))
    if cls == 3:                                   # fixed-length string
        return np.dtype("S%d" % size)
    raise NotImplementedError("HDF5 datatype class %d" % cls)


def _dtype_to_message(dt):
    dt = np.dtype(dt)
    if dt.kind in "iu":
        bits = 0x08 if dt.kind == "i" else 0x00
        return struct.pack("<BBBBIHH", 0x10 | 0, bits, 0, 0, dt.itemsize, 0, dt.itemsize * 8)
    if dt.kind == "f":
        spec = {2: (15, 10, 5, 0, 10, 15), 4: (31, 23, 8, 0, 23, 127), 8: (63, 52, 11, 0, 52, 1023)}[dt.itemsize]
        sign, eloc, esize, mloc, msize, bias = spec
        # byte order LE, padding 0, mantissa normalisation 2 (implied msb) in bits 4-5, sign location in byte 1
        return struct.pack("<BBBBIHHBBBBI", 0x10 | 1, 0x20, sign, 0, dt.itemsize, 0, dt.itemsize * 8, eloc, esize, mloc, msize, bias)
    if dt.kind == "S":
        return struct.pack("<BBBBI", 0x10 | 3, 0x00, 0, 0, dt.itemsize)      # null-terminated, ASCII
    raise NotImplementedError("dtype %s" % dt)


def _dataspace_to_message(shape):
    if shape == ():
        return struct.pack("<BBB5x", 1, 0, 0)
    return struct.pack("<BBB5x", 1, len(shape), 0) + b"".join(struct.pack("<Q", int(s)) for s in shape)


def _dataspace_from_message(buf):
    ver, rank = buf[0], buf[1]
    if ver == 1:
        off = 8
    elif ver == 2:
        off = 4
        if buf[3] == 2:                            # null dataspace
            return None
    else:
        raise NotImplementedError("dataspace version %d" % ver)
    return tuple(struct.unpack_from("<%dQ" % rank, buf, off)) if rank else ()


def _pad8(b):
    return b + b"\0" * (-len(b) % 8)


# ------------------------------------------------------------------------------------------------------------ reading
class _Reader:
    def __init__(self, path):
        # mapped, not read: a feature cache is tens of GB and a lazily filled FeatureManager touches a few patches of it
        with open(path, "rb") as fh:
            try:
                self.buf = mmap.mmap(fh.fileno(), 0, access=mmap.ACCESS_READ)
            except (ValueError, OSError):              # empty file / a file system without mmap
                self.buf = fh.read()
        b = self.buf
        base = b.find(_SIG)
        if base != 0:
            raise ValueError("not an HDF5 file (or a user block is present)")
        ver = b[8]
        if ver not in (0, 1):
            raise NotImplementedError("superblock version %d (written with libver='latest'?)" % ver)
        if b[13] != 8 or b[14] != 8:
            raise NotImplementedError("offset/length sizes other than 8")
        off = 24 + (4 if ver == 1 else 0)
        self.base, = struct.unpack_from("<Q", b, off)
        root = off + 32                             # base, free space, eof, driver -> root symbol table entry
        self.root_header, = struct.unpack_from("<Q", b, root + 8)

    # -- object headers
    def messages(self, addr):
        b = self.buf
        if b[addr:addr + 4] == b"OHDR":
            raise NotImplementedError("version-2 object headers (file written with libver='latest')")
        ver, _, nmsg, _refs, hsize = struct.unpack_from("<BBHII", b, addr)
        if ver != 1:
            raise NotImplementedError("object header version %d" % ver)
        out = []
        blocks = [(addr + 16, hsize)]
        while blocks and len(out) < nmsg:
            pos, size = blocks.pop(0)
            end = pos + size
            while pos + 8 <= end and len(out) < nmsg:
                mtype, msize, flags = struct.unpack_from("<HHB", b, pos)
                data = b[pos + 8:pos + 8 + msize]
                pos += 8 + msize
                if mtype == 0x0010:                # continuation
                    coff, clen = struct.unpack_from("<QQ", data, 0)
                    blocks.append((coff, clen))
                out.append((mtype, flags, data))
        return out

    def children(self, addr):
        """name -> object header address of a group's links"""
        for mtype, _, data in self.messages(addr):
            if mtype == 0x0011:
                btree, heap = struct.unpack_from("<QQ", data, 0)
                names = {}
                self._walk_group_btree(btree, self._heap_data(heap), names)
                return names
            if mtype in (0x0002, 0x0006):
                raise NotImplementedError("new-style group (link messages)")
        return None

    def _heap_data(self, addr):
        b = self.buf
        assert b[addr:addr + 4] == b"HEAP"
        size, _free, data = struct.unpack_from("<QQQ", b, addr + 8)
        return b[data:data + size]

    def _walk_group_btree(self, addr, heap, names):
        b = self.buf
        assert b[addr:addr + 4] == b"TREE"
        _ntype, level, used = struct.unpack_from("<BBH", b, addr + 4)
        pos = addr + 24
        for k in range(used):
            child, = struct.unpack_from("<Q", b, pos + 8)
            pos += 16
            if level > 0:
                self._walk_group_btree(child, heap, names)
            else:
                assert b[child:child + 4] == b"SNOD"
                nsym, = struct.unpack_from("<H", b, child + 6)
                for s in range(nsym):
                    e = child + 8 + 40 * s
                    noff, oaddr = struct.unpack_from("<QQ", b, e)
                    end = heap.index(b"\0", noff)
                    names[heap[noff:end].decode()] = oaddr

    # -- datasets
    def dataset(self, addr):
        shape = dtype = None
        layout = None
        filters = []
        attrs = {}
        for mtype, _, data in self.messages(addr):
            if mtype == 0x0001:
                shape = _dataspace_from_message(data)
            elif mtype == 0x0003:
                dtype = _dtype_from_message(data)
            elif mtype == 0x0008:
                layout = data
            elif mtype == 0x000B:
                filters = self._filters(data)
            elif mtype == 0x000C:
                k, v = self._attribute(data)
                attrs[k] = v
        return shape, dtype, layout, filters, attrs

    def attributes(self, addr):
        out = {}
        for mtype, _, data in self.messages(addr):
            if mtype == 0x000C:
                k, v = self._attribute(data)
                out[k] = v
        return out

    def _attribute(self, data):
        ver = data[0]
        if ver == 1:
            nsz, tsz, ssz = struct.unpack_from("<HHH", data, 2)
            pos = 8
            pad = lambda n: n + (-n % 8)                       # noqa: E731
        elif ver in (2, 3):
            nsz, tsz, ssz = struct.unpack_from("<HHH", data, 2)
            pos = 8 + (1 if ver == 3 else 0)
            pad = lambda n: n                                  # noqa: E731
        else:
            raise NotImplementedError("attribute message version %d" % ver)
        name = data[pos:pos + nsz].split(b"\0")[0].decode(); pos += pad(nsz)
        dt = _dtype_from_message(data[pos:pos + tsz]); pos += pad(tsz)
        shape = _dataspace_from_message(data[pos:pos + ssz]); pos += pad(ssz)
        n = int(np.prod(shape)) if shape else 1
        arr = np.frombuffer(data, dt, n, pos).reshape(shape if shape else ())
        if dt.kind == "S":
            arr = np.char.decode(arr, "utf-8") if arr.shape else arr.item().split(b"\0")[0].decode()
            return name, arr
        return name, (arr.copy() if arr.shape else arr.item())

    def _filters(self, data):
        ver, n = data[0], data[1]
        pos = 8 if ver == 1 else 2
        out = []
        for _ in range(n):
            fid, = struct.unpack_from("<H", data, pos)
            if ver == 1 or fid >= 256:
                nlen, _flags, ncd = struct.unpack_from("<HHH", data, pos + 2)
                pos += 8 + (nlen + (-nlen % 8) if ver == 1 else nlen)
            else:
                _flags, ncd = struct.unpack_from("<HH", data, pos + 2)
                pos += 6
            cd = struct.unpack_from("<%dI" % ncd, data, pos)
            pos += 4 * ncd + (4 if (ver == 1 and ncd % 2) else 0)
            out.append((fid, cd))
        return out

    def read(self, shape, dtype, layout, filters):
        ver, cls = layout[0], layout[1]
        if ver != 3:
            raise NotImplementedError("layout message version %d" % ver)
        n = int(np.prod(shape)) if shape else 1
        if cls == 0:                                # compact
            size, = struct.unpack_from("<H", layout, 2)
            return np.frombuffer(layout, dtype, n, 4).reshape(shape).copy()
        if cls == 1:                                # contiguous
            addr, size = struct.unpack_from("<QQ", layout, 2)
            if addr == UNDEF:
                return np.zeros(shape, dtype)
            return np.frombuffer(self.buf, dtype, n, addr).reshape(shape).copy()
        if cls == 2:                                # chunked, version-1 B-tree
            rank1 = layout[2]
            btree, = struct.unpack_from("<Q", layout, 3)
            cdims = struct.unpack_from("<%dI" % rank1, layout, 11)[:-1]
            out = np.zeros(shape, dtype)
            if btree != UNDEF:
                self._walk_chunk_btree(btree, rank1, cdims, dtype, filters, out)
            return out
        raise NotImplementedError("layout class %d" % cls)

    def _walk_chunk_btree(self, addr, rank1, cdims, dtype, filters, out):
        b = self.buf
        assert b[addr:addr + 4] == b"TREE"
        _ntype, level, used = struct.unpack_from("<BBH", b, addr + 4)
        ksize = 8 + 8 * rank1
        pos = addr + 24
        for _ in range(used):
            csize, _mask = struct.unpack_from("<II", b, pos)
            offs = struct.unpack_from("<%dQ" % rank1, b, pos + 8)[:-1]
            child, = struct.unpack_from("<Q", b, pos + ksize)
            pos += ksize + 8
            if level > 0:
                self._walk_chunk_btree(child, rank1, cdims, dtype, filters, out)
                continue
            raw = b[child:child + csize]
            for fid, cd in reversed(filters):
                if fid == 1:
                    raw = zlib.decompress(raw)
                elif fid == 2:
                    es = cd[0] if cd else dtype.itemsize
                    raw = np.frombuffer(raw, np.uint8).reshape(es, -1).T.tobytes()
                else:
                    raise NotImplementedError("HDF5 filter %d" % fid)
            chunk = np.frombuffer(raw, dtype, int(np.prod(cdims))).reshape(cdims)
            sel = tuple(slice(o, min(o + c, s)) for o, c, s in zip(offs, cdims, out.shape))
            out[sel] = chunk[tuple(slice(0, s.stop - s.start) for s in sel)]


# ------------------------------------------------------------------------------------------------------------ objects
class _Attrs(dict):
    pass


class Dataset:
    def __init__(self, file, name, addr=None, data=None, chunks=None):
        self.file, self.name = file, name
        self._addr, self._data, self.chunks = addr, data, chunks
        if addr is not None:
            self.shape, self.dtype, self._layout, self._filters, attrs = file._r.dataset(addr)
            self.attrs = _Attrs(attrs)
        else:
            self.shape, self.dtype = data.shape, data.dtype
            self.attrs = _Attrs()

    @property
    def parent(self):
        return self.file["/".join(self.name.strip("/").split("/")[:-1]) or "/"]

    def _load(self):
        if self._data is None:
            self._data = self.file._r.read(self.shape, self.dtype, self._layout, self._filters)
        return self._data

    def __array__(self, dtype=None, copy=None):
        a = self._load()
        return a if dtype is None else a.astype(dtype)

    def __getitem__(self, idx):
        return self._load()[idx]

    def __len__(self):
        return self.shape[0]


class Group:
    def __init__(self, file, name, addr=None):
        self.file, self.name = file, name
        self._addr = addr
        self._children = None if addr is not None else {}
        self.attrs = _Attrs(file._r.attributes(addr)) if addr is not None else _Attrs()

    @property
    def parent(self):
        return self.file["/".join(self.name.strip("/").split("/")[:-1]) or "/"]

    def _kids(self):
        if self._children is None:
            self._children = {}
            for k, a in (self.file._r.children(self._addr) or {}).items():
                self._children[k] = a                    # address until opened
        return self._children

    def _open(self, key):
        kids = self._kids()
        v = kids[key]
        if isinstance(v, int):
            full = (self.name.rstrip("/") + "/" + key)
            v = Group(self.file, full, v) if self.file._r.children(v) is not None else Dataset(self.file, full, v)
            kids[key] = v
        return v

    def __contains__(self, path):
        try:
            self[path]
            return True
        except KeyError:
            return False

    def __getitem__(self, path):
        node = self.file if path.startswith("/") and self is not self.file else self
        for part in [p for p in path.split("/") if p]:
            if not isinstance(node, Group) or part not in node._kids():
                raise KeyError(path)
            node = node._open(part)
        return node

    def keys(self):
        return sorted(self._kids().keys())

    def __iter__(self):
        return iter(self.keys())

    def items(self):
        return [(k, self._open(k)) for k in self.keys()]

    def _require_writable(self):
        if self.file.mode != "w":
            raise ValueError("file is open read-only")

    def create_group(self, path):
        self._require_writable()
        node = self
        for part in [p for p in path.split("/") if p]:
            kids = node._kids()
            if part not in kids:
                kids[part] = Group(self.file, node.name.rstrip("/") + "/" + part)
            node = kids[part]
            if not isinstance(node, Group):
                raise ValueError("%s is a dataset" % part)
        return node

    def require_group(self, path):
        return self.create_group(path)

    def create_dataset(self, path, data=None, chunks=None, **_):
        self._require_writable()
        parts = [p for p in path.split("/") if p]
        parent = self.create_group("/".join(parts[:-1])) if len(parts) > 1 else self
        arr = np.asarray(data)
        arr = arr if arr.flags["C_CONTIGUOUS"] else arr.copy()
        if arr.dtype.kind == "U":
            arr = np.char.encode(arr, "utf-8")
        if arr.dtype.kind not in "iufS":
            raise NotImplementedError("dataset dtype %s" % arr.dtype)
        if parts[-1] in parent._kids():
            raise ValueError("name already exists: %s" % path)
        d = Dataset(self.file, parent.name.rstrip("/") + "/" + parts[-1], data=arr, chunks=tuple(chunks) if chunks else None)
        parent._kids()[parts[-1]] = d
        return d

    def visititems(self, fn):
        def walk(g, prefix):
            for k, v in g.items():
                name = prefix + k
                r = fn(name, v)
                if r is not None:
                    return r
                if isinstance(v, Group):
                    r = walk(v, name + "/")
                    if r is not None:
                        return r
            return None
        return walk(self, "")


class File(Group):
    def __init__(self, path, mode="r"):
        if mode not in ("r", "w"):
            raise NotImplementedError("h5lite opens files for 'r' or 'w' (no append)")
        self.path, self.mode = str(path), mode
        self._r = _Reader(self.path) if mode == "r" else None
        Group.__init__(self, self, "/", self._r.root_header if mode == "r" else None)
        self._closed = False

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    def close(self):
        if not self._closed and self.mode == "w":
            _Writer(self).write(self.path)
        self._closed = True

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


# ------------------------------------------------------------------------------------------------------------ writing
class _Writer:
    """Lays the tree out in one pass: superblock (version 1: carries the chunk-B-tree fan-out) | per object: header,
    then its data.  Every group gets ONE level-0 B-tree node over SNODs of up to 2*leaf_K symbols; leaf_K / the
    internal fan-outs are chosen from the largest group / chunk count (they are file-wide parameters of the format)."""

    def __init__(self, root):
        self.root = root
        self.buf = bytearray()

    def alloc(self, data, align=8):
        pad = -len(self.buf) % align
        self.buf += b"\0" * pad
        addr = len(self.buf)
        self.buf += data
        return addr

    def write(self, path):
        def walk(g):
            n = len(g._kids())
            m = 0
            for v in g._kids().values():
                if isinstance(v, Group):
                    a, b = walk(v); n = max(n, a); m = max(m, b)
                elif v.chunks:
                    m = max(m, int(np.prod([-(-s // c) for s, c in zip(v.shape, v.chunks)])))
            return n, m
        nmax, cmax = walk(self.root)
        self.leaf_k = min(32767, max(4, -(-nmax // 2)))
        self.int_k = min(32767, max(16, -(-(-(-nmax // (2 * self.leaf_k))) // 2)))
        self.chunk_k = min(32767, max(32, -(-cmax // 2)))
        if cmax > 2 * self.chunk_k or nmax > 4 * self.leaf_k * self.int_k:
            raise NotImplementedError("too many links / chunks for single-level B-trees")
        self.buf += b"\0" * 104                           # superblock v1 (24 + 4 + 32 + 40 = 100, padded)
        root_addr, btree, heap = self.group(self.root)
        sb = _SIG + struct.pack("<BBBBBBBBHHIHH", 1, 0, 0, 0, 0, 8, 8, 0, self.leaf_k, self.int_k, 0, self.chunk_k, 0)
        sb += struct.pack("<QQQQ", 0, UNDEF, len(self.buf), UNDEF)
        sb += struct.pack("<QQII", 0, root_addr, 1, 0) + struct.pack("<QQ", btree, heap)
        self.buf[:len(sb)] = sb
        # a new inode, moved into place: a reader that still has the old file mapped keeps seeing the old bytes
        tmp = "%s.tmp%d" % (path, os.getpid())
        with open(tmp, "wb") as fh:
            fh.write(bytes(self.buf))
        os.replace(tmp, path)

    # -- messages / headers
    @staticmethod
    def msg(mtype, data, flags=0):
        data = _pad8(data)
        return struct.pack("<HHB3x", mtype, len(data), flags) + data

    def header(self, msgs):
        body = b"".join(msgs)
        return self.alloc(struct.pack("<BBHII4x", 1, 0, len(msgs), 1, len(body)) + body)

    def attr_msgs(self, attrs):
        out = []
        for k, v in attrs.items():
            arr = np.asarray(v)
            if arr.dtype.kind == "U":
                arr = np.char.encode(arr, "utf-8")
            if arr.dtype.kind == "b":
                arr = arr.astype(np.int8)
            if arr.dtype.kind == "i" and arr.dtype.itemsize == 8 and isinstance(v, (int, np.integer)):
                arr = arr.astype(np.int64)
            name = k.encode() + b"\0"
            dt, ds = _dtype_to_message(arr.dtype), _dataspace_to_message(arr.shape)
            body = struct.pack("<BxHHH", 1, len(name), len(dt), len(ds)) + _pad8(name) + _pad8(dt) + _pad8(ds) + arr.tobytes()
            out.append(self.msg(0x000C, body))
        return out

    def group(self, g):
        # children first (their header addresses go into the symbol nodes)
        entries = []
        for name in sorted(g._kids().keys(), key=lambda s: s.encode()):
            v = g._kids()[name]
            if isinstance(v, Group):
                addr, bt, hp = self.group(v)
                entries.append((name, addr, 1, struct.pack("<QQ", bt, hp)))
            else:
                entries.append((name, self.dataset(v), 0, b"\0" * 16))
        # local heap: "" at offset 0, then the names (8-byte aligned)
        heap = bytearray(b"\0" * 8)
        offs = []
        for name, *_ in entries:
            offs.append(len(heap))
            heap += _pad8(name.encode() + b"\0")
        free = len(heap)
        heap += struct.pack("<QQ", 1, 16) if True else b""   # one free block of 16 bytes (next = 1: last)
        data_addr = self.alloc(bytes(heap))
        heap_addr = self.alloc(b"HEAP" + struct.pack("<B3xQQQ", 0, len(heap), free, data_addr))
        # symbol nodes
        per = 2 * self.leaf_k
        snods, keys = [], [0]
        for s in range(0, max(len(entries), 1), per):
            part = entries[s:s + per]
            body = b"SNOD" + struct.pack("<BxH", 1, len(part))
            for (name, addr, ctype, scratch), o in zip(part, offs[s:s + per]):
                body += struct.pack("<QQII", o, addr, ctype, 0) + scratch
            body += b"\0" * (40 * (per - len(part)))
            snods.append(self.alloc(body))
            keys.append(offs[s + len(part) - 1] if part else 0)
        node = b"TREE" + struct.pack("<BBHQQ", 0, 0, len(snods), UNDEF, UNDEF)
        for k, child in zip(keys, snods):
            node += struct.pack("<QQ", k, child)
        node += struct.pack("<Q", keys[-1])
        node += b"\0" * (16 * (2 * self.int_k - len(snods)))
        btree_addr = self.alloc(node)
        msgs = [self.msg(0x0011, struct.pack("<QQ", btree_addr, heap_addr))] + self.attr_msgs(g.attrs)
        return self.header(msgs), btree_addr, heap_addr

    def dataset(self, d):
        arr = d._load() if d._data is None else d._data
        msgs = [self.msg(0x0001, _dataspace_to_message(arr.shape)), self.msg(0x0003, _dtype_to_message(arr.dtype), flags=1),
                self.msg(0x0005, struct.pack("<BBBB", 2, 2, 2, 0))]       # fill value: v2, alloc late, never written, undefined
        if d.chunks and arr.ndim:
            cd = tuple(int(c) for c in d.chunks)
            grid = [range(0, s, c) for s, c in zip(arr.shape, cd)]
            import itertools
            chunk_bytes = int(np.prod(cd)) * arr.dtype.itemsize
            recs = []
            for origin in itertools.product(*grid):
                block = np.zeros(cd, arr.dtype)
                sel = tuple(slice(o, min(o + c, s)) for o, c, s in zip(origin, cd, arr.shape))
                block[tuple(slice(0, s.stop - s.start) for s in sel)] = arr[sel]
                recs.append((origin, self.alloc(block.tobytes())))
            rank1 = arr.ndim + 1
            node = b"TREE" + struct.pack("<BBHQQ", 1, 0, len(recs), UNDEF, UNDEF)
            for origin, addr in recs:
                node += struct.pack("<II", chunk_bytes, 0) + struct.pack("<%dQ" % rank1, *origin, 0) + struct.pack("<Q", addr)
            node += struct.pack("<II", 0, 0) + struct.pack("<%dQ" % rank1, *arr.shape, 0)      # final key: one past the end
            node += b"\0" * ((8 + 8 * rank1 + 8) * (2 * self.chunk_k - len(recs)))
            bt = self.alloc(node)
            layout = struct.pack("<BBB", 3, 2, rank1) + struct.pack("<Q", bt) + struct.pack("<%dI" % rank1, *cd, arr.dtype.itemsize)
        else:
            raw = arr.tobytes()
            addr = self.alloc(raw) if raw else UNDEF
            layout = struct.pack("<BBQQ", 3, 1, addr, len(raw))
        msgs.append(self.msg(0x0008, layout))
        return self.header(msgs + self.attr_msgs(d.attrs))

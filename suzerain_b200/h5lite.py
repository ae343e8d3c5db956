"""Minimal read-only HDF5 reader for Suzerain restart files, NumPy only.

The image has neither libhdf5 nor h5py, and the reference's restart / fixture files
(``fields/*.h5``, written through ESIO: ``suzerain/support/support.cpp``, ``driver_base.cpp:1734-1795``)
are plain "version 0 superblock" HDF5: old-style groups (symbol-table B-trees + local heaps),
version-1 object headers, little-endian IEEE / integer atomic types, contiguous, compact or chunked
(optionally deflate / shuffle) layouts.  That subset is what this module reads; anything else
raises ``H5Error``.  (All of the reference's files are contiguous: the chunked / filter code follows the
specification but has no fixture to exercise it.)  Layout follows the public HDF5 File Format Specification (version 1.1/2.0).

    f = H5File(path); f.keys(); f["Dy0T"]  -> numpy array;  f.attrs("rho")  -> dict
"""
from __future__ import annotations

import struct
import zlib

import numpy as np


class H5Error(RuntimeError):
    pass


UNDEF = 0xFFFFFFFFFFFFFFFF


class _Dataset:
    def __init__(self):
        self.shape = None
        self.dtype = None
        self.layout = None          # ("contiguous", addr, size) | ("compact", bytes) | ("chunked", btree, chunk_dims)
        self.filters = []
        self.attrs = {}
        self.symtab = None          # (btree, heap) when the object is a group


class H5File:
    def __init__(self, path):
        with open(path, "rb") as fh:
            self.b = fh.read()
        b = self.b
        if b[:8] != b"\x89HDF\r\n\x1a\n":
            raise H5Error("not an HDF5 file")
        if b[8] != 0:
            raise H5Error(f"superblock version {b[8]} not supported")
        self.O, self.L = b[13], b[14]
        if self.O != 8 or self.L != 8:
            raise H5Error("only 8-byte offsets / lengths are supported")
        self.base = self._u(24, 8)
        root_entry = 24 + 4 * 8
        self.root = self._read_object(self._u(root_entry + 8, 8))
        if self.root.symtab is None:
            # cached in the scratch pad (cache type 1)
            if self._u(root_entry + 16, 4) == 1:
                self.root.symtab = (self._u(root_entry + 24, 8), self._u(root_entry + 32, 8))
            else:
                raise H5Error("root group has no symbol table")
        self._links = dict(self._group_links(*self.root.symtab))
        self._cache = {}

    # ---- primitives ----
    def _u(self, off, n):
        return int.from_bytes(self.b[off:off + n], "little")

    def keys(self):
        return sorted(self._links)

    def __contains__(self, name):
        return name in self._links

    def _object(self, name):
        if name not in self._cache:
            if name not in self._links:
                raise KeyError(name)
            self._cache[name] = self._read_object(self._links[name])
        return self._cache[name]

    def attrs(self, name):
        return dict(self._object(name).attrs)

    def shape(self, name):
        return tuple(self._object(name).shape or ())

    def __getitem__(self, name):
        d = self._object(name)
        if d.symtab is not None and d.layout is None:
            raise H5Error(f"{name} is a group")
        return self._read_data(d)

    # ---- groups ----
    def _heap_string(self, heap_addr, off):
        h = self.base + heap_addr
        if self.b[h:h + 4] != b"HEAP":
            raise H5Error("bad local heap")
        data = self.base + self._u(h + 8 + 2 * 8, 8)
        end = self.b.index(b"\0", data + off)
        return self.b[data + off:end].decode("ascii")

    def _group_links(self, btree, heap):
        out = []

        def node(addr):
            a = self.base + addr
            sig = self.b[a:a + 4]
            if sig == b"TREE":
                if self.b[a + 4] != 0:
                    raise H5Error("group B-tree expected")
                n = self._u(a + 6, 2)
                p = a + 8 + 16                       # past the sibling pointers
                for i in range(n):
                    child = self._u(p + 8 + i * 16, 8)      # key_i (8) child_i (8) ...
                    node(child)
            elif sig == b"SNOD":
                n = self._u(a + 6, 2)
                p = a + 8
                for i in range(n):
                    e = p + 40 * i
                    out.append((self._heap_string(heap, self._u(e, 8)), self._u(e + 8, 8)))
            else:
                raise H5Error(f"unexpected node signature {sig!r}")
        node(btree)
        return out

    # ---- object headers (version 1) ----
    def _read_object(self, addr):
        a = self.base + addr
        if self.b[a] != 1:
            raise H5Error(f"object header version {self.b[a]} not supported")
        nmsg = self._u(a + 2, 2)
        size = self._u(a + 8, 4)
        d = _Dataset()
        blocks = [(a + 16, size)]
        seen = 0
        while blocks and seen < nmsg:
            p, left = blocks.pop(0)
            end = p + left
            while p + 8 <= end and seen < nmsg:
                mtype, msize, mflags = self._u(p, 2), self._u(p + 2, 2), self.b[p + 4]
                body = p + 8
                seen += 1
                if mflags & 2:
                    raise H5Error("shared header messages are not supported")
                if mtype == 0x0010:                                   # continuation
                    blocks.append((self.base + self._u(body, 8), self._u(body + 8, 8)))
                elif mtype == 0x0001:
                    d.shape = self._dataspace(body)
                elif mtype == 0x0003:
                    d.dtype = self._datatype(body)[0]
                elif mtype == 0x0008:
                    d.layout = self._layout(body)
                elif mtype == 0x000B:
                    d.filters = self._filters(body)
                elif mtype == 0x000C:
                    try:
                        k, v = self._attribute(body)
                        d.attrs[k] = v
                    except H5Error:
                        pass
                elif mtype == 0x0011:
                    d.symtab = (self._u(body, 8), self._u(body + 8, 8))
                p = body + msize
        return d

    def _dataspace(self, p):
        ver, rank, flags = self.b[p], self.b[p + 1], self.b[p + 2]
        if ver == 1:
            q = p + 8
        elif ver == 2:
            q = p + 4
        else:
            raise H5Error(f"dataspace version {ver}")
        return tuple(self._u(q + 8 * i, 8) for i in range(rank))

    def _datatype(self, p):
        """-> (numpy dtype, bytes consumed)"""
        cls, ver = self.b[p] & 0x0F, self.b[p] >> 4
        bits0 = self.b[p + 1]
        size = self._u(p + 4, 4)
        if cls == 1:                                                  # floating point
            if bits0 & 1:
                raise H5Error("big-endian floats are not supported")
            if size not in (4, 8):
                raise H5Error(f"float size {size}")
            return np.dtype(f"<f{size}"), 8 + 12
        if cls == 0:                                                  # fixed point
            if bits0 & 1:
                raise H5Error("big-endian integers are not supported")
            return np.dtype(f"<{'i' if bits0 & 8 else 'u'}{size}"), 8 + 4
        if cls == 3:                                                  # fixed-length string
            return np.dtype(f"S{size}"), 8
        if cls == 10:                                                 # array (ESIO: complex = double[2])
            nd = self.b[p + 8]
            q = p + 9 + (3 if ver < 3 else 0)
            dims = tuple(self._u(q + 4 * i, 4) for i in range(nd))
            q += 4 * nd * (2 if ver < 3 else 1)                       # version 2 carries permutation indices
            base, used = self._datatype(q)
            return np.dtype((base, dims)), (q - p) + used
        raise H5Error(f"datatype class {cls} (version {ver}) not supported")

    def _layout(self, p):
        ver = self.b[p]
        if ver == 3:
            cls = self.b[p + 1]
            if cls == 0:
                n = self._u(p + 2, 2)
                return ("compact", self.b[p + 4:p + 4 + n])
            if cls == 1:
                return ("contiguous", self._u(p + 2, 8), self._u(p + 10, 8))
            if cls == 2:
                nd = self.b[p + 2]
                bt = self._u(p + 3, 8)
                dims = tuple(self._u(p + 11 + 4 * i, 4) for i in range(nd))
                return ("chunked", bt, dims)
            raise H5Error(f"layout class {cls}")
        if ver in (1, 2):
            nd, cls = self.b[p + 1], self.b[p + 2]
            q = p + 8
            addr = None
            if cls != 0:
                addr = self._u(q, 8)
                q += 8
            dims = tuple(self._u(q + 4 * i, 4) for i in range(nd))
            q += 4 * nd
            if cls == 1:
                return ("contiguous", addr, None)
            if cls == 2:
                esize = self._u(q, 4)
                return ("chunked", addr, dims + (esize,))
            n = self._u(q, 4)
            return ("compact", self.b[q + 4:q + 4 + n])
        raise H5Error(f"layout version {ver}")

    def _filters(self, p):
        ver, n = self.b[p], self.b[p + 1]
        out = []
        q = p + (8 if ver == 1 else 2)
        for _ in range(n):
            fid = self._u(q, 2)
            if ver == 1 or fid >= 256:
                namelen = self._u(q + 2, 2)
                ncv = self._u(q + 6, 2)
                q += 8 + ((namelen + 7) // 8 * 8 if ver == 1 else namelen)
            else:
                ncv = self._u(q + 4, 2)
                q += 6
            cv = [self._u(q + 4 * i, 4) for i in range(ncv)]
            q += 4 * ncv
            if ver == 1 and ncv % 2:
                q += 4
            out.append((fid, cv))
        return out

    def _attribute(self, p):
        ver = self.b[p]
        if ver not in (1, 2, 3):
            raise H5Error("attribute version")
        nlen, tlen, slen = self._u(p + 2, 2), self._u(p + 4, 2), self._u(p + 6, 2)
        q = p + 8 + (1 if ver == 3 else 0)
        pad = (lambda n: (n + 7) // 8 * 8) if ver == 1 else (lambda n: n)
        name = self.b[q:q + nlen].split(b"\0")[0].decode("ascii")
        q += pad(nlen)
        dt, _ = self._datatype(q)
        q += pad(tlen)
        shape = self._dataspace(q) if slen else ()
        q += pad(slen)
        count = int(np.prod(shape)) if shape else 1
        arr = np.frombuffer(self.b, dtype=dt, count=count, offset=q).reshape(shape)
        if dt.kind == "S":
            return name, arr.reshape(-1)[0].split(b"\0")[0].decode("ascii", "replace")
        return name, (arr.copy() if shape else arr.reshape(-1)[0])

    # ---- raw data ----
    def _read_data(self, d):
        if d.shape is None or d.dtype is None or d.layout is None:
            raise H5Error("not a dataset")
        if d.dtype.subdtype is not None:
            # array datatype: read the base type with the array dimensions appended to the shape
            base, sub = d.dtype.subdtype
            e = _Dataset()
            e.shape, e.dtype, e.filters = tuple(d.shape) + tuple(sub), base, d.filters
            if d.layout[0] == "chunked":
                # chunk dims are in elements of the array type: one more (full) dimension per array dim
                e.layout = ("chunked", d.layout[1], tuple(d.layout[2][:len(d.shape)]) + tuple(sub) + (base.itemsize,))
                e._array_rank = len(sub)
            else:
                e.layout = d.layout
            return self._read_data(e)
        count = int(np.prod(d.shape)) if d.shape else 1
        kind = d.layout[0]
        if kind == "compact":
            return np.frombuffer(d.layout[1], dtype=d.dtype, count=count).reshape(d.shape).copy()
        if kind == "contiguous":
            addr = d.layout[1]
            if addr == UNDEF:
                return np.zeros(d.shape, dtype=d.dtype)
            return np.frombuffer(self.b, dtype=d.dtype, count=count, offset=self.base + addr).reshape(d.shape).copy()
        # chunked
        _, bt, cdims = d.layout
        arank = getattr(d, "_array_rank", 0)
        rank = len(d.shape)
        chunk = cdims[:rank]
        krank = rank - arank                        # dimensions that appear in the B-tree keys
        out = np.zeros(d.shape, dtype=d.dtype)
        if bt == UNDEF:
            return out

        def node(addr):
            a = self.base + addr
            if self.b[a:a + 4] != b"TREE" or self.b[a + 4] != 1:
                raise H5Error("chunk B-tree expected")
            level, n = self.b[a + 5], self._u(a + 6, 2)
            p = a + 8 + 16
            ksz = 8 + 8 * (krank + 1)
            for i in range(n):
                k = p + i * (ksz + 8)
                nbytes, fmask = self._u(k, 4), self._u(k + 4, 4)
                offs = tuple(self._u(k + 8 + 8 * j, 8) for j in range(krank)) + (0,) * arank
                child = self._u(k + ksz, 8)
                if level > 0:
                    node(child)
                    continue
                raw = self.b[self.base + child:self.base + child + nbytes]
                for idx, (fid, cv) in reversed(list(enumerate(d.filters))):
                    if fmask & (1 << idx):
                        continue
                    if fid == 1:
                        raw = zlib.decompress(raw)
                    elif fid == 2:
                        es = cv[0] if cv else d.dtype.itemsize
                        raw = np.frombuffer(raw, dtype=np.uint8).reshape(es, -1).T.tobytes()
                    elif fid == 3:
                        raw = raw[:-4]                              # fletcher32 checksum
                    else:
                        raise H5Error(f"filter {fid} not supported")
                blk = np.frombuffer(raw, dtype=d.dtype, count=int(np.prod(chunk))).reshape(chunk)
                sl = tuple(slice(o, min(o + c, s)) for o, c, s in zip(offs, chunk, d.shape))
                out[sl] = blk[tuple(slice(0, s.stop - s.start) for s in sl)]
        node(bt)
        return out

"""Minimal reader for HOPR mesh files and PICLas state files (no HDF5 library in this image).

Handles what those files use: superblock version 0, old-style groups (B-tree v1 + symbol-table nodes + local heap), object
headers version 1 with continuation blocks, simple dataspaces, fixed-point / IEEE float / fixed-length string datatypes,
contiguous and (uncompressed) chunked layouts.  Used by hostmesh.from_hopr_file and tests/golden/make_reference_vectors.py."""
import struct

import numpy as np

SIG = b"\x89HDF\r\n\x1a\n"
UNDEF = 0xFFFFFFFFFFFFFFFF


class H5File:
    def __init__(self, path):
        self.b = open(path, "rb").read()
        self.base = self.b.find(SIG)
        if self.base < 0:
            raise ValueError("not an HDF5 file")
        s = self.base
        ver = self.b[s + 8]
        if ver != 0:
            raise NotImplementedError("superblock version %d" % ver)
        assert self.b[s + 13] == 8 and self.b[s + 14] == 8, "8-byte offsets and lengths expected"
        self.base_addr = struct.unpack_from("<Q", self.b, s + 24)[0]
        # root symbol table entry at s + 56: name offset, object header, cache type, reserved, scratch (btree, heap)
        _, ohdr, cache, _, btree, heap = struct.unpack_from("<QQIIQQ", self.b, s + 56)
        self.root = self._group(btree, heap) if cache == 1 else self._group_from_header(ohdr)

    def _a(self, rel):
        return self.base_addr + rel

    # ---- groups ----------------------------------------------------------------------------------------------------------------
    def _heap_data(self, heap):
        p = self._a(heap)
        assert self.b[p:p + 4] == b"HEAP"
        _, _, seg = struct.unpack_from("<QQQ", self.b, p + 8)
        return self._a(seg)

    def _group(self, btree, heap):
        names = {}
        hd = self._heap_data(heap)

        def walk(node):
            p = self._a(node)
            assert self.b[p:p + 4] == b"TREE"
            ntype, level, used = struct.unpack_from("<BBH", self.b, p + 4)
            q = p + 24
            for i in range(used):
                child = struct.unpack_from("<Q", self.b, q + 8 + 16 * i)[0]
                if level > 0:
                    walk(child)
                else:
                    sp = self._a(child)
                    assert self.b[sp:sp + 4] == b"SNOD"
                    nsym = struct.unpack_from("<H", self.b, sp + 6)[0]
                    for k in range(nsym):
                        e = sp + 8 + 40 * k
                        noff, oh = struct.unpack_from("<QQ", self.b, e)
                        end = self.b.index(b"\0", hd + noff)
                        names[self.b[hd + noff:end].decode()] = oh
        walk(btree)
        return names

    def _group_from_header(self, ohdr):
        for t, d in self._messages(ohdr):
            if t == 0x11:   # symbol table message
                btree, heap = struct.unpack_from("<QQ", d, 0)
                return self._group(btree, heap)
        raise ValueError("object is not a group")

    # ---- object headers ----------------------------------------------------------------------------------------------------------
    def _messages(self, ohdr):
        p = self._a(ohdr)
        ver, _, nmsg, _, size = struct.unpack_from("<BBHII", self.b, p)
        assert ver == 1, "object header version %d" % ver
        blocks = [(p + 16, size)]
        out = []
        while blocks:
            q, n = blocks.pop(0)
            end = q + n
            while q + 8 <= end and len(out) < nmsg:
                t, sz, fl = struct.unpack_from("<HHB", self.b, q)
                d = self.b[q + 8:q + 8 + sz]
                if t == 0x10:
                    off, ln = struct.unpack_from("<QQ", d, 0)
                    blocks.append((self._a(off), ln))
                out.append((t, d))
                q += 8 + sz
        return out

    # ---- datasets ------------------------------------------------------------------------------------------------------------------
    def keys(self):
        return sorted(self.root)

    def read(self, name):
        msgs = self._messages(self.root[name])
        shape = dtype = layout = None
        for t, d in msgs:
            if t == 0x01:
                ver, rank, flags = d[0], d[1], d[2]
                o = 8 if ver == 1 else 4
                shape = struct.unpack_from("<%dQ" % rank, d, o)
            elif t == 0x03:
                cls = d[0] & 0x0F
                size = struct.unpack_from("<I", d, 4)[0]
                if cls == 1:
                    dtype = np.dtype("<f%d" % size)
                elif cls == 0:
                    signed = (d[1] >> 3) & 1
                    dtype = np.dtype("<%s%d" % ("i" if signed else "u", size))
                elif cls == 3:
                    dtype = np.dtype("S%d" % size)
                else:
                    raise NotImplementedError("datatype class %d" % cls)
            elif t == 0x08:
                layout = d
        if shape is None or dtype is None or layout is None:
            raise ValueError("%s: not a simple dataset" % name)
        n = int(np.prod(shape)) if shape else 1
        ver = layout[0]
        if ver == 3:
            cls = layout[1]
            if cls == 1:
                addr, size = struct.unpack_from("<QQ", layout, 2)
                if addr == UNDEF:
                    return np.zeros(shape, dtype)
                return np.frombuffer(self.b, dtype, n, self._a(addr)).reshape(shape).copy()
            if cls == 2:
                nd = layout[2]
                btree = struct.unpack_from("<Q", layout, 3)[0]
                cdims = struct.unpack_from("<%dI" % nd, layout, 11)
                return self._read_chunked(btree, shape, dtype, cdims[:-1])
            if cls == 0:
                size = struct.unpack_from("<H", layout, 2)[0]
                return np.frombuffer(layout, dtype, n, 4).reshape(shape).copy()
        raise NotImplementedError("layout version %d" % ver)

    def _read_chunked(self, btree, shape, dtype, cdims):
        out = np.zeros(shape, dtype)
        if btree == UNDEF:
            return out
        rank = len(shape)

        def walk(node):
            p = self._a(node)
            assert self.b[p:p + 4] == b"TREE"
            ntype, level, used = struct.unpack_from("<BBH", self.b, p + 4)
            q = p + 24
            ksz = 8 + 8 * (rank + 1)
            for i in range(used):
                k = q + i * (ksz + 8)
                csize, fmask = struct.unpack_from("<II", self.b, k)
                offs = struct.unpack_from("<%dQ" % (rank + 1), self.b, k + 8)[:rank]
                child = struct.unpack_from("<Q", self.b, k + ksz)[0]
                if level > 0:
                    walk(child)
                else:
                    if fmask != 0 and csize != int(np.prod(cdims)) * dtype.itemsize:
                        raise NotImplementedError("filtered chunks")
                    c = np.frombuffer(self.b, dtype, int(np.prod(cdims)), self._a(child)).reshape(cdims)
                    sl = tuple(slice(o, min(o + cd, s)) for o, cd, s in zip(offs, cdims, shape))
                    out[sl] = c[tuple(slice(0, s.stop - s.start) for s in sl)]
        walk(btree)
        return out

"""Minimal pure-Python reader for the HDF5 files QMCTorch writes (h5py is not a dependency here).

QMCTorch stores molecules (``scf/molecule.py:305-350``, ``utils/hdf5_utils.py:27-98``) as small
"classic" HDF5 files: superblock version 0, version-1 object headers, groups as symbol tables
(v1 B-tree + local heap + ``SNOD`` nodes) and contiguous / compact datasets of fixed-point,
floating-point, fixed-length-string or variable-length-string type.  This reader covers exactly
that subset (HDF5 File Format Specification v1.1 §III-IV) and raises ``NotImplementedError`` on
anything else (chunked or filtered data, v2 B-trees, link messages, compound types).

    f = read_hdf5(path)            # nested dict: groups -> dict, datasets -> numpy array / str
    f["molecule"]["basis"]["mos"]  # [nao, nmo] float64
"""
import struct

import numpy as np

_SIG = b"\x89HDF\r\n\x1a\n"
_UNDEF = 0xFFFFFFFFFFFFFFFF


class Group(dict):
    """A group: members by name, HDF5 attributes in ``.attrs``."""

    def __init__(self, *a, **k):
        super().__init__(*a, **k)
        self.attrs = {}


class _File:
    def __init__(self, buf):
        self.b = buf
        if buf[:8] != _SIG:
            raise ValueError("not an HDF5 file")
        if buf[8] != 0:
            raise NotImplementedError("HDF5 superblock version %d (only 0)" % buf[8])
        if buf[13] != 8 or buf[14] != 8:
            raise NotImplementedError("offset/length sizes other than 8 bytes")
        self.base = self.u64(24)
        # root symbol table entry follows the four addresses (base, free space, EOF, driver info)
        self.root_header = self.u64(56 + 8)
        self._gheaps = {}

    def u16(self, o):
        return struct.unpack_from("<H", self.b, o)[0]

    def u32(self, o):
        return struct.unpack_from("<I", self.b, o)[0]

    def u64(self, o):
        return struct.unpack_from("<Q", self.b, o)[0]

    # ---- object headers -------------------------------------------------------------
    def messages(self, addr):
        """(type, flags, offset, size) of every message of the v1 object header at ``addr``."""
        a = self.base + addr
        if self.b[a] != 1:
            raise NotImplementedError("object header version %d (only 1)" % self.b[a])
        nmsg, size = self.u16(a + 2), self.u32(a + 8)
        blocks = [(a + 16, size)]
        out = []
        while blocks and len(out) < nmsg:
            o, n = blocks.pop(0)
            end = o + n
            while o + 8 <= end and len(out) < nmsg:
                mtype, msize, flags = self.u16(o), self.u16(o + 2), self.b[o + 4]
                body = o + 8
                if mtype == 0x10:      # continuation block
                    blocks.append((self.base + self.u64(body), self.u64(body + 8)))
                out.append((mtype, flags, body, msize))
                o = body + msize
        return out

    # ---- groups ---------------------------------------------------------------------
    def _heap_name(self, heap_addr, off):
        h = self.base + heap_addr
        if self.b[h:h + 4] != b"HEAP":
            raise ValueError("bad local heap signature")
        data = self.base + self.u64(h + 24)
        end = self.b.index(b"\0", data + off)
        return self.b[data + off:end].decode("utf-8")

    def _btree_entries(self, addr, heap_addr, out):
        a = self.base + addr
        if self.b[a:a + 4] != b"TREE":
            raise ValueError("bad B-tree signature")
        ntype, level, used = self.b[a + 4], self.b[a + 5], self.u16(a + 6)
        if ntype != 0:
            raise NotImplementedError("raw-data (chunk) B-tree")
        o = a + 24
        for i in range(used):
            child = self.u64(o + 8 + 16 * i)      # key_i (8) child_i (8) ... key_used
            if level > 0:
                self._btree_entries(child, heap_addr, out)
            else:
                s = self.base + child
                if self.b[s:s + 4] != b"SNOD":
                    raise ValueError("bad symbol node signature")
                for k in range(self.u16(s + 6)):
                    e = s + 8 + 40 * k
                    out.append((self._heap_name(heap_addr, self.u64(e)), self.u64(e + 8)))

    def read_object(self, addr):
        msgs = self.messages(addr)
        for mtype, _, o, _ in msgs:
            if mtype == 0x11:      # symbol table message: this object is a group
                entries = []
                self._btree_entries(self.u64(o), self.u64(o + 8), entries)
                grp = Group((name, self.read_object(child)) for name, child in entries)
                for mt, _, ao, _ in msgs:
                    if mt == 0x0C:
                        try:
                            name, val = self._attribute(ao)
                            grp.attrs[name] = val
                        except NotImplementedError:
                            pass
                return grp
        if any(m[0] == 0x02 or m[0] == 0x06 for m in msgs):
            raise NotImplementedError("link-message (new style) groups")
        return self._read_dataset(msgs)

    # ---- datasets -------------------------------------------------------------------
    def _dataspace(self, o):
        ver, rank, flags = self.b[o], self.b[o + 1], self.b[o + 2]
        if ver == 1:
            dims = o + 8
        elif ver == 2:
            if self.b[o + 3] == 2:     # null dataspace
                return None
            dims = o + 4
        else:
            raise NotImplementedError("dataspace version %d" % ver)
        return tuple(self.u64(dims + 8 * i) for i in range(rank))

    def _datatype(self, o):
        cls, size = self.b[o] & 0x0F, self.u32(o + 4)
        bits0 = self.b[o + 1]
        order = ">" if bits0 & 1 else "<"
        if cls == 0:
            return ("num", np.dtype("%s%s%d" % (order, "i" if bits0 & 0x08 else "u", size)))
        if cls == 1:
            return ("num", np.dtype("%sf%d" % (order, size)))
        if cls == 3:
            return ("str", size)
        if cls == 8:      # enumeration (h5py stores bool as an int8 enum): read the base type
            return self._datatype(o + 8)
        if cls == 9:
            if bits0 & 0x0F != 1:
                raise NotImplementedError("variable-length sequences")
            return ("vlen", size)
        raise NotImplementedError("datatype class %d" % cls)

    def _global_heap_object(self, addr, index):
        if addr not in self._gheaps:
            g = self.base + addr
            if self.b[g:g + 4] != b"GCOL":
                raise ValueError("bad global heap signature")
            end = g + self.u64(g + 8)
            objs, o = {}, g + 16
            while o + 16 <= end:
                idx, n = self.u16(o), self.u64(o + 8)
                if idx == 0:
                    break
                objs[idx] = self.b[o + 16:o + 16 + n]
                o += 16 + ((n + 7) // 8) * 8
            self._gheaps[addr] = objs
        return self._gheaps[addr][index]

    def _attribute(self, o):
        """Version-1 attribute message -> (name, value)."""
        if self.b[o] != 1:
            raise NotImplementedError("attribute message version %d" % self.b[o])
        nn, nt, ns = self.u16(o + 2), self.u16(o + 4), self.u16(o + 6)
        p8 = lambda n: (n + 7) // 8 * 8
        name = self.b[o + 8:o + 8 + nn].split(b"\0")[0].decode("utf-8")
        t = o + 8 + p8(nn)
        d = t + p8(nt)
        raw = d + p8(ns)
        shape, kind = self._dataspace(d), self._datatype(t)
        count = int(np.prod(shape)) if shape else 1
        return name, self._decode(kind, shape, count, self.b[raw:raw + count * (kind[1].itemsize if kind[0] == "num" else kind[1])])

    def _decode(self, kind, shape, count, raw):
        if kind[0] == "num":
            arr = np.frombuffer(raw, dtype=kind[1], count=count).astype(kind[1].newbyteorder("="))
            return arr.reshape(shape) if shape else arr[0]
        if kind[0] == "str":
            vals = [raw[i * kind[1]:(i + 1) * kind[1]].split(b"\0")[0].decode("utf-8") for i in range(count)]
        else:  # vlen string: length (4) + global heap collection address (8) + object index (4)
            vals = []
            for i in range(count):
                n, gaddr, gidx = struct.unpack_from("<IQI", raw, 16 * i)
                vals.append(self._global_heap_object(gaddr, gidx)[:n].decode("utf-8") if n else "")
        return np.array(vals).reshape(shape) if shape else vals[0]

    def _read_dataset(self, msgs):
        shape = kind = raw = None
        for mtype, _, o, msize in msgs:
            if mtype == 0x01:
                shape = self._dataspace(o)
            elif mtype == 0x03:
                kind = self._datatype(o)
            elif mtype == 0x0B:
                raise NotImplementedError("filtered datasets")
            elif mtype == 0x08:
                ver, lclass = self.b[o], self.b[o + 1]
                if ver != 3:
                    raise NotImplementedError("data layout version %d" % ver)
                if lclass == 0:
                    n = self.u16(o + 2)
                    raw = self.b[o + 4:o + 4 + n]
                elif lclass == 1:
                    addr, n = self.u64(o + 2), self.u64(o + 10)
                    raw = b"" if addr == _UNDEF else self.b[self.base + addr:self.base + addr + n]
                else:
                    raise NotImplementedError("chunked datasets")
        if shape is None or kind is None or raw is None:
            raise ValueError("dataset without dataspace / datatype / layout")
        count = int(np.prod(shape)) if shape else 1
        return self._decode(kind, shape, count, raw)


def read_hdf5(path):
    """Whole file as a nested dict (groups) of numpy arrays / numpy scalars / str (datasets)."""
    with open(path, "rb") as fh:
        f = _File(fh.read())
    return f.read_object(f.root_header)

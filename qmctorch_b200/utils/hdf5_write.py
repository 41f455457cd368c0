"""Minimal pure-Python HDF5 WRITER for the files QMCTorch dumps (h5py is not a dependency here).

Counterpart of ``utils/hdf5_min.py``.  ``dump_to_hdf5(obj, fname, root_name)`` follows
``qmctorch/utils/hdf5_utils.py:128-160`` (+ ``insert_object`` ``:163-372``): the object becomes a
group named ``root_name`` (renamed ``<root>_<n>`` when that name exists), attributes / dict items
become sub-groups or datasets, lists and tuples become arrays, tensors and parameters numpy arrays,
names starting with ``_`` and ``None`` values are skipped, a module's ``state_dict`` entries are
stored next to its attributes.  ``add_group_attr`` (``:...``) attaches string attributes to a group.

The file is the "classic" layout h5py's default driver writes for small files and the one the
reader understands (HDF5 File Format Specification v1.1): superblock version 0, version-1 object
headers, groups as symbol tables (one v1 B-tree leaf + local heap + ``SNOD`` nodes, names sorted),
contiguous datasets of little-endian fixed-point / IEEE floating-point / fixed-length-string type,
version-1 attribute messages.  Appending re-reads the existing file with ``read_hdf5`` and rewrites
it (result files are small: observables and parameter sets).
"""
import os
import struct
from types import SimpleNamespace

import numpy as np

from .hdf5_min import Group, read_hdf5

_SIG = b"\x89HDF\r\n\x1a\n"
_UNDEF = 0xFFFFFFFFFFFFFFFF
_LEAF_K, _INT_K = 4, 16           # symbols per SNOD = 2 * _LEAF_K, children per B-tree node = 2 * _INT_K


_Attrs = Group      # dict of group members that also carries HDF5 attributes (``.attrs``)


def _pad8(b):
    return b + b"\0" * (-len(b) % 8)


class _Writer:
    def __init__(self):
        self.buf = bytearray(96)          # superblock, filled in at the end

    def alloc(self, data):
        """Appends 8-byte aligned data, returns its address."""
        self.buf += b"\0" * (-len(self.buf) % 8)
        addr = len(self.buf)
        self.buf += data
        return addr

    # ---- messages ---------------------------------------------------------------------
    @staticmethod
    def _msg(mtype, body, flags=0):
        body = _pad8(body)
        return struct.pack("<HHB3x", mtype, len(body), flags) + body

    def _header(self, msgs):
        body = b"".join(msgs)
        return self.alloc(struct.pack("<BxHII4x", 1, len(msgs), 1, len(body)) + body)

    @staticmethod
    def _dataspace(shape):
        return struct.pack("<BBB5x", 1, len(shape), 0) + b"".join(struct.pack("<Q", int(n)) for n in shape)

    @staticmethod
    def _datatype(arr):
        """(datatype message body, raw little-endian bytes) of a numpy array."""
        k = arr.dtype.kind
        if k in "SU":
            a = np.char.encode(arr, "utf-8") if k == "U" else arr
            size = max(int(a.dtype.itemsize), 1)
            a = a.astype("S%d" % size)
            # class 3 (string), version 1; null-padded, UTF-8
            return struct.pack("<BBBBI", 0x13, 0x11, 0, 0, size), a.tobytes()
        if k == "b":
            arr, k = arr.astype(np.int8), "i"
        if k in "iu":
            a = arr.astype(arr.dtype.newbyteorder("<"))
            size = a.dtype.itemsize
            bits0 = 0x08 if k == "i" else 0x00
            return struct.pack("<BBBBIHH", 0x10, bits0, 0, 0, size, 0, 8 * size), a.tobytes()
        if k == "f":
            a = arr.astype(arr.dtype.newbyteorder("<"))
            size = a.dtype.itemsize
            if size == 8:
                prop = struct.pack("<HHBBBBI", 0, 64, 52, 11, 0, 52, 1023)
                sign = 63
            elif size == 4:
                prop = struct.pack("<HHBBBBI", 0, 32, 23, 8, 0, 23, 127)
                sign = 31
            else:
                raise NotImplementedError("float%d" % (8 * size))
            # class 1 (floating point), version 1; little-endian, implied-msb mantissa, sign bit location
            return struct.pack("<BBBBI", 0x11, 0x20, sign, 0, size) + prop, a.tobytes()
        raise NotImplementedError("dtype %s" % arr.dtype)

    def _attr_msgs(self, attrs):
        out = []
        for name, val in attrs.items():
            arr = np.asarray(val)
            dt, raw = self._datatype(arr)
            ds = self._dataspace(arr.shape)
            nm = name.encode("utf-8") + b"\0"
            body = struct.pack("<BxHHH", 1, len(nm), len(dt), len(ds)) + _pad8(nm) + _pad8(dt) + _pad8(ds) + raw
            out.append(self._msg(0x000C, body))
        return out

    # ---- objects ----------------------------------------------------------------------
    def dataset(self, value):
        arr = np.asarray(value)
        if arr.dtype == object:
            raise TypeError("object arrays cannot be stored")
        dt, raw = self._datatype(arr)
        addr = self.alloc(raw) if raw else _UNDEF
        layout = struct.pack("<BBQQ", 3, 1, addr, len(raw))
        return self._header([self._msg(0x0001, self._dataspace(arr.shape)), self._msg(0x0003, dt, flags=1),
                             self._msg(0x0008, layout)])

    def group(self, members, attrs=None):
        """members: name -> dict (sub-group) | array-like (dataset).  Returns (header, btree, heap) addresses."""
        names = sorted(members)                    # symbol nodes hold their entries in name order
        if len(names) > 2 * _LEAF_K * 2 * _INT_K:
            raise NotImplementedError("more than %d members in one group" % (2 * _LEAF_K * 2 * _INT_K))
        children = {}
        for n in names:
            v = members[n]
            if isinstance(v, dict):
                children[n] = ("g",) + self.group(v, getattr(v, "attrs", None))
            else:
                children[n] = ("d", self.dataset(v))
        # local heap: offset 0 is the empty name; names null-terminated, 8-byte aligned
        heap, offs = bytearray(8), {}
        for n in names:
            offs[n] = len(heap)
            heap += _pad8(n.encode("utf-8") + b"\0")
        heap_data = self.alloc(bytes(heap) if len(heap) > 8 else bytes(heap) + bytes(8))
        heap_addr = self.alloc(b"HEAP" + struct.pack("<B3xQQQ", 0, max(len(heap), 16), 1, heap_data))
        # symbol nodes of up to 2K entries, then ONE B-tree leaf over them
        snods, keys = [], [0]
        per = 2 * _LEAF_K
        for i in range(0, len(names), per):
            chunk = names[i:i + per]
            body = b"SNOD" + struct.pack("<BxH", 1, len(chunk))
            for n in chunk:
                c = children[n]
                if c[0] == "g":
                    body += struct.pack("<QQII", offs[n], c[1], 1, 0) + struct.pack("<QQ", c[2], c[3])
                else:
                    body += struct.pack("<QQII", offs[n], c[1], 0, 0) + bytes(16)
            body += bytes(40 * (per - len(chunk)))
            snods.append(self.alloc(body))
            keys.append(offs[chunk[-1]])
        node = b"TREE" + struct.pack("<BBHQQ", 0, 0, len(snods), _UNDEF, _UNDEF)
        for i, s in enumerate(snods):
            node += struct.pack("<QQ", keys[i], s)
        node += struct.pack("<Q", keys[len(snods)])
        node += bytes(24 + (2 * _INT_K + 1) * 8 + 2 * _INT_K * 8 - len(node))
        btree = self.alloc(node)
        msgs = [self._msg(0x0011, struct.pack("<QQ", btree, heap_addr))] + self._attr_msgs(attrs or {})
        return self._header(msgs), btree, heap_addr

    def finish(self, root):
        hdr, btree, heap = root
        self.buf += b"\0" * (-len(self.buf) % 8)
        sb = _SIG + struct.pack("<BBBxBBBxHHI", 0, 0, 0, 0, 8, 8, _LEAF_K, _INT_K, 0)
        sb += struct.pack("<QQQQ", 0, _UNDEF, len(self.buf), _UNDEF)
        sb += struct.pack("<QQII", 0, hdr, 1, 0) + struct.pack("<QQ", btree, heap)
        assert len(sb) == 96
        self.buf[:96] = sb
        return bytes(self.buf)


def write_hdf5(path, tree):
    """Writes a nested dict (groups; ``.attrs`` of an ``_Attrs`` dict become attributes) of array-likes
    / scalars / strings (datasets) as an HDF5 file."""
    w = _Writer()
    data = w.finish(w.group(tree, getattr(tree, "attrs", None)))
    tmp = path + ".tmp%d" % os.getpid()
    with open(tmp, "wb") as f:
        f.write(data)
    os.replace(tmp, path)


# ---- object -> tree (hdf5_utils.py:163-460) -------------------------------------------
def _is_leaf(obj):
    import torch
    if isinstance(obj, (torch.Tensor, np.ndarray, str, bytes, int, float, bool, np.generic, list, tuple)):
        return True
    return not (hasattr(obj, "__dict__") or hasattr(obj, "keys"))


def _leaf(obj):
    """Dataset value of a leaf, or None when it is not stored (None, devices, callables, ragged lists)."""
    import torch
    if obj is None or isinstance(obj, torch.device) or callable(obj):
        return None
    if isinstance(obj, torch.Tensor):
        return obj.detach().cpu().numpy()
    if isinstance(obj, (list, tuple)):
        items = [o.detach().cpu().numpy() if isinstance(o, torch.Tensor) else o for o in obj]
        if any(o is None for o in items):
            return None
        try:
            arr = np.array(items)
        except ValueError:
            return None
        return None if arr.dtype == object else arr
    try:
        arr = np.asarray(obj)
    except Exception:
        return None
    return None if arr.dtype == object else arr


def _to_tree(obj, depth=0):
    if depth > 12:
        return None
    if _is_leaf(obj):
        return _leaf(obj)
    out = _Attrs()
    if hasattr(obj, "__dict__"):
        items = list(vars(obj).items())
    else:
        items = list(obj.items())
    for name in getattr(obj, "__extra_attr__", []):
        items.append((name, getattr(obj, name)))
    if hasattr(obj, "state_dict") and callable(obj.state_dict):
        items += list(obj.state_dict().items())
    for name, child in items:
        name = str(name)
        if name.startswith("_") or name in out:
            continue
        if isinstance(child, (list, tuple)) and _leaf(child) is None and len(child):
            # ragged / heterogeneous list: one entry per element, like insert_list's fallback
            for i, el in enumerate(child):
                sub = _to_tree(el, depth + 1)
                if sub is not None:
                    out["%s_%d" % (name, i)] = sub
            continue
        sub = _to_tree(child, depth + 1)
        if sub is not None and not (isinstance(sub, dict) and not sub):
            out[name] = sub
    return out


def _load_tree(fname):
    if not os.path.isfile(fname):
        return _Attrs()

    return read_hdf5(fname)


def dump_to_hdf5(obj, fname, root_name=None):
    """hdf5_utils.py:128-160.  Returns the name of the group the object went to."""
    tree = _load_tree(fname)
    if root_name is None:
        root_name = obj.__class__.__name__
    if root_name in tree:
        n = sum(1 for k in tree if k.startswith(root_name)) + 1
        root_name = "%s_%d" % (root_name, n)
    sub = _to_tree(obj)
    tree[root_name] = sub if isinstance(sub, dict) else _Attrs(value=sub)
    write_hdf5(fname, tree)
    return root_name


def add_group_attr(fname, grp_name, attr):
    """hdf5_utils.py: attaches ``attr`` (dict of strings / numbers) to group ``grp_name``."""
    tree = _load_tree(fname)
    if grp_name not in tree or not isinstance(tree[grp_name], dict):
        raise KeyError("no group %s in %s" % (grp_name, fname))
    tree[grp_name].attrs.update(attr)
    write_hdf5(fname, tree)

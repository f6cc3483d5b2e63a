"""Minimal pure-Python HDF5 writer / reader for the chain sink (h5py is not installed in this image).

EMPEROR's parent process reloads a reddemcee run from HDF5 files (emp.py:722-762 writes them, emp.py:781-789
reads them back through `reddemcee.hdf.PTHDFBackend / HDFBackend_plus`, which extend emcee's `HDFBackend`): one
file `<name>.h5` for the ladder-level histories and one `<name>_<t>.h5` per temperature, each with ONE group
('mcmc') that carries a few scalar attributes and a handful of numeric datasets.  That subset of the format is
what this module writes, following the HDF5 File Format Specification version 1.x structures that every
libhdf5 release reads:

  superblock version 0  ->  root group (object header v1 + symbol-table message: B-tree v1 'TREE' node, 'SNOD'
  symbol-table node, local 'HEAP')  ->  one sub-group of the same kind  ->  datasets = object header v1 with
  dataspace (v1), datatype (v1: IEEE f64 / two's-complement integers, little endian), fill-value (v2), contiguous
  data-layout (v3) messages and attribute (v1) messages; at most 8 links per group (one SNOD leaf).

`read_h5` parses the same structures (it does not share code paths with the writer beyond the struct layouts) and
is what `postproc.load_backend` uses where h5py is absent.  Where h5py IS installed, tests/test_h5min.py
cross-checks both directions against it; in this image the check is structural (round trip + field-by-field
layout assertions against the specification's offsets).
"""
from __future__ import annotations

import struct
from typing import Dict

import numpy as np

UNDEF = 0xFFFFFFFFFFFFFFFF
SIGNATURE = b"\x89HDF\r\n\x1a\n"
LEAF_K, INTERNAL_K = 4, 16   # group B-tree ranks stored in the superblock: <= 2*LEAF_K links per SNOD


def _pad8(b: bytes) -> bytes:
    return b + b"\0" * (-len(b) % 8)


# ---- datatype / dataspace messages -----------------------------------------------------------------------
def _datatype_msg(dt: np.dtype) -> bytes:
    dt = np.dtype(dt)
    if dt.kind == "f" and dt.itemsize == 8:
        # class 1 (floating point), version 1; bit field: little endian, mantissa normalisation 2 (implied msb),
        # sign bit 63; properties: bit offset 0, precision 64, exponent at 52 (11 bits), mantissa at 0 (52 bits),
        # bias 1023
        return struct.pack("<BBBBI", 0x11, 0x20, 63, 0, 8) + struct.pack("<HHBBBBI", 0, 64, 52, 11, 0, 52, 1023)
    if dt.kind in "iu" and dt.itemsize in (1, 2, 4, 8):
        signed = 0x08 if dt.kind == "i" else 0x00
        return struct.pack("<BBBBI", 0x10, signed, 0, 0, dt.itemsize) + struct.pack("<HH", 0, 8 * dt.itemsize)
    if dt.kind == "S":
        # class 3 (string), null-terminated, ASCII
        return struct.pack("<BBBBI", 0x13, 0x00, 0, 0, dt.itemsize)
    raise TypeError(f"h5min cannot store dtype {dt}")


def _dataspace_msg(shape) -> bytes:
    shape = tuple(int(x) for x in shape)
    head = struct.pack("<BBBB4x", 1, len(shape), 0, 0)   # version 1, rank, flags (no max dims), reserved
    return head + b"".join(struct.pack("<Q", n) for n in shape)


def _message(mtype: int, body: bytes, flags: int = 0) -> bytes:
    body = _pad8(body)
    return struct.pack("<HHB3x", mtype, len(body), flags) + body


def _attribute_msg(name: str, value) -> bytes:
    if isinstance(value, str):
        raw = value.encode("ascii") + b"\0"
        arr = np.frombuffer(raw, dtype=f"S{len(raw)}")
        shape = ()
    else:
        arr = np.asarray(value)
        if arr.dtype == np.bool_:
            arr = arr.astype(np.int8)
        elif arr.dtype.kind == "i" and arr.dtype.itemsize != 8:
            arr = arr.astype(np.int64)
        elif arr.dtype.kind == "f":
            arr = arr.astype(np.float64)
        shape = arr.shape
    nm = name.encode("ascii") + b"\0"
    dtm, dsm = _datatype_msg(arr.dtype), _dataspace_msg(shape)
    body = struct.pack("<BxHHH", 1, len(nm), len(dtm), len(dsm)) + _pad8(nm) + _pad8(dtm) + _pad8(dsm)
    body += np.ascontiguousarray(arr).tobytes()
    return _message(0x000C, body)


def _object_header(messages) -> bytes:
    messages = list(messages)
    if sum(len(m) for m in messages) < 40:  # keep small headers above libhdf5's minimum chunk size with a NIL message
        messages.append(_message(0x0000, b"\0" * 16))
    data = b"".join(messages)
    # version 1, reserved, number of messages, reference count 1, header data size; 4 bytes pad to 8-alignment
    return struct.pack("<BxHII4x", 1, len(messages), 1, len(data)) + data


# ---- writer ---------------------------------------------------------------------------------------------
class _Writer:
    def __init__(self):
        self.buf = bytearray(96)   # superblock (56 bytes) + root symbol-table entry (40 bytes), filled last

    def alloc(self, data: bytes) -> int:
        self.buf += b"\0" * (-len(self.buf) % 8)
        addr = len(self.buf)
        self.buf += data
        return addr

    def dataset(self, arr: np.ndarray, attrs: Dict) -> int:
        arr = np.ascontiguousarray(arr)
        if arr.dtype == np.bool_:
            arr = arr.astype(np.uint8)
        if arr.dtype.byteorder == ">":
            arr = arr.astype(arr.dtype.newbyteorder("<"))
        raw = arr.tobytes()
        data_addr = self.alloc(raw) if len(raw) else UNDEF
        msgs = [_message(0x0001, _dataspace_msg(arr.shape)),
                _message(0x0003, _datatype_msg(arr.dtype), flags=1),          # constant message
                _message(0x0005, struct.pack("<BBBB", 2, 1, 0, 0)),           # fill value v2: early alloc, undefined
                _message(0x0008, struct.pack("<BBQQ", 3, 1, data_addr, len(raw)))]  # layout v3, contiguous
        msgs += [_attribute_msg(k, v) for k, v in attrs.items()]
        return self.alloc(_object_header(msgs))

    def group(self, links: Dict[str, tuple], attrs: Dict):
        """links: name -> (object header address, cache type, scratch bytes).  Returns (header address, B-tree
        address, heap address)."""
        if len(links) > 2 * LEAF_K:
            raise ValueError(f"h5min groups hold at most {2 * LEAF_K} links")
        names = sorted(links)  # symbol-table entries are ordered by name
        heap_data = bytearray(b"\0" * 8)  # offset 0: the empty string
        offs = {}
        for n in names:
            offs[n] = len(heap_data)
            heap_data += _pad8(n.encode("ascii") + b"\0")
        free_off = len(heap_data)
        heap_data += struct.pack("<QQ", 1, 16)  # one free block at the end: next = 1 (none), size 16
        heap_data_addr = self.alloc(bytes(heap_data))
        heap_addr = self.alloc(b"HEAP" + struct.pack("<B3xQQQ", 0, len(heap_data), free_off, heap_data_addr))
        entries = b"".join(struct.pack("<QQI4x", offs[n], links[n][0], links[n][1]) + links[n][2].ljust(16, b"\0")
                           for n in names)
        snod = b"SNOD" + struct.pack("<BxH", 1, len(names)) + entries.ljust(2 * LEAF_K * 40, b"\0")
        snod_addr = self.alloc(snod)
        keys_children = struct.pack("<QQQ", 0, snod_addr, offs[names[-1]] if names else 0)
        tree = b"TREE" + struct.pack("<BBHQQ", 0, 0, 1, UNDEF, UNDEF) + keys_children
        tree = tree.ljust(24 + (2 * INTERNAL_K + 1) * 8 + 2 * INTERNAL_K * 8, b"\0")
        tree_addr = self.alloc(tree)
        msgs = [_message(0x0011, struct.pack("<QQ", tree_addr, heap_addr))]
        msgs += [_attribute_msg(k, v) for k, v in attrs.items()]
        return self.alloc(_object_header(msgs)), tree_addr, heap_addr

    def finish(self, root_header: int, root_tree: int, root_heap: int) -> bytes:
        eof = len(self.buf)
        sb = SIGNATURE + struct.pack("<BBBBBBBBHHI", 0, 0, 0, 0, 0, 8, 8, 0, LEAF_K, INTERNAL_K, 0)
        sb += struct.pack("<QQQQ", 0, UNDEF, eof, UNDEF)
        sb += struct.pack("<QQI4xQQ", 0, root_header, 1, root_tree, root_heap)  # root symbol-table entry
        assert len(sb) == 96
        self.buf[:96] = sb
        return bytes(self.buf)


def write_h5(path: str, groups: Dict[str, Dict]):
    """groups: {group name: {"attrs": {name: scalar | str | small array}, "datasets": {name: ndarray}}}.
    Every group hangs off the root group; datasets are contiguous, little endian."""
    w = _Writer()
    root_links = {}
    for gname, g in groups.items():
        links = {}
        for dname, arr in g.get("datasets", {}).items():
            links[dname] = (w.dataset(np.asarray(arr), {}), 0, b"")
        hdr, tree, heap = w.group(links, g.get("attrs", {}))
        root_links[gname] = (hdr, 1, struct.pack("<QQ", tree, heap))
    hdr, tree, heap = w.group(root_links, {})
    with open(path, "wb") as fh:
        fh.write(w.finish(hdr, tree, heap))
    return path


# ---- reader ---------------------------------------------------------------------------------------------
class _Reader:
    def __init__(self, data: bytes):
        self.d = data
        if data[:8] != SIGNATURE:
            raise ValueError("not an HDF5 file")
        ver, = struct.unpack_from("<B", data, 8)
        if ver != 0:
            raise ValueError(f"h5min reads superblock version 0 files only (found {ver}); use h5py")
        so, sl = struct.unpack_from("<BB", data, 13)
        if (so, sl) != (8, 8):
            raise ValueError("h5min expects 8-byte offsets and lengths")
        self.root_header, = struct.unpack_from("<Q", data, 56 + 8)

    def messages(self, addr):
        ver, nmsg, _ref, size = struct.unpack_from("<BxHII", self.d, addr)
        if ver != 1:
            raise ValueError("h5min reads version-1 object headers only")
        pos, end, out = addr + 16, addr + 16 + size, []
        while pos < end and len(out) < nmsg:
            mtype, msize, _flags = struct.unpack_from("<HHB", self.d, pos)
            out.append((mtype, self.d[pos + 8:pos + 8 + msize]))
            pos += 8 + msize
        return out

    @staticmethod
    def dtype_of(msg):
        cv, b0, _b1, _b2, size = struct.unpack_from("<BBBBI", msg, 0)
        cls = cv & 0x0F
        if cls == 1 and size == 8:
            return np.dtype("<f8")
        if cls == 1 and size == 4:
            return np.dtype("<f4")
        if cls == 0:
            return np.dtype(("<i" if b0 & 0x08 else "<u") + str(size))
        if cls == 3:
            return np.dtype(f"S{size}")
        raise TypeError(f"h5min cannot read datatype class {cls}")

    @staticmethod
    def shape_of(msg):
        ver, rank, flags = struct.unpack_from("<BBB", msg, 0)
        if ver != 1:
            raise ValueError("dataspace version")
        return tuple(struct.unpack_from("<Q", msg, 8 + 8 * i)[0] for i in range(rank))

    def attribute(self, msg):
        ver, nsz, tsz, ssz = struct.unpack_from("<BxHHH", msg, 0)
        p = 8
        name = msg[p:p + nsz].split(b"\0")[0].decode()
        p += (nsz + 7) // 8 * 8
        dt = self.dtype_of(msg[p:p + tsz])
        p += (tsz + 7) // 8 * 8
        shape = self.shape_of(msg[p:p + ssz])
        p += (ssz + 7) // 8 * 8
        n = int(np.prod(shape)) if shape else 1
        val = np.frombuffer(msg[p:p + n * dt.itemsize], dtype=dt).reshape(shape)
        if dt.kind == "S":
            return name, val.reshape(-1)[0].split(b"\0")[0].decode()
        return name, (val.reshape(-1)[0].item() if shape == () else val.copy())

    def links(self, tree_addr, heap_addr):
        d = self.d
        assert d[heap_addr:heap_addr + 4] == b"HEAP"
        _size, _free, heap_data = struct.unpack_from("<QQQ", d, heap_addr + 8)
        out = {}

        def walk(addr):
            assert d[addr:addr + 4] == b"TREE"
            ntype, level, used = struct.unpack_from("<BBH", d, addr + 4)
            for i in range(used):
                child, = struct.unpack_from("<Q", d, addr + 24 + 8 + 16 * i)
                if level > 0:
                    walk(child)
                    continue
                assert d[child:child + 4] == b"SNOD"
                nsym, = struct.unpack_from("<H", d, child + 6)
                for k in range(nsym):
                    off, hdr, cache = struct.unpack_from("<QQI", d, child + 8 + 40 * k)
                    end = d.index(b"\0", heap_data + off)
                    out[d[heap_data + off:end].decode()] = hdr
        walk(tree_addr)
        return out

    def node(self, addr):
        """-> ("group", attrs, {name: addr}) or ("dataset", attrs, ndarray)."""
        msgs = self.messages(addr)
        attrs = dict(self.attribute(m) for t, m in msgs if t == 0x000C)
        sym = [m for t, m in msgs if t == 0x0011]
        if sym:
            tree, heap = struct.unpack_from("<QQ", sym[0], 0)
            return "group", attrs, self.links(tree, heap)
        dt = self.dtype_of(next(m for t, m in msgs if t == 0x0003))
        shape = self.shape_of(next(m for t, m in msgs if t == 0x0001))
        lay = next(m for t, m in msgs if t == 0x0008)
        ver, cls = struct.unpack_from("<BB", lay, 0)
        if ver != 3 or cls != 1:
            raise ValueError("h5min reads contiguous (layout v3) datasets only; use h5py")
        daddr, dsize = struct.unpack_from("<QQ", lay, 2)
        n = int(np.prod(shape)) if shape else 1
        if daddr == UNDEF or n == 0:
            return "dataset", attrs, np.zeros(shape, dtype=dt)
        return "dataset", attrs, np.frombuffer(self.d[daddr:daddr + n * dt.itemsize], dtype=dt).reshape(shape).copy()


def read_h5(path: str) -> Dict[str, Dict]:
    """Inverse of `write_h5`: {group: {"attrs": {...}, "datasets": {...}}} (datasets in the root group are returned
    under the group name '/')."""
    with open(path, "rb") as fh:
        r = _Reader(fh.read())
    kind, attrs, links = r.node(r.root_header)
    out = {}
    for name, addr in links.items():
        k, a, body = r.node(addr)
        if k == "group":
            ds = {}
            for dn, daddr in body.items():
                dk, _da, arr = r.node(daddr)
                if dk == "dataset":
                    ds[dn] = arr
            out[name] = {"attrs": a, "datasets": ds}
        else:
            out.setdefault("/", {"attrs": attrs, "datasets": {}})["datasets"][name] = body
    return out

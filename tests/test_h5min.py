"""The minimal HDF5 writer / reader of the chain sink (astroemperor_b200/h5min.py).

h5py is not installed in this image, so the checks are (i) a round trip through the independent reader,
(ii) field-by-field assertions of the bytes against the HDF5 File Format Specification (version-0 superblock,
version-1 object headers, symbol-table groups) and (iii), wherever h5py IS importable, a cross-check in both
directions against libhdf5."""
import struct

import numpy as np
import pytest

from astroemperor_b200.h5min import SIGNATURE, UNDEF, read_h5, write_h5


def _example():
    rng = np.random.default_rng(0)
    return {"mcmc": {"attrs": {"iteration": 40, "ntemps": 3, "version": "b200-0.1.0", "x": 2.5, "has_blobs": False},
                     "datasets": {"chain": rng.normal(size=(40, 24, 7)), "log_like": rng.normal(size=(40, 24)),
                                  "accepted": np.arange(24), "beta_history": rng.uniform(size=40),
                                  "mask": np.arange(6, dtype=np.int32).reshape(2, 3), "empty": np.zeros((0, 2))}}}


def test_round_trip(tmp_path):
    g = _example()
    p = write_h5(str(tmp_path / "a.h5"), g)
    r = read_h5(p)
    assert r["mcmc"]["attrs"] == {"iteration": 40, "ntemps": 3, "version": "b200-0.1.0", "x": 2.5, "has_blobs": 0}
    for k, v in g["mcmc"]["datasets"].items():
        got = r["mcmc"]["datasets"][k]
        assert got.shape == v.shape and got.dtype == v.dtype and np.array_equal(got, v), k


def test_bytes_follow_the_specification(tmp_path):
    p = write_h5(str(tmp_path / "a.h5"), _example())
    d = open(p, "rb").read()
    # superblock version 0 (spec III.A): signature, versions, 8-byte offsets / lengths, group K values, addresses
    assert d[:8] == SIGNATURE and d[8] == 0 and d[9] == 0 and d[10] == 0 and d[12] == 0
    assert d[13] == 8 and d[14] == 8
    leaf_k, internal_k, flags = struct.unpack_from("<HHI", d, 16)
    assert (leaf_k, internal_k, flags) == (4, 16, 0)
    base, free, eof, driver = struct.unpack_from("<QQQQ", d, 24)
    assert base == 0 and free == UNDEF and eof == len(d) and driver == UNDEF
    # root symbol-table entry: cached B-tree / heap addresses (cache type 1)
    name_off, root_hdr, cache, tree, heap = struct.unpack_from("<QQI4xQQ", d, 56)
    assert name_off == 0 and cache == 1
    assert d[tree:tree + 4] == b"TREE" and d[heap:heap + 4] == b"HEAP"
    # root object header (version 1): its symbol-table message (type 0x11) repeats the two addresses
    ver, nmsg, ref, size = struct.unpack_from("<BxHII", d, root_hdr)
    assert ver == 1 and ref == 1 and root_hdr % 8 == 0
    mtype, msize = struct.unpack_from("<HH", d, root_hdr + 16)
    assert mtype == 0x11 and struct.unpack_from("<QQ", d, root_hdr + 24) == (tree, heap)
    # B-tree node: group node (type 0), leaf level, one child = a SNOD whose single entry is the 'mcmc' group
    ntype, level, used, left, right = struct.unpack_from("<BBHQQ", d, tree + 4)
    assert (ntype, level, used, left, right) == (0, 0, 1, UNDEF, UNDEF)
    key0, child, key1 = struct.unpack_from("<QQQ", d, tree + 24)
    assert d[child:child + 4] == b"SNOD" and d[child + 4] == 1 and struct.unpack_from("<H", d, child + 6)[0] == 1
    hsize, hfree, hdata = struct.unpack_from("<QQQ", d, heap + 8)
    off, ghdr, gcache = struct.unpack_from("<QQI", d, child + 8)
    assert d[hdata:hdata + 8] == b"\0" * 8 and d[hdata + off:hdata + off + 5] == b"mcmc\0" and gcache == 1
    assert key0 == 0 and key1 == off
    # inside the group: links sorted by name, every dataset header carries dataspace / datatype / layout messages
    gtree, gheap = struct.unpack_from("<QQ", d, ghdr + 24)
    gchild, = struct.unpack_from("<Q", d, gtree + 32)
    n, = struct.unpack_from("<H", d, gchild + 6)
    _, _, ghdata = struct.unpack_from("<QQQ", d, gheap + 8)
    names = []
    for k in range(n):
        o, hdr, c = struct.unpack_from("<QQI", d, gchild + 8 + 40 * k)
        names.append(d[ghdata + o:d.index(b"\0", ghdata + o)].decode())
        assert c == 0 and hdr % 8 == 0
        nm = struct.unpack_from("<H", d, hdr + 2)[0]
        pos, types = hdr + 16, []
        for _ in range(nm):
            t, sz = struct.unpack_from("<HH", d, pos)
            types.append(t)
            if t == 0x0003 and names[-1] == "chain":   # IEEE f64 little endian
                assert d[pos + 8] == 0x11 and d[pos + 9] == 0x20 and d[pos + 10] == 63
                assert struct.unpack_from("<I", d, pos + 12)[0] == 8
                assert struct.unpack_from("<HHBBBBI", d, pos + 16) == (0, 64, 52, 11, 0, 52, 1023)
            if t == 0x0001 and names[-1] == "chain":   # dataspace v1, rank 3
                assert d[pos + 8] == 1 and d[pos + 9] == 3
                assert struct.unpack_from("<QQQ", d, pos + 16) == (40, 24, 7)
            if t == 0x0008 and names[-1] == "chain":   # layout v3, contiguous, address + size
                assert d[pos + 8] == 3 and d[pos + 9] == 1
                addr, nbytes = struct.unpack_from("<QQ", d, pos + 10)
                assert nbytes == 40 * 24 * 7 * 8 and addr % 8 == 0
            assert sz % 8 == 0
            pos += 8 + sz
        assert {0x0001, 0x0003, 0x0008} <= set(types)
    assert names == sorted(names) and "chain" in names


def test_group_capacity_and_dtypes(tmp_path):
    with pytest.raises(ValueError):
        write_h5(str(tmp_path / "b.h5"), {"g": {"datasets": {f"d{i}": np.zeros(2) for i in range(9)}}})
    with pytest.raises(TypeError):
        write_h5(str(tmp_path / "c.h5"), {"g": {"datasets": {"c": np.zeros(2, dtype=np.complex128)}}})
    with pytest.raises(ValueError):
        open(str(tmp_path / "x.h5"), "wb").write(b"not hdf5 at all")
        read_h5(str(tmp_path / "x.h5"))


def test_against_h5py_when_available(tmp_path):
    h5py = pytest.importorskip("h5py")
    g = _example()
    p = write_h5(str(tmp_path / "a.h5"), g)
    with h5py.File(p, "r") as f:   # libhdf5 reads what h5min wrote
        assert int(f["mcmc"].attrs["iteration"]) == 40
        for k, v in g["mcmc"]["datasets"].items():
            assert np.array_equal(f["mcmc"][k][...], v)
    q = str(tmp_path / "b.h5")
    with h5py.File(q, "w", libver="earliest") as f:   # and h5min reads libhdf5's earliest-format files
        gg = f.create_group("mcmc")
        gg.attrs["iteration"] = 7
        gg.create_dataset("chain", data=g["mcmc"]["datasets"]["chain"])
    r = read_h5(q)
    assert r["mcmc"]["attrs"]["iteration"] == 7 and np.array_equal(r["mcmc"]["datasets"]["chain"], g["mcmc"]["datasets"]["chain"])

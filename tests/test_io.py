"""CPU tests of the on-disk format (vpm.save / vpm.read!): byte layout of the hand-written HDF5 container against the file
format specification's field tables, round trip through the reader, and the XDMF wrapper."""
import os
import struct
import types
import xml.etree.ElementTree as ET

import numpy as np
import pytest

from flowunsteady_b200 import h5min, vpm


def _fake_field(n=37, maxp=64):
    rng = np.random.default_rng(3)
    P = np.zeros((maxp, 43))
    P[:n, 0:3] = rng.random((n, 3))
    P[:n, 3:6] = rng.standard_normal((n, 3))
    P[:n, 6] = 0.1 + rng.random(n)
    P[:n, 7] = rng.random(n)
    P[:n, 8] = rng.random(n)
    P[:n, 42] = rng.random(n) < 0.2
    return types.SimpleNamespace(particles=P, np=n, nt=12, t=0.345, maxparticles=maxp)


def test_save_read_roundtrip(tmp_path):
    pf = _fake_field()
    out = vpm.save(pf, "run_pfield", path=str(tmp_path))
    assert out == "run_pfield.12.xmf;"
    assert sorted(os.listdir(tmp_path)) == ["run_pfield.12.h5", "run_pfield.12.xmf"]
    d = h5min.read(str(tmp_path / "run_pfield.12.h5"))
    assert set(d) == {"np", "nt", "t", "X", "Gamma", "sigma", "circulation", "vol", "static", "i"}
    assert d["np"].shape == () and int(d["np"]) == 37 and int(d["nt"]) == 12 and float(d["t"]) == 0.345
    assert d["X"].shape == (37, 3) and np.array_equal(d["X"], pf.particles[:37, 0:3])
    assert d["i"].dtype == np.int64 and np.array_equal(d["i"], np.arange(1, 38))
    # restart: vpm.read!(pfield, file; overwrite=true, load_time=false)  (simulation.jl:263-265)
    pf2 = types.SimpleNamespace(particles=np.full((64, 43), 9.0), np=5, nt=0, t=0.0, maxparticles=64)
    vpm.read_(pf2, "run_pfield.12.h5", path=str(tmp_path), overwrite=True, load_time=False)
    assert pf2.np == 37 and (pf2.t, pf2.nt) == (0.0, 0)
    for sl in (slice(0, 7), slice(7, 9), slice(42, 43)):
        assert np.array_equal(pf2.particles[:37, sl], pf.particles[:37, sl])
    assert np.all(pf2.particles[:37, 9:42] == 0)
    pf3 = types.SimpleNamespace(particles=np.zeros((80, 43)), np=0, nt=0, t=0.0, maxparticles=80)
    vpm.read_(pf3, "run_pfield.12.h5", path=str(tmp_path))
    vpm.read_(pf3, "run_pfield.12.h5", path=str(tmp_path), overwrite=False, load_time=True)   # append
    assert pf3.np == 74 and (pf3.t, pf3.nt) == (0.345, 12)
    assert np.array_equal(pf3.particles[37:74, 0:7], pf.particles[:37, 0:7])


def test_overflow_on_read(tmp_path):
    pf = _fake_field()
    vpm.save(pf, "f", path=str(tmp_path), add_num=False)
    small = types.SimpleNamespace(particles=np.zeros((10, 43)), np=0, nt=0, t=0.0, maxparticles=10)
    with pytest.raises(RuntimeError, match="PARTICLE OVERFLOW"):
        vpm.read_(small, "f.h5", path=str(tmp_path))


def test_hdf5_byte_layout(tmp_path):
    """Field-by-field check of the structures the HDF5 File Format Specification defines (superblock v0, v1 object header,
    local heap, v1 group B-tree, symbol-table node, dataspace / datatype / layout messages)."""
    path = str(tmp_path / "t.h5")
    h5min.write(path, {"b": np.arange(6, dtype=np.float64).reshape(2, 3), "a": np.int64(7)})
    b = open(path, "rb").read()
    assert b[:8] == b"\x89HDF\r\n\x1a\n"
    assert b[8:16] == bytes([0, 0, 0, 0, 0, 8, 8, 0])                       # versions, offset/length sizes
    leaf_k, int_k, flags = struct.unpack_from("<HHI", b, 16)
    base, free, eof, drv = struct.unpack_from("<QQQQ", b, 24)
    assert (base, free, drv) == (0, h5min.UNDEF, h5min.UNDEF) and eof == len(b) and flags == 0
    name0, root_ohdr, cache, _, btree, heap = struct.unpack_from("<QQIIQQ", b, 56)
    assert name0 == 0 and cache == 1 and root_ohdr == 96
    ver, _, nmsg, refc, hsize = struct.unpack_from("<BBHII", b, root_ohdr)
    assert (ver, nmsg, refc) == (1, 1, 1) and hsize == 24
    mtype, msize = struct.unpack_from("<HH", b, root_ohdr + 16)
    assert mtype == 0x0011 and msize == 16 and struct.unpack_from("<QQ", b, root_ohdr + 24) == (btree, heap)
    assert b[heap:heap + 4] == b"HEAP" and b[btree:btree + 4] == b"TREE"
    dseg_size, free_head, dseg = struct.unpack_from("<QQQ", b, heap + 8)
    assert free_head == 1 and dseg_size % 8 == 0 and b[dseg:dseg + 8] == b"\x00" * 8
    ntype, level, used = struct.unpack_from("<BBH", b, btree + 4)
    assert (ntype, level, used) == (0, 0, 1)
    key0, snod, key1 = struct.unpack_from("<QQQ", b, btree + 24)
    assert key0 == 0 and b[snod:snod + 4] == b"SNOD" and struct.unpack_from("<BBH", b, snod + 4) == (1, 0, 2)
    names = []
    for e in range(2):
        noff, oaddr = struct.unpack_from("<QQ", b, snod + 8 + 40 * e)
        names.append(b[dseg + noff:b.index(b"\x00", dseg + noff)].decode())
    assert names == ["a", "b"]                                               # sorted by name
    assert b[dseg + key1:dseg + key1 + 1] == b"b"                            # right key = largest name in the node
    assert eof % 8 == 0
    d = h5min.read(path)
    assert int(d["a"]) == 7 and np.array_equal(d["b"], np.arange(6.0).reshape(2, 3))
    # datatype message of "b": IEEE binary64 little-endian exactly as libhdf5 encodes H5T_IEEE_F64LE
    _, oaddr_b = struct.unpack_from("<QQ", b, snod + 8 + 40)
    msgs = h5min._Reader(b).messages(oaddr_b)
    dt = [m for m in msgs if m[0] == 0x0003][0][2]
    assert dt[:20] == bytes([0x11, 0x20, 0x3F, 0x00, 8, 0, 0, 0, 0, 0, 64, 0, 52, 11, 0, 52, 0xFF, 0x03, 0, 0])
    sp = [m for m in msgs if m[0] == 0x0001][0][2]
    assert sp[:2] == bytes([1, 2]) and struct.unpack_from("<QQ", sp, 8) == (2, 3)
    lay = [m for m in msgs if m[0] == 0x0008][0][2]
    assert lay[:2] == bytes([3, 1]) and struct.unpack_from("<Q", lay, 10)[0] == 48


def test_xdmf_wrapper(tmp_path):
    pf = _fake_field()
    vpm.save(pf, "w", path=str(tmp_path), num=3, overwrite_time=1.5)
    root = ET.parse(tmp_path / "w.3.xmf").getroot()
    grid = root.find("Domain/Grid")
    assert grid.find("Time").get("Value") == "1.5"
    assert grid.find("Topology").get("Type") == "Polyvertex" and grid.find("Topology").get("Dimensions") == "37"
    geo = grid.find("Geometry/DataItem")
    assert geo.text == "w.3.h5:X" and geo.get("Dimensions") == "37 3" and geo.get("Format") == "HDF"
    attrs = {a.get("Name"): a for a in grid.findall("Attribute")}
    assert set(attrs) == {"Gamma", "sigma", "circulation", "vol", "static", "i"} and attrs["Gamma"].get("Type") == "Vector"


def test_file_opens_the_way_libhdf5_opens_it(tmp_path):
    """the hand-written container read through an independent from-spec reader that follows libhdf5's open
    path (tests/h5spec.py: superblock checks, root group, NAME lookup by bisecting B-tree keys and the symbol-table node,
    object-header and datatype validation) — every dataset of a saved particle field must be found by name and decode to the
    saved values; and the reader's checks are live: breaking one rule at a time makes it refuse the file."""
    from flowunsteady_b200 import h5min
    from tests import h5spec
    rng = np.random.default_rng(0)
    n = 777
    data = {"X": rng.random((n, 3)), "Gamma": rng.standard_normal((n, 3)), "sigma": rng.random(n), "circulation": rng.random(n),
            "vol": rng.random(n), "static": (rng.random(n) < 0.1).astype(np.int64), "i": np.arange(1, n + 1, dtype=np.int64),
            "np": np.int64(n), "nt": np.int64(12), "t": np.float64(0.125)}
    path = str(tmp_path / "pfield.12.h5")
    h5min.write(path, data)
    f = h5spec.File(path)
    for name, ref in data.items():
        got = f.dataset(name)
        assert got.shape == np.shape(ref) and np.array_equal(got, np.asarray(ref)), name
    with pytest.raises(KeyError):
        f.lookup("no_such_dataset")
    raw = bytearray(open(path, "rb").read())

    def broken(mutate, rule_fragment):
        b = bytearray(raw)
        mutate(b)
        q = str(tmp_path / "broken.h5")
        open(q, "wb").write(b)
        with pytest.raises(h5spec.H5SpecError) as ei:
            g = h5spec.File(q)
            for name in data:
                g.dataset(name)
        assert rule_fragment in str(ei.value), str(ei.value)

    broken(lambda b: b.__setitem__(slice(40, 48), (len(raw) + 8).to_bytes(8, "little")), "end-of-file address")
    broken(lambda b: b.__setitem__(8, 2), "superblock version")
    snod = raw.index(b"SNOD")
    e0, e1 = bytes(raw[snod + 8:snod + 48]), bytes(raw[snod + 48:snod + 88])
    broken(lambda b: b.__setitem__(slice(snod + 8, snod + 88), e1 + e0), "sorted by name")       # libhdf5's bisection would miss
    tree = raw.index(b"TREE")
    broken(lambda b: b.__setitem__(slice(tree + 40, tree + 48), (0).to_bytes(8, "little")), "keys strictly increasing")
    broken(lambda b: b.__delitem__(slice(len(raw) - 8, len(raw))), "end-of-file address")          # truncated file

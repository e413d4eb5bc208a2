"""Minimal HDF5 writer / reader for the particle-field files FLOWVPM's `vpm.save` / `vpm.read!` exchange
(`<run>_pfield.<nt>.h5`, /root/reference/src/FLOWUnsteady_simulation.jl:263-265,436-440).

No libhdf5 / h5py exists in this image, so the container format is written by hand from the HDF5 File Format
Specification (version 1.1 structures, the ones every libhdf5 release reads): superblock v0, a root group stored as a
v1 B-tree + local heap + one symbol-table node, v1 object headers, contiguous little-endian datasets of IEEE float64 /
int64.  The reader understands exactly that subset (plus object-header continuation blocks and the extra messages
libhdf5 adds), which covers files HDF5.jl writes for plain arrays with default properties.

NOT VALIDATED AGAINST libhdf5 HERE (none available offline): the tests check the byte layout against the specification's
field tables and round-trip through the reader.  DESIGN.md §7 says so.
"""
from __future__ import annotations

import struct
from typing import Dict

import numpy as np

SIG = b"\x89HDF\r\n\x1a\n"
UNDEF = 0xFFFFFFFFFFFFFFFF
LEAF_K = 16        # symbol-table node holds 2 * LEAF_K entries (enough for a particle field's datasets)
INTERNAL_K = 16


def _pad8(b: bytes) -> bytes:
    return b + b"\x00" * (-len(b) % 8)


def _msg(mtype: int, data: bytes, flags: int = 0) -> bytes:
    data = _pad8(data)
    return struct.pack("<HHB3x", mtype, len(data), flags) + data


def _dataspace(shape) -> bytes:
    return struct.pack("<BBB5x", 1, len(shape), 0) + b"".join(struct.pack("<Q", int(d)) for d in shape)


def _datatype(dtype: np.dtype) -> bytes:
    if dtype == np.float64:
        # class 1 (floating point), version 1; bit field: little-endian, mantissa normalisation = implied msb (2 << 4),
        # sign bit location 63; properties: bit offset 0, precision 64, exponent at 52 (11 bits), mantissa at 0 (52 bits), bias 1023
        return struct.pack("<BBBBI", 0x11, 0x20, 0x3F, 0x00, 8) + struct.pack("<HHBBBBI", 0, 64, 52, 11, 0, 52, 1023)
    if dtype == np.int64:
        # class 0 (fixed point), version 1; bit field: little-endian, signed (bit 3)
        return struct.pack("<BBBBI", 0x10, 0x08, 0x00, 0x00, 8) + struct.pack("<HH", 0, 64)
    raise TypeError(f"unsupported dtype {dtype}")


def _object_header(messages) -> bytes:
    body = b"".join(messages)
    return struct.pack("<BBHII4x", 1, 0, len(messages), 1, len(body)) + body


def write(path: str, datasets: Dict[str, np.ndarray]) -> None:
    """Write `datasets` (name -> float64 / int64 array or scalar) as the root group's members."""
    names = sorted(datasets)                      # symbol-table entries are ordered by name
    if len(names) > 2 * LEAF_K:
        raise ValueError("too many datasets for one symbol-table node")
    arrays = {}
    for n in names:
        a = np.asarray(datasets[n])
        if a.dtype.kind == "f":
            a = a.astype("<f8", order="C")
        elif a.dtype.kind in "iub":
            a = a.astype("<i8", order="C")
        else:
            raise TypeError(f"dataset {n!r}: unsupported dtype {a.dtype}")
        arrays[n] = a                                # 0-d stays 0-d (scalar dataspace), as HDF5.jl writes scalars

    # ---- local heap data: "" at offset 0, then every name, 8-byte padded
    heap = bytearray(b"\x00" * 8)
    name_off = {}
    for n in names:
        name_off[n] = len(heap)
        heap += _pad8(n.encode() + b"\x00")
    heap_data = bytes(heap)

    # ---- addresses
    pos = 96                                        # superblock v0
    root_ohdr = _object_header([_msg(0x0011, struct.pack("<QQ", 0, 0))])   # placeholder to get the size
    a_root = pos; pos += len(root_ohdr)
    a_heap = pos; pos += 32
    a_heapdata = pos; pos += len(heap_data)
    btree_size = 24 + (2 * INTERNAL_K + 1) * 8 + 2 * INTERNAL_K * 8
    a_btree = pos; pos += btree_size
    snod_size = 8 + 2 * LEAF_K * 40
    a_snod = pos; pos += snod_size

    def dataset_header(a: np.ndarray, data_addr: int) -> bytes:
        msgs = [
            _msg(0x0001, _dataspace(a.shape)),
            _msg(0x0003, _datatype(a.dtype), flags=1),                      # constant message
            _msg(0x0005, struct.pack("<BBBBI", 2, 2, 2, 1, 0)),             # fill value v2: alloc late, write if-set, default (size 0)
            _msg(0x0008, struct.pack("<BBQQ", 3, 1, data_addr, a.nbytes)),  # layout v3, contiguous
        ]
        return _object_header(msgs)

    a_ohdr, a_data = {}, {}
    for n in names:                                 # headers first (sizes do not depend on the addresses)
        a_ohdr[n] = pos
        pos += len(dataset_header(arrays[n], 0))
    for n in names:
        a_data[n] = pos if arrays[n].nbytes else UNDEF
        pos += (arrays[n].nbytes + 7) // 8 * 8
    eof = pos

    # ---- emit
    out = bytearray()
    root_entry = struct.pack("<QQII", 0, a_root, 1, 0) + struct.pack("<QQ", a_btree, a_heap)
    out += SIG + struct.pack("<BBBBBBBBHHI", 0, 0, 0, 0, 0, 8, 8, 0, LEAF_K, INTERNAL_K, 0)
    out += struct.pack("<QQQQ", 0, UNDEF, eof, UNDEF) + root_entry
    assert len(out) == 96
    out += _object_header([_msg(0x0011, struct.pack("<QQ", a_btree, a_heap))])
    # free-list head: libhdf5's H5HL_FREE_NULL (= 1) marks "no free block" on disk
    out += b"HEAP" + struct.pack("<B3xQQQ", 0, len(heap_data), 1, a_heapdata)
    out += heap_data
    bt = b"TREE" + struct.pack("<BBHQQ", 0, 0, 1 if names else 0, UNDEF, UNDEF)
    bt += struct.pack("<QQQ", 0, a_snod, name_off[names[-1]] if names else 0)
    out += bt + b"\x00" * (btree_size - len(bt))
    sn = b"SNOD" + struct.pack("<BBH", 1, 0, len(names))
    for n in names:
        sn += struct.pack("<QQII16x", name_off[n], a_ohdr[n], 0, 0)
    out += sn + b"\x00" * (snod_size - len(sn))
    for n in names:
        assert len(out) == a_ohdr[n]
        out += dataset_header(arrays[n], a_data[n])
    for n in names:
        if arrays[n].nbytes:
            assert len(out) == a_data[n]
            out += _pad8(arrays[n].tobytes())
    assert len(out) == eof
    with open(path, "wb") as f:
        f.write(out)


# ---------------------------------------------------------------------------------------------------------------------
class _Reader:
    def __init__(self, buf: bytes):
        self.b = buf
        if buf[:8] != SIG:
            raise ValueError("not an HDF5 file")
        ver = buf[8]
        if ver not in (0, 1):
            raise NotImplementedError(f"superblock version {ver} is not supported (only the 1.x-compatible layout)")
        self.so, self.sl = buf[13], buf[14]
        if (self.so, self.sl) != (8, 8):
            raise NotImplementedError("only 8-byte offsets/lengths are supported")
        p = 24 if ver == 0 else 28
        self.base = struct.unpack_from("<Q", buf, p)[0]
        root = p + 32
        self.root_ohdr = struct.unpack_from("<Q", buf, root + 8)[0]

    def messages(self, addr: int):
        """(type, flags, data) of every message of a v1 object header, following continuation blocks."""
        b = self.b
        ver, _, nmsg, _, size = struct.unpack_from("<BBHII", b, addr)
        if ver != 1:
            raise NotImplementedError("only version-1 object headers are supported")
        blocks = [(addr + 16, size)]
        out = []
        while blocks and len(out) < nmsg:
            p, left = blocks.pop(0)
            end = p + left
            while p + 8 <= end and len(out) < nmsg:
                mtype, msize, flags = struct.unpack_from("<HHB", b, p)
                data = b[p + 8:p + 8 + msize]
                if mtype == 0x0010:                              # continuation
                    off, ln = struct.unpack_from("<QQ", data, 0)
                    blocks.append((self.base + off, ln))
                out.append((mtype, flags, data))
                p += 8 + msize
        return out

    def group_members(self, ohdr: int) -> Dict[str, int]:
        st = [d for t, _, d in self.messages(ohdr) if t == 0x0011]
        if not st:
            raise NotImplementedError("root group is not a symbol-table group (new-style links are not supported)")
        btree, heap = struct.unpack_from("<QQ", st[0], 0)
        b = self.b
        if b[heap:heap + 4] != b"HEAP":
            raise ValueError("bad local heap")
        heap_data = struct.unpack_from("<Q", b, heap + 24)[0] + self.base
        members = {}

        def walk(node):
            if b[node:node + 4] != b"TREE":
                raise ValueError("bad B-tree node")
            _, level, used = struct.unpack_from("<BBH", b, node + 4)
            p = node + 24
            for k in range(used):
                child = struct.unpack_from("<Q", b, p + 8 + 16 * k)[0] + self.base
                if level > 0:
                    walk(child)
                else:
                    if b[child:child + 4] != b"SNOD":
                        raise ValueError("bad symbol-table node")
                    nsym = struct.unpack_from("<H", b, child + 6)[0]
                    for e in range(nsym):
                        noff, oaddr = struct.unpack_from("<QQ", b, child + 8 + 40 * e)
                        q = heap_data + noff
                        name = b[q:b.index(b"\x00", q)].decode()
                        members[name] = oaddr + self.base
        walk(btree + self.base)
        return members

    def dataset(self, ohdr: int) -> np.ndarray:
        shape = dtype = layout = None
        for t, _, d in self.messages(ohdr):
            if t == 0x0001:
                ver, rank = d[0], d[1]
                off = 8 if ver == 1 else 4
                shape = tuple(struct.unpack_from("<Q", d, off + 8 * k)[0] for k in range(rank))
            elif t == 0x0003:
                cls, size = d[0] & 0x0F, struct.unpack_from("<I", d, 4)[0]
                if d[1] & 1:
                    raise NotImplementedError("big-endian data")
                if cls == 1 and size == 8:
                    dtype = np.dtype("<f8")
                elif cls == 1 and size == 4:
                    dtype = np.dtype("<f4")
                elif cls == 0 and size in (1, 2, 4, 8):
                    dtype = np.dtype(("<i" if d[1] & 0x08 else "<u") + str(size))
                else:
                    raise NotImplementedError(f"datatype class {cls} size {size}")
            elif t == 0x0008:
                ver = d[0]
                if ver == 3 and d[1] == 1:
                    layout = struct.unpack_from("<QQ", d, 2)
                elif ver == 3 and d[1] == 0:                      # compact: data inside the message
                    n = struct.unpack_from("<H", d, 2)[0]
                    layout = ("compact", d[4:4 + n])
                else:
                    raise NotImplementedError("only contiguous / compact version-3 layouts are supported (no chunking)")
        if shape is None or dtype is None or layout is None:
            raise ValueError("incomplete dataset header")
        count = int(np.prod(shape)) if shape else 1
        if layout[0] == "compact":
            raw = layout[1]
        else:
            addr, size = layout
            raw = b"" if addr == UNDEF else self.b[addr + self.base:addr + self.base + count * dtype.itemsize]
        a = np.frombuffer(raw, dtype=dtype, count=count if raw else 0)
        return a.reshape(shape).copy() if raw else np.zeros(shape, dtype)


def read(path: str) -> Dict[str, np.ndarray]:
    """All datasets of the root group as numpy arrays (scalars come back 0-dimensional)."""
    with open(path, "rb") as f:
        r = _Reader(f.read())
    return {name: r.dataset(addr) for name, addr in r.group_members(r.root_ohdr).items()}

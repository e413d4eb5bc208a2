"""A second, independent reading of the HDF5 File Format Specification (v1.1 structures) — TEST INFRASTRUCTURE.

flowunsteady_b200/h5min.py writes `<run>_pfield.<nt>.h5` by hand because no libhdf5 exists in this image.  This module opens
such a file the way libhdf5 does and applies the validity checks libhdf5 applies on that path, so that "libhdf5 would open
it" is tested from the READER's side of the specification rather than by round-tripping through h5min's own reader:

  H5F open        signature, superblock v0 field by field, sizes of offsets / lengths, B-tree K values, end-of-file address
                  against the real file size (libhdf5: "truncated file" when eof < stored eoa)
  root group      symbol-table entry with cache type 1 whose scratch pad must equal the object header's symbol-table message
  object header   v1 prefix, messages 8-byte aligned, sizes multiples of 8, messages must fill the chunk exactly
  H5G lookup      by NAME: binary search of the group B-tree's keys (heap strings, strcmp) then of the symbol-table node —
                  this only works if keys and entries are sorted the way libhdf5 expects
  H5D open        dataspace v1, datatype v1 (class, size, byte order, IEEE field layout), layout v3 contiguous, fill value v2;
                  data extent inside the file

Everything that deviates raises H5SpecError with the name of the violated rule.
"""
from __future__ import annotations

import struct

import numpy as np

SIG = b"\x89HDF\r\n\x1a\n"
UNDEF = 0xFFFFFFFFFFFFFFFF


class H5SpecError(ValueError):
    pass


def _req(cond, rule):
    if not cond:
        raise H5SpecError(rule)


class File:
    def __init__(self, path):
        with open(path, "rb") as f:
            self.b = f.read()
        b = self.b
        _req(b[:8] == SIG, "format signature at offset 0")
        (sb_ver, fs_ver, root_ver, r0, shm_ver, so, sl, r1, leaf_k, int_k, flags) = struct.unpack_from("<BBBBBBBBHHI", b, 8)
        _req(sb_ver == 0, "superblock version 0")
        _req(fs_ver == 0 and root_ver == 0 and shm_ver == 0, "free-space / root-group / shared-header versions are 0")
        _req(r0 == 0 and r1 == 0, "superblock reserved bytes are zero")
        _req(so == 8 and sl == 8, "size of offsets and lengths = 8")
        _req(leaf_k > 0 and int_k > 0, "group leaf / internal node K > 0")
        _req(flags == 0, "file consistency flags clear (file was closed)")
        self.leaf_k, self.int_k = leaf_k, int_k
        base, fsaddr, eof, drv = struct.unpack_from("<QQQQ", b, 24)
        _req(base == 0, "base address 0")
        _req(fsaddr == UNDEF and drv == UNDEF, "no free-space info / driver info block")
        _req(eof <= len(b), "end-of-file address beyond the file (libhdf5: truncated file)")
        _req(eof == len(b), "end-of-file address equals the file size")
        self.eof = eof
        name_off, ohdr, cache, rsv = struct.unpack_from("<QQII", b, 56)
        _req(name_off == 0 and rsv == 0, "root symbol-table entry: link name offset 0, reserved 0")
        _req(cache == 1, "root entry cache type 1 (group: B-tree and heap addresses cached)")
        self.root_btree, self.root_heap = struct.unpack_from("<QQ", b, 80)
        msgs = self.object_header(ohdr)
        st = [d for t, f, d in msgs if t == 0x0011]
        _req(len(st) == 1, "root object header carries one symbol-table message")
        _req(struct.unpack_from("<QQ", st[0]) == (self.root_btree, self.root_heap), "scratch pad == symbol-table message")
        self.heap_data, self.heap_size = self.local_heap(self.root_heap)

    # ---- v1 object header ---------------------------------------------------------------------------------------------
    def object_header(self, addr):
        b = self.b
        _req(addr % 8 == 0 and addr + 16 <= self.eof, "object header address aligned and inside the file")
        ver, rsv, nmsg, refc, size = struct.unpack_from("<BBHII", b, addr)
        _req(ver == 1 and rsv == 0, "object header version 1, reserved 0")
        _req(refc >= 1, "object reference count >= 1")
        _req(b[addr + 12:addr + 16] == b"\x00" * 4, "object header prefix padded to 16 bytes")
        p, end = addr + 16, addr + 16 + size
        _req(end <= self.eof, "object header chunk inside the file")
        out = []
        while p < end:
            _req(p + 8 <= end, "message header inside the chunk")
            mtype, msize, flags, r3 = struct.unpack_from("<HHB3s", b, p)
            _req(r3 == b"\x00\x00\x00", "message header reserved bytes zero")
            _req(msize % 8 == 0, "message size is a multiple of 8")
            _req(p + 8 + msize <= end, "message inside the chunk")
            _req(mtype != 0x0010, "no continuation blocks expected in these files")
            out.append((mtype, flags, b[p + 8:p + 8 + msize]))
            p += 8 + msize
        _req(p == end, "messages fill the chunk exactly (a gap needs a NIL message)")
        _req(len(out) == nmsg, "number of messages matches the prefix")
        return out

    # ---- local heap ----------------------------------------------------------------------------------------------------
    def local_heap(self, addr):
        b = self.b
        _req(b[addr:addr + 4] == b"HEAP", "local heap signature")
        ver, r0, r1, r2 = struct.unpack_from("<BBBB", b, addr + 4)
        _req(ver == 0 and (r0, r1, r2) == (0, 0, 0), "local heap version 0, reserved 0")
        size, free, data = struct.unpack_from("<QQQ", b, addr + 8)
        _req(data + size <= self.eof, "heap data segment inside the file")
        _req(free == 1 or free + 16 <= size, "free-list head: H5HL_FREE_NULL (1) or a block inside the segment")
        _req(size % 8 == 0, "heap data segment size aligned")
        return data, size

    def heap_string(self, off):
        _req(off < self.heap_size, "heap offset inside the data segment")
        q = self.heap_data + off
        e = self.b.index(b"\x00", q)
        _req(e < self.heap_data + self.heap_size, "heap string terminated inside the segment")
        return self.b[q:e]

    # ---- H5G lookup by name (H5B_find + H5G__node_found) ----------------------------------------------------------------
    def lookup(self, name: str) -> int:
        b, target = self.b, name.encode()
        node = self.root_btree
        while True:
            _req(b[node:node + 4] == b"TREE", "B-tree node signature")
            ntype, level, used, left, right = struct.unpack_from("<BBHQQ", b, node + 4)
            _req(ntype == 0, "B-tree node type 0 (group)")
            _req(0 < used <= 2 * self.int_k, "entries used within 1 .. 2K")
            _req(left == UNDEF and right == UNDEF, "single node per level: no siblings")
            keys = [struct.unpack_from("<Q", b, node + 24 + 16 * k)[0] for k in range(used + 1)]
            kids = [struct.unpack_from("<Q", b, node + 32 + 16 * k)[0] for k in range(used)]
            ks = [self.heap_string(k) for k in keys]
            _req(all(x < y for x, y in zip(ks, ks[1:])), "B-tree keys strictly increasing (strcmp)")
            # child i holds names with key[i] < name <= key[i + 1]
            lo, hi, idx = 0, used, -1
            while lo < hi:
                mid = (lo + hi) // 2
                if target <= ks[mid]:
                    hi = mid
                elif target > ks[mid + 1]:
                    lo = mid + 1
                else:
                    idx = mid
                    break
            if idx < 0:
                raise KeyError(name)
            node = kids[idx]
            if level == 0:
                break
        _req(b[node:node + 4] == b"SNOD", "symbol-table node signature")
        ver, rsv, nsym = struct.unpack_from("<BBH", b, node + 4)
        _req(ver == 1 and rsv == 0, "symbol-table node version 1")
        _req(0 < nsym <= 2 * self.leaf_k, "symbols within 1 .. 2 leaf K")
        ents = [struct.unpack_from("<QQII", b, node + 8 + 40 * e) for e in range(nsym)]
        names = [self.heap_string(e[0]) for e in ents]
        _req(all(x < y for x, y in zip(names, names[1:])), "symbol-table entries sorted by name (strcmp)")
        _req(names[-1] == ks[idx + 1], "right key of the node == its largest name")
        lo, hi = 0, nsym
        while lo < hi:                                     # libhdf5 bisects the node too
            mid = (lo + hi) // 2
            if names[mid] == target:
                _req(ents[mid][2] == 0 and ents[mid][3] == 0, "dataset entry: cache type 0, reserved 0")
                return ents[mid][1]
            if names[mid] < target:
                lo = mid + 1
            else:
                hi = mid
        raise KeyError(name)

    # ---- H5D open ------------------------------------------------------------------------------------------------------
    def dataset(self, name: str) -> np.ndarray:
        msgs = self.object_header(self.lookup(name))
        by = {}
        for t, f, d in msgs:
            _req(t not in by, "one message of each type")
            by[t] = (f, d)
        _req(0x0001 in by and 0x0003 in by and 0x0008 in by, "dataspace, datatype and layout messages present")
        d = by[0x0001][1]
        ver, rank, flags = d[0], d[1], d[2]
        _req(ver == 1 and flags == 0 and d[3:8] == b"\x00" * 5, "dataspace version 1, no max dims, reserved 0")
        shape = tuple(struct.unpack_from("<Q", d, 8 + 8 * k)[0] for k in range(rank))
        d = by[0x0003][1]
        cls, tver = d[0] & 0x0F, d[0] >> 4
        size = struct.unpack_from("<I", d, 4)[0]
        _req(tver == 1, "datatype version 1")
        _req(d[1] & 1 == 0, "little-endian")
        if cls == 1:
            _req(size == 8, "8-byte float")
            _req((d[1] >> 4) & 3 == 2, "mantissa normalisation: msb implied")
            _req(d[2] == 63, "sign bit at 63")
            off, prec, eloc, esize, mloc, msize, bias = struct.unpack_from("<HHBBBBI", d, 8)
            _req((off, prec, eloc, esize, mloc, msize, bias) == (0, 64, 52, 11, 0, 52, 1023), "IEEE 754 binary64 field layout")
            dtype = np.dtype("<f8")
        else:
            _req(cls == 0 and size == 8, "8-byte fixed point")
            _req(d[1] & 0x08, "signed")
            off, prec = struct.unpack_from("<HH", d, 8)
            _req((off, prec) == (0, 64), "bit offset 0, precision 64")
            dtype = np.dtype("<i8")
        d = by[0x0008][1]
        _req(d[0] == 3 and d[1] == 1, "layout version 3, contiguous")
        addr, nbytes = struct.unpack_from("<QQ", d, 2)
        count = int(np.prod(shape)) if shape else 1
        _req(nbytes == count * 8, "layout size == elements x 8")
        if 0x0005 in by:
            f = by[0x0005][1]
            _req(f[0] == 2, "fill value message version 2")
            _req(f[1] in (1, 2, 3) and f[2] in (0, 1, 2), "space allocation / fill write time in range")
            _req(f[3] in (0, 1), "fill value defined flag")
        if nbytes == 0:
            return np.zeros(shape, dtype)
        _req(addr != UNDEF and addr % 8 == 0 and addr + nbytes <= self.eof, "raw data allocated, aligned and inside the file")
        return np.frombuffer(self.b, dtype=dtype, count=count, offset=addr).reshape(shape).copy()

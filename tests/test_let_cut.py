"""The cut of the Morton curve that shards UJ_fmm over the ranks (vpmb200_let_cut: host arithmetic only, so it runs here without
a GPU) and the collectives wrapper of flowunsteady_b200/dist.py over a world_size-2 gloo group on CPU.

What must hold for the local essential tree to reproduce the one-GPU tree (flowunsteady_b200/csrc/fmm_let.cuh):
every rank derives the SAME splitters from the all-reduced histogram; a splitter never falls inside a top cell that holds
<= ncrit particles (a leaf of the global tree has one owner); counts (or counted work) are equal up to one unit."""
import ctypes as C
import os
import socket

import numpy as np
import pytest

LC = 3
BINS = 8 ** LC
SHIFT = 3 * (21 - LC)


def _cut(hist, nparts, ncrit=50, work=None, lc=LC):
    from flowunsteady_b200 import _lib
    h = np.ascontiguousarray(hist, dtype=np.int32)
    sp = (C.c_uint64 * (nparts + 1))()
    wp = None
    if work is not None:
        w = np.ascontiguousarray(work, dtype=np.int64)
        wp = w.ctypes.data_as(C.POINTER(C.c_int64))
    rc = _lib.lib().vpmb200_let_cut(h.ctypes.data_as(C.POINTER(C.c_int32)), wp, lc, ncrit, nparts, sp)
    assert rc == 0
    return [int(v) for v in sp]


def _owned(hist, sp):
    """particles (and bins) of every rank for splitters given as keys"""
    edges = [min(s >> SHIFT, BINS) for s in sp]
    return [int(np.sum(hist[a:b])) for a, b in zip(edges, edges[1:])], edges


def test_equal_counts_within_one_unit_and_monotone_splitters():
    rng = np.random.default_rng(0)
    hist = rng.integers(0, 400, BINS).astype(np.int32)
    for nparts in (1, 2, 3, 8, 13):
        sp = _cut(hist, nparts)
        assert sp[0] == 0 and sp[-1] == 1 << 63 and all(a <= b for a, b in zip(sp, sp[1:]))
        assert all(s % (1 << SHIFT) == 0 for s in sp[:-1])                      # bin-aligned
        counts, _ = _owned(hist, sp)
        assert sum(counts) == int(hist.sum())
        assert max(counts) - min(counts) <= 2 * int(hist.max())
    assert _cut(hist, 8) == _cut(hist.copy(), 8)                                 # deterministic


def test_a_leaf_of_the_global_top_tree_has_one_owner():
    """A level-1 octant with <= ncrit particles spread over many level-3 bins is ONE unit: no splitter inside it, even when the
    equal-count target falls there."""
    hist = np.zeros(BINS, dtype=np.int32)
    hist[:64] = 100                      # octant 0: dense (64 level-3 bins)
    hist[64:128:2] = 1                   # octant 1: 32 particles in 32 bins -> a leaf of the global tree at level 1
    hist[128:192] = 100                  # octant 2: dense
    for nparts in (2, 3, 5, 7):
        sp = _cut(hist, nparts, ncrit=50)
        for s in sp[1:-1]:
            b = s >> SHIFT
            assert not (64 < b < 128), (nparts, b)
    # with ncrit below its count the octant is split like any other
    sp = _cut(np.where(np.arange(BINS) // 64 == 1, 7, 0).astype(np.int32), 4, ncrit=50)
    assert any(64 < (s >> SHIFT) < 128 for s in sp[1:-1])


def test_work_weighted_cut_equalises_work_not_counts():
    hist = np.full(BINS, 100, dtype=np.int32)
    work = np.where(np.arange(BINS) < BINS // 4, 9000, 1000).astype(np.int64)     # the first quarter costs 9x per particle
    by_count, _ = _owned(hist, _cut(hist, 4))
    assert max(by_count) - min(by_count) <= 200
    sp = _cut(hist, 4, work=work)
    counts, edges = _owned(hist, sp)
    blended = work + 0.25 * work.sum() / hist.sum() * hist                         # the weight vpmb200_let_cut uses
    loads = [float(blended[a:b].sum()) for a, b in zip(edges, edges[1:])]
    assert max(loads) / (sum(loads) / 4) < 1.05
    assert counts[0] < 0.6 * counts[-1]                                            # fewer particles where they cost more
    assert _cut(hist, 4, work=np.zeros(BINS, dtype=np.int64)) == _cut(hist, 4)     # no work counted yet: cut by count


def test_more_ranks_than_units_and_empty_field():
    hist = np.zeros(BINS, dtype=np.int32)
    hist[5] = 30
    sp = _cut(hist, 6)
    counts, _ = _owned(hist, sp)
    assert sorted(counts) == [0, 0, 0, 0, 0, 30]
    assert _cut(np.zeros(BINS, dtype=np.int32), 3) == [0, 1 << 63, 1 << 63, 1 << 63]


# ---- the collectives wrapper over gloo (2 processes on CPU) -----------------------------------------------------------------
def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, out_dir):
    import torch
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from flowunsteady_b200.dist import TorchCollectives
        c = TorchCollectives()
        rng = np.random.default_rng(100 + rank)
        keys = np.sort(rng.integers(0, 1 << 63, 5000 + 700 * rank, dtype=np.uint64))      # this rank's sorted Morton keys
        hist = torch.from_numpy(np.bincount((keys >> np.uint64(SHIFT)).astype(np.int64), minlength=BINS).astype(np.int32))
        c.all_reduce_(hist, "sum")
        sp = _cut(hist.numpy(), world)
        send = [int(np.searchsorted(keys, np.uint64(min(b, (1 << 63) - 1)), "left")) for b in sp]
        send[-1] = len(keys)
        send_counts = [b - a for a, b in zip(send, send[1:])]
        counts = c.all_gather_ints(send_counts, "cpu")
        recv_counts = [counts[q][rank] for q in range(world)]
        rows = torch.from_numpy(np.stack([keys.astype(np.float64), np.full(len(keys), float(rank))], 1).copy())
        got = c.all_to_all_rows(rows, send_counts, recv_counts)
        back = c.all_to_all_rows(got * 2.0, recv_counts, send_counts)                       # the inverse exchange
        lo = torch.tensor([float(len(keys)), -float(rank)])
        c.all_reduce_(lo, "max")
        np.savez(os.path.join(out_dir, f"r{rank}.npz"), sp=np.array(sp, dtype=np.uint64), got=got.numpy(), back=back.numpy(),
                 rows=rows.numpy(), counts=np.array(counts), lo=lo.numpy())
    finally:
        dist.destroy_process_group()


def test_partition_and_row_exchange_over_gloo(tmp_path):
    import torch.multiprocessing as mp
    mp.spawn(_worker, args=(2, _free_port(), str(tmp_path)), nprocs=2, join=True)
    r = [np.load(tmp_path / f"r{k}.npz") for k in range(2)]
    assert np.array_equal(r[0]["sp"], r[1]["sp"]) and np.array_equal(r[0]["counts"], r[1]["counts"])   # same cut everywhere
    allkeys = np.concatenate([r[0]["rows"][:, 0], r[1]["rows"][:, 0]])
    for k in range(2):
        lo, hi = float(r[0]["sp"][k]), float(r[0]["sp"][k + 1])
        got = r[k]["got"]
        assert np.all((got[:, 0] >= lo) & (got[:, 0] < hi))                        # every received particle is in the owner's range
        assert got.shape[0] == int(np.sum((allkeys >= lo) & (allkeys < hi)))        # and none is missing
        assert np.array_equal(got[:, 1], np.sort(got[:, 1]))                        # grouped by source rank
        assert np.array_equal(r[k]["back"], 2.0 * r[k]["rows"])                     # the inverse all-to-all restores the home order
        assert r[k]["lo"][0] == 5700 and r[k]["lo"][1] == 0.0
    assert abs(r[0]["got"].shape[0] - r[1]["got"].shape[0]) < 0.1 * len(allkeys)

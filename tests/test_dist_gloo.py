"""world_size-2 gloo tests (CPU) of the sharded direct-path orchestration in flowunsteady_b200.dist.

The CUDA engine is replaced by tests/fake_backend.py (numpy + oracle pair sums); what is under test is the host
logic: partitioning, tile all-gather, own-tiles-first ordering, accumulate flags and the stage sequence of
pfield.SFS / vpm.nextstep.  Reference result: the single-process oracle on the whole field.
"""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from tests.util import mixed_field, relmax


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, case, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import flowunsteady_b200 as fb
        from flowunsteady_b200.dist import ShardedField, partition
        from oracle import oracle as o
        from tests.fake_backend import FakeBackend
        n = case["n"]
        x, g, s, static = mixed_field(n, seed=17)
        g = g * 50.0
        static = np.where(np.all(g == 0, axis=1), 1.0, static)
        P = fb.new_particles(x, g, s, static=static)
        kw = case["schemes"]
        se = fb.default_schemes(**kw)
        so = o.default_schemes(**{k: v for k, v in kw.items() if k != "uj"})      # the oracle has one U/J evaluator: the exact sum
        lo, hi = partition(n, world)[rank]
        be = FakeBackend(P[lo:hi].copy(), se, so)
        sf = ShardedField(be, max_local=hi - lo + 5, device="cpu")
        if case["op"] == "uj":
            sf.uj(True, True, True)
        else:
            for _ in range(2):
                sf.nextstep(2e-3, (1.0, -0.5, 0.25), relax=True)
        np.save(os.path.join(out_dir, f"shard{rank}.npy"), be.P)
        np.save(os.path.join(out_dir, f"time{rank}.npy"), np.array(be.get_time()))
    finally:
        dist.destroy_process_group()


CASES = {
    "uj_estr_ragged": dict(n=777, op="uj", schemes=dict(sfs="constant")),
    "rk3_pedrizzetti": dict(n=600, op="step", schemes=dict(integration="rungekutta3")),
    "euler_dynamic_sfs": dict(n=515, op="step", schemes=dict(integration="euler", sfs="dynamic", force_positive=1, clippings=1)),
    "rk3_constant_sfs_tiny": dict(n=3, op="step", schemes=dict(integration="rungekutta3", sfs="constant", clippings=1)),
    # vpm_UJ = UJ_fmm over the sharded field (ShardedField._uj_fmm: gather of (X, Gamma, sigma), share split, all-reduce of the
    # U / J / E_str rows).  The stand-in backend's "FMM" is the exact sum, so the single-process oracle is the truth here too.
    "fmm_uj_estr_ragged": dict(n=333, op="uj", schemes=dict(sfs="constant", uj="fmm")),
    "fmm_rk3_dynamic_sfs": dict(n=301, op="step", schemes=dict(integration="rungekutta3", sfs="dynamic", force_positive=1,
                                                                clippings=1, uj="fmm")),
}


@pytest.mark.parametrize("name", sorted(CASES))
def test_sharded_matches_single_process(name, tmp_path):
    import flowunsteady_b200 as fb
    from flowunsteady_b200.dist import partition
    from oracle import oracle as o
    case = CASES[name]
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), case, str(tmp_path)), nprocs=world, join=True)
    n = case["n"]
    x, g, s, static = mixed_field(n, seed=17)
    g = g * 50.0
    static = np.where(np.all(g == 0, axis=1), 1.0, static)
    Po = fb.new_particles(x, g, s, static=static)
    so = o.default_schemes(**{k: v for k, v in case["schemes"].items() if k != "uj"})
    if case["op"] == "uj":
        o.field_uj(Po, so, reset=True, reset_sfs=True, sfs=True)
        t_expect = (0.0, 0)
    else:
        t, nt = 0.0, 0
        for _ in range(2):
            t, nt = o.nextstep(Po, so, 2e-3, (1.0, -0.5, 0.25), relax=True, t=t, nt=nt)
        t_expect = (t, nt)
    got = np.concatenate([np.load(tmp_path / f"shard{r}.npy") for r in range(world)])
    assert got.shape == Po.shape
    tol = 1e-9 if case["schemes"].get("sfs") == "dynamic" else 1e-11
    for name_, sl in dict(X=slice(0, 3), Gamma=slice(3, 6), sigma=slice(6, 7), U=slice(9, 12), J=slice(15, 24),
                          SFS=slice(39, 42)).items():
        assert relmax(got[:, sl], Po[:, sl]) < tol, name_
    for r in range(world):
        tt = np.load(tmp_path / f"time{r}.npy")
        assert tt[0] == pytest.approx(t_expect[0]) and int(tt[1]) == t_expect[1]
    assert [hi - lo for lo, hi in partition(n, world)] == [np.load(tmp_path / f"shard{r}.npy").shape[0] for r in range(world)]


def test_partition_properties():
    from flowunsteady_b200.dist import partition
    for n in (0, 1, 7, 8, 1_000_000, 999_999):
        for w in (1, 2, 3, 8):
            p = partition(n, w)
            assert p[0][0] == 0 and p[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(p, p[1:]))
            sizes = [hi - lo for lo, hi in p]
            assert max(sizes) - min(sizes) <= 1

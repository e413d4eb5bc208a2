"""world_size-2/3 gloo tests (CPU) of the local-essential-tree orchestration in flowunsteady_b200.dist — both exchange modes
(`fmm="let"`: all-gather of skeletons / multipoles / records; `fmm="let_halo"`: skeletons only + request / reply).

The CUDA engine's `vpmb200_let_*` phases are replaced by tests/fake_let_backend.py (numpy + the oracle's pair sums over a
trivial bin tree); what is under test is the host logic between them: which collective carries what, with which counts and
in which order, far-field reuse between DynamicSFS's two evaluations, ranks without particles.  The same choreography runs
against the real engine in tests/test_gpu_let.py (in-process ranks on one GPU) and tests/test_gpu_dist.py (NCCL)."""
import os
import socket

import numpy as np
import pytest
import torch.distributed as dist
import torch.multiprocessing as mp

from tests.util import mixed_field, relmax

MODES = ("let", "let_halo")
GROUPS = dict(X=slice(0, 3), Gamma=slice(3, 6), sigma=slice(6, 7), U=slice(9, 12), J=slice(15, 24), SFS=slice(39, 42))


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _field(n):
    import flowunsteady_b200 as fb
    x, g, s, static = mixed_field(n, seed=17)
    g = g * 50.0
    static = np.where(np.all(g == 0, axis=1), 1.0, static)
    return fb.new_particles(x, g, s, static=static)


def _run(case, mode, rank, world, coll=None):
    """one rank's share of the case on the stand-in backend -> (its particle rows afterwards, the phase log)"""
    import flowunsteady_b200 as fb
    from flowunsteady_b200.dist import ShardedField, partition
    from oracle import oracle as o
    from tests.fake_let_backend import FakeLetBackend
    P = _field(case["n"])
    kw = case["schemes"]
    se = fb.default_schemes(uj="fmm", **kw)
    so = o.default_schemes(**kw)
    bounds = case.get("bounds", {}).get(world) or partition(case["n"], world)
    lo, hi = bounds[rank]
    be = FakeLetBackend(P[lo:hi].copy(), se, so, far=case.get("far"))
    sf = ShardedField(be, max_local=max(hi - lo, 1) + 5, device="cpu", coll=coll, fmm=mode, let_level=case.get("level", 2))
    assert sf.fmm_mode == mode
    if case["op"] == "uj":
        sf.uj(True, True, True)
    elif case["op"] == "uj_twice":
        sf.uj(True, False, False)
        sf.uj(False, False, False)
    else:
        for _ in range(2):
            sf.nextstep(2e-3, (1.0, -0.5, 0.25), relax=True)
    return be.P, be.calls


def _worker(rank, world, port, case, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        for mode in MODES:
            P, calls = _run(case, mode, rank, world)
            np.save(os.path.join(out_dir, f"shard_{mode}_{rank}.npy"), P)
            with open(os.path.join(out_dir, f"calls_{mode}_{rank}.txt"), "w") as f:
                f.write(" ".join(calls))
    finally:
        dist.destroy_process_group()


CASES = {
    # everything near field: the sharded result must equal the single-process ORACLE
    "uj_estr_all_near": dict(n=333, op="uj", schemes=dict(sfs="constant"), far=None, truth="oracle"),
    # with a far field (the multipole channel carries data): equal to the same stand-in on ONE rank
    "uj_estr_far": dict(n=250, op="uj", schemes=dict(sfs="constant"), far=2, level=3, truth="one_rank"),
    "rk3_dynamic_sfs_far": dict(n=180, op="step", far=2, level=3, truth="one_rank",
                                schemes=dict(integration="rungekutta3", sfs="dynamic", force_positive=1, clippings=1)),
    "accumulate_far": dict(n=150, op="uj_twice", schemes=dict(), far=2, level=2, truth="one_rank"),
    # three ranks, one of them without home particles; 7 particles on 3 ranks (owners without particles)
    "empty_home_rank": dict(n=120, op="uj", schemes=dict(sfs="constant"), far=2, level=2, truth="one_rank", world=3,
                            bounds={3: [(0, 70), (70, 70), (70, 120)]}),
    "more_ranks_than_bins": dict(n=7, op="uj", schemes=dict(sfs="constant"), far=None, level=1, truth="oracle", world=3),
}


@pytest.mark.parametrize("name", sorted(CASES))
def test_let_orchestration_over_gloo(name, tmp_path):
    from oracle import oracle as o
    from tests.loopback import run_ranks
    case = CASES[name]
    world = case.get("world", 2)
    mp.spawn(_worker, args=(world, _free_port(), case, str(tmp_path)), nprocs=world, join=True)
    if case["truth"] == "oracle":
        ref = _field(case["n"])
        so = o.default_schemes(**case["schemes"])
        if case["op"] == "uj":
            o.field_uj(ref, so, reset=True, reset_sfs=True, sfs=True)
        else:
            t, nt = 0.0, 0
            for _ in range(2):
                t, nt = o.nextstep(ref, so, 2e-3, (1.0, -0.5, 0.25), relax=True, t=t, nt=nt)
    else:   # the same stand-in on ONE rank (in-process collectives of world size 1)
        ref = run_ranks(1, lambda rank, coll: _run(case, "let", 0, 1, coll)[0])[0]
    tol = 1e-9 if case["schemes"].get("sfs") == "dynamic" else 1e-11
    for mode in MODES:
        got = np.concatenate([np.load(tmp_path / f"shard_{mode}_{r}.npy") for r in range(world)])
        assert got.shape == ref.shape
        for g, sl in GROUPS.items():
            assert relmax(got[:, sl], ref[:, sl]) < tol, (mode, g)
        calls = [open(tmp_path / f"calls_{mode}_{r}.txt").read().split() for r in range(world)]
        for c in calls:
            if mode == "let_halo":
                assert "attach_tree" not in c and "attach_records" not in c
                assert "attach_skeleton" in c and "halo_plan" in c
            else:
                assert "attach_skeleton" not in c and "halo_plan" not in c and "attach_tree" in c
        if case["op"] == "step" and case["schemes"].get("sfs") == "dynamic":
            # DynamicSFS's second filter evaluation of the first substep reuses tree, lists and far field: records only
            assert any("build_reuse" in c for c in calls)
            if mode == "let_halo":
                assert any("halo_serve_rec" in c for c in calls)

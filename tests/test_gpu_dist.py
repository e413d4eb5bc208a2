"""2-GPU NCCL tests of the sharded field against the single-GPU engine (skipped when < 2 GPUs are visible): the direct path
(all-gather of source tiles) and UJ_fmm with the local essential tree (all-to-all of particle rows, all-gather of skeletons /
multipoles / records — or, "let_halo", of the skeletons only with multipoles / records on request — inverse all-to-all) and
with round 1's replicated tree.  The same LET logic runs on ONE GPU with
in-process collectives in tests/test_gpu_let.py."""
import os
import socket

import numpy as np
import pytest
import torch

from tests.util import mixed_field, relmax

pytestmark = pytest.mark.gpu


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, n, kw, out_dir, fmm_mode="let"):
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        import flowunsteady_b200 as fb
        from flowunsteady_b200.dist import ShardedField, partition
        x, g, s, static = mixed_field(n, seed=23)
        g = g * 50.0
        static = np.where(np.all(g == 0, axis=1), 1.0, static)
        P = fb.new_particles(x, g, s, static=static)
        lo, hi = partition(n, world)[rank]
        eng = fb.Engine(hi - lo + 8, device=rank, schemes=fb.default_schemes(**kw))
        eng.upload(P[lo:hi].copy())
        sf = ShardedField(eng, max_local=hi - lo + 8, device=f"cuda:{rank}", fmm=fmm_mode)
        for _ in range(2):
            sf.nextstep(2e-3, (1.0, -0.5, 0.25), relax=True)
        eng.synchronize()
        np.save(os.path.join(out_dir, f"shard{rank}.npy"), eng.download(np.zeros((hi - lo, 43))))
        eng.close()
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("kw,fmm_mode", [
    (dict(integration="rungekutta3"), "let"),
    (dict(integration="rungekutta3", sfs="dynamic", force_positive=1, clippings=1), "let"),
    (dict(integration="rungekutta3", uj="fmm", fmm_nonzero_sigma=1, sfs="constant", clippings=1), "let"),
    (dict(integration="rungekutta3", uj="fmm", sfs="dynamic", force_positive=1, clippings=1), "let"),
    (dict(integration="rungekutta3", uj="fmm", fmm_nonzero_sigma=1, sfs="constant", clippings=1), "let_halo"),
    (dict(integration="rungekutta3", uj="fmm", sfs="dynamic", force_positive=1, clippings=1), "let_halo"),
    (dict(integration="rungekutta3", uj="fmm", fmm_nonzero_sigma=1, sfs="constant", clippings=1), "replicated")])
def test_two_gpu_matches_one_gpu(kw, fmm_mode, tmp_path):
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    import torch.multiprocessing as mp

    import flowunsteady_b200 as fb
    n = 3001
    mp.spawn(_worker, args=(2, _free_port(), n, kw, str(tmp_path), fmm_mode), nprocs=2, join=True)
    x, g, s, static = mixed_field(n, seed=23)
    g = g * 50.0
    static = np.where(np.all(g == 0, axis=1), 1.0, static)
    P = fb.new_particles(x, g, s, static=static)
    with fb.Engine(n, schemes=fb.default_schemes(**kw)) as eng:
        eng.upload(P)
        for _ in range(2):
            eng.nextstep(2e-3, (1.0, -0.5, 0.25), relax=True)
        ref = eng.download(np.zeros_like(P))
    got = np.concatenate([np.load(tmp_path / f"shard{r}.npy") for r in range(2)])
    # sharded UJ_fmm builds the same tree on every rank and only adds zeros in the all-reduce: identical to one GPU
    tol = 1e-9 if kw.get("sfs") == "dynamic" else 1e-12
    for name, sl in dict(X=slice(0, 3), Gamma=slice(3, 6), sigma=slice(6, 7), U=slice(9, 12), J=slice(15, 24)).items():
        assert relmax(got[:, sl], ref[:, sl]) < tol, name

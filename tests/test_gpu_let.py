"""Multi-GPU UJ_fmm with a local essential tree (fmm_let.cuh + dist.py) against the single-GPU UJ_fmm.

`world` ranks run as threads on ONE GPU with in-process collectives (tests/loopback.py): the per-rank phases and the
orchestration are the product's.  The union of the ranks' trees is the one-GPU tree and every interaction is the same
interaction, so U, J and E_str must agree with one GPU to summation round-off: 1e-12 (max-norm relative), whatever the
expansion order or theta.  Edge cases: more ranks than occupied top cells (empty owners), a rank without home particles,
histogram levels 1..5, nonzero_sigma (global sigma_max of the top cells), dynamic-SFS far-field reuse.
"""
import numpy as np
import pytest

from tests.loopback import run_ranks
from tests.util import mixed_field, relmax

pytestmark = pytest.mark.gpu

GROUPS = dict(X=slice(0, 3), Gamma=slice(3, 6), sigma=slice(6, 7), U=slice(9, 12), J=slice(15, 24), C=slice(36, 39),
              SFS=slice(39, 42))


def _field(n, kind, seed=23):
    import flowunsteady_b200 as fb
    from flowunsteady_b200 import fields
    if kind == "cloud":
        x, g, s, static = mixed_field(n, seed=seed)
        g = g * 50.0
        static = np.where(np.all(g == 0, axis=1), 1.0, static)
        return fb.new_particles(x, g, s, static=static)
    if kind == "rings":
        x, g, s = fields.vortex_rings(n)
    else:
        x, g, s = fields.rotor_wake(n)
    return fb.new_particles(x, fields.floor_gamma(g), s)


def _single(P, kw, action):
    import flowunsteady_b200 as fb
    with fb.Engine(P.shape[0], schemes=fb.default_schemes(**kw)) as eng:
        eng.upload(P)
        action(eng)
        return eng.download(np.zeros_like(P))


def _sharded(P, kw, world, action, bounds=None, let_level=5, mode="let"):
    """`action(field_or_engine)` on a ShardedField of `world` engines sharing cuda:0."""
    import flowunsteady_b200 as fb
    from flowunsteady_b200.dist import ShardedField, partition
    n = P.shape[0]
    parts = bounds if bounds is not None else partition(n, world)

    def body(rank, coll):
        import torch
        torch.cuda.set_device(0)
        lo, hi = parts[rank]
        eng = fb.Engine(max(hi - lo, 1) + 8, device=0, schemes=fb.default_schemes(**kw))
        eng.upload(P[lo:hi].copy())
        sf = ShardedField(eng, max_local=max(hi - lo, 1) + 8, device="cuda:0", coll=coll, fmm=mode, let_level=let_level)
        action(sf)
        eng.synchronize()
        out = eng.download(np.zeros((hi - lo, 43)))
        eng.close()
        return out

    return np.concatenate(run_ranks(world, body))


@pytest.fixture(params=["let", "let_halo"])
def let_mode(request):
    """all-gathered essential tree / demand-driven halo (dist.py: ShardedField(fmm=...)): same tree, same lists, same results"""
    return request.param


def _assert_same(got, ref, names, tol):
    for name in names:
        err = relmax(got[:, GROUPS[name]], ref[:, GROUPS[name]])
        assert err < tol, f"{name}: {err:.3e} >= {tol}"


@pytest.mark.parametrize("world,kind,n", [(2, "cloud", 3001), (3, "cloud", 20000), (8, "rings", 60000), (4, "rotor", 40000)])
def test_let_uj_estr_matches_one_gpu(world, kind, n, let_mode):
    P = _field(n, kind)
    kw = dict(uj="fmm", sfs="constant")
    ref = _single(P, kw, lambda e: e.uj(True, True, True))
    got = _sharded(P, kw, world, lambda f: f.uj(True, True, True), mode=let_mode)
    _assert_same(got, ref, ["U", "J", "SFS"], 1e-12)


@pytest.mark.parametrize("let_level", [1, 2, 4])
def test_let_histogram_level_does_not_matter(let_level, let_mode):
    P = _field(12000, "cloud", seed=5)
    kw = dict(uj="fmm", fmm_p=3, fmm_theta=0.5, fmm_ncrit=20)
    ref = _single(P, kw, lambda e: e.uj())
    got = _sharded(P, kw, 4, lambda f: f.uj(), let_level=let_level, mode=let_mode)
    _assert_same(got, ref, ["U", "J"], 1e-12)


def test_let_more_ranks_than_particles_and_empty_home_rank(let_mode):
    """6 ranks, 40 particles (a single leaf: one owner, five empty owners), one rank holding no home particles at all."""
    P = _field(40, "cloud", seed=9)
    kw = dict(uj="fmm")
    ref = _single(P, kw, lambda e: e.uj())
    bounds = [(0, 10), (10, 10), (10, 25), (25, 30), (30, 39), (39, 40)]
    got = _sharded(P, kw, 6, lambda f: f.uj(), bounds=bounds, mode=let_mode)
    _assert_same(got, ref, ["U", "J"], 1e-12)


def test_let_nonzero_sigma_matches_one_gpu(let_mode):
    """nonzero_sigma = true: the acceptance uses every cell's largest core size; for the partial top cells that maximum is
    taken over ALL ranks (per-bin sigma max, all-reduced), so the lists equal the one-GPU lists."""
    P = _field(20000, "cloud", seed=31)
    P[::7, 6] *= 2.5                      # uneven cores so that sigma_max matters
    kw = dict(uj="fmm", fmm_nonzero_sigma=1)
    ref = _single(P, kw, lambda e: e.uj())
    got = _sharded(P, kw, 4, lambda f: f.uj(), mode=let_mode)
    _assert_same(got, ref, ["U", "J"], 1e-12)


@pytest.mark.parametrize("world", [2, 5])
def test_let_rk3_dynamic_sfs_steps_match_one_gpu(world, let_mode):
    """Two whole RK3 + DynamicSFS + pedrizzetti steps (far-field reuse between the two filter evaluations on both sides)."""
    P = _field(9000, "cloud", seed=12)
    kw = dict(uj="fmm", integration="rungekutta3", relaxation="pedrizzetti", sfs="dynamic", alpha=0.999, force_positive=1,
              clippings=1)

    def steps(f):
        for _ in range(2):
            f.nextstep(2e-3, (1.0, -0.5, 0.25), relax=True)

    ref = _single(P, kw, steps)
    got = _sharded(P, kw, world, steps, mode=let_mode)
    _assert_same(got, ref, ["X", "U", "J"], 1e-11)
    # through the dynamic procedure: 1e-12 x 1/(1 - alpha) = 1e-9 per evaluation, two steps (measured 1.1e-9 on C)
    _assert_same(got, ref, ["Gamma", "sigma", "C"], 5e-9)


def test_let_accumulate_flag(let_mode):
    """reset = false adds the new U, J to the rows the home rank holds."""
    P = _field(5000, "cloud", seed=2)
    kw = dict(uj="fmm")

    def twice(f):
        f.uj(True, False, False)
        f.uj(False, False, False)

    ref = _single(P, kw, twice)
    got = _sharded(P, kw, 3, twice, mode=let_mode)
    _assert_same(got, ref, ["U", "J"], 1e-12)


def test_let_halo_moves_less_than_the_all_gather_and_detaches_cleanly():
    """Demand-driven halo: (1) a rank receives only the multipoles / records its lists name — well under what the all-gather
    variant receives; (2) a plain one-GPU UJ_fmm on the same engine afterwards does not see the halo buffers."""
    import flowunsteady_b200 as fb
    from flowunsteady_b200.dist import ShardedField, partition
    n, world = 60000, 8
    P = _field(n, "rings")
    kw = dict(uj="fmm")
    parts = partition(n, world)

    def body(rank, coll):
        import torch
        torch.cuda.set_device(0)
        lo, hi = parts[rank]
        eng = fb.Engine(hi - lo + 8, device=0, schemes=fb.default_schemes(**kw))
        eng.upload(P[lo:hi].copy())
        sf = ShardedField(eng, max_local=hi - lo + 8, device="cuda:0", coll=coll, fmm="let_halo")
        sf.uj()
        eng.synchronize()
        L = sf._let
        nm3 = 3 * (eng.get_schemes().fmm_p * (eng.get_schemes().fmm_p + 1) * (eng.get_schemes().fmm_p + 2)) // 6
        full = 8 * (sum(L["np"]) - L["np"][rank]) * 10 + 8 * (sum(L["nc"]) - L["nc"][rank]) * nm3
        got = L["halo_bytes"]
        eng.uj()                                   # one-GPU evaluation of the home particles alone
        eng.synchronize()
        alone = eng.download(np.zeros((hi - lo, 43)))
        eng.close()
        return got, full, alone

    res = run_ranks(world, body)
    for rank, (got, full, alone) in enumerate(res):
        assert got < 0.8 * full, (rank, got, full)
        lo, hi = parts[rank]
        ref = _single(P[lo:hi].copy(), kw, lambda e: e.uj())
        _assert_same(alone, ref, ["U", "J"], 1e-12)

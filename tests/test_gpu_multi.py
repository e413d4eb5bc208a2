"""Several GPUs behind ONE handle and ONE host thread (include/vpmb200.h: vpmb200_multi_*; flowunsteady_b200/csrc/multi.inl),
driven through the header alone (ctypes), against the single-GPU engine and the reference's index semantics.

The shards map to the GPUs that are visible: on the one-GPU box of the round-end test run all shards share cuda:0 (peer copies
become device-to-device copies, the logic is the same); on a multi-GPU box they spread over the devices.
"""
import numpy as np
import pytest

from tests.util import mixed_field, relmax

pytestmark = pytest.mark.gpu

DYN = dict(integration="rungekutta3", relaxation="pedrizzetti", sfs="dynamic", alpha=0.999, force_positive=1, clippings=1)


def _devices(k):
    import torch
    nd = max(torch.cuda.device_count(), 1)
    return [i % nd for i in range(k)]


def _field(n, seed=23):
    import flowunsteady_b200 as fb
    x, g, s, static = mixed_field(n, seed=seed)
    g = g * 50.0
    static = np.where(np.all(g == 0, axis=1), 1.0, static)
    return fb.new_particles(x, g, s, static=static)


@pytest.mark.parametrize("ngpus", [1, 2, 3, 8])
@pytest.mark.parametrize("kw", [dict(integration="rungekutta3"), DYN, dict(integration="euler", sfs="constant", clippings=1)])
def test_multi_steps_match_one_gpu(ngpus, kw):
    import flowunsteady_b200 as fb
    P = _field(3001)
    with fb.Engine(P.shape[0], schemes=fb.default_schemes(**kw)) as eng:
        eng.upload(P)
        for _ in range(2):
            eng.nextstep(2e-3, (1.0, -0.5, 0.25), relax=True)
        ref = eng.download(np.zeros_like(P))
    with fb.MultiEngine(P.shape[0], ngpus, _devices(ngpus), schemes=fb.default_schemes(**kw)) as me:
        me.upload(P)
        assert sum(me.shard_sizes()) == P.shape[0] and max(me.shard_sizes()) - min(me.shard_sizes()) <= 1
        for _ in range(2):
            me.nextstep(2e-3, (1.0, -0.5, 0.25), relax=True)
        got = me.download(np.zeros_like(P))
        assert me.get_time() == (pytest.approx(4e-3), 2)
    tol = 5e-9 if kw.get("sfs") == "dynamic" else 1e-12
    for sl in (slice(0, 3), slice(3, 6), slice(6, 7), slice(9, 12), slice(15, 24)):
        assert relmax(got[:, sl], ref[:, sl]) < tol
    assert np.array_equal(got[:, 42], ref[:, 42])


@pytest.mark.parametrize("ngpus", [2, 5])
@pytest.mark.parametrize("kw", [dict(uj="fmm", sfs="constant", clippings=1), dict(uj="fmm", fmm_nonzero_sigma=1),
                                dict(uj="fmm", integration="rungekutta3", relaxation="pedrizzetti", sfs="dynamic", alpha=0.999,
                                     force_positive=1, clippings=1)])
def test_multi_fmm_local_essential_tree_matches_one_gpu(ngpus, kw):
    """vpm_UJ = UJ_fmm on the multi handle: the local-essential-tree phases with peer copies between them (multi.inl:
    multi_uj_fmm) against ONE engine — U, J, E_str of an evaluation to 1e-12, two whole steps to the dynamic procedure's
    budget."""
    import flowunsteady_b200 as fb
    P = _field(9000, seed=12)
    dyn = kw.get("sfs") == "dynamic"
    outs = []
    for make in (lambda: fb.Engine(P.shape[0], schemes=fb.default_schemes(**kw)),
                 lambda: fb.MultiEngine(P.shape[0], ngpus, _devices(ngpus), schemes=fb.default_schemes(**kw))):
        with make() as e:
            e.upload(P)
            if dyn:
                for _ in range(2):
                    e.nextstep(2e-3, (1.0, -0.5, 0.25), relax=True)
            else:
                e.uj(True, True, True)
                e.uj(False, False, False)          # accumulate on top
            outs.append(e.download(np.zeros_like(P)))
    ref, got = outs
    for name, sl, tol in (("X", slice(0, 3), 1e-11), ("U", slice(9, 12), 1e-11 if dyn else 1e-12), ("J", slice(15, 24), 1e-11 if dyn else 1e-12),
                          ("SFS", slice(39, 42), 1e-9 if dyn else 1e-11), ("Gamma", slice(3, 6), 5e-9), ("sigma", slice(6, 7), 5e-9)):
        assert relmax(got[:, sl], ref[:, sl]) < tol, name


def test_multi_mutation_follows_the_reference_order():
    """add_particle / remove_particle / wake treatment on a sharded field leave the host-visible order exactly as the same
    calls on ONE engine do (the reference's swap-with-last and removal-loop semantics), whatever shard a particle lives on;
    appended particles go to the least-loaded shard; rebalance moves particles without changing the order."""
    import flowunsteady_b200 as fb
    P = _field(2500, seed=4)
    extra = _field(700, seed=5)
    extra[:, 0:3] += 0.3
    rng = np.random.default_rng(0)
    with fb.Engine(4000, schemes=fb.default_schemes()) as one, fb.MultiEngine(4000, 4, _devices(4), schemes=fb.default_schemes()) as me:
        for e in (one, me):
            e.upload(P)
        for e in (one, me):
            for i in (17, 0, 2400, 1234):                       # swap-with-last across shard boundaries
                e.remove_particle(i)
            e.remove_particle(e.np - 1)
        for e in (one, me):
            e.add_particles(extra[:300])
        sizes = me.shard_sizes()
        assert sum(sizes) == one.np == 2795
        assert max(sizes) - min(sizes) <= 300                    # the 300 newcomers went to the least-loaded shard(s)
        for e in (one, me):
            removed = e.remove_where(fb.Engine.REMOVE_SPHERE, [0.45 ** 2, 0.5, 0.5, 0.5])
            assert removed > 500
        for e in (one, me):
            e.add_particles(extra[300:])
        assert one.np == me.np
        for i in rng.integers(0, one.np - 60, 40):               # the same indices on both (np shrinks by one per removal)
            for e in (one, me):
                e.remove_particle(int(i))
        a = one.download(np.zeros((one.np, 43)))
        b = me.download(np.zeros((me.np, 43)))
        assert a.shape == b.shape and np.array_equal(a, b)
        before = me.shard_sizes()
        moved = me.rebalance(0.02)
        after = me.shard_sizes()
        assert moved > 0 and max(after) - min(after) <= max(0.02 * me.np / 4, 1) + 1 and max(before) - min(before) > max(after) - min(after)
        assert np.array_equal(me.download(np.zeros((me.np, 43))), a)
        # and the field still steps like the single engine after all that
        for e in (one, me):
            e.nextstep(1e-3, (1.0, 0.0, 0.0), relax=True)
        a2, b2 = one.download(np.zeros((one.np, 43))), me.download(np.zeros((me.np, 43)))
        for sl in (slice(0, 3), slice(3, 6), slice(6, 7), slice(9, 12), slice(15, 24)):
            assert relmax(b2[:, sl], a2[:, sl]) < 1e-12
        Xp = rng.random((25, 3))
        Ua, Ja = one.uj_probe(Xp, want_J=True)
        Ub, Jb = me.uj_probe(Xp, want_J=True)
        assert relmax(Ub, Ua) < 1e-12 and relmax(Jb, Ja) < 1e-12


def test_multi_simulation_loop_order():
    """The order of calls of FLOWUnsteady's loop (simulation.jl:339-447): statics appended -> nextstep -> statics removed from
    the end -> shedding -> probes -> wake treatment, on 3 shards vs one engine."""
    import flowunsteady_b200 as fb
    rng = np.random.default_rng(7)
    kw = dict(integration="rungekutta3", relaxation="pedrizzetti")

    def cols(X, G, s, static):
        c = np.zeros((X.shape[0], 43))
        c[:, 0:3], c[:, 3:6], c[:, 6], c[:, 8], c[:, 42] = X, G, s, np.linalg.norm(G, axis=1), static
        return c

    script = [dict(statics=cols(rng.random((12, 3)) * 0.2 + [0, 0.4, 0.4], rng.standard_normal((12, 3)) * 0.02, np.full(12, 0.06), 1.0),
                   shed=cols(rng.random((60, 3)) * [0.1, 1.0, 0.2] + [0.1 * k, 0.0, 0.4], rng.standard_normal((60, 3)) * 0.05, np.full(60, 0.08), 0.0),
                   probes=rng.random((9, 3))) for k in range(4)]
    outs = []
    for make, fast in ((lambda: fb.Engine(1000, schemes=fb.default_schemes(**kw)), False),
                       (lambda: fb.MultiEngine(1000, 3, _devices(3), schemes=fb.default_schemes(**kw)), False),
                       (lambda: fb.MultiEngine(1000, 3, _devices(3), schemes=fb.default_schemes(**kw)), True)):
        V = []
        with make() as e:
            for k, st in enumerate(script):
                org = e.np
                if k > 0 and fast:                       # static-particle fast path on the sharded field
                    e.set_statics(st["statics"])
                    assert e.np == org and sum(e.shard_sizes()) == org + 12
                    e.nextstep(0.02, (1.0, 0.0, 0.1), relax=True)
                    assert e.np == org and sum(e.shard_sizes()) == org
                elif k > 0:
                    e.add_particles(st["statics"])
                    e.nextstep(0.02, (1.0, 0.0, 0.1), relax=True)
                    for i in range(e.np - 1, org - 1, -1):
                        e.remove_particle(i)
                e.add_particles(st["shed"])
                V.append(e.uj_probe(st["probes"]))
                e.remove_where(fb.Engine.REMOVE_SPHERE, [1.2 ** 2, 0.3, 0.5, 0.5])
            outs.append((e.download(np.zeros((e.np, 43))), np.array(V), e.get_time()))
    (a, Va, ta) = outs[0]
    for (b, Vb, tb) in outs[1:]:
        assert a.shape == b.shape and ta == tb
        assert relmax(Vb, Va) < 1e-12
        for sl in (slice(0, 3), slice(3, 6), slice(6, 7), slice(9, 12), slice(15, 24)):
            assert relmax(b[:, sl], a[:, sl]) < 1e-12
        assert np.array_equal(a[:, 42], b[:, 42])


def test_multi_errors():
    import flowunsteady_b200 as fb
    with fb.MultiEngine(100, 2, _devices(2)) as me:
        with pytest.raises(fb.EngineError) as ei:
            me.upload(np.zeros((101, 43)))
        assert ei.value.code == -4
        me.upload(_field(64))
        with pytest.raises(fb.EngineError) as ei:
            me.set_schemes(fb.default_schemes(viscous="corespreading", nu=1e-5, cs_sgm0=0.1))
        assert ei.value.code == -5                                  # VPMB200_ENOTSUP, stated loudly (RBF re-fit is single-GPU)
        with pytest.raises(fb.EngineError):
            me.remove_particle(64)
    with pytest.raises(fb.EngineError):
        fb.MultiEngine(100, 0)

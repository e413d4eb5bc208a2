"""The drop-in boundary exercised in FLOWUnsteady's call order (/root/reference/src/FLOWUnsteady_simulation.jl:339-447):
static particles appended -> vpm.nextstep -> statics removed from the end -> wake shed -> probes evaluated by appending
zero-strength particles, calling pfield.UJ(pfield) and reading get_U (Vvpm_on_Xs, :494-570) -> wake treatment.
The same loop is replayed on a plain numpy matrix with the CPU oracle; the particle matrices must agree."""
import types

import numpy as np
import pytest

from tests.util import relmax

pytestmark = pytest.mark.gpu


def _script(rng, nsteps=3, nshed=40, nstatic=12, nprobe=9):
    """Deterministic inputs of a run: per step the static (bound) particles, the shed particles and the probe points."""
    steps = []
    for k in range(nsteps):
        steps.append(dict(
            statics=(rng.random((nstatic, 3)) * 0.2 + [0.0, 0.4, 0.4], rng.standard_normal((nstatic, 3)) * 0.02, np.full(nstatic, 0.06)),
            shed=(rng.random((nshed, 3)) * [0.1, 1.0, 0.2] + [0.1 * k, 0.0, 0.4], rng.standard_normal((nshed, 3)) * 0.05,
                  np.full(nshed, 0.08)),
            probes=rng.random((nprobe, 3))))
    return steps


def _static_cols(X, G, s):
    c = np.zeros((X.shape[0], 43))
    c[:, 0:3], c[:, 3:6], c[:, 6], c[:, 8], c[:, 42] = X, G, s, np.linalg.norm(G, axis=1), 1.0
    return c


def _run_engine(steps, dt, Uinf, sfs, integration, fast_statics=False, sync="always"):
    from flowunsteady_b200 import vpm, wake
    pf = vpm.ParticleField(2000, Uinf=lambda t: Uinf, UJ=vpm.UJ_direct, SFS=sfs, integration=integration,
                           relaxation=vpm.pedrizzetti, sync=sync)
    sim = types.SimpleNamespace(nt=0, vehicle=types.SimpleNamespace(system=types.SimpleNamespace(O=np.zeros(3))))
    treatment = wake.remove_particles_sphere(1.2 ** 2, 1, Xoff=[0.3, 0.5, 0.5])
    V = []
    for k, st in enumerate(steps):
        org_np = vpm.get_np(pf)
        if k > 0 and fast_statics:
            pf.set_statics(_static_cols(*st["statics"]))               # parked behind the field, consumed by nextstep
            vpm.nextstep(pf, dt, relax=True)
            assert vpm.get_np(pf) == org_np and pf.engine.get_statics()[0] == 0
        elif k > 0:
            for X, G, s in zip(*st["statics"]):
                vpm.add_particle(pf, X, G, s, vol=0, circulation=np.linalg.norm(G), static=True)
            vpm.nextstep(pf, dt, relax=True)
            for i in range(vpm.get_np(pf) - 1, org_np - 1, -1):        # simulation.jl:361-365
                vpm.remove_particle(pf, i)
        for X, G, s in zip(*st["shed"]):
            vpm.add_particle(pf, X, G, s, vol=1e-3, circulation=np.linalg.norm(G))
        # Vvpm_on_Xs the reference way: probes are particles
        sta_np = vpm.get_np(pf)
        for X in st["probes"]:
            vpm.add_particle(pf, X, np.zeros(3), 1e-6, vol=0)
        pf.UJ(pf)
        if sync == "lazy":
            pf.pull()                                                  # lazy mode: results stay on the device until asked for
        Vref = np.array([vpm.get_U(P).copy() for P in vpm.iterator(pf, start_i=sta_np, include_static=True)])
        for i in range(vpm.get_np(pf) - 1, sta_np - 1, -1):
            vpm.remove_particle(pf, i)
        Vfast = pf.U_at(st["probes"])                                  # the probe fast path must give the same numbers
        assert relmax(Vfast, Vref) < 1e-13
        V.append(Vref)
        sim.nt = k
        treatment(sim, pf, pf.t, dt)
    if sync == "lazy":
        pf.pull()
    return pf.particles[:pf.np].copy(), np.array(V), pf.t, pf.nt


def _run_oracle(steps, dt, Uinf, kw):
    from oracle import oracle as o
    sch = o.default_schemes(**kw)
    P = np.zeros((0, 43))
    t, nt = 0.0, 0
    V = []

    def cols(X, G, s, vol, circ, static):
        c = np.zeros((X.shape[0], 43))
        c[:, 0:3], c[:, 3:6], c[:, 6], c[:, 7], c[:, 8], c[:, 42] = X, G, s, vol, circ, static
        return c

    for k, st in enumerate(steps):
        org_np = P.shape[0]
        if k > 0:
            X, G, s = st["statics"]
            P = np.ascontiguousarray(np.concatenate([P, cols(X, G, s, 0.0, np.linalg.norm(G, axis=1), 1.0)]))
            t, nt = o.nextstep(P, sch, dt, Uinf, relax=True, t=t, nt=nt)
            P = P[:org_np]
        X, G, s = st["shed"]
        P = np.ascontiguousarray(np.concatenate([P, cols(X, G, s, 1e-3, np.linalg.norm(G, axis=1), 0.0)]))
        U, _ = o.uj_direct(sch.kernel, P[:, 0:3], P[:, 3:6], P[:, 6], st["probes"], accum=1)
        V.append(U)
        # pfield.UJ(pfield) also refreshed U, J of the field itself (probes contribute nothing: Gamma = 0)
        Pq = np.ascontiguousarray(np.concatenate([P, cols(st["probes"], np.zeros_like(st["probes"]), np.full(len(st["probes"]), 1e-6), 0, 1.0, 0)]))
        o.field_uj(Pq, sch)
        P = np.ascontiguousarray(Pq[:P.shape[0]])
        # remove_particles_sphere, replayed literally
        n = P.shape[0]
        c = np.array([0.3, 0.5, 0.5])
        for i in range(n - 1, -1, -1):
            if ((P[i, 0:3] - c) ** 2).sum() > 1.2 ** 2:
                if i != n - 1:
                    P[i] = P[n - 1]
                n -= 1
        P = np.ascontiguousarray(P[:n])
    return P, np.array(V), t, nt


@pytest.mark.parametrize("variant", ["rk3_nosfs", "euler_dynamic"])
def test_simulation_loop_matches_oracle(variant):
    from flowunsteady_b200 import vpm
    steps = _script(np.random.default_rng(11))
    dt, Uinf = 0.02, (1.0, 0.0, 0.1)
    if variant == "rk3_nosfs":
        got = _run_engine(steps, dt, Uinf, vpm.SFS_none, vpm.rungekutta3)
        exp = _run_oracle(steps, dt, Uinf, dict(integration="rungekutta3"))
        tol = 1e-11
    else:
        got = _run_engine(steps, dt, Uinf, vpm.SFS_Cd_twolevel_nobackscatter, vpm.euler)
        exp = _run_oracle(steps, dt, Uinf, dict(integration="euler", sfs="dynamic", alpha=0.999, force_positive=1, clippings=1))
        tol = 1e-8
    Pg, Vg, tg, ntg = got
    Po, Vo, to, nto = exp
    assert Pg.shape == Po.shape and (tg, ntg) == (pytest.approx(to), nto)
    assert relmax(Vg, Vo) < 1e-11
    for sl in (slice(0, 3), slice(3, 6), slice(6, 7), slice(9, 12), slice(15, 24)):
        assert relmax(Pg[:, sl], Po[:, sl]) < tol
    assert np.allclose(Pg[:, 7:9], Po[:, 7:9], rtol=1e-14) and np.array_equal(Pg[:, 42], Po[:, 42])


@pytest.mark.parametrize("sync", ["always", "lazy"])
def test_static_particle_fast_path_matches_add_nextstep_remove(sync):
    """vpmb200_set_statics (the static set parked device-side, consumed by nextstep) against the reference's own sequence
    add_particle x n -> nextstep -> remove_particle x n (simulation.jl:355-365) and against the oracle replay of that loop."""
    from flowunsteady_b200 import vpm
    steps = _script(np.random.default_rng(11))
    dt, Uinf = 0.02, (1.0, 0.0, 0.1)
    slow = _run_engine(steps, dt, Uinf, vpm.SFS_none, vpm.rungekutta3)
    fast = _run_engine(steps, dt, Uinf, vpm.SFS_none, vpm.rungekutta3, fast_statics=True, sync=sync)
    assert fast[0].shape == slow[0].shape and fast[2:] == slow[2:]
    for sl in (slice(0, 9), slice(9, 12), slice(15, 24), slice(42, 43)):
        assert np.array_equal(fast[0][:, sl], slow[0][:, sl])           # same kernels, same order of sources: same bits
    assert np.array_equal(fast[1], slow[1])
    exp = _run_oracle(steps, dt, Uinf, dict(integration="rungekutta3"))
    for sl in (slice(0, 3), slice(3, 6), slice(6, 7), slice(9, 12), slice(15, 24)):
        assert relmax(fast[0][:, sl], exp[0][:, sl]) < 1e-11


def test_probe_statics_fsgm_quirk_and_mirror_vs_oracle_replay():
    """Vvpm_on_Xs(pfield, Xs; static_particles_fun, fsgm, mirror) (simulation.jl:494-570) through vpmb200_uj_probe_ex:
    the reference's sequence replayed literally on a numpy matrix with the oracle's UJ — including the single-subscript
    `particles[SIGMA_INDEX] *= fsgm` that scales the FIRST particle's core once per static particle in the field."""
    import flowunsteady_b200 as fb
    from oracle import oracle as o
    from tests.util import mixed_field
    x, g, s, static = mixed_field(700, seed=8)
    g = g * 30 + 1e-9
    static[:] = 0
    static[[5, 90, 400]] = 1                                            # three statics IN the field: sigma[0] *= fsgm^3
    P = fb.new_particles(x, g, s, static=static)
    rng = np.random.default_rng(3)
    statics = _static_cols(rng.random((20, 3)), rng.standard_normal((20, 3)) * 0.01, np.full(20, 0.07))
    probes = rng.random((33, 3))
    X0, nrm, fsgm = np.array([0.5, 0.5, -0.2]), np.array([0.0, 0.6, 0.8]), 5.5

    def images(Q):                                                      # vehicle_vlm_unsteady.jl:248-258, simulation.jl:520-533
        out = np.zeros_like(Q)
        for i in range(Q.shape[0]):
            X, G = Q[i, 0:3], Q[i, 3:6]
            a = 2 * (X - X0)
            d = a[0] * nrm[0] + a[1] * nrm[1] + a[2] * nrm[2]
            gn = G[0] * nrm[0] + G[1] * nrm[1] + G[2] * nrm[2]
            out[i, 0:3] = X - d * nrm
            out[i, 3:6] = 2 * gn * G / np.sqrt(G[0] * G[0] + G[1] * G[1] + G[2] * G[2]) - G
            out[i, 6:9], out[i, 36:39], out[i, 42] = Q[i, 6:9], Q[i, 36:39], 1.0
        return out

    for static_mirror in (False, True):
        for probe_mirror in (False, True):
            Q = P.copy()
            k = int((Q[:, 42] > 0).sum())
            for _ in range(k):
                Q[0, 6] *= fsgm
            S = np.concatenate([Q, statics])
            if static_mirror:
                S = np.concatenate([S, images(S)])
            if probe_mirror:
                S = np.concatenate([S, images(S)])
            Uo, Jo = o.uj_direct("gaussianerf", S[:, 0:3], S[:, 3:6], S[:, 6], probes, accum=1)
            with fb.Engine(4 * (P.shape[0] + 20) + 8, schemes=fb.default_schemes()) as eng:
                eng.upload(P)
                eng.set_mirror(static_mirror, X0, nrm)
                eng.set_statics(statics)
                Ug, Jg = eng.uj_probe_ex(probes, fsgm=fsgm, mirror=probe_mirror, want_J=True)
                after = eng.download(np.zeros_like(P))
                assert eng.np == P.shape[0] and eng.get_statics()[0] == 20        # probes do not consume the set
            assert relmax(Ug, Uo) < 1e-12 and relmax(Jg, Jo) < 1e-12
            s0 = P[0, 6]
            for _ in range(k):
                s0 *= fsgm
            for _ in range(k):
                s0 /= fsgm
            assert after[0, 6] == s0 and np.array_equal(after[1:, 6], P[1:, 6])    # the reference's round trip, bit for bit


def test_stale_static_set_is_ignored_and_mutation_drops_it():
    import flowunsteady_b200 as fb
    from tests.util import mixed_field
    x, g, s, _ = mixed_field(300, seed=1)
    P = fb.new_particles(x, g * 30 + 1e-9, s)
    statics = _static_cols(x[:10] + 0.01, g[:10], s[:10])
    with fb.Engine(400, schemes=fb.default_schemes()) as eng:
        eng.upload(P)
        eng.uj()
        base = eng.download(np.zeros_like(P))
        eng.set_statics(statics, generation=7)                                    # belongs to another step: never used
        eng.uj()
        assert np.array_equal(eng.download(np.zeros_like(P)), base) and eng.get_statics()[0] == 0
        eng.set_statics(statics)
        eng.uj()
        assert not np.array_equal(eng.download(np.zeros_like(P))[:, 9:12], base[:, 9:12])
        eng.add_particles(P[:1])                                                   # the columns behind np change hands
        assert eng.get_statics() == (0, -1)
        with pytest.raises(fb.EngineError):
            eng.set_statics(np.zeros((200, 43)))                                   # capacity


def test_lazy_sync_keeps_field_on_device():
    """sync='lazy': no host traffic between calls until pull(); results equal the always-synchronised field."""
    import flowunsteady_b200 as fb
    from flowunsteady_b200 import vpm
    from tests.util import mixed_field
    x, g, s, static = mixed_field(800, seed=2)
    static = np.where(np.all(g == 0, axis=1), 1.0, static)
    P = fb.new_particles(x, 50 * g, s, static=static)
    outs = []
    for mode in ("always", "lazy"):
        pf = vpm.ParticleField(800, UJ=vpm.UJ_direct, sync=mode)
        pf.particles[:800] = P
        pf.np = 800
        for _ in range(3):
            vpm.nextstep(pf, 1e-3, relax=True)
        if mode == "lazy":
            assert pf.d2h_bytes == 0 and pf.h2d_bytes == 800 * 8 * 22   # ONE upload (state + M rows), nothing downloaded yet
            pf.pull()
        outs.append(pf.particles[:800].copy())
    assert np.array_equal(outs[0], outs[1])


def test_lazy_sync_with_shedding_and_removal_between_steps():
    """sync='lazy' + vpm.add_particle / remove_particle / _reset_particles between steps.  The appended
    rows alone cross the bus (vpmb200_add_particles); the device-newer X / Gamma / sigma of the previous step must survive
    (they used to be overwritten by the stale host copy).  Same bits as sync='always'."""
    import flowunsteady_b200 as fb
    from flowunsteady_b200 import vpm
    from tests.util import mixed_field
    x, g, s, _ = mixed_field(900, seed=4)
    g = 50 * g + 1e-12
    outs, traffic = [], []
    for mode in ("always", "lazy"):
        pf = vpm.ParticleField(1000, UJ=vpm.UJ_direct, SFS=vpm.SFS_Cd_twolevel_nobackscatter, sync=mode)
        for i in range(600):
            vpm.add_particle(pf, x[i], g[i], s[i], vol=1e-3, circulation=1.0)
        for k in range(3):
            vpm.nextstep(pf, 2e-3, relax=True)
            for i in range(600 + 100 * k, 700 + 100 * k):         # shed 100 particles (simulation.jl:363)
                vpm.add_particle(pf, x[i], g[i], s[i], vol=1e-3, circulation=1.0)
            vpm.remove_particle(pf, 17 + k)                       # a wake treatment removing one particle
            vpm.remove_particle(pf, pf.np - 1)
            if k == 1:
                vpm._reset_particles(pf)
        vpm.nextstep(pf, 2e-3, relax=True)
        if mode == "lazy":
            # one full upload of the first 600 rows, then only the appended rows
            assert pf.h2d_bytes == 600 * 8 * 22 + 3 * 100 * 8 * 43
            pf.pull()
        outs.append(pf.particles[:pf.np].copy())
        traffic.append(pf.h2d_bytes)
    assert outs[0].shape == outs[1].shape == (894, 43)      # 600 + 3 x 100 shed - 3 x 2 removed
    assert np.array_equal(outs[0], outs[1])
    assert traffic[1] < traffic[0] / 2


def test_page_locked_host_matrix_gives_the_same_field():
    """pinned=True page-locks `pfield.particles` in place through the ABI (vpmb200_host_register — what the Julia stub
    does with the reference's matrix): same bits as the pageable matrix, registering twice is harmless, and the lock is
    released with the field."""
    import flowunsteady_b200 as fb
    from flowunsteady_b200 import _lib, vpm
    from tests.util import mixed_field
    x, g, s, static = mixed_field(1500, seed=6)
    P = fb.new_particles(x, 50 * g + 1e-12, s)
    outs = []
    for pinned in (False, True):
        pf = vpm.ParticleField(2000, UJ=vpm.UJ_fmm, SFS=vpm.SFS_Cd_twolevel_nobackscatter, pinned=pinned)
        pf.particles[:1500] = P
        pf.np = 1500
        for _ in range(2):
            vpm.nextstep(pf, 1e-3, relax=True)
        if pinned:
            L = _lib.lib()
            assert L.vpmb200_host_register(pf.particles.ctypes.data, pf.particles.nbytes) == 0     # already registered: OK
        outs.append(pf.particles[:1500].copy())
        del pf
    assert np.all(np.isfinite(outs[0])) and np.array_equal(outs[0], outs[1])
    L = _lib.lib()
    buf = np.zeros(4096)
    assert L.vpmb200_host_register(buf.ctypes.data, buf.nbytes) == 0
    assert L.vpmb200_host_unregister(buf.ctypes.data) == 0
    assert L.vpmb200_host_unregister(buf.ctypes.data) == 0                                           # not registered: no-op
    assert L.vpmb200_host_register(None, 16) != 0

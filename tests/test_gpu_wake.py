"""GPU tests of the device-side wake treatments and monitors against a line-by-line Python replay of the reference's
loops (/root/reference/src/FLOWUnsteady_processing.jl:50-187): same survivors AND same resulting order."""
import types

import numpy as np
import pytest

from tests.util import mixed_field

pytestmark = pytest.mark.gpu


def _replay(P, keep_fn):
    """for i in np:-1:1; if !keep(P_i) remove_particle(i) end — remove_particle moves the last particle into slot i."""
    P = P.copy()
    n = P.shape[0]
    for i in range(n - 1, -1, -1):
        if not keep_fn(P[i]):
            if i != n - 1:
                P[i] = P[n - 1]
            n -= 1
    return P[:n]


def _pfield(P):
    from flowunsteady_b200 import vpm
    pf = vpm.ParticleField(P.shape[0] + 10, UJ=vpm.UJ_direct)
    pf.particles[:P.shape[0]] = P
    pf.np = P.shape[0]
    return pf


SIM = types.SimpleNamespace(nt=4, vehicle=types.SimpleNamespace(system=types.SimpleNamespace(O=np.array([0.2, 0.1, -0.1]))))


@pytest.mark.parametrize("n", [1, 7, 1000, 20_000])
def test_wake_treatments_match_reference_loop(n):
    import flowunsteady_b200 as fb
    from flowunsteady_b200 import wake
    x, g, s, static = mixed_field(n, seed=n)
    P = fb.new_particles(x, g, s, static=static)
    P[:, 36] = np.arange(n)                                    # tag to check the order
    g2 = (P[:, 3:6] ** 2).sum(1)
    lo, hi = np.quantile(g2, 0.2) if n > 1 else 0.0, np.quantile(g2, 0.9) if n > 1 else 1.0
    cases = [
        (wake.remove_particles_strength(lo, hi), lambda p: lo <= p[3] ** 2 + p[4] ** 2 + p[5] ** 2 <= hi),
        (wake.remove_particles_lowstrength(lo, 2), lambda p: lo <= p[3] ** 2 + p[4] ** 2 + p[5] ** 2 <= np.inf),
        (wake.remove_particles_sigma(np.quantile(s, 0.3), np.quantile(s, 0.8)),
         lambda p: np.quantile(s, 0.3) <= p[6] <= np.quantile(s, 0.8)),
        (wake.remove_particles_box([-0.1, 0.0, 0.2], [0.6, 0.7, 0.9], 4),
         lambda p: not ((p[0] - 0.2 < -0.1 or p[0] - 0.2 > 0.6) or (p[1] - 0.1 < 0.0 or p[1] - 0.1 > 0.7)
                        or (p[2] + 0.1 < 0.2 or p[2] + 0.1 > 0.9))),
        (wake.remove_particles_sphere(0.16, 1, Xoff=[0.3, 0.3, 0.6]),
         lambda p: not ((p[0] - 0.5) ** 2 + (p[1] - 0.4) ** 2 + (p[2] - 0.5) ** 2 > 0.16)),
    ]
    for treatment, keep in cases:
        pf = _pfield(P)
        assert treatment(SIM, pf, 0.0, 0.1) is False
        expect = _replay(P, keep)
        assert pf.np == expect.shape[0]
        assert np.array_equal(pf.particles[:pf.np], expect)


def test_every_nsteps_gate():
    import flowunsteady_b200 as fb
    from flowunsteady_b200 import wake
    x, g, s, _ = mixed_field(100, seed=1)
    pf = _pfield(fb.new_particles(x, g, s))
    sim = types.SimpleNamespace(nt=3, vehicle=SIM.vehicle)
    wake.remove_particles_sigma(1e9, 2e9, every_nsteps=2)(sim, pf, 0, 0)       # 3 % 2 != 0: nothing happens
    assert pf.np == 100
    sim.nt = 4
    wake.remove_particles_sigma(1e9, 2e9, every_nsteps=2)(sim, pf, 0, 0)       # removes everything
    assert pf.np == 0


def test_monitors_match_numpy():
    import flowunsteady_b200 as fb
    from flowunsteady_b200 import vpm
    x, g, s, static = mixed_field(5000, seed=3)
    P = fb.new_particles(x, g, s, static=static)
    P[:, 36] = np.where(np.arange(5000) % 3 == 0, 0.0, np.random.default_rng(0).random(5000))
    pf = _pfield(P)
    pf.UJ(pf)
    m = pf.monitors()
    assert m["enstrophy"] == pytest.approx(vpm.monitor_enstrophy_value(pf), rel=1e-12)
    C = pf.particles[:pf.np, 36]
    nz = C[C != 0]
    assert m["Cd_count"] == nz.size and m["Cd_mean"] == pytest.approx(nz.mean(), rel=1e-12)
    assert m["Cd_std"] == pytest.approx(nz.std(ddof=1), rel=1e-10)
    assert m["n_static"] == (P[:, 42] > 0).sum()


def test_fluiddomain_probes():
    import flowunsteady_b200 as fb
    from oracle import oracle as o
    x, g, s, _ = mixed_field(3000, seed=6)
    pf = _pfield(fb.new_particles(x, g, s))
    nodes = np.random.default_rng(1).random((500, 3))
    U, W = pf.fluiddomain(nodes)
    Uo, Jo = o.uj_direct("gaussianerf", x, g, s, nodes, accum=1)
    Wo = np.stack([Jo[:, 5] - Jo[:, 7], Jo[:, 6] - Jo[:, 2], Jo[:, 1] - Jo[:, 3]], -1)
    assert np.abs(U - Uo).max() < 1e-12 * np.abs(Uo).max()
    assert np.abs(W - Wo).max() < 1e-12 * np.abs(Wo).max()


def test_fluiddomain_through_fmm():
    """Grid nodes evaluated the reference's way (zero-strength probe particles + UJ_fmm) agree with the direct probes to
    the FMM's truncation error."""
    import flowunsteady_b200 as fb
    from flowunsteady_b200 import fields, vpm
    x, g, s = fields.vortex_rings(40_000)
    pf = vpm.ParticleField(x.shape[0], UJ=vpm.UJ_fmm, fmm=vpm.FMM(p=4, ncrit=50, theta=0.4, nonzero_sigma=True))
    pf.particles[:] = fb.new_particles(x, g, s)
    pf.np = x.shape[0]
    gx = np.linspace(-1.5, 1.5, 24)
    nodes = np.stack(np.meshgrid(gx, gx, np.linspace(-0.5, 1.5, 16), indexing="ij"), -1).reshape(-1, 3)
    Ud, Wd = pf.fluiddomain(nodes, method="direct")
    Uf, Wf = pf.fluiddomain(nodes, method="fmm")
    assert np.linalg.norm(Uf - Ud) / np.linalg.norm(Ud) < 2e-3
    assert np.linalg.norm(Wf - Wd) / np.linalg.norm(Wd) < 2e-2


def test_fp32_engine_runs_a_time_step():
    """vpm_floattype = Float32 (simulation.jl:137): FP32 pair arithmetic, FP64 state — a whole RK3 step stays within
    single-precision distance of the FP64 engine."""
    import flowunsteady_b200 as fb
    x, g, s, static = mixed_field(4000, seed=9)
    static = np.where(np.all(g == 0, axis=1), 1.0, static)
    P = fb.new_particles(x, 50 * g, s, static=static)
    outs = []
    for bits in (64, 32):
        with fb.Engine(4000, float_bits=bits, schemes=fb.default_schemes()) as eng:
            eng.upload(P)
            eng.nextstep(2e-3, (1.0, 0.0, 0.0), relax=True)
            outs.append(eng.download(np.zeros_like(P)))
    d = outs[1][:, 0:3] - P[:, 0:3]
    d64 = outs[0][:, 0:3] - P[:, 0:3]
    assert np.abs(d - d64).max() < 1e-4 * np.abs(d64).max()
    assert np.abs(outs[1][:, 3:6] - outs[0][:, 3:6]).max() < 1e-4 * np.abs(outs[0][:, 3:6]).max()

"""CPU tests of the oracle (oracle/vpm_oracle.c) against the committed mpmath golden vectors, analytic identities and the
one in-tree reference P2P, /root/reference/src/FLOWUnsteady_processing_force.jl:879-929 (restated as vpmo_ffv_direct).

The reference's own implementation of the hot path cannot run here (un-vendored Julia dependency): PARITY UNPINNED.
These pins are the substitute.
"""
import numpy as np
import pytest

from oracle import oracle as o
from tests.util import mixed_field, relmax

KERNELS = ["gaussianerf", "winckelmans", "gaussian", "singular"]


@pytest.mark.parametrize("kernel", KERNELS)
def test_oracle_uj_vs_mpmath(golden, kernel):
    """The reference's expression form (g = erf(..) - .., aux = g'/(sigma r) - 3 g/r^2) cancels for pairs much closer
    than sigma: relative error ~ eps/s^3 in g/s^3, i.e. ~1e-16/s^2 in J.  The golden set contains one such pair
    (particles 0 and 1, s = 3.5e-7) on purpose: everything else must match mpmath to 1e-12, and that pair only to the
    bound the reference form itself can reach.  (The CUDA path evaluates g/s^3 from a series-exact table and matches
    mpmath to 1e-12 everywhere: tests/test_gpu_uj.py::test_golden_mpmath.)"""
    x, g, s, probes = golden["x"], golden["gamma"], golden["sigma"], golden["probes"]
    xt = np.concatenate([x, probes])
    ok = np.ones(xt.shape[0], dtype=bool)
    ok[[0, 1]] = False
    for accum in (0, 1):
        U, J = o.uj_direct(kernel, x, g, s, xt, accum=accum)
        Ug, Jg = golden[f"U_{kernel}"], golden[f"J_{kernel}"]
        assert np.abs(U - Ug)[ok].max() < 1e-12 * np.abs(Ug).max()
        assert np.abs(J - Jg)[ok].max() < 1e-12 * np.abs(Jg).max()
        loose_U, loose_J = (1e-12, 1e-12) if kernel in ("winckelmans", "singular") else (1e-6, 1.0)
        assert relmax(U, Ug) < loose_U and relmax(J, Jg) < loose_J


@pytest.mark.parametrize("kernel", ["gaussianerf", "winckelmans", "gaussian"])
def test_oracle_estr_vs_mpmath(golden, kernel):
    x, g, s = golden["x"], golden["gamma"], golden["sigma"]
    n = x.shape[0]
    J = golden[f"J_{kernel}"][:n]
    for transposed, tag in ((1, "T"), (0, "C")):
        E = o.estr_direct(kernel, transposed, x, g, s, J, x, J, accum=1)
        assert relmax(E, golden[f"E_{kernel}_{tag}"]) < 1e-11


@pytest.mark.parametrize("kernel", ["gaussianerf", "winckelmans", "gaussian"])
def test_kernel_functions_vs_mpmath(golden, kernel):
    rh = golden["rhat"]
    got = np.array([o.g_dgdr(kernel, r) for r in rh])
    zz = np.array([o.zeta(kernel, r) for r in rh])
    gg, dg, zt = golden[f"g_{kernel}"], golden[f"dg_{kernel}"], golden[f"zeta_{kernel}"]
    # g suffers the reference's own cancellation at tiny r_hat: compare absolutely there, relatively elsewhere
    assert np.all(np.abs(got[:, 0] - gg) <= 4e-16 + 1e-13 * np.abs(gg))
    # golden g' comes from numerical differentiation at 50 digits: absolute noise ~1e-40; exp(-r^2/2) in double carries
    # the rounding of its argument (|arg| eps relative)
    assert np.all(np.abs(got[:, 1] - dg) <= 1e-30 + 1e-12 * np.abs(dg))
    assert np.all(np.abs(zz - zt) <= 1e-300 + 1e-12 * np.abs(zt))
    # zeta = g'/(4 pi r^2): the radial basis is consistent with the regularising function (A.9)
    big = rh > 1e-2
    assert np.allclose(zz[big], got[big, 1] / (4 * np.pi * rh[big] ** 2), rtol=1e-12, atol=1e-300)


def test_zeta0_values():
    assert o.zeta("gaussianerf", 0.0) == pytest.approx(0.0634936359342, rel=1e-12)
    assert o.zeta("winckelmans", 0.0) == pytest.approx(0.5968310365946, rel=1e-12)
    assert o.zeta("gaussian", 0.0) == pytest.approx(0.2387324146378, rel=1e-12)


@pytest.mark.parametrize("kernel", ["gaussianerf", "winckelmans"])
def test_ffv_direct_identity(kernel):
    """REF pin: _Ffv_direct's M[1:3] = sum_f U_b(x_f) x Gamma_f with U_b from the same kernel form UJ_direct uses, and
    M[4:6] = sum_f g K(x_b - x_f) x (Gamma_b x Gamma_f)."""
    rng = np.random.default_rng(7)
    nb, nf = 9, 40
    xb, gb, sb = rng.random((nb, 3)), rng.standard_normal((nb, 3)), 0.2 + 0.2 * rng.random(nb)
    xf, gf = rng.random((nf, 3)), rng.standard_normal((nf, 3))
    xf[3] = xb[2]                       # an exactly coincident pair is skipped (r != 0 test)
    M = o.ffv_direct(kernel, xb, gb, sb, xf, gf)
    for b in range(nb):
        Ub, _ = o.uj_direct(kernel, xb[b:b + 1], gb[b:b + 1], sb[b:b + 1], xf, accum=1)
        assert relmax(M[b, 0:3], np.cross(Ub, gf).sum(0)) < 1e-12
        # second block: velocity at x_b induced by "particles" with strength (Gamma_b x Gamma_f) and sigma_b, sign flipped
        tot = np.zeros(3)
        for f in range(nf):
            Uf, _ = o.uj_direct(kernel, xf[f:f + 1], np.cross(gb[b], gf[f])[None], sb[b:b + 1], xb[b:b + 1], accum=1)
            tot += Uf[0]
        assert relmax(M[b, 3:6], tot) < 1e-12


@pytest.mark.parametrize("kernel", KERNELS)
def test_jacobian_is_gradient_and_divergence_free(kernel):
    x, g, s, _ = mixed_field(60, seed=1)
    probes = np.random.default_rng(2).random((10, 3))
    U, J = o.uj_direct(kernel, x, g, s, probes, accum=1)
    h = 1e-6
    for jdir in range(3):
        d = np.zeros(3)
        d[jdir] = h
        Up, _ = o.uj_direct(kernel, x, g, s, probes + d, accum=1)
        Um, _ = o.uj_direct(kernel, x, g, s, probes - d, accum=1)
        fd = (Up - Um) / (2 * h)
        assert relmax(J[:, 3 * jdir:3 * jdir + 3], fd) < 1e-7
    assert np.abs(J[:, 0] + J[:, 4] + J[:, 8]).max() < 1e-13 * np.abs(J).max()


def test_far_field_is_singular_biot_savart():
    x, g, s, _ = mixed_field(30, seed=3)
    far = np.array([[30.0, -20.0, 10.0], [0.0, 50.0, 5.0]])
    Ua, Ja = o.uj_direct("gaussianerf", x, g, s, far, accum=1)
    Ub, Jb = o.uj_direct("singular", x, g, s, far, accum=1)
    assert np.array_equal(Ua, Ub) and np.array_equal(Ja, Jb)


def test_field_uj_matches_raw_and_reset_flags():
    x, g, s, static = mixed_field(200, seed=4)
    P = o.new_field(x, g, s, static=static)
    sch = o.default_schemes()
    o.field_uj(P, sch)
    U, J = o.uj_direct("gaussianerf", x, g, s, x)
    assert np.array_equal(P[:, o.U:o.U + 3], U) and np.array_equal(P[:, o.J:o.J + 9], J)
    o.field_uj(P, sch, reset=False)
    assert relmax(P[:, o.U:o.U + 3], 2 * U) < 1e-13


def test_nextstep_statics_and_time():
    x, g, s, static = mixed_field(150, seed=5)
    static = np.where(np.all(g == 0, axis=1), 1.0, static)
    P = o.new_field(x, 50 * g, s, static=static)
    P0 = P.copy()
    for integ in ("euler", "rungekutta3"):
        P = P0.copy()
        t, nt = o.nextstep(P, o.default_schemes(integration=integ), 1e-3, (1.0, 0.0, 0.0), relax=True)
        assert (t, nt) == (1e-3, 1)
        st = P0[:, o.STATIC] > 0
        assert np.array_equal(P[st, 0:7], P0[st, 0:7])
        assert np.all(P[~st, 0] != P0[~st, 0])
        assert np.all(np.isfinite(P))


def test_rk3_is_third_order_on_freestream_and_consistent_with_euler():
    """A single particle (no self-induction) just advects with Uinf: low-storage RK3 must land exactly on x + dt U."""
    P = o.new_field(np.array([[0.1, 0.2, 0.3]]), np.array([[0.0, 0.0, 1.0]]), np.array([0.1]))
    o.nextstep(P, o.default_schemes(relaxation="none"), 0.5, (2.0, -1.0, 0.5), relax=False)
    assert np.allclose(P[0, 0:3], [1.1, -0.3, 0.55], rtol=0, atol=1e-15)


def test_pedrizzetti_aligns_and_corrected_preserves_norm():
    rng = np.random.default_rng(6)
    p = np.zeros(43)
    p[o.GAMMA:o.GAMMA + 3] = rng.standard_normal(3)
    p[o.J:o.J + 9] = rng.standard_normal(9)
    L = o.lib()
    a = p.copy()
    L.vpmo_relax_particle(a, 1, 0.3)
    b = p.copy()
    L.vpmo_relax_particle(b, 2, 0.3)
    Jm = p[o.J:o.J + 9]
    w = np.array([Jm[5] - Jm[7], Jm[6] - Jm[2], Jm[1] - Jm[3]])
    G = p[o.GAMMA:o.GAMMA + 3]
    expect = 0.7 * G + 0.3 * np.linalg.norm(G) * w / np.linalg.norm(w)
    assert np.allclose(a[o.GAMMA:o.GAMMA + 3], expect, rtol=1e-14)
    assert np.linalg.norm(b[o.GAMMA:o.GAMMA + 3]) == pytest.approx(np.linalg.norm(G), rel=1e-13)


def test_rvpm_sigma_closure():
    """rVPM (f=0, g=1/5): d sigma/dt = -(1/5) sigma (S.Gamma)/|Gamma|^2 ; cVPM keeps sigma (rvpm.md:197-239)."""
    p = np.zeros(43)
    p[o.GAMMA:o.GAMMA + 3] = [0.0, 0.0, 2.0]
    p[o.SIGMA] = 0.5
    p[o.J + 8] = 3.0            # du_z/dz = 3  -> S = (0, 0, 6), S.Gamma/|Gamma|^2 = 3
    L = o.lib()
    import ctypes as C
    for g_, expect in ((0.2, 0.5 - 1e-3 * 0.5 * 0.2 * 3.0), (0.0, 0.5)):
        q = p.copy()
        sch = o.default_schemes(g=g_)
        L.vpmo_update_particle(q, C.byref(sch), 0.0, 1.0, 1e-3, np.zeros(3), o.zeta("gaussianerf", 0.0))
        assert q[o.SIGMA] == pytest.approx(expect, rel=1e-15)
        # Gamma: dGamma/dt = S - 3 Z Gamma, Z = g * 3  -> z-component 6 - 3 (3 g) 2
        assert q[o.GAMMA + 2] == pytest.approx(2.0 + 1e-3 * (6.0 - 18.0 * g_), rel=1e-15)


def _golden_step():
    import os
    return np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "step_oracle.npz"))


def test_oracle_matches_its_frozen_step_fixture():
    """tests/golden/step_oracle.npz (tools/gen_golden_step.py) freezes two-step results of every scheme family, so a
    later edit of the oracle cannot silently move the goal posts of the GPU parity tests."""
    import sys, os
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tools"))
    import gen_golden_step as gg
    G = _golden_step()
    assert np.array_equal(G["P0"], gg.field())
    for name, kw in gg.CASES.items():
        P = gg.field()
        if "corespreading" in name:
            P[:, o.SIGMA] = 0.2
        t, nt = 0.0, 0
        for _ in range(2):
            t, nt = o.nextstep(P, o.default_schemes(**kw), 5e-3, (1.0, -0.5, 0.25), relax=True, t=t, nt=nt)
        assert np.allclose(P, G[name], rtol=1e-12, atol=1e-14), name


def test_thin_vortex_ring_self_induced_velocity():
    """a thin ring of circulation Gamma and radius R discretised along its centreline moves along its axis
    with U = Gamma/(4 pi R) [ln(8R/sigma) - C]: direction by the right-hand rule, 1/(4 pi R) in front of the logarithm, a
    constant C that does not depend on sigma, R or Gamma — and within a few percent of Saffman's speed of a Gaussian-core ring
    (core a = sqrt(2) sigma, constant 0.558), which is the velocity of the vorticity centroid rather than of the centreline."""
    def ring(N, R, G, sig):
        phi = 2 * np.pi * (np.arange(N) + 0.5) / N
        x = np.stack([R * np.cos(phi), R * np.sin(phi), np.zeros(N)], -1)
        t = np.stack([-np.sin(phi), np.cos(phi), np.zeros(N)], -1)
        return x, G * (2 * np.pi * R / N) * t, np.full(N, sig)

    consts = []
    for R, G, sig in ((1.0, 1.0, 0.05), (1.0, 1.0, 0.02), (2.0, 3.0, 0.04)):
        x, g, s = ring(4000, R, G, sig)
        U, J = o.uj_direct("gaussianerf", x, g, s, x[:4], accum=0)
        assert np.abs(U[:, :2]).max() < 1e-12 and np.all(U[:, 2] > 0)          # along +z for a counter-clockwise ring
        consts.append(np.log(8 * R / sig) - 4 * np.pi * R * U[0, 2] / G)
        saffman = G / (4 * np.pi * R) * (np.log(8 * R / (np.sqrt(2) * sig)) - 0.558)
        assert abs(U[0, 2] / saffman - 1) < 0.05
        assert abs(J[0, 0] + J[0, 4] + J[0, 8]) < 1e-12 * np.abs(J[0]).max()
    assert max(consts) - min(consts) < 3e-3, consts

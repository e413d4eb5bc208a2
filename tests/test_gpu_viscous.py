"""GPU parity of the zeta pass and CoreSpreading's spatial adaptation (RBF conjugate gradient) against the oracle."""
import numpy as np
import pytest

from tests.util import mixed_field, relmax

pytestmark = pytest.mark.gpu


def _field(n, seed, spread):
    import flowunsteady_b200 as fb
    x, g, s, static = mixed_field(n, seed=seed, sigma_jitter=False)
    g = np.stack([np.sin(3 * x[:, 0]), np.cos(2 * x[:, 1]), x[:, 2]], -1) / n
    return fb.new_particles(x, g, s * spread, static=static), s[0]


def test_zeta_pass_vs_oracle():
    import flowunsteady_b200 as fb
    from oracle import oracle as o
    P, _ = _field(3000, 2, 1.0)
    with fb.Engine(P.shape[0]) as eng:
        eng.upload(P)
        eng.zeta()
        W = eng.download(np.zeros_like(P))[:, 12:15]
    Wo = o.zeta_direct("gaussianerf", P[:, 0:3], P[:, 3:6], P[:, 6], P[:, 0:3])
    assert relmax(W, Wo) < 1e-12


@pytest.mark.parametrize("itmax", [3, 15])
def test_corespreading_reset_vs_oracle(itmax):
    import flowunsteady_b200 as fb
    from oracle import oracle as o
    P, sgm0 = _field(1200, 4, 1.7)
    Po = P.copy()
    it_o, res_o = o.corespreading_reset(Po, "gaussianerf", sgm0, beta=1.5, itmax=itmax, tol=1e-3)
    sch = fb.default_schemes(viscous="corespreading", nu=1e-5, cs_sgm0=sgm0, cs_beta=1.5, cs_itmax=itmax, cs_tol=1e-3)
    with fb.Engine(P.shape[0], schemes=sch) as eng:
        eng.upload(P)
        it_g, res_g = eng.corespreading_reset()
        Pg = eng.download(np.zeros_like(P))
    assert it_g == it_o == itmax
    static = P[:, 42] > 0
    assert np.array_equal(Pg[static, 3:7], P[static, 3:7])                 # statics keep Gamma and sigma
    assert np.all(Pg[~static, 6] == sgm0)
    assert relmax(Pg[:, 12:15], Po[:, 12:15]) < 1e-12                      # target vorticity (spread cores)
    assert relmax(Pg[:, 3:6], Po[:, 3:6]) < 1e-8                           # CG amplifies round-off; same iterates
    assert np.allclose(res_g, res_o, rtol=1e-6)
    # the re-fit reproduces the spread-core vorticity better than the un-fitted strengths do
    W_before = o.zeta_direct("gaussianerf", P[:, 0:3], P[:, 3:6], Pg[:, 6], P[:, 0:3])
    W_after = o.zeta_direct("gaussianerf", P[:, 0:3], Pg[:, 3:6], Pg[:, 6], P[:, 0:3])
    assert relmax(W_after, Po[:, 12:15]) < 0.2 * relmax(W_before, Po[:, 12:15])


def test_no_reset_below_beta():
    import flowunsteady_b200 as fb
    P, sgm0 = _field(500, 5, 1.2)
    sch = fb.default_schemes(viscous="corespreading", nu=1e-5, cs_sgm0=sgm0, cs_beta=1.5)
    with fb.Engine(P.shape[0], schemes=sch) as eng:
        eng.upload(P)
        it, _ = eng.corespreading_reset()
        Pg = eng.download(np.zeros_like(P))
    assert it == 0 and np.array_equal(Pg[:, 0:9], P[:, 0:9])


def test_nextstep_with_corespreading_reset_vs_oracle():
    """vpm.nextstep with CoreSpreading(nu, sgm0; beta): sigma grows every substep and the reset fires after the last one."""
    import flowunsteady_b200 as fb
    from oracle import oracle as o
    P, sgm0 = _field(700, 7, 1.499)
    P[:, 3:6] *= 50.0
    kw = dict(integration="rungekutta3", viscous="corespreading", nu=2e-3, cs_sgm0=sgm0, cs_beta=1.5, cs_itmax=6, cs_tol=1e-6)
    Po = P.copy()
    o.nextstep(Po, o.default_schemes(**kw), 5e-2, (0.3, 0.0, 0.0), relax=True)
    with fb.Engine(P.shape[0], schemes=fb.default_schemes(**kw)) as eng:
        eng.upload(P)
        eng.nextstep(5e-2, (0.3, 0.0, 0.0), relax=True)
        Pg = eng.download(np.zeros_like(P))
    live = ~(P[:, 42] > 0)
    assert np.all(Pg[live, 6] == sgm0) and np.all(Po[live, 6] == sgm0)     # the reset did fire in both
    for sl in (slice(0, 3), slice(3, 6), slice(9, 12), slice(15, 24)):
        assert relmax(Pg[:, sl], Po[:, sl]) < 1e-8

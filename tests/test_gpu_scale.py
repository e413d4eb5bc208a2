"""Parity at the sizes BASELINE.json names, where the tile-level shortcuts actually fire.

The small-N tests (tests/test_gpu_step.py) use clouds whose sigma spans the whole box, so K2's T_FAR tile skip, K1's all-far
tiles and the FP32 far path never run there.  Here the fields are the wake geometries of the BASELINE configurations:
rotor-hover stand-in at N = 70,000 (configs[1]) — every target of one RK3 + DynamicSFS step against the oracle — and the
vortex-ring field at N = 1,000,000 (configs[2]) — 2048 sampled targets of U, J and E_str against oracle slabs.
Tolerances: 1e-12 U/J (north_star), 1e-11 E_str, 1e-9 for what passes through the dynamic procedure.
"""
import numpy as np
import pytest

from tests.util import relmax

pytestmark = pytest.mark.gpu

DYN = dict(kernel="gaussianerf", integration="rungekutta3", relaxation="pedrizzetti", sfs="dynamic", alpha=0.999,
           force_positive=1, clippings=1)


def _sample(n, m=2048):
    return np.sort(np.random.default_rng(1234).choice(n, m, replace=False))


def _estr_sampled(o, P, idx, transposed=1):
    """Oracle E_str at the sampled targets, fed with the J the device produced for ALL particles (J's own parity is asserted
    on the same sample)."""
    X, G, S = (np.ascontiguousarray(P[:, 0:3]), np.ascontiguousarray(P[:, 3:6]), np.ascontiguousarray(P[:, 6]))
    J = np.ascontiguousarray(P[:, 15:24])
    return o.estr_direct("gaussianerf", transposed, X, G, S, J, X[idx], J[idx], accum=1)


def test_tile_skip_fires_on_wake_fields():
    """The premise of this file: on the wake geometries (block, tile) pairs beyond T_FAR exist — 6 % of them on the rotor
    stand-in at 70k (measured), 90+ % on the rings — on the unit-cube cloud of the small tests none does."""
    import flowunsteady_b200 as fb
    from flowunsteady_b200 import fields
    from tests.util import mixed_field
    for (x, g, s), lo, hi in ((fields.rotor_wake(70_000), 0.03, 1.0), (fields.vortex_rings(200_000), 0.5, 1.0),
                              (mixed_field(1500, seed=5)[:3], 0.0, 0.0)):
        with fb.Engine(x.shape[0], schemes=fb.default_schemes()) as eng:
            eng.upload(fb.new_particles(x, g, s))
            st = eng.direct_tile_stats()
        assert st["all_pairs"] == st["target_blocks"] * st["source_tiles"]
        assert lo <= st["tile_far_fraction"] <= hi, st


def test_rotor70k_uj_estr_sampled_parity():
    """configs[1] geometry, N = 70k: U, J, E_str of `pfield.UJ(pfield; sfs=true)` at 2048 targets vs the oracle (FP64)."""
    import flowunsteady_b200 as fb
    from flowunsteady_b200 import fields
    from oracle import oracle as o
    x, g, s = fields.rotor_wake(70_000)
    P = fb.new_particles(x, g, s)
    with fb.Engine(P.shape[0], schemes=fb.default_schemes(**DYN)) as eng:
        eng.upload(P)
        eng.uj(True, True, True)
        Pg = eng.download(np.zeros_like(P))
    idx = _sample(P.shape[0])
    Uo, Jo = o.uj_direct("gaussianerf", x, g, s, x[idx], accum=1)
    assert relmax(Pg[idx, 9:12], Uo) < 1e-12
    assert relmax(Pg[idx, 15:24], Jo) < 1e-12
    assert relmax(Pg[idx, 39:42], _estr_sampled(o, Pg, idx)) < 1e-11


def test_rotor70k_nextstep_dynamic_sfs_full_parity():
    """One whole RK3 + DynamicSFS (alpha = 0.999, pseudo3level_positive, backscatter clipping) + pedrizzetti step on the
    rotor-hover stand-in at N = 70k (rotorhover.jl:53-55): EVERY particle's X, Gamma, sigma, C, U, J against the oracle's
    step (5 U/J + 4 E_str full N^2 evaluations on the host: about a minute on 16 cores)."""
    import flowunsteady_b200 as fb
    from flowunsteady_b200 import fields
    from oracle import oracle as o
    x, g, s = fields.rotor_wake(70_000)
    P = fb.new_particles(x, fields.floor_gamma(g), s)
    dt, Uinf = 1e-4, (0.0, 0.0, -1.0)
    Po = P.copy()
    o.nextstep(Po, o.default_schemes(**DYN), dt, Uinf, relax=True)
    with fb.Engine(P.shape[0], schemes=fb.default_schemes(**DYN)) as eng:
        eng.upload(P)
        eng.nextstep(dt, Uinf, relax=True)
        Pg = eng.download(np.zeros_like(P))
    assert relmax(Pg[:, 0:3], Po[:, 0:3]) < 1e-12
    assert relmax(Pg[:, 9:12], Po[:, 9:12]) < 1e-11
    assert relmax(Pg[:, 15:24], Po[:, 15:24]) < 1e-11
    # Through the dynamic procedure the test-filter minus domain-filter difference (alpha = 0.999) amplifies the 1e-12 of
    # U/J/E_str by 1/(1 - alpha) = 1e3 in BOTH implementations; the max norm then picks the worst of 70,000 particles
    # (measured: Gamma 1.3e-9, sigma 3.5e-13; the 2,000-particle cases of test_gpu_step.py stay under 1e-9).  Budget for the
    # integrated state: 5 x 1e-12 x 1e3.  The coefficient C itself is a RATIO of two such differences, <Gamma.L> / <Gamma.m>,
    # and where the denominator is small it amplifies once more (measured 1.2e-8 at the worst of the 70,000 particles, relative
    # to max C = 1); it enters Gamma only multiplied by dt E_str, which is why Gamma holds 1.3e-9.  Budget: 1e-7.
    errs = {name: relmax(Pg[:, sl], Po[:, sl]) for name, sl in (("Gamma", slice(3, 6)), ("sigma", slice(6, 7)), ("C", slice(36, 39)))}
    assert errs["Gamma"] < 5e-9 and errs["sigma"] < 5e-9 and errs["C"] < 1e-7, errs
    assert np.abs(Po[:, 36]).max() > 0                        # the coefficient is active on this field


def test_rings1m_uj_estr_sampled_parity_fp64_and_fp32():
    """configs[2] at full size: `pfield.UJ(pfield; sfs=true)` on 10^6 vortex-ring particles; 2048 sampled targets of U, J,
    E_str vs oracle slabs over all 10^6 sources.  Then the FP32 variant (vpm_floattype = Float32) on the same field, whose
    far path only runs at sizes like this one."""
    import flowunsteady_b200 as fb
    from flowunsteady_b200 import engine as E, fields
    from oracle import oracle as o
    x, g, s = fields.vortex_rings(1_000_000)
    P = fb.new_particles(x, g, s)
    idx = _sample(P.shape[0])
    Uo, Jo = o.uj_direct("gaussianerf", x, g, s, x[idx], accum=1)
    mask = E.FM_X | E.FM_GAMMA | E.FM_SIGMA | E.FM_U | E.FM_J | E.FM_SFS
    with fb.Engine(P.shape[0], schemes=fb.default_schemes(**DYN)) as eng:
        eng.upload(P)
        eng.uj(True, True, True)
        Pg = eng.download(np.zeros_like(P), field_mask=mask)
        st = eng.direct_tile_stats()
    assert st["tile_far_fraction"] > 0.8
    assert relmax(Pg[idx, 9:12], Uo) < 1e-12
    assert relmax(Pg[idx, 15:24], Jo) < 1e-12
    assert relmax(Pg[idx, 39:42], _estr_sampled(o, Pg, idx)) < 1e-11
    with fb.Engine(P.shape[0], float_bits=32, schemes=fb.default_schemes(**DYN)) as eng:
        eng.upload(P)
        eng.uj(True, True, True)
        P32 = eng.download(np.zeros_like(P), field_mask=mask)
    assert relmax(P32[idx, 9:12], Uo) < 2e-5
    assert relmax(P32[idx, 15:24], Jo) < 4e-4
    assert relmax(P32[idx, 39:42], Pg[idx, 39:42]) < 4e-3     # E_str from FP32 J: differences of J lose a further digit


# (field generator, N, [U, J] error bounds at the reference's defaults = 2 x measured, profiles/r02r_fmm_error_table.md)
FMM_CASES = {
    "rotor_200k": (lambda f: f.rotor_wake(200_000, nfil=101, nsteps_per_rev=72), (6.5e-3, 1.2e-1)),
    "rings_1m": (lambda f: f.vortex_rings(1_000_000), (3.1e-3, 2.8e-2)),
    "random_2m": (lambda f: f.random_field(2_000_000), (1.1e-1, 3.8e-1)),
    "vahana_5m": (lambda f: f.vahana_wake(5_000_000), (3.8e-3, 5.0e-2)),
}


@pytest.mark.parametrize("case", list(FMM_CASES))
def test_fmm_error_vs_direct_at_baseline_sizes(case):
    """UJ_fmm at the reference's defaults (p = 4, ncrit = 50, theta = 0.4, nonzero_sigma = false) against the direct kernel on
    2048 sampled particles of BASELINE configs[1]-[4] at their full sizes.  The bounds are twice the measured relative L2 errors
    (VERDICT r1 next #5): they are dominated by the regularisation error of the singular far field, which the reference's method
    shares (oracle/fmm_oracle.c); the expansion truncation alone is 10x - 1000x smaller (same table)."""
    import flowunsteady_b200 as fb
    from flowunsteady_b200 import engine as E, fields
    gen, (bU, bJ) = FMM_CASES[case]
    x, g, s = gen(fields)
    g = fields.floor_gamma(g)
    P = fb.new_particles(x, g, s)
    idx = _sample(P.shape[0])
    with fb.Engine(P.shape[0], schemes=fb.default_schemes(uj="direct")) as e:
        e.upload(P)
        Ud, Jd = e.uj_probe(x[idx], want_J=True)
        e.set_schemes(fb.default_schemes(uj="fmm", fmm_p=4, fmm_ncrit=50, fmm_theta=0.4))
        e.uj()
        F = e.download(np.zeros_like(P), field_mask=E.FM_U | E.FM_J)
        st = e.fmm_stats()
    eU = float(np.linalg.norm(F[idx, 9:12] - Ud) / np.linalg.norm(Ud))
    eJ = float(np.linalg.norm(F[idx, 15:24] - Jd) / np.linalg.norm(Jd))
    assert st["m2l_pairs"] > st["p2p_pairs"] > 0
    assert eU < bU and eJ < bJ, (case, eU, eJ)
    assert eU > bU / 20 and eJ > bJ / 20, (case, eU, eJ)          # the bounds are tight: a silent change of method would show

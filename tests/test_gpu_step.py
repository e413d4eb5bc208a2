"""GPU parity of the SFS pass (K2), the dynamic coefficient (K4) and the euler / rungekutta3 update (K5) against the
CPU oracle, through the C ABI (vpmb200_uj(sfs=1), vpmb200_sfs, vpmb200_nextstep).

Tolerances (max-norm relative, FP64): 1e-12 for U/J, 1e-11 for quantities built from differences of J (E_str, C_d)
and for fields after whole time steps (round-off of two evaluations compounds through the update).
"""
import ctypes as C

import numpy as np
import pytest

from tests.util import mixed_field, relmax

pytestmark = pytest.mark.gpu


def _schemes(**kw):
    """Matching (engine, oracle) scheme structs from one description."""
    import flowunsteady_b200 as fb
    from oracle import oracle as o
    se = fb.default_schemes(**kw)
    so = o.default_schemes(**kw)
    return se, so


def _field(n, seed, C0=None):
    import flowunsteady_b200 as fb
    x, g, s, static = mixed_field(n, seed=seed)
    g = g * 50.0  # strong enough that stretching changes Gamma visibly in one step
    # zero-strength particles are probes; the reference never integrates them (they are removed before nextstep,
    # src/FLOWUnsteady_simulation.jl:550-552) and 0/|Gamma|^2 is NaN in its update, so mark them static here
    static = np.where(np.all(g == 0, axis=1), 1.0, static)
    P = fb.new_particles(x, g, s, static=static)
    if C0 is not None:
        P[:, 36:39] = C0
    return P


GROUPS = {"X": slice(0, 3), "Gamma": slice(3, 6), "sigma": slice(6, 7), "U": slice(9, 12), "J": slice(15, 24),
          "M": slice(27, 36), "C": slice(36, 39), "SFS": slice(39, 42)}


def _compare(Pg, Po, tol, groups):
    for name in groups:
        sl = GROUPS[name]
        err = relmax(Pg[:, sl], Po[:, sl])
        assert err < tol, f"{name}: {err:.3e} >= {tol}"


@pytest.mark.parametrize("kernel", ["gaussianerf", "winckelmans", "gaussian"])
@pytest.mark.parametrize("transposed", [1, 0])
def test_estr_vs_oracle(kernel, transposed):
    import flowunsteady_b200 as fb
    from oracle import oracle as o
    se, so = _schemes(kernel=kernel, transposed=transposed, sfs="constant")
    P = _field(1500, seed=5)
    Po = P.copy()
    o.field_uj(Po, so, reset=True, reset_sfs=True, sfs=True)
    with fb.Engine(P.shape[0], schemes=se) as eng:
        eng.upload(P)
        eng.uj(reset=True, reset_sfs=True, sfs=True)
        Pg = eng.download(np.zeros_like(P))
    _compare(Pg, Po, 1e-12, ["U", "J"])
    _compare(Pg, Po, 1e-11, ["SFS"])


def test_estr_golden_mpmath(golden):
    """E_str against the 50-digit evaluation (both stretching schemes)."""
    import flowunsteady_b200 as fb
    x, g, s = golden["x"], golden["gamma"], golden["sigma"]
    for kernel in ("gaussianerf", "winckelmans"):
        for transposed, tag in ((1, "T"), (0, "C")):
            with fb.Engine(x.shape[0], schemes=fb.default_schemes(kernel=kernel, transposed=transposed)) as eng:
                eng.upload(fb.new_particles(x, g, s))
                eng.uj(reset=True, reset_sfs=True, sfs=True)
                Pg = eng.download(np.zeros((x.shape[0], 43)))
            assert relmax(Pg[:, 39:42], golden[f"E_{kernel}_{tag}"]) < 1e-10


@pytest.mark.parametrize("cfg", [
    dict(sfs="dynamic", alpha=0.999, force_positive=1, clippings=1),                       # rotorhover high fidelity
    dict(sfs="dynamic", alpha=0.667, clippings=1, controls=3, minC=0.0, maxC=1.0),          # three-level + controls (vahana)
    dict(sfs="constant", Cs=1.0, clippings=1),                                              # SFS_Cs_nobackscatter
    dict(sfs="none"),
])
def test_sfs_call_vs_oracle(cfg):
    """pfield.SFS(pfield; a=0) — the first-substep path (dynamic procedure = two filtered evaluations)."""
    import flowunsteady_b200 as fb
    from oracle import oracle as o
    se, so = _schemes(**cfg)
    P = _field(900, seed=9, C0=np.array([0.3, 1e-3, 4e-3]))
    Po = P.copy()
    o.field_sfs(Po, so, a=0.0, b=1.0 / 3.0, t=0.02, nt=4)
    with fb.Engine(P.shape[0], schemes=se) as eng:
        eng.upload(P)
        eng.set_time(0.02, 4)
        eng.sfs(0.0, 1.0 / 3.0)
        Pg = eng.download(np.zeros_like(P))
    _compare(Pg, Po, 1e-12, ["U", "J", "sigma"])
    _compare(Pg, Po, 1e-10, ["SFS", "C", "M"])


STEP_CASES = {
    "rk3_rvpm_pedrizzetti": dict(integration="rungekutta3", relaxation="pedrizzetti"),
    "euler_rvpm_pedrizzetti": dict(integration="euler", relaxation="pedrizzetti"),
    "rk3_dynamic_sfs": dict(integration="rungekutta3", sfs="dynamic", force_positive=1, clippings=1),
    "euler_dynamic_controls": dict(integration="euler", sfs="dynamic", alpha=0.667, clippings=1, controls=3),
    "rk3_cvpm_classic_corrected": dict(integration="rungekutta3", g=0.0, transposed=0, relaxation="correctedpedrizzetti"),
    "rk3_winckelmans_norelax": dict(integration="rungekutta3", kernel="winckelmans", relaxation="none"),
    "rk3_corespreading": dict(integration="rungekutta3", viscous="corespreading", nu=1.5e-5),
    "euler_corespreading_constsfs": dict(integration="euler", viscous="corespreading", nu=1.5e-5, sfs="constant", clippings=1),
}


@pytest.mark.parametrize("case", sorted(STEP_CASES))
def test_nextstep_vs_oracle(case):
    """Two vpm.nextstep calls with a freestream; every state group compared (statics must not move)."""
    import flowunsteady_b200 as fb
    from oracle import oracle as o
    se, so = _schemes(**STEP_CASES[case])
    P = _field(600, seed=21)
    Po = P.copy()
    dt, Uinf = 2.0e-3, (1.0, -0.5, 0.25)
    t, nt = 0.0, 0
    for _ in range(2):
        t, nt = o.nextstep(Po, so, dt, Uinf, relax=True, t=t, nt=nt)
    with fb.Engine(P.shape[0], schemes=se) as eng:
        eng.upload(P)
        for _ in range(2):
            eng.nextstep(dt, Uinf, relax=True)
        Pg = eng.download(np.zeros_like(P))
        assert eng.get_time() == (t, nt)
    # the dynamic procedure differences two evaluations whose filters differ by (1 - alpha) = 1e-3 .. 0.3, so C_d
    # carries round-off amplified by 1/(1 - alpha) in BOTH implementations; the state inherits it through C_d E_str
    tol = 1e-9 if STEP_CASES[case].get("sfs") == "dynamic" else 1e-11
    _compare(Pg, Po, tol, ["X", "Gamma", "sigma", "U", "J"])
    _compare(Pg, Po, 1e-8, ["C", "SFS"])
    static = P[:, 42] > 0
    assert np.array_equal(Pg[static, 0:7], P[static, 0:7])


def test_add_remove_roundtrip():
    """vpm.add_particle / remove_particle semantics: append, swap-remove from the end downward."""
    import flowunsteady_b200 as fb
    P = _field(300, seed=2)
    with fb.Engine(400) as eng:
        eng.upload(P[:200])
        eng.add_particles(P[200:300])
        assert eng.np == 300
        assert np.array_equal(eng.download(np.zeros((300, 43))), P)
        for i in range(299, 249, -1):      # the static-removal loop of simulation.jl:361-365
            eng.remove_particle(i)
        assert eng.np == 250
        eng.remove_particle(10)            # swap-remove: last particle lands in slot 10
        got = eng.download(np.zeros((249, 43)))
        expect = P[:249].copy()
        expect[10] = P[249]
        assert np.array_equal(got, expect)
        with pytest.raises(fb.EngineError):
            eng.add_particles(np.zeros((200, 43)))


def test_partial_download_keeps_host_rows():
    import flowunsteady_b200 as fb
    from flowunsteady_b200 import engine as E
    P = _field(100, seed=4)
    with fb.Engine(100) as eng:
        eng.upload(P)
        eng.uj()
        host = np.full((100, 43), 7.0)
        eng.download(host, field_mask=E.FM_U)
    assert np.all(host[:, :9] == 7.0) and np.all(host[:, 12:] == 7.0)
    assert not np.any(host[:, 9:12] == 7.0)


def test_masked_transfers_move_only_the_selected_column_runs():
    """Masked upload / download cross the bus as strided 2-D copies of the selected column runs (engine.cu: mask_runs):
    every group mask, a padded host matrix (ld > 43), and columns outside the mask untouched on both sides."""
    import flowunsteady_b200 as fb
    from flowunsteady_b200 import engine as E
    n = 777
    rng = np.random.default_rng(9)
    A = rng.random((n, 43))
    B = rng.random((n, 43)) + 10.0
    big = np.zeros((n, 48))                       # ld = 48: rows are not contiguous
    big[:, :43] = B
    cols = {E.FM_X: (0, 3), E.FM_GAMMA: (3, 6), E.FM_SIGMA: (6, 7), E.FM_VOL: (7, 8), E.FM_CIRCULATION: (8, 9), E.FM_U: (9, 12),
            E.FM_VORTICITY: (12, 15), E.FM_J: (15, 24), E.FM_PSE: (24, 27), E.FM_M: (27, 36), E.FM_C: (36, 39),
            E.FM_SFS: (39, 42), E.FM_STATIC: (42, 43)}
    for mask in (E.FM_X | E.FM_GAMMA | E.FM_J, E.FM_STATE, E.FM_SIGMA | E.FM_STATIC, E.FM_ALL & ~E.FM_M, E.FM_ALL):
        sel = np.zeros(43, dtype=bool)
        for bit, (c0, c1) in cols.items():
            if mask & bit:
                sel[c0:c1] = True
        with fb.Engine(n) as eng:
            eng.upload(A)                          # device = A everywhere
            eng.upload(big[:, :43], field_mask=mask)   # selected columns <- B (strided host rows)
            got = eng.download(np.zeros((n, 43)))
            assert np.array_equal(got[:, sel], B[:, sel]) and np.array_equal(got[:, ~sel], A[:, ~sel])
            host = np.full((n, 48), -1.0)
            eng.download(host[:, :43], field_mask=mask)
            assert np.array_equal(host[:, :43][:, sel], got[:, sel])
            assert np.all(host[:, :43][:, ~sel] == -1.0) and np.all(host[:, 43:] == -1.0)


def test_impulse_conserved_large():
    """Size-independent property at N = 100k: the linear impulse 0.5 sum x × Gamma is conserved by an inviscid
    rVPM step to integration accuracy, and no NaN appears."""
    import flowunsteady_b200 as fb
    from flowunsteady_b200 import fields
    x, g, s = fields.vortex_rings(100_000)
    P = fb.new_particles(x, g, s)
    I0 = fields.ring_impulse(x, g)
    with fb.Engine(P.shape[0], schemes=fb.default_schemes(relaxation="none")) as eng:
        eng.upload(P)
        eng.nextstep(1e-2, relax=False)
        assert eng.count_nonfinite() == 0
        Pg = eng.download(np.zeros_like(P))
    I1 = fields.ring_impulse(Pg[:, 0:3], Pg[:, 3:6])
    assert np.abs(I1 - I0).max() < 1e-4 * np.abs(I0).max()
    assert np.abs(Pg[:, 0:3] - x).max() > 0


def test_nextstep_vs_frozen_oracle_fixture():
    """GPU vs tests/golden/step_oracle.npz — the committed two-step results of every scheme family (96 particles)."""
    import os
    import sys
    import flowunsteady_b200 as fb
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tools"))
    import gen_golden_step as gg
    G = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "step_oracle.npz"))
    for name, kw in gg.CASES.items():
        P = G["P0"].copy()
        if "corespreading" in name:
            P[:, 6] = 0.2
        with fb.Engine(P.shape[0], schemes=fb.default_schemes(**kw)) as eng:
            eng.upload(P)
            for _ in range(2):
                eng.nextstep(5e-3, (1.0, -0.5, 0.25), relax=True)
            Pg = eng.download(np.zeros_like(P))
        tol = 1e-8 if ("dynamic" in name or "corespreading" in name) else 1e-11
        _compare(Pg, G[name], tol, ["X", "Gamma", "sigma", "U", "J"])

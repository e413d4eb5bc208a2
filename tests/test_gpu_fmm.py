"""GPU tests of K3 (UJ_fmm) through the C ABI.

The reference's FMM backend (ExaFMM / FastMultipole.jl) is not in the tree and cannot run here, so "parity" for this row
means: same algorithm class and parameters (p, ncrit, theta), exact agreement with the direct sum in the theta -> 0
limit (everything becomes near field, evaluated by the same pair functions), and a STATED truncation error versus the
direct sum at the reference defaults p = 4, ncrit = 50, theta = 0.4 (src/FLOWUnsteady_simulation.jl:43).
"""
import numpy as np
import pytest

from tests.util import mixed_field, relmax

pytestmark = pytest.mark.gpu


def rel_l2(a, b):
    return float(np.linalg.norm(a - b) / np.linalg.norm(b))


def _eval(P, sfs=False, options=None, **kw):
    import flowunsteady_b200 as fb
    with fb.Engine(P.shape[0], schemes=fb.default_schemes(**kw)) as eng:
        for name, value in (options or {}).items():
            eng.set_option(name, value)
        eng.upload(P)
        eng.uj(True, True, sfs)
        out = eng.download(np.zeros_like(P))
        stats = eng.fmm_stats()
    return out, stats


def _field(n, seed=31):
    import flowunsteady_b200 as fb
    x, g, s, static = mixed_field(n, seed=seed)
    return fb.new_particles(x, g, s, static=static)


@pytest.mark.parametrize("kernel", ["gaussianerf", "winckelmans", "singular"])
@pytest.mark.parametrize("n", [1, 2, 63, 3000])
def test_theta_to_zero_is_the_direct_sum(kernel, n):
    """With theta -> 0 no cell pair is ever well separated: the FMM degenerates to leaf-pair P2P and must reproduce the
    direct kernel to round-off (different summation order only)."""
    P = _field(n)
    D, _ = _eval(P, kernel=kernel, uj="direct", sfs=True)
    F, st = _eval(P, kernel=kernel, uj="fmm", fmm_theta=1e-6, fmm_ncrit=50, sfs=True)
    assert st["m2l_pairs"] == 0 and st["p2p_pairs"] == st["leaves"] ** 2
    assert relmax(F[:, 9:12], D[:, 9:12]) < 1e-12 or np.abs(D[:, 9:12]).max() == 0
    assert relmax(F[:, 15:24], D[:, 15:24]) < 1e-12 or np.abs(D[:, 15:24]).max() == 0
    if kernel != "singular" and n > 2:
        assert relmax(F[:, 39:42], D[:, 39:42]) < 1e-11


@pytest.mark.parametrize("copies", [1, 8])
@pytest.mark.parametrize("kernel", ["gaussianerf", "winckelmans"])
@pytest.mark.parametrize("ncrit", [5, 18, 33, 128])
def test_near_field_lane_mapping_every_leaf_size(kernel, ncrit, copies):
    """The near-field kernel splits a leaf's targets into passes of T = 4..32 targets x 32/T ways and pads the last source
    batch to whole groups of 2 x ways records (fmm.cuh: leaf_pass, pad_batch).  ncrit = 5 / 18 / 33 / 128 on a field with a
    dense cluster produces leaves of every count from 1 to ncrit, i.e. every (T, ways) pair, multi-pass leaves and leaves
    larger than a warp; theta -> 0 makes all of it near field, which must equal the direct sum to round-off."""
    P = _field(1500, seed=5)
    P[:400, 0:3] = P[0, 0:3] + 0.02 * (P[:400, 0:3] - P[0, 0:3])     # cluster: deep, unevenly filled leaves
    D, _ = _eval(P, kernel=kernel, uj="direct", sfs=True)
    F, st = _eval(P, kernel=kernel, uj="fmm", fmm_theta=1e-6, fmm_ncrit=ncrit, sfs=True, options={"fmm_table_copies": copies})
    assert st["m2l_pairs"] == 0
    assert relmax(F[:, 9:12], D[:, 9:12]) < 1e-12
    assert relmax(F[:, 15:24], D[:, 15:24]) < 1e-12
    assert relmax(F[:, 39:42], D[:, 39:42]) < 1e-11


def test_reference_defaults_error_vs_direct():
    """p = 4, ncrit = 50, theta = 0.4, nonzero_sigma = false (src/FLOWUnsteady_simulation.jl:43): the far field is the
    SINGULAR kernel wherever the acceptance criterion holds, even inside the regularised range, so the error against the
    direct sum depends on leaf size / sigma and does not improve with p (measured: 2e-3 .. 1e-2 in U on ring wakes).
    With nonzero_sigma = true the criterion also keeps 5 sigma of clearance and the error is the expansion truncation."""
    import flowunsteady_b200 as fb
    from flowunsteady_b200 import fields
    x, g, s = fields.vortex_rings(60_000)
    P = fb.new_particles(x, g, s)
    D, _ = _eval(P, uj="direct", sfs=True)
    F, st = _eval(P, uj="fmm", fmm_p=4, fmm_ncrit=50, fmm_theta=0.4, sfs=True)
    assert st["m2l_pairs"] > 0 and st["levels"] >= 3
    assert rel_l2(F[:, 9:12], D[:, 9:12]) < 3e-2
    assert rel_l2(F[:, 15:24], D[:, 15:24]) < 1e-1
    F, st = _eval(P, uj="fmm", fmm_p=4, fmm_ncrit=50, fmm_theta=0.4, fmm_nonzero_sigma=1, sfs=True)
    assert rel_l2(F[:, 9:12], D[:, 9:12]) < 2e-3
    assert rel_l2(F[:, 15:24], D[:, 15:24]) < 4e-3
    assert rel_l2(F[:, 39:42], D[:, 39:42]) < 5e-2      # E_str: near field only, as Estr_fmm


@pytest.mark.parametrize("field", ["rings", "rotor", "random"])
def test_cuda_fmm_vs_reference_style_fmm_oracle(field):
    """Row a3 of the scope table: the CUDA FMM against the CPU restatement of the REFERENCE's method (oracle/fmm_oracle.c:
    ExaFMM-style solid harmonics of degree < p for multipoles and locals, same octree, same acceptance, same traversal) at
    equal (p, ncrit, theta), both measured against the direct sum on the same field.
      * pure truncation (singular kernel everywhere): the CUDA expansions (Cartesian Taylor, multipoles to order p - 1, locals to
        order p + 1) carry the same multipole information and two more local orders, so their error must not exceed the
        oracle's (x 1.2 slack for the different tree of the sparse-leaf refinement);
      * reference defaults (gaussianerf near field, singular far field, nonzero_sigma = false): both methods carry the same
        regularisation error — the singular far field is used inside the regularised range wherever the acceptance holds —
        on top of their truncation error, so the CUDA error never exceeds the oracle's (measured at p = 3 on the rings:
        CUDA 2.6e-3 / 6.6e-3 in U / J, oracle 1.3e-2 / 4.5e-2, where the oracle's degree-2 locals still dominate): the
        1e-2 .. 1e-1 level in J at the defaults is a property of the reference's method, not of this implementation
        (profiles/r02r_fmm_error_table.md splits it into truncation and regularisation parts at full size)."""
    import flowunsteady_b200 as fb
    from flowunsteady_b200 import fields
    from oracle import oracle as o
    x, g, s = {"rings": lambda: fields.vortex_rings(30_000), "rotor": lambda: fields.rotor_wake(20_000),
               "random": lambda: fields.random_field(20_000)}[field]()
    g = fields.floor_gamma(g)
    P = fb.new_particles(x, g, s)
    for kernel in ("singular", "gaussianerf"):
        D, _ = _eval(P, kernel=kernel, uj="direct")
        Ud, Jd = D[:, 9:12], D[:, 15:24]
        for p in (3, 4, 5):
            F, st = _eval(P, kernel=kernel, uj="fmm", fmm_p=p, fmm_ncrit=50, fmm_theta=0.4)
            Uo, Jo, so = o.fmm_uj(kernel, x, g, s, p=p, ncrit=50, theta=0.4, leaf_sigmas=4.0)
            assert abs(st["cells"] - so["cells"]) <= 0.02 * so["cells"] + 2            # the same octree
            eg = (rel_l2(F[:, 9:12], Ud), rel_l2(F[:, 15:24], Jd))
            eo = (rel_l2(Uo, Ud), rel_l2(Jo, Jd))
            if kernel == "singular":
                assert eg[0] <= 1.2 * eo[0] and eg[1] <= 1.2 * eo[1], (field, p, eg, eo)
                assert eo[0] < 0.1 and eg[0] < 0.1
            else:
                assert eg[0] <= 1.25 * eo[0] and eg[1] <= 1.25 * eo[1], (field, p, eg, eo)


def test_error_decreases_with_order_and_theta():
    P = _field(20_000, seed=8)
    D, _ = _eval(P, uj="direct")
    errs = {}
    for p in (2, 4, 6):
        F, _ = _eval(P, uj="fmm", fmm_p=p, fmm_theta=0.4, fmm_nonzero_sigma=1)
        errs[p] = rel_l2(F[:, 9:12], D[:, 9:12])
    assert errs[6] < 0.5 * errs[4], errs
    assert errs[4] < 0.5 * errs[2], errs
    F3, _ = _eval(P, uj="fmm", fmm_p=4, fmm_theta=0.25, fmm_nonzero_sigma=1)
    assert rel_l2(F3[:, 9:12], D[:, 9:12]) <= errs[4]   # equal when the 5-sigma clearance, not theta, decides
    assert errs[6] < 1e-3


def test_accumulate_and_clustered_points():
    """reset = False accumulates; 300 coincident particles exceed ncrit at the deepest level (a 21-level chain of
    single-child cells, a leaf larger than a warp) and must not break the tree.  Singular kernel: isolates the tree from
    the regularisation error of nonzero_sigma = false (sigma is 17 % of the box in this tiny field)."""
    import flowunsteady_b200 as fb
    P = _field(2000, seed=12)
    P[:300, 0:3] = P[0, 0:3]
    with fb.Engine(2000, schemes=fb.default_schemes(uj="fmm", fmm_theta=0.4, kernel="singular")) as eng:
        eng.upload(P)
        eng.uj()
        a = eng.download(np.zeros_like(P)).copy()
        eng.uj(reset=False)
        b = eng.download(np.zeros_like(P)).copy()
    D, _ = _eval(P, uj="direct", kernel="singular")
    assert np.all(np.isfinite(a))
    assert rel_l2(a[:, 9:12], D[:, 9:12]) < 5e-3
    assert relmax(b[:, 9:12], 2 * a[:, 9:12]) < 1e-14


def test_nextstep_with_fmm_tracks_direct():
    """One RK3 + dynamic-SFS + pedrizzetti step driven by UJ_fmm stays within the FMM truncation error of the direct path."""
    import flowunsteady_b200 as fb
    from flowunsteady_b200 import fields
    x, g, s = fields.vortex_rings(30_000)
    P = fb.new_particles(x, g, s)
    res = {}
    for uj in ("direct", "fmm"):
        sch = fb.default_schemes(uj=uj, sfs="dynamic", force_positive=1, clippings=1, fmm_nonzero_sigma=1)
        with fb.Engine(P.shape[0], schemes=sch) as eng:
            eng.upload(P)
            eng.nextstep(5e-3, relax=True)
            assert eng.count_nonfinite() == 0
            res[uj] = eng.download(np.zeros_like(P))
    dx_d = res["direct"][:, 0:3] - P[:, 0:3]
    dx_f = res["fmm"][:, 0:3] - P[:, 0:3]
    assert rel_l2(dx_f, dx_d) < 5e-3
    assert rel_l2(res["fmm"][:, 3:6], res["direct"][:, 3:6]) < 1e-3


def test_dynamic_sfs_far_field_reuse_is_bit_exact():
    """DynamicSFS evaluates twice at the same positions (test filter, domain filter).  vpmb200_nextstep lets the second
    UJ_fmm evaluation reuse the tree, the lists and the local expansions of the first (engine.cu do_sfs / fmm_hint); the
    same step driven stage by stage through the ABI rebuilds everything.  Both must give identical bits."""
    import flowunsteady_b200 as fb
    from flowunsteady_b200 import engine as E, fields
    from flowunsteady_b200.dist import RK3
    x, g, s = fields.vortex_rings(40_000)
    P = fb.new_particles(x, g, s)
    sch = dict(uj="fmm", sfs="dynamic", alpha=0.999, force_positive=1, clippings=1, controls=3, fmm_nonzero_sigma=0)
    dt, Uinf = 2e-3, (0.1, 0.0, 0.0)
    with fb.Engine(P.shape[0], schemes=fb.default_schemes(**sch)) as eng:
        eng.upload(P)
        eng.nextstep(dt, Uinf, relax=True)
        A = eng.download(np.zeros_like(P)).copy()
    with fb.Engine(P.shape[0], schemes=fb.default_schemes(**sch)) as eng:
        eng.upload(P)
        eng.stage(E.STAGE_ZERO_M)
        for a, b in RK3:
            if a == 0.0:
                eng.stage(E.STAGE_SCALE_SIGMA_TEST)
                eng.uj(True, True, True)
                eng.stage(E.STAGE_STORE_TEST)
                eng.stage(E.STAGE_SCALE_SIGMA_DOMAIN)
                eng.uj(True, True, True)
                eng.stage(E.STAGE_DYNAMIC_COEFF)
                eng.stage(E.STAGE_CLIP_CONTROL)
            else:
                eng.uj(True, True, True)
            eng.stage(E.STAGE_UPDATE, a, b, dt, Uinf)
        eng.uj(True, False, False)
        eng.stage(E.STAGE_RELAX)
        B = eng.download(np.zeros_like(P)).copy()
    assert np.all(np.isfinite(A))
    assert np.array_equal(A, B)


def test_fmm_parameter_validation():
    import flowunsteady_b200 as fb
    P = _field(100)
    with fb.Engine(100) as eng:
        eng.upload(P)
        for bad in (dict(fmm_p=9), dict(fmm_ncrit=1000), dict(fmm_theta=1.5)):
            eng.set_schemes(fb.default_schemes(uj="fmm", **bad))
            with pytest.raises(fb.EngineError):
                eng.uj()


def test_table_layouts_agree_bit_for_bit():
    """The replicated (bank-conflict-free) and the single-copy G table hold the same coefficients and the kernels run the
    same arithmetic in the same order: identical bits at the reference defaults, every expansion order."""
    import flowunsteady_b200 as fb
    from flowunsteady_b200 import fields
    x, g, s = fields.vortex_rings(30_000)
    P = fb.new_particles(x, g, s)
    for p in (3, 4, 5):
        a, _ = _eval(P, uj="fmm", fmm_p=p, sfs=True, options={"fmm_table_copies": 1})
        b, _ = _eval(P, uj="fmm", fmm_p=p, sfs=True, options={"fmm_table_copies": 8})
        assert np.array_equal(a, b)
    with fb.Engine(10) as eng, pytest.raises(fb.EngineError):
        eng.set_option("fmm_table_copies", 4)


def test_fmm_is_deterministic():
    """Two evaluations of the same field give bit-identical U, J, SFS (sorted pair lists, fixed-order reductions)."""
    import flowunsteady_b200 as fb
    from flowunsteady_b200 import fields
    x, g, s = fields.random_field(30_000)
    P = fb.new_particles(x, g, s)
    outs = []
    for _ in range(2):
        with fb.Engine(P.shape[0], schemes=fb.default_schemes(uj="fmm")) as eng:
            eng.upload(P)
            eng.uj(True, True, True)
            outs.append(eng.download(np.zeros_like(P)))
    assert np.array_equal(outs[0], outs[1])

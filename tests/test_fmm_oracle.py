"""The FMM oracle (oracle/fmm_oracle.c): the reference's ExaFMM-style method restated on the CPU — solid harmonics of degree
< p for multipoles and locals, ncrit octree, dual tree traversal with (R_i + R_j) < theta |c_i - c_j|, regularised near field.
No golden vector exists for the reference's FMM (PARITY UNPINNED, oracle header), so the oracle is pinned by what an FMM must
satisfy: theta -> 0 is the direct sum to round-off, the truncation error falls geometrically with p, U and J are consistent
(J = grad U), and the identities it is built on hold (tools/sh_identities.py)."""
import numpy as np
import pytest

from flowunsteady_b200 import fields
from oracle import oracle as o


def rel_l2(a, b):
    return float(np.linalg.norm(a - b) / np.linalg.norm(b))


@pytest.fixture(scope="module")
def ring_field():
    x, g, s = fields.vortex_rings(6000)
    return x, fields.floor_gamma(g), s


def test_theta_to_zero_is_the_direct_sum(ring_field):
    x, g, s = ring_field
    for kernel in ("gaussianerf", "singular"):
        Ud, Jd = o.uj_direct(kernel, x, g, s, x, accum=0)
        U, J, st = o.fmm_uj(kernel, x, g, s, p=4, ncrit=50, theta=1e-3)
        assert st["m2l_pairs"] == 0 and st["p2p_pairs"] == st["leaves"] ** 2
        assert rel_l2(U, Ud) < 1e-13 and rel_l2(J, Jd) < 1e-13


def test_truncation_error_falls_geometrically_with_p(ring_field):
    """Pure truncation (singular kernel near and far): one order of magnitude per two degrees at theta = 0.4."""
    x, g, s = ring_field
    Ud, Jd = o.uj_direct("singular", x, g, s, x, accum=0)
    eU, eJ = [], []
    for p in (2, 4, 6, 8):
        U, J, _ = o.fmm_uj("singular", x, g, s, p=p, ncrit=50, theta=0.4)
        eU.append(rel_l2(U, Ud))
        eJ.append(rel_l2(J, Jd))
    assert all(b < 0.25 * a for a, b in zip(eU, eU[1:])), eU
    assert all(b < 0.25 * a for a, b in zip(eJ, eJ[1:])), eJ
    assert eU[1] < 1e-2 and eU[3] < 1e-4           # p = 4 (the reference's default) and p = 8
    # a tighter acceptance converges faster
    U, J, _ = o.fmm_uj("singular", x, g, s, p=4, ncrit=50, theta=0.25)
    assert rel_l2(U, Ud) < 0.3 * eU[1]


def test_far_field_J_is_the_gradient_of_the_far_field_U():
    """The local expansion's second derivatives against central differences of its first derivatives: two probe particles of
    zero strength a step apart see the same sources through the same local expansion."""
    rng = np.random.default_rng(3)
    n = 3000
    xs = rng.random((n, 3))
    gs = rng.standard_normal((n, 3)) / n
    ss = np.full(n, 0.02)
    h = 1e-5
    base = np.array([3.0, 2.5, 2.8]) + 0.01 * rng.random((40, 3))          # far from the cloud: pure far field
    pts = [base] + [base + h * e for e in np.eye(3)] + [base - h * e for e in np.eye(3)]
    X = np.concatenate([xs] + pts)
    G = np.concatenate([gs, np.full((7 * 40, 3), 1e-300)])
    S = np.concatenate([ss, np.full(7 * 40, 0.02)])
    U, J, st = o.fmm_uj("singular", X, G, S, p=6, ncrit=50, theta=0.4)
    assert st["m2l_pairs"] > 0
    U0, J0 = U[n:n + 40], J[n:n + 40]
    for jd in range(3):
        dU = (U[n + 40 * (1 + jd):n + 40 * (2 + jd)] - U[n + 40 * (4 + jd):n + 40 * (5 + jd)]) / (2 * h)
        assert np.abs(dU - J0[:, 3 * jd:3 * jd + 3]).max() < 2e-3 * np.abs(J0).max()


def test_sparse_leaf_refinement_only_moves_work_to_the_far_field():
    """leaf_sigmas = 4 (what the CUDA path does): a stray particle in an empty octant next to a dense cloud becomes a small
    leaf; the result stays within the method's error of the plain tree and the near-field list shrinks."""
    rng = np.random.default_rng(5)
    n = 4000
    x = np.concatenate([rng.random((n, 3)) * 0.25, [[0.9, 0.9, 0.9], [0.6, 0.1, 0.1]]])
    g = np.concatenate([rng.standard_normal((n, 3)) / n, np.full((2, 3), 1e-6)])
    s = np.full(n + 2, 0.01)
    Ud, Jd = o.uj_direct("gaussianerf", x, g, s, x, accum=0)
    U0, J0, st0 = o.fmm_uj("gaussianerf", x, g, s, p=4, ncrit=50, theta=0.4)
    U1, J1, st1 = o.fmm_uj("gaussianerf", x, g, s, p=4, ncrit=50, theta=0.4, leaf_sigmas=4.0)
    assert st1["cells"] > st0["cells"] and st1["p2p_pairs"] < st0["p2p_pairs"]
    assert rel_l2(U1, Ud) < 2 * rel_l2(U0, Ud) + 1e-4

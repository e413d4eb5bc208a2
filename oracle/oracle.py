"""ctypes binding of libvpm_oracle.so (the C/OpenMP restatement) — TEST INFRASTRUCTURE ONLY.

PARITY UNPINNED: the reference's implementation of this path lives in the un-vendored Julia package
FLOWVPM; nothing in /root/reference can be executed for it.  The oracle is validated against analytic
identities, 50-digit mpmath evaluations (tests/golden/) and the one in-tree P2P,
src/FLOWUnsteady_processing_force.jl:879-929 (restated as `ffv_direct`).
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "libvpm_oracle.so")

NFIELDS = 43
X, GAMMA, SIGMA, VOL, CIRC, U, W, J, PSE, M, CC, SFS, STATIC = 0, 3, 6, 7, 8, 9, 12, 15, 24, 27, 36, 39, 42

KERNELS = {"gaussianerf": 0, "winckelmans": 1, "gaussian": 2, "singular": 3}
RELAX = {"none": 0, "pedrizzetti": 1, "correctedpedrizzetti": 2}
SFS_IDS = {"none": 0, "constant": 1, "dynamic": 2}
CLIP_BACKSCATTER = 1
CTRL_DIRECTIONAL, CTRL_MAGNITUDE, CTRL_SIGMASENSOR = 1, 2, 4


class Schemes(C.Structure):
    _fields_ = [
        ("kernel", C.c_int32), ("f", C.c_double), ("g", C.c_double), ("transposed", C.c_int32),
        ("relaxation", C.c_int32), ("rlxf", C.c_double), ("sfs", C.c_int32), ("alpha", C.c_double),
        ("sfs_rlxf", C.c_double), ("minC", C.c_double), ("maxC", C.c_double), ("Cs", C.c_double),
        ("force_positive", C.c_int32), ("clippings", C.c_int32), ("controls", C.c_int32),
        ("viscous", C.c_int32), ("nu", C.c_double), ("integration", C.c_int32),
        ("cs_sgm0", C.c_double), ("cs_beta", C.c_double), ("cs_itmax", C.c_int32), ("cs_tol", C.c_double),
    ]


def build(force: bool = False) -> str:
    """Compile the oracle with the committed Makefile (gcc + OpenMP)."""
    srcs = [os.path.join(_HERE, f) for f in ("vpm_oracle.c", "fmm_oracle.c", "vpm_oracle.h")]
    if force or not os.path.exists(_LIB_PATH) or os.path.getmtime(_LIB_PATH) < max(os.path.getmtime(f) for f in srcs):
        subprocess.check_call(["make", "-B", "-C", _HERE, "libvpm_oracle.so"], stdout=subprocess.DEVNULL)
    return _LIB_PATH


_lib = None
_dp = np.ctypeslib.ndpointer(dtype=np.float64, flags="C_CONTIGUOUS")


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_LIB_PATH):
            build()
        L = C.CDLL(_LIB_PATH)
        L.vpmo_default_schemes.argtypes = [C.POINTER(Schemes)]
        L.vpmo_g_dgdr.argtypes = [C.c_int32, C.c_double, C.POINTER(C.c_double), C.POINTER(C.c_double)]
        L.vpmo_zeta.argtypes = [C.c_int32, C.c_double]
        L.vpmo_zeta.restype = C.c_double
        L.vpmo_uj_direct.argtypes = [C.c_int32, C.c_int64, _dp, _dp, _dp, C.c_int64, _dp, _dp, _dp, C.c_int32]
        L.vpmo_estr_direct.argtypes = [C.c_int32, C.c_int32, C.c_int64, _dp, _dp, _dp, _dp, C.c_int64, _dp, _dp,
                                       _dp, C.c_int32]
        L.vpmo_ffv_direct.argtypes = [C.c_int32, C.c_int64, _dp, _dp, _dp, C.c_int64, _dp, _dp, _dp]
        L.vpmo_reset_particles.argtypes = [_dp, C.c_int64]
        L.vpmo_reset_particles_sfs.argtypes = [_dp, C.c_int64]
        L.vpmo_field_uj.argtypes = [_dp, C.c_int64, C.POINTER(Schemes), C.c_int32, C.c_int32, C.c_int32]
        L.vpmo_field_sfs.argtypes = [_dp, C.c_int64, C.POINTER(Schemes), C.c_double, C.c_double, C.c_double,
                                     C.c_int64]
        L.vpmo_nextstep.argtypes = [_dp, C.c_int64, C.POINTER(Schemes), C.c_double, _dp, C.c_int32,
                                    C.POINTER(C.c_double), C.POINTER(C.c_int64)]
        L.vpmo_relax_particle.argtypes = [_dp, C.c_int32, C.c_double]
        L.vpmo_update_particle.argtypes = [_dp, C.POINTER(Schemes), C.c_double, C.c_double, C.c_double, _dp,
                                           C.c_double]
        L.vpmo_zeta_direct.argtypes = [C.c_int32, C.c_int64, _dp, _dp, _dp, C.c_int64, _dp, _dp]
        L.vpmo_corespreading_reset.argtypes = [_dp, C.c_int64, C.c_int32, C.c_double, C.c_double, C.c_int32, C.c_double, _dp]
        L.vpmo_corespreading_reset.restype = C.c_int32
        L.vpmo_fmm_uj.argtypes = [C.c_int32, C.c_int32, C.c_int32, C.c_double, C.c_double, C.c_int64, _dp, _dp, _dp, _dp, _dp,
                                  C.POINTER(C.c_int64)]
        L.vpmo_fmm_uj.restype = C.c_int32
        L.vpmo_num_threads.restype = C.c_int32
        L.vpmo_set_num_threads.argtypes = [C.c_int32]
        _lib = L
    return _lib


def default_schemes(**kw) -> Schemes:
    s = Schemes()
    lib().vpmo_default_schemes(C.byref(s))
    for k, v in kw.items():
        if k == "kernel" and isinstance(v, str):
            v = KERNELS[v]
        if k == "relaxation" and isinstance(v, str):
            v = RELAX[v]
        if k == "sfs" and isinstance(v, str):
            v = SFS_IDS[v]
        if k == "integration" and isinstance(v, str):
            v = {"euler": 0, "rungekutta3": 1}[v]
        if k == "viscous" and isinstance(v, str):
            v = {"inviscid": 0, "corespreading": 1}[v]
        if not hasattr(s, k):
            raise AttributeError(k)
        setattr(s, k, v)
    return s


def _kid(kernel) -> int:
    return KERNELS[kernel] if isinstance(kernel, str) else int(kernel)


def g_dgdr(kernel, r: float):
    g, dg = C.c_double(), C.c_double()
    lib().vpmo_g_dgdr(_kid(kernel), float(r), C.byref(g), C.byref(dg))
    return g.value, dg.value


def zeta(kernel, r: float) -> float:
    return lib().vpmo_zeta(_kid(kernel), float(r))


def _c(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def uj_direct(kernel, xs, gs, sig, xt, accum: int = 0):
    """U (nt,3) and J (nt,9; J[i + 3 j] = du_i/dx_j) at targets xt from sources (xs, gs, sig)."""
    xs, gs, sig, xt = _c(xs), _c(gs), _c(sig), _c(xt)
    ns, nt = xs.shape[0], xt.shape[0]
    Uo = np.zeros((nt, 3))
    Jo = np.zeros((nt, 9))
    lib().vpmo_uj_direct(_kid(kernel), ns, xs, gs, sig, nt, xt, Uo, Jo, accum)
    return Uo, Jo


def estr_direct(kernel, transposed, xs, gs, sig, Js, xt, Jt, accum: int = 0):
    xs, gs, sig, Js, xt, Jt = _c(xs), _c(gs), _c(sig), _c(Js), _c(xt), _c(Jt)
    E = np.zeros((xt.shape[0], 3))
    lib().vpmo_estr_direct(_kid(kernel), int(bool(transposed)), xs.shape[0], xs, gs, sig, Js, xt.shape[0], xt, Jt,
                           E, accum)
    return E


def ffv_direct(kernel, xb, gb, sb, xf, gf):
    xb, gb, sb, xf, gf = _c(xb), _c(gb), _c(sb), _c(xf), _c(gf)
    M6 = np.zeros((xb.shape[0], 6))
    lib().vpmo_ffv_direct(_kid(kernel), xb.shape[0], xb, gb, sb, xf.shape[0], xf, gf, M6)
    return M6


def new_field(x, gamma, sigma, static=None, vol=None, circulation=None) -> np.ndarray:
    """(np, 43) C-contiguous array == the reference's 43 x np column-major particle matrix."""
    x = np.asarray(x, dtype=np.float64)
    n = x.shape[0]
    P = np.zeros((n, NFIELDS))
    P[:, X:X + 3] = x
    P[:, GAMMA:GAMMA + 3] = gamma
    P[:, SIGMA] = sigma
    if vol is not None:
        P[:, VOL] = vol
    if circulation is not None:
        P[:, CIRC] = circulation
    if static is not None:
        P[:, STATIC] = np.asarray(static, dtype=np.float64)
    return P


def field_uj(P, schemes, reset=True, reset_sfs=False, sfs=False):
    assert P.flags.c_contiguous and P.shape[1] == NFIELDS
    lib().vpmo_field_uj(P, P.shape[0], C.byref(schemes), int(reset), int(reset_sfs), int(sfs))
    return P


def field_sfs(P, schemes, a=1.0, b=1.0, t=0.0, nt=0):
    lib().vpmo_field_sfs(P, P.shape[0], C.byref(schemes), a, b, t, nt)
    return P


def nextstep(P, schemes, dt, Uinf=(0.0, 0.0, 0.0), relax=True, t=0.0, nt=0):
    tt, nn = C.c_double(t), C.c_int64(nt)
    lib().vpmo_nextstep(P, P.shape[0], C.byref(schemes), dt, _c(Uinf), int(relax), C.byref(tt), C.byref(nn))
    return tt.value, nn.value


def zeta_direct(kernel, xs, vs, sig, xt):
    xs, vs, sig, xt = _c(xs), _c(vs), _c(sig), _c(xt)
    out = np.zeros((xt.shape[0], 3))
    lib().vpmo_zeta_direct(_kid(kernel), xs.shape[0], xs, vs, sig, xt.shape[0], xt, out)
    return out


def corespreading_reset(P, kernel, sgm0, beta=1.5, itmax=15, tol=1e-3):
    res = np.zeros(3)
    it = lib().vpmo_corespreading_reset(P, P.shape[0], _kid(kernel), sgm0, beta, itmax, tol, res)
    return it, res


def fmm_uj(kernel, x, g, sig, p=4, ncrit=50, theta=0.4, leaf_sigmas=0.0):
    """The reference's FMM restated (fmm_oracle.c): U (n,3), J (n,9) at every particle; also returns tree statistics."""
    x, g, sig = _c(x), _c(g), _c(sig)
    n = x.shape[0]
    Uo, Jo = np.zeros((n, 3)), np.zeros((n, 9))
    st = (C.c_int64 * 4)()
    rc = lib().vpmo_fmm_uj(_kid(kernel), int(p), int(ncrit), float(theta), float(leaf_sigmas), n, x, g, sig, Uo, Jo, st)
    if rc != 0:
        raise ValueError("vpmo_fmm_uj: bad arguments")
    return Uo, Jo, dict(zip(("cells", "leaves", "m2l_pairs", "p2p_pairs"), [int(v) for v in st]))


def num_threads() -> int:
    return lib().vpmo_num_threads()


def set_num_threads(n: int) -> None:
    lib().vpmo_set_num_threads(int(n))

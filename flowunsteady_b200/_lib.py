"""ctypes loader for libvpmb200.so — the C-ABI engine declared in include/vpmb200.h.

There is no CPU fallback: if the library is missing this raises, and `vpmb200_create` itself fails with
VPMB200_ENODEVICE when no sm_100 GPU is present.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libvpmb200.so")

NFIELDS = 43

# error codes (include/vpmb200.h)
OK, EINVAL, ENODEVICE, ECUDA, ECAPACITY, ENOTSUP = 0, -1, -2, -3, -4, -5


class Schemes(C.Structure):
    """vpmb200_schemes (include/vpmb200.h)."""
    _fields_ = [
        ("kernel", C.c_int32), ("f", C.c_double), ("g", C.c_double), ("transposed", C.c_int32),
        ("relaxation", C.c_int32), ("rlxf", C.c_double), ("sfs", C.c_int32), ("alpha", C.c_double),
        ("sfs_rlxf", C.c_double), ("minC", C.c_double), ("maxC", C.c_double), ("Cs", C.c_double),
        ("force_positive", C.c_int32), ("clippings", C.c_int32), ("controls", C.c_int32),
        ("viscous", C.c_int32), ("nu", C.c_double), ("integration", C.c_int32),
        ("cs_sgm0", C.c_double), ("cs_beta", C.c_double), ("cs_itmax", C.c_int32), ("cs_tol", C.c_double),
        ("uj", C.c_int32), ("fmm_p", C.c_int32), ("fmm_ncrit", C.c_int32), ("fmm_theta", C.c_double),
        ("fmm_nonzero_sigma", C.c_int32),
    ]


# every symbol include/vpmb200.h declares: name -> (restype, argtypes)
_H = C.c_void_p
_dp = C.POINTER(C.c_double)
SYMBOLS = {
    "vpmb200_version": (C.c_char_p, []),
    "vpmb200_default_schemes": (C.c_int32, [C.POINTER(Schemes)]),
    "vpmb200_create": (C.c_int32, [C.c_int64, C.c_int32, C.c_int32, C.c_int32, C.POINTER(_H)]),
    "vpmb200_destroy": (C.c_int32, [_H]),
    "vpmb200_last_error": (C.c_char_p, [_H]),
    "vpmb200_set_schemes": (C.c_int32, [_H, C.POINTER(Schemes)]),
    "vpmb200_get_schemes": (C.c_int32, [_H, C.POINTER(Schemes)]),
    "vpmb200_set_time": (C.c_int32, [_H, C.c_double, C.c_int64]),
    "vpmb200_get_time": (C.c_int32, [_H, _dp, C.POINTER(C.c_int64)]),
    "vpmb200_upload": (C.c_int32, [_H, C.c_void_p, C.c_int64, C.c_int64, C.c_uint32]),
    "vpmb200_download": (C.c_int32, [_H, C.c_void_p, C.c_int64, C.c_int64, C.c_uint32]),
    "vpmb200_get_np": (C.c_int32, [_H, C.POINTER(C.c_int64)]),
    "vpmb200_host_register": (C.c_int32, [C.c_void_p, C.c_uint64]),
    "vpmb200_host_unregister": (C.c_int32, [C.c_void_p]),
    "vpmb200_add_particles": (C.c_int32, [_H, C.c_void_p, C.c_int64, C.c_int64]),
    "vpmb200_remove_particle": (C.c_int32, [_H, C.c_int64]),
    "vpmb200_remove_where": (C.c_int32, [_H, C.c_int32, C.c_void_p, C.POINTER(C.c_int64)]),
    "vpmb200_zeta": (C.c_int32, [_H]),
    "vpmb200_corespreading_reset": (C.c_int32, [_H, C.POINTER(C.c_int32), _dp]),
    "vpmb200_monitors": (C.c_int32, [_H, _dp]),
    "vpmb200_reset_particles": (C.c_int32, [_H]),
    "vpmb200_reset_particles_sfs": (C.c_int32, [_H]),
    "vpmb200_uj": (C.c_int32, [_H, C.c_int32, C.c_int32, C.c_int32]),
    "vpmb200_uj_probe": (C.c_int32, [_H, C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p]),
    "vpmb200_sfs": (C.c_int32, [_H, C.c_double, C.c_double]),
    "vpmb200_nextstep": (C.c_int32, [_H, C.c_double, C.c_void_p, C.c_int32]),
    "vpmb200_count_nonfinite": (C.c_int32, [_H, C.POINTER(C.c_int64)]),
    "vpmb200_device_field": (C.c_int32, [_H, C.c_int32, C.POINTER(C.c_void_p), C.POINTER(C.c_int64)]),
    "vpmb200_stream": (C.c_int32, [_H, C.POINTER(C.c_void_p)]),
    "vpmb200_synchronize": (C.c_int32, [_H]),
    "vpmb200_fmm_global": (C.c_int32, [_H, C.c_void_p, C.c_int64, C.c_int64, C.c_int32, C.c_int32, C.c_int32]),
    "vpmb200_set_option": (C.c_int32, [_H, C.c_char_p, C.c_int64]),
    "vpmb200_fmm_stats": (C.c_int32, [_H, C.POINTER(C.c_int64)]),
    "vpmb200_direct_tile_stats": (C.c_int32, [_H, C.POINTER(C.c_int64)]),
    "vpmb200_uj_probe_ex": (C.c_int32, [_H, C.c_void_p, C.c_int64, C.c_double, C.c_int32, C.c_void_p, C.c_void_p]),
    "vpmb200_set_statics": (C.c_int32, [_H, C.c_void_p, C.c_int64, C.c_int64, C.c_int64]),
    "vpmb200_get_statics": (C.c_int32, [_H, C.POINTER(C.c_int64), C.POINTER(C.c_int64)]),
    "vpmb200_set_mirror": (C.c_int32, [_H, C.c_int32, _dp, _dp]),
    "vpmb200_multi_create": (C.c_int32, [C.c_int64, C.c_int32, C.c_int32, C.c_int32, C.POINTER(C.c_int32), C.POINTER(_H)]),
    "vpmb200_multi_destroy": (C.c_int32, [_H]),
    "vpmb200_multi_last_error": (C.c_char_p, [_H]),
    "vpmb200_multi_set_schemes": (C.c_int32, [_H, C.POINTER(Schemes)]),
    "vpmb200_multi_set_time": (C.c_int32, [_H, C.c_double, C.c_int64]),
    "vpmb200_multi_get_time": (C.c_int32, [_H, _dp, C.POINTER(C.c_int64)]),
    "vpmb200_multi_get_np": (C.c_int32, [_H, C.POINTER(C.c_int64)]),
    "vpmb200_multi_shard_sizes": (C.c_int32, [_H, C.POINTER(C.c_int64)]),
    "vpmb200_multi_upload": (C.c_int32, [_H, C.c_void_p, C.c_int64, C.c_int64, C.c_uint32]),
    "vpmb200_multi_download": (C.c_int32, [_H, C.c_void_p, C.c_int64, C.c_int64, C.c_uint32]),
    "vpmb200_multi_add_particles": (C.c_int32, [_H, C.c_void_p, C.c_int64, C.c_int64]),
    "vpmb200_multi_remove_particle": (C.c_int32, [_H, C.c_int64]),
    "vpmb200_multi_remove_where": (C.c_int32, [_H, C.c_int32, C.c_void_p, C.POINTER(C.c_int64)]),
    "vpmb200_multi_rebalance": (C.c_int32, [_H, C.c_double, C.POINTER(C.c_int64)]),
    "vpmb200_multi_set_statics": (C.c_int32, [_H, C.c_void_p, C.c_int64, C.c_int64, C.c_int64]),
    "vpmb200_multi_uj": (C.c_int32, [_H, C.c_int32, C.c_int32, C.c_int32]),
    "vpmb200_multi_sfs": (C.c_int32, [_H, C.c_double, C.c_double]),
    "vpmb200_multi_nextstep": (C.c_int32, [_H, C.c_double, _dp, C.c_int32]),
    "vpmb200_multi_uj_probe": (C.c_int32, [_H, C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p]),
    "vpmb200_multi_synchronize": (C.c_int32, [_H]),
    "vpmb200_multi_engine": (C.c_int32, [_H, C.c_int32, C.POINTER(_H)]),
    "vpmb200_let_cell_bytes": (C.c_int32, []),
    "vpmb200_let_bounds": (C.c_int32, [_H, _dp]),
    "vpmb200_let_keys": (C.c_int32, [_H, _dp, C.c_int32, C.POINTER(C.c_void_p), C.POINTER(C.c_void_p)]),
    "vpmb200_let_partition": (C.c_int32, [_H, C.c_int32, C.c_int32, C.c_int32, C.POINTER(C.c_int64)]),
    "vpmb200_let_work": (C.c_int32, [_H, C.POINTER(C.c_void_p)]),
    "vpmb200_let_cut": (C.c_int32, [C.POINTER(C.c_int32), C.POINTER(C.c_int64), C.c_int32, C.c_int32, C.c_int32, C.POINTER(C.c_uint64)]),
    "vpmb200_let_pack": (C.c_int32, [_H, C.c_void_p]),
    "vpmb200_let_build": (C.c_int32, [_H, C.c_void_p, C.c_int64, C.c_int64, C.c_int32, C.POINTER(C.c_int64)]),
    "vpmb200_let_ptrs": (C.c_int32, [_H, C.POINTER(C.c_void_p)]),
    "vpmb200_let_attach_tree": (C.c_int32, [_H, C.c_void_p, C.c_void_p, C.c_int64, C.POINTER(C.c_int64), C.POINTER(C.c_int64)]),
    "vpmb200_let_attach_records": (C.c_int32, [_H, C.c_void_p, C.c_int64, C.POINTER(C.c_int64)]),
    "vpmb200_let_attach_skeleton": (C.c_int32, [_H, C.c_void_p, C.c_int64, C.POINTER(C.c_int64), C.POINTER(C.c_int64), C.POINTER(C.c_int64)]),
    "vpmb200_let_halo_plan": (C.c_int32, [_H, C.POINTER(C.c_int64), C.POINTER(C.c_void_p), C.POINTER(C.c_void_p)]),
    "vpmb200_let_halo_serve": (C.c_int32, [_H, C.c_void_p, C.c_int64, C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p]),
    "vpmb200_let_halo_set": (C.c_int32, [_H, C.c_void_p, C.c_void_p]),
    "vpmb200_let_evaluate": (C.c_int32, [_H, C.c_void_p, C.c_int32, C.c_int32]),
    "vpmb200_let_estr_records": (C.c_int32, [_H]),
    "vpmb200_let_estr_evaluate": (C.c_int32, [_H, C.c_void_p]),
    "vpmb200_let_finish": (C.c_int32, [_H, C.c_void_p, C.c_int32, C.c_int32]),
    "vpmb200_fmm_times": (C.c_int32, [_H, _dp]),
    "vpmb200_launch_count": (C.c_int32, [_H, C.POINTER(C.c_uint64)]),
    "vpmb200_tiles_for": (C.c_int64, [C.c_int64]),
    "vpmb200_tile_doubles": (C.c_int64, []),
    "vpmb200_pack_uj_records": (C.c_int32, [_H, C.c_void_p]),
    "vpmb200_pack_estr_records": (C.c_int32, [_H, C.c_void_p]),
    "vpmb200_uj_from_records": (C.c_int32, [_H, C.c_void_p, C.c_int64, C.c_int32]),
    "vpmb200_estr_from_records": (C.c_int32, [_H, C.c_void_p, C.c_int64]),
    "vpmb200_measure_fp64_peak": (C.c_int32, [C.c_int32, C.c_int32, C.c_int32, _dp, _dp]),
    "vpmb200_measure_fp64_peak2": (C.c_int32, [C.c_int32, C.c_int32, C.c_int32, _dp]),
    "vpmb200_stage": (C.c_int32, [_H, C.c_int32, C.c_double, C.c_double, C.c_double, C.c_void_p]),
}


def build(force: bool = False) -> str:
    """Compile the engine with nvcc for sm_100a (flowunsteady_b200/csrc/Makefile)."""
    cs = os.path.join(_HERE, "csrc")
    args = ["make", "-C", cs] + (["-B"] if force else [])
    subprocess.check_call(args, stdout=subprocess.DEVNULL)
    return LIB_PATH


_lib = None


def lib():
    """The loaded engine library.  Raises (never falls back) when it has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'`. "
                "flowunsteady_b200 has no CPU fallback.")
        L = C.CDLL(LIB_PATH)
        for name, (res, args) in SYMBOLS.items():
            fn = getattr(L, name)  # AttributeError if the .so does not export a declared symbol
            fn.restype = res
            fn.argtypes = args
        _lib = L
    return _lib


class EngineError(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__(f"vpmb200 error {code}: {msg}")
        self.code = code

"""Thin object wrapper over the C ABI (include/vpmb200.h).  One Engine = one GPU-resident particle field.

Host arrays use the reference's particle matrix: `particles[i, :]` is particle i's 43-field column, i.e. a
C-contiguous (np, 43) numpy array is byte-identical to FLOWVPM's 43 x np column-major `pfield.particles`
(/root/reference/src/FLOWUnsteady_simulation.jl:509-510).
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib
from ._lib import NFIELDS, EngineError, Schemes

# 0-based field offsets (include/vpmb200.h)
X, GAMMA, SIGMA, VOL, CIRCULATION, U, VORTICITY, J, PSE, M, CC, SFS, STATIC = 0, 3, 6, 7, 8, 9, 12, 15, 24, 27, 36, 39, 42

FM_X, FM_GAMMA, FM_SIGMA, FM_VOL, FM_CIRCULATION, FM_U, FM_VORTICITY, FM_J, FM_PSE, FM_M, FM_C, FM_SFS, FM_STATIC = (
    1 << k for k in range(13))
FM_ALL = 0x1FFF
FM_STATE = FM_X | FM_GAMMA | FM_SIGMA | FM_VOL | FM_CIRCULATION | FM_C | FM_STATIC

KERNEL_IDS = {"gaussianerf": 0, "winckelmans": 1, "gaussian": 2, "singular": 3}
RELAX_IDS = {"none": 0, "pedrizzetti": 1, "correctedpedrizzetti": 2}
SFS_IDS = {"none": 0, "constant": 1, "dynamic": 2}
VISCOUS_IDS = {"inviscid": 0, "corespreading": 1}
INTEGRATION_IDS = {"euler": 0, "rungekutta3": 1}
UJ_IDS = {"direct": 0, "fmm": 1}
CLIP_BACKSCATTER = 1
CTRL_DIRECTIONAL, CTRL_MAGNITUDE = 1, 2

STAGE_SCALE_SIGMA_TEST, STAGE_STORE_TEST, STAGE_SCALE_SIGMA_DOMAIN, STAGE_DYNAMIC_COEFF = 1, 2, 3, 4
STAGE_CONSTANT_COEFF, STAGE_CLIP_CONTROL, STAGE_ZERO_M, STAGE_UPDATE, STAGE_RELAX, STAGE_UPDATE_EULER_RELAX = 5, 6, 7, 8, 9, 10

_ENUMS = {"kernel": KERNEL_IDS, "relaxation": RELAX_IDS, "sfs": SFS_IDS, "viscous": VISCOUS_IDS,
          "integration": INTEGRATION_IDS, "uj": UJ_IDS}


def default_schemes(**kw) -> Schemes:
    s = Schemes()
    _lib.lib().vpmb200_default_schemes(C.byref(s))
    for k, v in kw.items():
        if isinstance(v, str):
            v = _ENUMS[k][v]
        if not hasattr(s, k):
            raise AttributeError(f"vpmb200_schemes has no member {k!r}")
        setattr(s, k, v)
    return s


def _as_matrix(a: np.ndarray) -> np.ndarray:
    if isinstance(a, np.ndarray) and a.ndim == 2 and a.shape[0] == 0 and a.dtype == np.float64 and a.shape[1] >= NFIELDS:
        return a                                   # an empty shard (strides of a 0-row array carry no information)
    if not (isinstance(a, np.ndarray) and a.dtype == np.float64 and a.ndim == 2 and a.shape[1] >= NFIELDS
            and a.strides[1] == 8 and a.strides[0] >= 8 * NFIELDS and a.strides[0] % 8 == 0):
        raise ValueError("particles must be a float64 (np, >=43) array with contiguous rows")
    return a


class Engine:
    """RAII wrapper of a vpmb200_handle."""

    def __init__(self, max_particles: int, float_bits: int = 64, device: int = 0, schemes: Schemes | None = None):
        self._L = _lib.lib()
        self._h = C.c_void_p()
        rc = self._L.vpmb200_create(int(max_particles), NFIELDS, int(float_bits), int(device), C.byref(self._h))
        if rc != 0:
            msg = self._L.vpmb200_last_error(None).decode()
            self._h = None
            raise EngineError(rc, msg)
        self.max_particles = int(max_particles)
        self.float_bits = int(float_bits)
        self.device = int(device)
        if schemes is not None:
            self.set_schemes(schemes)

    # ---- lifetime -------------------------------------------------------------------------------------
    def close(self):
        if getattr(self, "_h", None):
            self._L.vpmb200_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    def _check(self, rc: int):
        if rc != 0:
            raise EngineError(rc, self._L.vpmb200_last_error(self._h).decode())

    # ---- schemes / time -------------------------------------------------------------------------------
    def set_schemes(self, s: Schemes):
        self._check(self._L.vpmb200_set_schemes(self._h, C.byref(s)))

    def get_schemes(self) -> Schemes:
        s = Schemes()
        self._check(self._L.vpmb200_get_schemes(self._h, C.byref(s)))
        return s

    def set_time(self, t: float, nt: int):
        self._check(self._L.vpmb200_set_time(self._h, float(t), int(nt)))

    def get_time(self):
        t, nt = C.c_double(), C.c_int64()
        self._check(self._L.vpmb200_get_time(self._h, C.byref(t), C.byref(nt)))
        return t.value, nt.value

    # ---- data movement --------------------------------------------------------------------------------
    @property
    def np(self) -> int:
        n = C.c_int64()
        self._check(self._L.vpmb200_get_np(self._h, C.byref(n)))
        return n.value

    def upload(self, particles: np.ndarray, np_: int | None = None, field_mask: int = FM_ALL):
        P = _as_matrix(particles)
        n = P.shape[0] if np_ is None else int(np_)
        ld = P.strides[0] // 8 if P.shape[0] else NFIELDS
        self._check(self._L.vpmb200_upload(self._h, P.ctypes.data if n else None, ld, n, field_mask))

    def download(self, particles: np.ndarray, np_: int | None = None, field_mask: int = FM_ALL):
        P = _as_matrix(particles)
        n = self.np if np_ is None else int(np_)
        if n > P.shape[0]:
            raise ValueError("host matrix is too small")
        self._check(self._L.vpmb200_download(self._h, P.ctypes.data, P.strides[0] // 8, n, field_mask))
        return P

    def add_particles(self, cols: np.ndarray):
        P = _as_matrix(np.atleast_2d(cols))
        self._check(self._L.vpmb200_add_particles(self._h, P.ctypes.data, P.strides[0] // 8, P.shape[0]))

    def remove_particle(self, i: int):
        self._check(self._L.vpmb200_remove_particle(self._h, int(i)))

    REMOVE_STRENGTH, REMOVE_SIGMA, REMOVE_BOX, REMOVE_SPHERE = 1, 2, 3, 4

    def remove_where(self, criterion: int, params) -> int:
        """Wake treatment on the device (include/vpmb200.h: vpmb200_remove_where); returns the number removed."""
        p = np.ascontiguousarray(params, dtype=np.float64)
        r = C.c_int64()
        self._check(self._L.vpmb200_remove_where(self._h, int(criterion), p.ctypes.data, C.byref(r)))
        return r.value

    def zeta(self):
        """W rows <- sum_q Gamma_q zeta_sigma_q(x_p - x_q)  (vpm.zeta_direct / zeta_fmm)."""
        self._check(self._L.vpmb200_zeta(self._h))

    def corespreading_reset(self):
        """CoreSpreading's spatial adaptation (sigma <- sgm0 + RBF conjugate-gradient re-fit of Gamma); returns
        (CG iterations, residual per component).  0 iterations = no particle had sigma/sgm0 > beta."""
        it = C.c_int32()
        res = (C.c_double * 3)()
        self._check(self._L.vpmb200_corespreading_reset(self._h, C.byref(it), res))
        return it.value, np.array(res[:])

    def monitors(self) -> dict:
        out = (C.c_double * 6)()
        self._check(self._L.vpmb200_monitors(self._h, out))
        return dict(zip(("enstrophy", "Cd_mean", "Cd_std", "Cd_count", "n_static", "sum_abs_Gamma"), [float(v) for v in out]))

    # ---- hot path -------------------------------------------------------------------------------------
    def reset_particles(self):
        self._check(self._L.vpmb200_reset_particles(self._h))

    def reset_particles_sfs(self):
        self._check(self._L.vpmb200_reset_particles_sfs(self._h))

    def uj(self, reset: bool = True, reset_sfs: bool = False, sfs: bool = False):
        self._check(self._L.vpmb200_uj(self._h, int(reset), int(reset_sfs), int(sfs)))

    def uj_probe(self, Xp: np.ndarray, want_J: bool = False):
        Xp = np.ascontiguousarray(Xp, dtype=np.float64).reshape(-1, 3)
        m = Xp.shape[0]
        Uo = np.zeros((m, 3))
        Jo = np.zeros((m, 9)) if want_J else None
        self._check(self._L.vpmb200_uj_probe(self._h, Xp.ctypes.data, m, Uo.ctypes.data,
                                             Jo.ctypes.data if want_J else None))
        return (Uo, Jo) if want_J else Uo

    def uj_probe_ex(self, Xp: np.ndarray, fsgm: float = 1.0, mirror: bool = False, want_J: bool = False):
        """Vvpm_on_Xs with its `fsgm` (reference quirk reproduced) and `mirror` options (include/vpmb200.h)."""
        Xp = np.ascontiguousarray(Xp, dtype=np.float64).reshape(-1, 3)
        m = Xp.shape[0]
        Uo = np.zeros((m, 3))
        Jo = np.zeros((m, 9)) if want_J else None
        self._check(self._L.vpmb200_uj_probe_ex(self._h, Xp.ctypes.data, m, float(fsgm), int(mirror), Uo.ctypes.data,
                                                Jo.ctypes.data if want_J else None))
        return (Uo, Jo) if want_J else Uo

    def set_statics(self, cols: np.ndarray, generation: int | None = None):
        """Park the step's static particles behind the field (vpmb200_set_statics); generation defaults to the current nt."""
        P = _as_matrix(np.atleast_2d(cols)) if len(cols) else np.zeros((0, NFIELDS))
        gen = self.get_time()[1] if generation is None else int(generation)
        self._check(self._L.vpmb200_set_statics(self._h, P.ctypes.data if P.shape[0] else None, P.strides[0] // 8 if P.shape[0] else NFIELDS,
                                                P.shape[0], gen))

    def get_statics(self):
        n, g = C.c_int64(), C.c_int64()
        self._check(self._L.vpmb200_get_statics(self._h, C.byref(n), C.byref(g)))
        return n.value, g.value

    def set_mirror(self, enabled: bool, X0=(0.0, 0.0, 0.0), normal=(0.0, 0.0, 1.0)):
        a, b = (C.c_double * 3)(*[float(v) for v in X0]), (C.c_double * 3)(*[float(v) for v in normal])
        self._check(self._L.vpmb200_set_mirror(self._h, int(enabled), a, b))

    def sfs(self, a: float = 1.0, b: float = 1.0):
        self._check(self._L.vpmb200_sfs(self._h, float(a), float(b)))

    def nextstep(self, dt: float, Uinf=(0.0, 0.0, 0.0), relax: bool = True):
        u = (C.c_double * 3)(*[float(v) for v in Uinf])
        self._check(self._L.vpmb200_nextstep(self._h, float(dt), u, int(relax)))

    def stage(self, stage: int, a: float = 0.0, b: float = 0.0, dt: float = 0.0, Uinf=None):
        u = (C.c_double * 3)(*[float(v) for v in Uinf]) if Uinf is not None else None
        self._check(self._L.vpmb200_stage(self._h, int(stage), float(a), float(b), float(dt), u))

    def count_nonfinite(self) -> int:
        c = C.c_int64()
        self._check(self._L.vpmb200_count_nonfinite(self._h, C.byref(c)))
        return c.value

    # ---- device-level hooks ---------------------------------------------------------------------------
    def device_field(self, field: int):
        p, ld = C.c_void_p(), C.c_int64()
        self._check(self._L.vpmb200_device_field(self._h, int(field), C.byref(p), C.byref(ld)))
        return p.value, ld.value

    @property
    def stream(self) -> int:
        s = C.c_void_p()
        self._check(self._L.vpmb200_stream(self._h, C.byref(s)))
        return s.value or 0

    def fmm_global(self, G_ptr: int, ldg: int, ntot: int, part: int, nparts: int, pass_: int):
        self._check(self._L.vpmb200_fmm_global(self._h, C.c_void_p(G_ptr), int(ldg), int(ntot), int(part), int(nparts), int(pass_)))

    # ---- multi-GPU UJ_fmm: local-essential-tree phases (include/vpmb200.h: vpmb200_let_*) ----------------
    def let_bounds(self) -> np.ndarray:
        o = (C.c_double * 6)()
        self._check(self._L.vpmb200_let_bounds(self._h, o))
        return np.array(o[:])

    def let_keys(self, lohi_global, Lc: int):
        """-> (device pointer of the int32 histogram [8^Lc], device pointer of the per-bin sigma max or None)."""
        g = (C.c_double * 6)(*[float(v) for v in lohi_global])
        h, m = C.c_void_p(), C.c_void_p()
        self._check(self._L.vpmb200_let_keys(self._h, g, int(Lc), C.byref(h), C.byref(m)))
        return h.value, m.value

    def let_partition(self, nparts: int, part: int, use_work: bool = False):
        sc = (C.c_int64 * nparts)()
        self._check(self._L.vpmb200_let_partition(self._h, int(nparts), int(part), int(use_work), sc))
        return [int(v) for v in sc]

    def let_work(self) -> int:
        """Device pointer of the int64 per-bin work counts of the last evaluation (0 before the first let_keys)."""
        p = C.c_void_p()
        self._check(self._L.vpmb200_let_work(self._h, C.byref(p)))
        return p.value or 0

    def let_pack(self, rows_ptr: int):
        self._check(self._L.vpmb200_let_pack(self._h, C.c_void_p(rows_ptr)))

    def let_build(self, rows_ptr: int, n_own: int, n_all: int, reuse: bool = False):
        info = (C.c_int64 * 4)()
        self._check(self._L.vpmb200_let_build(self._h, C.c_void_p(rows_ptr), int(n_own), int(n_all), int(reuse), info))
        return [int(v) for v in info]

    def let_ptrs(self):
        p = (C.c_void_p * 3)()
        self._check(self._L.vpmb200_let_ptrs(self._h, p))
        return [v or 0 for v in p]

    def let_attach_tree(self, cells_ptr: int, M_ptr: int, slot_cells: int, ncells, nparticles):
        g = len(ncells)
        a, b = (C.c_int64 * g)(*[int(v) for v in ncells]), (C.c_int64 * g)(*[int(v) for v in nparticles])
        self._check(self._L.vpmb200_let_attach_tree(self._h, C.c_void_p(cells_ptr), C.c_void_p(M_ptr), int(slot_cells), a, b))

    def let_attach_skeleton(self, cells_ptr: int, slot_cells: int, ncells, nparticles, nleaves):
        g = len(ncells)
        a, b, c = ((C.c_int64 * g)(*[int(v) for v in x]) for x in (ncells, nparticles, nleaves))
        self._check(self._L.vpmb200_let_attach_skeleton(self._h, C.c_void_p(cells_ptr), int(slot_cells), a, b, c))

    def let_halo_plan(self, nparts: int):
        """-> (counts3 [3 * nparts], device pointer of the requested cell ids, of the requested (start, count) leaf pairs)"""
        cnt = (C.c_int64 * (3 * nparts))()
        pc, pl = C.c_void_p(), C.c_void_p()
        self._check(self._L.vpmb200_let_halo_plan(self._h, cnt, C.byref(pc), C.byref(pl)))
        return [int(v) for v in cnt], pc.value or 0, pl.value or 0

    def let_halo_serve(self, req_cells_ptr: int, ncell: int, req_leaf_ptr: int, nleaf: int, M_out_ptr: int, rec_out_ptr: int):
        self._check(self._L.vpmb200_let_halo_serve(self._h, C.c_void_p(req_cells_ptr), int(ncell), C.c_void_p(req_leaf_ptr), int(nleaf),
                                                   C.c_void_p(M_out_ptr) if M_out_ptr else None,
                                                   C.c_void_p(rec_out_ptr) if rec_out_ptr else None))

    def let_halo_set(self, M2_ptr: int, rec2_ptr: int):
        self._check(self._L.vpmb200_let_halo_set(self._h, C.c_void_p(M2_ptr) if M2_ptr else None,
                                                 C.c_void_p(rec2_ptr) if rec2_ptr else None))

    def let_attach_records(self, rec_ptr: int, slot_n: int, nparticles):
        b = (C.c_int64 * len(nparticles))(*[int(v) for v in nparticles])
        self._check(self._L.vpmb200_let_attach_records(self._h, C.c_void_p(rec_ptr), int(slot_n), b))

    def let_evaluate(self, out_ptr: int, reuse: bool = False, stage: int = 0):
        self._check(self._L.vpmb200_let_evaluate(self._h, C.c_void_p(out_ptr), int(reuse), int(stage)))

    def let_estr_records(self):
        self._check(self._L.vpmb200_let_estr_records(self._h))

    def let_estr_evaluate(self, out_ptr: int):
        self._check(self._L.vpmb200_let_estr_evaluate(self._h, C.c_void_p(out_ptr)))

    def let_finish(self, res_ptr: int, what: int, reset: bool):
        self._check(self._L.vpmb200_let_finish(self._h, C.c_void_p(res_ptr), int(what), int(reset)))

    @staticmethod
    def let_cell_bytes() -> int:
        return _lib.lib().vpmb200_let_cell_bytes()

    def set_option(self, name: str, value: int):
        self._check(self._L.vpmb200_set_option(self._h, name.encode(), int(value)))

    def fmm_stats(self) -> dict:
        a = (C.c_int64 * 5)()
        self._check(self._L.vpmb200_fmm_stats(self._h, a))
        return dict(zip(("cells", "leaves", "levels", "m2l_pairs", "p2p_pairs"), [int(v) for v in a]))

    def direct_tile_stats(self) -> dict:
        """(target block, source tile) classification of the direct path for the current field (instrumentation)."""
        a = (C.c_int64 * 4)()
        self._check(self._L.vpmb200_direct_tile_stats(self._h, a))
        d = dict(zip(("target_blocks", "source_tiles", "far_pairs", "all_pairs"), [int(v) for v in a]))
        d["tile_far_fraction"] = d["far_pairs"] / d["all_pairs"] if d["all_pairs"] else 0.0
        return d

    def fmm_times(self) -> dict:
        """Device ms of the sections of the last UJ_fmm evaluation (vpmb200_fmm_times)."""
        a = (C.c_double * 6)()
        self._check(self._L.vpmb200_fmm_times(self._h, a))
        return dict(zip(("tree", "lists", "upward", "m2l_l2l", "l2p_near", "estr_near"), [round(float(v), 3) for v in a]))

    @property
    def launch_count(self) -> int:
        c = C.c_uint64()
        self._check(self._L.vpmb200_launch_count(self._h, C.byref(c)))
        return c.value

    def synchronize(self):
        self._check(self._L.vpmb200_synchronize(self._h))

    @staticmethod
    def tiles_for(n: int) -> int:
        return _lib.lib().vpmb200_tiles_for(int(n))

    @staticmethod
    def tile_doubles() -> int:
        return _lib.lib().vpmb200_tile_doubles()

    def pack_uj_records(self, dst_ptr: int):
        self._check(self._L.vpmb200_pack_uj_records(self._h, C.c_void_p(dst_ptr)))

    def pack_estr_records(self, dst_ptr: int):
        self._check(self._L.vpmb200_pack_estr_records(self._h, C.c_void_p(dst_ptr)))

    def uj_from_records(self, tiles_ptr: int, ntiles: int, accumulate: bool):
        self._check(self._L.vpmb200_uj_from_records(self._h, C.c_void_p(tiles_ptr), int(ntiles), int(accumulate)))

    def estr_from_records(self, tiles_ptr: int, ntiles: int):
        self._check(self._L.vpmb200_estr_from_records(self._h, C.c_void_p(tiles_ptr), int(ntiles)))


class MultiEngine:
    """RAII wrapper of a vpmb200_multi_handle: ONE host thread, several GPUs (include/vpmb200.h: vpmb200_multi_*).

    Same call shapes as `Engine` on the GLOBAL particle order of the host's matrix; particles are sharded over `ngpus`
    per-device engines (devices=None: 0 .. ngpus-1; an ordinal may repeat — several shards on one GPU)."""

    def __init__(self, max_particles: int, ngpus: int, devices=None, float_bits: int = 64, schemes: Schemes | None = None):
        self._L = _lib.lib()
        self._h = C.c_void_p()
        dv = (C.c_int32 * ngpus)(*[int(d) for d in devices]) if devices is not None else None
        rc = self._L.vpmb200_multi_create(int(max_particles), NFIELDS, int(float_bits), int(ngpus), dv, C.byref(self._h))
        if rc != 0:
            msg = self._L.vpmb200_multi_last_error(None).decode()
            self._h = None
            raise EngineError(rc, msg)
        self.ngpus = int(ngpus)
        if schemes is not None:
            self.set_schemes(schemes)

    def close(self):
        if getattr(self, "_h", None):
            self._L.vpmb200_multi_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    def _check(self, rc: int):
        if rc != 0:
            raise EngineError(rc, self._L.vpmb200_multi_last_error(self._h).decode())

    def set_schemes(self, s: Schemes):
        self._check(self._L.vpmb200_multi_set_schemes(self._h, C.byref(s)))

    def set_time(self, t: float, nt: int):
        self._check(self._L.vpmb200_multi_set_time(self._h, float(t), int(nt)))

    def get_time(self):
        t, nt = C.c_double(), C.c_int64()
        self._check(self._L.vpmb200_multi_get_time(self._h, C.byref(t), C.byref(nt)))
        return t.value, nt.value

    @property
    def np(self) -> int:
        n = C.c_int64()
        self._check(self._L.vpmb200_multi_get_np(self._h, C.byref(n)))
        return n.value

    def shard_sizes(self):
        a = (C.c_int64 * self.ngpus)()
        self._check(self._L.vpmb200_multi_shard_sizes(self._h, a))
        return [int(v) for v in a]

    def upload(self, particles: np.ndarray, np_: int | None = None, field_mask: int = FM_ALL):
        P = _as_matrix(particles)
        n = P.shape[0] if np_ is None else int(np_)
        ld = P.strides[0] // 8 if P.shape[0] else NFIELDS
        self._check(self._L.vpmb200_multi_upload(self._h, P.ctypes.data if n else None, ld, n, field_mask))

    def download(self, particles: np.ndarray, np_: int | None = None, field_mask: int = FM_ALL):
        P = _as_matrix(particles)
        n = self.np if np_ is None else int(np_)
        if n > P.shape[0]:
            raise ValueError("host matrix is too small")
        ld = P.strides[0] // 8 if P.shape[0] else NFIELDS
        self._check(self._L.vpmb200_multi_download(self._h, P.ctypes.data if n else None, ld, n, field_mask))
        return P

    def add_particles(self, cols: np.ndarray):
        P = _as_matrix(np.atleast_2d(cols))
        self._check(self._L.vpmb200_multi_add_particles(self._h, P.ctypes.data, P.strides[0] // 8, P.shape[0]))

    def remove_particle(self, i: int):
        self._check(self._L.vpmb200_multi_remove_particle(self._h, int(i)))

    def remove_where(self, criterion: int, params) -> int:
        p = np.ascontiguousarray(params, dtype=np.float64)
        r = C.c_int64()
        self._check(self._L.vpmb200_multi_remove_where(self._h, int(criterion), p.ctypes.data, C.byref(r)))
        return r.value

    def set_statics(self, cols: np.ndarray, generation: int | None = None):
        P = _as_matrix(np.atleast_2d(cols)) if len(cols) else np.zeros((0, NFIELDS))
        gen = self.get_time()[1] if generation is None else int(generation)
        self._check(self._L.vpmb200_multi_set_statics(self._h, P.ctypes.data if P.shape[0] else None,
                                                      P.strides[0] // 8 if P.shape[0] else NFIELDS, P.shape[0], gen))

    def rebalance(self, tolerance: float = 0.05) -> int:
        moved = C.c_int64()
        self._check(self._L.vpmb200_multi_rebalance(self._h, float(tolerance), C.byref(moved)))
        return moved.value

    def uj(self, reset: bool = True, reset_sfs: bool = False, sfs: bool = False):
        self._check(self._L.vpmb200_multi_uj(self._h, int(reset), int(reset_sfs), int(sfs)))

    def sfs(self, a: float = 1.0, b: float = 1.0):
        self._check(self._L.vpmb200_multi_sfs(self._h, float(a), float(b)))

    def nextstep(self, dt: float, Uinf=(0.0, 0.0, 0.0), relax: bool = True):
        u = (C.c_double * 3)(*[float(v) for v in Uinf])
        self._check(self._L.vpmb200_multi_nextstep(self._h, float(dt), u, int(relax)))

    def uj_probe(self, Xp: np.ndarray, want_J: bool = False):
        Xp = np.ascontiguousarray(Xp, dtype=np.float64).reshape(-1, 3)
        m = Xp.shape[0]
        Uo = np.zeros((m, 3))
        Jo = np.zeros((m, 9)) if want_J else None
        self._check(self._L.vpmb200_multi_uj_probe(self._h, Xp.ctypes.data, m, Uo.ctypes.data, Jo.ctypes.data if want_J else None))
        return (Uo, Jo) if want_J else Uo

    def synchronize(self):
        self._check(self._L.vpmb200_multi_synchronize(self._h))


def new_particles(x, gamma, sigma, static=None, vol=None, circulation=None, C_=None) -> np.ndarray:
    """(n, 43) particle matrix with X, Gamma, sigma (and optional vol, circulation, C, static) filled in."""
    x = np.asarray(x, dtype=np.float64).reshape(-1, 3)
    n = x.shape[0]
    P = np.zeros((n, NFIELDS))
    P[:, X:X + 3] = x
    P[:, GAMMA:GAMMA + 3] = gamma
    P[:, SIGMA] = sigma
    if vol is not None:
        P[:, VOL] = vol
    if circulation is not None:
        P[:, CIRCULATION] = circulation
    if C_ is not None:
        P[:, CC:CC + 3] = C_
    if static is not None:
        P[:, STATIC] = np.asarray(static, dtype=np.float64)
    return P

"""Host-side mirror of the FLOWVPM interface that FLOWUnsteady drives, backed by the GPU engine.

FLOWUnsteady selects the hot path through scheme *objects* stored in `vpm.ParticleField` and calls it as
`pfield.UJ(pfield)` (/root/reference/src/FLOWUnsteady_simulation.jl:544, src/FLOWUnsteady_processing_force.jl:238) and
`vpm.nextstep(pfield, dt; relax)` (simulation.jl:358).  This module keeps those names, argument meanings and error
behaviour, with Python standing in for Julia (no julia binary exists in this image; the ccall glue a maintainer would
add is in INTEGRATION.md and julia/FLOWVPMB200.jl).  Differences forced by the language: indices are 0-based, keyword
`ε_tol` is spelled `eps_tol`, and `pfield.particles` is a (maxparticles, 43) row-major numpy array — byte-identical to
the reference's 43 x maxparticles column-major matrix.

Every numerical call goes to libvpmb200.so (CUDA); there is no CPU implementation here.  The Kernel objects expose
`g_dgdr` etc. as small host functions only because the reference's API has them (processing_force.jl:866-868).
"""
from __future__ import annotations

import math
import time
from dataclasses import dataclass, field as _dc_field
from typing import Callable, Iterator, Optional, Sequence

import numpy as np

from . import engine as _E
from .engine import Engine, EngineError

# ---- particle-matrix row indices (vpm.X_INDEX ... in FLOWVPM v4; 0-based here) -------------------------------
X_INDEX = slice(_E.X, _E.X + 3)
GAMMA_INDEX = slice(_E.GAMMA, _E.GAMMA + 3)
SIGMA_INDEX = _E.SIGMA
VOL_INDEX = _E.VOL
CIRCULATION_INDEX = _E.CIRCULATION
U_INDEX = slice(_E.U, _E.U + 3)
VORTICITY_INDEX = slice(_E.VORTICITY, _E.VORTICITY + 3)
J_INDEX = slice(_E.J, _E.J + 9)
PSE_INDEX = slice(_E.PSE, _E.PSE + 3)
M_INDEX = slice(_E.M, _E.M + 9)
C_INDEX = slice(_E.CC, _E.CC + 3)
SFS_INDEX = slice(_E.SFS, _E.SFS + 3)
STATIC_INDEX = _E.STATIC
NFIELDS = _E.NFIELDS

_C1 = 1.0 / (2.0 * math.pi) ** 1.5
_C2 = math.sqrt(2.0 / math.pi)
_C4 = 1.0 / (4.0 * math.pi)


# ---- kernels (vpm.Kernel) ------------------------------------------------------------------------
@dataclass(frozen=True)
class Kernel:
    name: str
    id: int
    zeta: Callable[[float], float]
    g: Callable[[float], float]
    dgdr: Callable[[float], float]
    g_dgdr: Callable[[float], tuple]


def _gauserf(r):
    aux = _C2 * r * math.exp(-r * r / 2)
    return math.erf(r / math.sqrt(2)) - aux, r * aux


def _wnk(r):
    a0 = (r * r + 1) ** 2.5
    return r ** 3 * (r * r + 2.5) / a0, 7.5 * r * r / (a0 * (r * r + 1))


def _gaus(r):
    e = math.exp(-r ** 3)
    return 1 - e, 3 * r * r * e


gaussianerf = Kernel("gaussianerf", 0, lambda r: _C1 * math.exp(-r * r / 2), lambda r: _gauserf(r)[0],
                     lambda r: _gauserf(r)[1], _gauserf)
winckelmans = Kernel("winckelmans", 1, lambda r: _C4 * 7.5 / (r * r + 1) ** 3.5, lambda r: _wnk(r)[0],
                     lambda r: _wnk(r)[1], _wnk)
gaussian = Kernel("gaussian", 2, lambda r: 3 * _C4 * math.exp(-r ** 3), lambda r: _gaus(r)[0], lambda r: _gaus(r)[1], _gaus)
singular = Kernel("singular", 3, lambda r: 1.0 if r == 0 else 0.0, lambda r: 1.0, lambda r: 0.0, lambda r: (1.0, 0.0))
kernel_default = gaussianerf


# ---- formulations (vpm.rVPM / vpm.cVPM; rvpm.md:197-239) --------------------------------------------------------
@dataclass(frozen=True)
class Formulation:
    f: float
    g: float


formulation_rVPM = rVPM = Formulation(0.0, 1.0 / 5.0)
formulation_cVPM = cVPM = Formulation(0.0, 0.0)
formulation_default = rVPM


# ---- viscous schemes ---------------------------------------------------------------------------------------------
class ViscousScheme:
    pass


@dataclass
class Inviscid(ViscousScheme):
    nu: float = 0.0


@dataclass
class CoreSpreading(ViscousScheme):
    """vpm.CoreSpreading(nu, sgm0, zeta; beta, itmax, tol) (examples/rotorhover/rotorhover.jl:170).  The engine applies
    the sigma update of SURVEY.md A.8 every substep and, once the last substep is done, the spatial adaptation: when any
    sigma/sgm0 > beta the cores are reset to sgm0 and Gamma is re-fitted by RBF conjugate gradients (itmax, tol) so the
    particle-approximated vorticity is preserved."""
    nu: float
    sgm0: float
    zeta: object = None
    beta: float = 1.5
    itmax: int = 15
    tol: float = 1e-3


@dataclass
class ParticleStrengthExchange(ViscousScheme):
    """vpm.ParticleStrengthExchange(nu) — named in FLOWVPM's API but used by none of FLOWUnsteady's
    examples (they run Inviscid(), or CoreSpreading in commented lines).  The engine has no PSE kernel: a field constructed
    with it raises NotImplementedError instead of silently running inviscid."""
    nu: float = 0.0


def isinviscid(v) -> bool:
    return isinstance(v, Inviscid)


def iscorespreading(v) -> bool:
    return isinstance(v, CoreSpreading)


def zeta_fmm(pfield):
    """vpm.zeta_fmm (passed as CoreSpreading's `zeta`, simulation.jl:232): vorticity sum_q Gamma_q zeta_sigma_q(x_p - x_q) at every
    particle, stored in the W rows.  Runs the near-field zeta pass of the engine."""
    pfield._call_zeta()


zeta_direct = zeta_fmm


def _kernel_compatibility(viscous) -> tuple:
    """Kernels a viscous scheme accepts (src/FLOWUnsteady_simulation.jl:308-314)."""
    return (gaussianerf,) if iscorespreading(viscous) else (gaussianerf, winckelmans, gaussian, singular)


# ---- relaxation --------------------------------------------------------------------------------------------------
@dataclass(frozen=True)
class Relaxation:
    name: str
    id: int
    nsteps_relax: int = 1
    rlxf: float = 0.3


pedrizzetti = relaxation_pedrizzeti = Relaxation("pedrizzetti", 1)
correctedpedrizzetti = relaxation_correctedpedrizzeti = Relaxation("correctedpedrizzetti", 2)
norelaxation = relaxation_none = Relaxation("none", 0, nsteps_relax=-1, rlxf=0.0)
relaxation_default = pedrizzetti


# ---- FMM settings (vpm.FMM; simulation.jl:43) ---------------------------------------------------------------------
@dataclass
class FMM:
    p: int = 4
    ncrit: int = 50
    theta: float = 0.4
    nonzero_sigma: bool = False
    eps_tol: Optional[float] = None


# ---- SFS schemes (examples/rotorhover/rotorhover.jl:53-55,172-182) ----------------------------------
def Estr_direct(pfield):
    pfield.UJ(pfield, sfs=True, reset=True, reset_sfs=True)


def Estr_fmm(pfield):
    pfield.UJ(pfield, sfs=True, reset=True, reset_sfs=True)


def pseudo3level(*a, **k):  # marker objects: the procedure itself runs on the GPU (K4)
    raise RuntimeError("pseudo3level is evaluated inside the engine; pass it to DynamicSFS")


def pseudo3level_positive(*a, **k):
    raise RuntimeError("pseudo3level_positive is evaluated inside the engine; pass it to DynamicSFS")


def clipping_backscatter(*a, **k):
    raise RuntimeError("clipping_backscatter is evaluated inside the engine; pass it in `clippings`")


def control_directional(*a, **k):
    raise RuntimeError("control_directional is evaluated inside the engine; pass it in `controls`")


def control_magnitude(*a, **k):
    raise RuntimeError("control_magnitude is evaluated inside the engine; pass it in `controls`")


def control_sigmasensor(*a, **k):
    raise RuntimeError("control_sigmasensor is not implemented (upstream form unverified)")


class SubFilterScale:
    id = 0

    def __call__(self, pfield, a: float = 1.0, b: float = 1.0):
        pfield._call_sfs(a, b)


class NoSFS(SubFilterScale):
    id = 0


@dataclass
class ConstantSFS(SubFilterScale):
    model: Callable = Estr_fmm
    Cs: float = 1.0
    clippings: Sequence = ()
    controls: Sequence = ()
    id = 1


@dataclass
class DynamicSFS(SubFilterScale):
    model: Callable = Estr_fmm
    procedure: Callable = pseudo3level_positive
    alpha: float = 0.667
    rlxf: float = 0.005
    minC: float = 0.0
    maxC: float = 1.0
    clippings: Sequence = ()
    controls: Sequence = ()
    id = 2


def isSFSenabled(sfs) -> bool:
    return not isinstance(sfs, NoSFS)


SFS_none = NoSFS()
SFS_Cs_nobackscatter = ConstantSFS(Estr_fmm, Cs=1.0, clippings=(clipping_backscatter,))
SFS_Cd_twolevel_nobackscatter = DynamicSFS(Estr_fmm, pseudo3level_positive, alpha=0.999, clippings=(clipping_backscatter,))
SFS_Cd_threelevel_nobackscatter = DynamicSFS(Estr_fmm, pseudo3level_positive, alpha=0.667, clippings=(clipping_backscatter,))
SFS_default = SFS_none


# ---- UJ schemes ----------------------------------------------------------------------------------------------------
def UJ_direct(pfield, reset: bool = True, reset_sfs: bool = False, sfs: bool = False, rbf: bool = False, **_):
    """pfield.UJ(pfield; reset, reset_sfs, sfs): U and J at every particle by direct P2P on the GPU (K1 [+ K2])."""
    if rbf:
        raise NotImplementedError("rbf=True (vorticity by zeta) is not built yet")
    pfield._call_uj(_E.UJ_IDS["direct"], reset, reset_sfs, sfs)


def UJ_fmm(pfield, reset: bool = True, reset_sfs: bool = False, sfs: bool = False, rbf: bool = False, sort: bool = True, **_):
    if rbf:
        raise NotImplementedError("rbf=True (vorticity by zeta) is not built yet")
    pfield._call_uj(_E.UJ_IDS["fmm"], reset, reset_sfs, sfs)


# ---- time integration ----------------------------------------------------------------------------------------------
def euler(pfield, dt: float, relax: bool = False, custom_UJ=None):
    pfield._call_nextstep(_E.INTEGRATION_IDS["euler"], dt, relax, custom_UJ)


def rungekutta3(pfield, dt: float, relax: bool = False, custom_UJ=None):
    pfield._call_nextstep(_E.INTEGRATION_IDS["rungekutta3"], dt, relax, custom_UJ)


# ---- the particle field --------------------------------------------------------------------------------------------
class Particle:
    """View of one particle column (vpm.get_particle): fields are numpy views into pfield.particles."""
    __slots__ = ("_c",)

    def __init__(self, col: np.ndarray):
        self._c = col

    X = property(lambda s: s._c[X_INDEX])
    Gamma = property(lambda s: s._c[GAMMA_INDEX])
    sigma = property(lambda s: s._c[SIGMA_INDEX:SIGMA_INDEX + 1])
    vol = property(lambda s: s._c[VOL_INDEX:VOL_INDEX + 1])
    circulation = property(lambda s: s._c[CIRCULATION_INDEX:CIRCULATION_INDEX + 1])
    U = property(lambda s: s._c[U_INDEX])
    J = property(lambda s: s._c[J_INDEX].reshape(3, 3).T)   # J[i, j] = du_i/dx_j
    M = property(lambda s: s._c[M_INDEX].reshape(3, 3).T)
    C = property(lambda s: s._c[C_INDEX])
    SFS = property(lambda s: s._c[SFS_INDEX])
    static = property(lambda s: bool(s._c[STATIC_INDEX] > 0))


def get_X(P): return P.X
def get_Gamma(P): return P.Gamma
def get_sigma(P): return P.sigma
def get_vol(P): return P.vol
def get_circulation(P): return P.circulation
def get_U(P): return P.U
def get_C(P): return P.C
def get_J(P): return P.J
def get_SFS(P): return P.SFS


class ParticleField:
    """vpm.ParticleField(maxparticles, R; Uinf, formulation, viscous, kernel, UJ, SFS, integration, transposed,
    relaxation, fmm)  — constructor call at src/FLOWUnsteady_simulation.jl:239-253.

    `particles` is host memory owned by the caller's side of the boundary (the reference's Julia matrix).  Each hot-path
    call uploads the groups the engine reads and downloads the groups it wrote (`sync="always"`, the drop-in
    behaviour), or keeps the field device-resident between calls (`sync="lazy"`: call `pull()` before reading and
    `mark_dirty()` after writing `particles` by hand).
    """

    def __init__(self, maxparticles: int, R=np.float64, *, Uinf: Callable[[float], Sequence[float]] = lambda t: (0.0, 0.0, 0.0),
                 formulation: Formulation = formulation_default, viscous: ViscousScheme = None, kernel: Kernel = kernel_default,
                 UJ: Callable = UJ_fmm, SFS: SubFilterScale = SFS_default, integration: Callable = rungekutta3,
                 transposed: bool = True, relaxation: Relaxation = relaxation_default, fmm: FMM = None,
                 device: int = 0, sync: str = "always", pinned: bool = False):
        if R not in (np.float64, np.float32, float):
            raise TypeError("R must be Float64 or Float32 (vpm_floattype, simulation.jl:137)")
        viscous = Inviscid() if viscous is None else viscous
        if isinstance(viscous, ParticleStrengthExchange):
            raise NotImplementedError("ParticleStrengthExchange is not available in the GPU engine (use Inviscid or CoreSpreading)")
        if kernel not in _kernel_compatibility(viscous):
            raise ValueError(f"Kernel {kernel.name} is not compatible with viscous scheme {type(viscous).__name__}")
        if sync not in ("always", "lazy"):
            raise ValueError("sync must be 'always' or 'lazy'")
        self.maxparticles = int(maxparticles)
        self.R = R
        self.particles = np.zeros((self.maxparticles, NFIELDS))
        self._registered = False
        if pinned:   # page-lock the matrix in place (what the Julia stub does with pfield.particles): full-rate DMA, no staging
            from . import _lib as _L
            rc = _L.lib().vpmb200_host_register(self.particles.ctypes.data, self.particles.nbytes)
            if rc != 0:
                raise EngineError(rc, _L.lib().vpmb200_last_error(None).decode())
            self._registered = True
        self.np = 0
        self.nt = 0
        self.t = 0.0
        self.Uinf = Uinf
        self.formulation, self.viscous, self.kernel = formulation, viscous, kernel
        self.UJ, self.SFS, self.integration = UJ, SFS, integration
        self.transposed, self.relaxation = bool(transposed), relaxation
        self.fmm = FMM() if fmm is None else fmm
        self.sync = sync
        self._engine = Engine(self.maxparticles, float_bits=32 if R is np.float32 else 64, device=device)
        self._host_dirty = True     # host matrix holds changes (to rows the device already has) the device has not seen
        self._dev_dirty = 0         # field-group mask the device holds newer than the host
        self._dev_np = 0            # particles the device holds; rows [_dev_np, np) are host-side appends not yet sent
        self._statics = None        # (columns, nt) of the static-particle fast path (set_statics)
        self.h2d_bytes = 0
        self.d2h_bytes = 0
        self.h2d_seconds = 0.0      # wall time inside upload calls
        self.d2h_seconds = 0.0      # wall time inside download calls (includes waiting for the enqueued step)

    # ---- scheme translation ------------------------------------------------------------------------------------
    def _schemes(self, uj_id: int, integration_id: Optional[int] = None):
        sfs = self.SFS
        if getattr(self.fmm, "eps_tol", None) is not None:
            raise NotImplementedError("vpm.FMM(eps_tol=...) (error-controlled acceptance) is not available in the GPU engine; "
                                      "use theta / p")
        kw = dict(kernel=self.kernel.id, f=self.formulation.f, g=self.formulation.g, transposed=int(self.transposed),
                  relaxation=self.relaxation.id, rlxf=self.relaxation.rlxf, sfs=sfs.id, uj=uj_id,
                  fmm_p=self.fmm.p, fmm_ncrit=self.fmm.ncrit, fmm_theta=self.fmm.theta,
                  fmm_nonzero_sigma=int(self.fmm.nonzero_sigma))
        if isinstance(sfs, (ConstantSFS, DynamicSFS)):
            clip = 0
            for c in sfs.clippings:
                if c is clipping_backscatter:
                    clip |= _E.CLIP_BACKSCATTER
                else:
                    raise NotImplementedError(f"clipping {getattr(c, '__name__', c)} is not available in the engine")
            ctrl = 0
            for c in sfs.controls:
                if c is control_directional:
                    ctrl |= _E.CTRL_DIRECTIONAL
                elif c is control_magnitude:
                    ctrl |= _E.CTRL_MAGNITUDE
                else:
                    raise NotImplementedError(f"control {getattr(c, '__name__', c)} is not available in the engine")
            kw.update(clippings=clip, controls=ctrl)
        if isinstance(sfs, ConstantSFS):
            kw.update(Cs=sfs.Cs)
        if isinstance(sfs, DynamicSFS):
            if sfs.procedure not in (pseudo3level, pseudo3level_positive):
                raise NotImplementedError("only the pseudo3level procedures are available in the engine")
            kw.update(alpha=sfs.alpha, sfs_rlxf=sfs.rlxf, minC=sfs.minC, maxC=sfs.maxC,
                      force_positive=int(sfs.procedure is pseudo3level_positive))
        if iscorespreading(self.viscous):
            kw.update(viscous=1, nu=self.viscous.nu, cs_sgm0=self.viscous.sgm0, cs_beta=self.viscous.beta,
                      cs_itmax=self.viscous.itmax, cs_tol=self.viscous.tol)
        if integration_id is not None:
            kw.update(integration=integration_id)
        return _E.default_schemes(**kw)

    def _uj_id(self) -> int:
        if self.UJ is UJ_direct:
            return _E.UJ_IDS["direct"]
        if self.UJ is UJ_fmm:
            return _E.UJ_IDS["fmm"]
        raise NotImplementedError("custom UJ callables cannot run inside the GPU engine; use vpm.UJ_direct or vpm.UJ_fmm")

    # ---- host <-> device synchronisation -------------------------------------------------------------------------
    def mark_dirty(self):
        """Tell the field that `particles` was modified on the host."""
        self._host_dirty = True

    def _device_is_truth(self) -> bool:
        """Lazy mode with nothing pending on the host side for the rows the device holds: rows [0, _dev_np) live on the GPU
        and the host copy of the `_dev_dirty` groups is stale until `pull()`."""
        return self.sync == "lazy" and not self._host_dirty

    def _flush_appends(self):
        """Send the particles appended on the host since the last call (vpm.add_particle) — only those rows cross the bus."""
        if self.np > self._dev_np:
            t0 = time.perf_counter()
            self._engine.add_particles(self.particles[self._dev_np:self.np])
            self.h2d_seconds += time.perf_counter() - t0
            self.h2d_bytes += (self.np - self._dev_np) * 8 * NFIELDS
            self._dev_np = self.np

    def _push(self, mask: int = _E.FM_ALL):
        if self.sync == "always" or self._host_dirty:
            if self.sync == "lazy" and self._dev_dirty:
                # the host copy of the device-newer groups is stale: they must not go back up (a shed between two
                # lazy steps used to overwrite X/Gamma/sigma with the previous step's values)
                mask &= ~self._dev_dirty
                if self._dev_np < self.np:       # rows the device never saw need every group
                    self.pull()
                    mask = _E.FM_ALL
            t0 = time.perf_counter()
            self._engine.upload(self.particles, self.np, mask)     # returns after the copy (borrowed pointer)
            self.h2d_seconds += time.perf_counter() - t0
            self.h2d_bytes += self.np * 8 * _rows(mask)            # exactly what crosses the bus (selected column runs)
            self._host_dirty = False
            self._dev_np = self.np
        else:
            self._flush_appends()
        self._engine.set_time(self.t, self.nt)

    def _pulled(self, mask: int):
        if self.sync == "always":
            self.pull(mask)
        else:
            self._dev_dirty |= mask

    def pull(self, mask: Optional[int] = None):
        """Bring device results back into `particles`."""
        mask = self._dev_dirty if mask is None else mask
        n = min(self.np, self._dev_np)      # rows [_dev_np, np) are host-side appends the device has not seen yet
        if mask and n > 0:
            t0 = time.perf_counter()
            self._engine.download(self.particles, n, mask)         # waits for the step, then copies
            self.d2h_seconds += time.perf_counter() - t0
            self.d2h_bytes += n * 8 * _rows(mask)
        self._dev_dirty &= ~mask

    # ---- static-particle fast path (simulation.jl:355-365 without add_particle / remove_particle) ------------------
    def set_statics(self, cols):
        """The embedded (static) particles of the CURRENT step — what `static_particles_function(pfield, t, dt)` would append
        (simulation.jl:355) — as an (n, 43) block of particle columns.  They are parked device-side behind the field for the
        next `nextstep` / `UJ` / `U_at` of this step and never enter `particles` (include/vpmb200.h: vpmb200_set_statics)."""
        cols = np.ascontiguousarray(np.atleast_2d(np.asarray(cols, dtype=np.float64)))
        if cols.shape[0] and cols.shape[1] != NFIELDS:
            raise ValueError("static particles must be (n, 43) particle columns")
        if self.np + cols.shape[0] > self.maxparticles:
            raise RuntimeError(f"PARTICLE OVERFLOW. Max number of particles {self.maxparticles} has been reached")
        self._statics = (cols.copy(), self.nt) if cols.shape[0] else None

    def _apply_statics(self):
        """After the field itself is on the device: park the step's statics behind it (a stale set is dropped)."""
        if self._statics is not None and self._statics[1] == self.nt:
            self._engine.set_statics(self._statics[0], self.nt)
            self.h2d_bytes += self._statics[0].shape[0] * 8 * NFIELDS
        else:
            self._statics = None

    # ---- hot-path entry points (called by the scheme objects) -----------------------------------------------------
    def _call_uj(self, uj_id: int, reset: bool, reset_sfs: bool, sfs: bool):
        self._engine.set_schemes(self._schemes(uj_id))
        self._push(_E.FM_ALL if not (reset and reset_sfs) else _E.FM_STATE | _E.FM_SFS | _E.FM_U | _E.FM_J)
        self._apply_statics()
        self._engine.uj(reset, reset_sfs, sfs)
        self._pulled(_E.FM_U | _E.FM_J | _E.FM_PSE | (_E.FM_SFS if (sfs or reset_sfs) else 0))

    def _call_zeta(self):
        self._engine.set_schemes(self._schemes(_E.UJ_IDS["direct"]))
        self._push(_E.FM_STATE)
        self._engine.zeta()
        self._pulled(_E.FM_VORTICITY)

    def _call_sfs(self, a: float, b: float):
        self._engine.set_schemes(self._schemes(self._uj_id()))
        self._push(_E.FM_ALL)
        self._engine.sfs(a, b)
        self._pulled(_E.FM_U | _E.FM_J | _E.FM_PSE | _E.FM_SFS | _E.FM_C | _E.FM_M | _E.FM_SIGMA)

    def _call_nextstep(self, integration_id: int, dt: float, relax: bool, custom_UJ):
        if custom_UJ is not None:
            raise NotImplementedError("custom_UJ cannot run inside the GPU engine")
        self._engine.set_schemes(self._schemes(self._uj_id(), integration_id))
        self._push(_E.FM_STATE | _E.FM_M)
        self._apply_statics()
        self._engine.nextstep(dt, tuple(self.Uinf(self.t)), relax)     # consumes the static set
        self._statics = None
        self._pulled(_E.FM_ALL & ~(_E.FM_VOL | _E.FM_CIRCULATION | _E.FM_STATIC))

    # ---- wake treatments / monitors on the device ---------------------------------------------------------------
    def remove_where(self, criterion: int, params) -> int:
        """Device-side compaction with the reference's resulting particle order (vpmb200_remove_where)."""
        if self.np == 0:
            return 0
        self._push(_E.FM_ALL)
        removed = self._engine.remove_where(criterion, params)
        self.np = self._dev_np = self._engine.np
        if removed:
            self._pulled(_E.FM_ALL)
        return removed

    def monitors(self) -> dict:
        """Enstrophy and C_d statistics reduced on the device (vpm.monitor_enstrophy / vpm.monitor_Cd)."""
        self._push(_E.FM_ALL)
        return self._engine.monitors()

    # ---- probes: Vvpm_on_Xs without evaluating every target (simulation.jl:494-570) -------------------------------
    def U_at(self, Xs, want_J: bool = False, fsgm: float = 1.0, mirror: bool = False):
        """Vvpm_on_Xs(pfield, Xs; fsgm, mirror) (simulation.jl:494-570) on the probe fast path; the static set given with
        `set_statics` stands in for `static_particles_fun`."""
        self._engine.set_schemes(self._schemes(_E.UJ_IDS["direct"]))
        self._push(_E.FM_STATE)
        self._apply_statics()
        return self._engine.uj_probe_ex(np.asarray(Xs, dtype=np.float64), fsgm, mirror, want_J)

    def set_mirror(self, enabled: bool, X0=(0.0, 0.0, 0.0), normal=(0.0, 0.0, 1.0)):
        """run_simulation's mirror / mirror_X / mirror_normal (simulation.jl:149-151): images join the static set on the device."""
        self._engine.set_mirror(enabled, X0, normal)

    def fluiddomain(self, Xs, method: str = "direct"):
        """U and W = curl u at arbitrary nodes (what vpm.computefluiddomain evaluates on its grids,
        examples/rotorhover/rotorhover_fluiddomain.jl:93-104).
        method="direct": m probes x np sources with the pair kernel (exact; right for m up to ~1e5 or small fields);
        method="fmm":    the reference's own way — the nodes join a scratch field as zero-strength particles (add_probe:
                         Gamma = 0, sigma = 1e-6, simulation.jl:572), UJ_fmm runs once on np + m particles, and U, J are
                         read at the nodes.  O(np + m): grids of 1e6..1e7 nodes (examples/vahana: 2.4e7) stay cheap."""
        Xs = np.ascontiguousarray(Xs, dtype=np.float64).reshape(-1, 3)
        if method == "direct":
            Uo, Jo = self.U_at(Xs, want_J=True)
        elif method == "fmm":
            if self._dev_dirty:
                self.pull()
            m, n = Xs.shape[0], self.np
            cols = np.zeros((n + m, NFIELDS))
            cols[:n] = self.particles[:n]
            cols[n:, X_INDEX] = Xs
            cols[n:, SIGMA_INDEX] = 1e-6
            with Engine(n + m, device=self._engine.device, schemes=self._schemes(_E.UJ_IDS["fmm"])) as scratch:
                scratch.upload(cols)
                scratch.uj()
                scratch.download(cols, field_mask=_E.FM_U | _E.FM_J)
            Uo, Jo = cols[n:, U_INDEX].copy(), cols[n:, J_INDEX].copy()
        else:
            raise ValueError("method must be 'direct' or 'fmm'")
        W = np.stack([Jo[:, 5] - Jo[:, 7], Jo[:, 6] - Jo[:, 2], Jo[:, 1] - Jo[:, 3]], -1)
        return Uo, W

    @property
    def engine(self) -> Engine:
        return self._engine

    def __del__(self):
        # release the page lock before numpy frees the matrix (the Julia stub's finalizer does the same)
        try:
            if getattr(self, "_registered", False):
                from . import _lib as _L
                _L.lib().vpmb200_host_unregister(self.particles.ctypes.data)
                self._registered = False
        except Exception:
            pass


def _rows(mask: int) -> int:
    counts = (3, 3, 1, 1, 1, 3, 3, 9, 3, 9, 3, 3, 1)
    return sum(c for k, c in enumerate(counts) if (mask >> k) & 1)


# ---- free functions of the FLOWVPM API ---------------------------------------------------------------------------------
def get_np(pfield: ParticleField) -> int:
    return pfield.np


def get_particle(pfield: ParticleField, i: int) -> Particle:
    if i < 0 or i >= pfield.np:
        raise IndexError(f"Requested invalid particle index {i}")
    return Particle(pfield.particles[i])


def iterator(pfield: ParticleField, start_i: int = 0, end_i: Optional[int] = None, include_static: bool = False) -> Iterator[Particle]:
    end_i = pfield.np if end_i is None else end_i
    for i in range(start_i, end_i):
        if include_static or not pfield.particles[i, STATIC_INDEX] > 0:
            yield Particle(pfield.particles[i])


iterate = iterator


def add_particle(pfield: ParticleField, X, Gamma=None, sigma=None, *, vol=0.0, circulation=1.0, C=0.0, static=False, index=-1):
    """vpm.add_particle(pfield, X, Gamma, sigma; vol, circulation, C, static, index) (simulation.jl:486,572) or
    vpm.add_particle(pfield, P) with a Particle."""
    if pfield.np == pfield.maxparticles:
        raise RuntimeError(f"PARTICLE OVERFLOW. Max number of particles {pfield.maxparticles} has been reached")
    col = pfield.particles[pfield.np]
    if isinstance(X, Particle):
        col[:] = X._c
    else:
        col[:] = 0.0
        col[X_INDEX] = X
        col[GAMMA_INDEX] = Gamma
        col[SIGMA_INDEX] = float(np.ravel(sigma)[0])
        col[VOL_INDEX] = float(np.ravel(vol)[0])
        col[CIRCULATION_INDEX] = abs(float(np.ravel(circulation)[0]))
        col[C_INDEX] = C
        col[STATIC_INDEX] = 1.0 if static else 0.0
    pfield.np += 1
    if not pfield._device_is_truth():
        pfield.mark_dirty()
    # lazy + device-resident field: the new row is a pending append (rows [_dev_np, np)); the next hot-path call sends
    # exactly those rows through vpmb200_add_particles and leaves the device-newer rows alone


def remove_particle(pfield: ParticleField, i: int):
    """vpm.remove_particle(pfield, i): the last particle is moved into slot i (0-based here)."""
    if i < 0 or i >= pfield.np:
        raise IndexError(f"Requested removal of invalid particle index {i}")
    if pfield._device_is_truth():
        # same swap-remove on both sides; the host copy of the device-newer groups stays stale until pull()
        pfield._flush_appends()
        pfield._engine.remove_particle(i)
        pfield._dev_np -= 1
    elif pfield._dev_dirty:
        pfield.pull()
        pfield.mark_dirty()
    else:
        pfield.mark_dirty()
    if i != pfield.np - 1:
        pfield.particles[i] = pfield.particles[pfield.np - 1]
    pfield.np -= 1


def _reset_particles(pfield: ParticleField):
    """vpm._reset_particles (processing_force.jl:237): U, J, PSE <- 0."""
    pfield.particles[:pfield.np, U_INDEX] = 0.0
    pfield.particles[:pfield.np, J_INDEX] = 0.0
    pfield.particles[:pfield.np, PSE_INDEX] = 0.0
    if pfield._device_is_truth():
        pfield._flush_appends()
        pfield._engine.reset_particles()
        pfield._dev_dirty &= ~(_E.FM_U | _E.FM_J | _E.FM_PSE)     # both sides hold zeros now
    else:
        pfield.mark_dirty()


def _reset_particles_sfs(pfield: ParticleField):
    pfield.particles[:pfield.np, SFS_INDEX] = 0.0
    if pfield._device_is_truth():
        pfield._flush_appends()
        pfield._engine.reset_particles_sfs()
        pfield._dev_dirty &= ~_E.FM_SFS
    else:
        pfield.mark_dirty()


def nextstep(pfield: ParticleField, dt: float, relax: bool = False, custom_UJ=None):
    """vpm.nextstep(pfield, dt; relax, custom_UJ) (simulation.jl:358)."""
    if pfield.np > 0:
        pfield.integration(pfield, dt, relax=relax, custom_UJ=custom_UJ)
    else:
        pass
    pfield.t += dt
    pfield.nt += 1


# ---- on-disk format: <file_name>[.<num>].h5 + .xmf  (vpm.save / vpm.read!, simulation.jl:263-265,436-440) ----------------
_XMF = """<?xml version="1.0" ?>
<!DOCTYPE Xdmf SYSTEM "Xdmf.dtd" []>
<Xdmf Version="3.0">
  <Domain>
    <Grid Name="particles" GridType="Uniform">
      <Time Value="{t}" />
      <Geometry GeometryType="XYZ">
        <DataItem DataType="Float" Dimensions="{np} 3" Format="HDF" Precision="8">{h5}:X</DataItem>
      </Geometry>
      <Topology Dimensions="{np}" Type="Polyvertex"/>
{attrs}    </Grid>
  </Domain>
</Xdmf>
"""
_XMF_ATTR = """      <Attribute Center="Node" ElementCell="" ElementDegree="0" ElementFamily="" ItemType="" Name="{name}" Type="{kind}">
        <DataItem DataType="{dtype}" Dimensions="{dims}" Format="HDF" Precision="8">{h5}:{name}</DataItem>
      </Attribute>
"""


def save(pfield, file_name: str, path: str = "", add_num: bool = True, num: Optional[int] = None,
         overwrite_time: Optional[float] = None) -> str:
    """vpm.save(pfield, file_name; path, add_num, num, overwrite_time): writes `<file_name>.<num>.h5` (np, nt, t, X, Gamma,
    sigma, circulation, vol, static, i) and the XDMF wrapper ParaView opens.  Returns the name of the .xmf file."""
    import os

    from . import h5min
    if getattr(pfield, "_dev_dirty", 0):
        pfield.pull()
    n = int(pfield.np)
    P = pfield.particles[:n]
    fname = file_name + (f".{pfield.nt if num is None else num}" if add_num else "")
    h5name = fname + ".h5"
    t = pfield.t if overwrite_time is None else overwrite_time
    data = {
        "np": np.int64(n), "nt": np.int64(pfield.nt), "t": np.float64(t),
        "X": P[:, X_INDEX], "Gamma": P[:, GAMMA_INDEX], "sigma": P[:, SIGMA_INDEX], "circulation": P[:, CIRCULATION_INDEX],
        "vol": P[:, VOL_INDEX], "static": P[:, STATIC_INDEX], "i": np.arange(1, n + 1, dtype=np.int64),
    }
    h5min.write(os.path.join(path, h5name), data)
    attrs = ""
    for name, kind, dtype, dims in (("Gamma", "Vector", "Float", f"{n} 3"), ("sigma", "Scalar", "Float", f"{n}"),
                                    ("circulation", "Scalar", "Float", f"{n}"), ("vol", "Scalar", "Float", f"{n}"),
                                    ("static", "Scalar", "Float", f"{n}"), ("i", "Scalar", "Int", f"{n}")):
        attrs += _XMF_ATTR.format(name=name, kind=kind, dtype=dtype, dims=dims, h5=h5name)
    with open(os.path.join(path, fname + ".xmf"), "w") as f:
        f.write(_XMF.format(t=repr(float(t)), np=n, h5=h5name, attrs=attrs))
    return fname + ".xmf;"


def read_(pfield, h5_fname: str, path: str = "", overwrite: bool = True, load_time: bool = True):
    """vpm.read!(pfield, h5_fname; path, overwrite, load_time) — restart from a saved field (simulation.jl:263-265 calls it
    with overwrite=true, load_time=false)."""
    import os

    from . import h5min
    d = h5min.read(os.path.join(path, h5_fname))
    n = int(d["np"])
    if getattr(pfield, "_dev_dirty", 0):
        if overwrite:
            pfield._dev_dirty = 0          # the device's results are discarded with the particles
        else:
            pfield.pull()                  # keep the device-newer rows of the particles that stay
    if overwrite:
        pfield.np = 0
    if pfield.np + n > pfield.maxparticles:
        raise RuntimeError(f"PARTICLE OVERFLOW. Max number of particles {pfield.maxparticles} has been reached")
    cols = pfield.particles[pfield.np:pfield.np + n]
    cols[:] = 0.0
    cols[:, X_INDEX] = np.asarray(d["X"]).reshape(n, 3)
    cols[:, GAMMA_INDEX] = np.asarray(d["Gamma"]).reshape(n, 3)
    cols[:, SIGMA_INDEX] = np.asarray(d["sigma"]).reshape(n)
    for key, row in (("circulation", CIRCULATION_INDEX), ("vol", VOL_INDEX), ("static", STATIC_INDEX)):
        if key in d:
            cols[:, row] = np.asarray(d[key]).reshape(n)
    pfield.np += n
    if load_time:
        pfield.t = float(d["t"])
        pfield.nt = int(d["nt"])
    if hasattr(pfield, "mark_dirty"):
        pfield.mark_dirty()
    return pfield


read = read_   # `read!` in Julia


def monitor_enstrophy_value(pfield: ParticleField) -> float:
    """0.5 sum Gamma_p . omega(x_p) with omega = curl u from J (vpm.monitor_enstrophy, monitors.jl:614)."""
    P = pfield.particles[:pfield.np]
    Jm = P[:, J_INDEX]
    w = np.stack([Jm[:, 5] - Jm[:, 7], Jm[:, 6] - Jm[:, 2], Jm[:, 1] - Jm[:, 3]], -1)
    return 0.5 * float(np.einsum("ij,ij->", P[:, GAMMA_INDEX], w))


# ---- host-side conveniences of the FLOWVPM API that FLOWUnsteady calls around the hot path ------------------------------
# (SURVEY.md Appendix B "I/O & misc").  Pure host code: they touch a field only through `particles`, `np`, `nt`, `t`,
# `monitors()` and `pull()`, so the CPU tests drive them with a stand-in object.
utilities_path = __import__("os").path.dirname(__import__("os").path.abspath(__file__))   # vpm.utilities_path (FLOWUnsteady.jl:73)


def cd_statistics(C: np.ndarray):
    """(ratio of zeros, mean, std, skewness, kurtosis, min, max) of the dynamic coefficient over the particles where it is
    non-zero — the tuple vpm.monitor_Cd appends after `t` (unpacked at src/FLOWUnsteady_monitors.jl:702 as
    `t, rationzero, mean, stddev, skew, kurt, minC, maxC`).  Sample standard deviation; skewness and kurtosis are the
    standardised third and fourth central moments (kurtosis of a normal distribution = 3)."""
    C = np.asarray(C, dtype=np.float64).ravel()
    n = C.size
    nz = C[C != 0.0]
    if n == 0 or nz.size == 0:
        return (1.0 if n else 0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0)
    mean = float(nz.mean())
    d = nz - mean
    m2 = float(np.mean(d * d))
    std = float(np.sqrt(np.sum(d * d) / (nz.size - 1))) if nz.size > 1 else 0.0
    skew = float(np.mean(d ** 3) / m2 ** 1.5) if m2 > 0 else 0.0
    kurt = float(np.mean(d ** 4) / (m2 * m2)) if m2 > 0 else 0.0
    return (1.0 - nz.size / n, mean, std, skew, kurt, float(nz.min()), float(nz.max()))


def _log_row(save_path, fname, header, row, first):
    import os
    with open(os.path.join(save_path, fname), "w" if first else "a") as f:
        if first:
            f.write(header + "\n")
        f.write(",".join(repr(v) if isinstance(v, float) else str(v) for v in row) + "\n")


def monitor_enstrophy(pfield, t, dt, save_path=None, run_name="", suff="enstrophy.log", vprintln=None, out=None) -> bool:
    """vpm.monitor_enstrophy(pfield, t, dt; save_path, run_name, vprintln, out) (src/FLOWUnsteady_monitors.jl:614):
    global enstrophy 0.5 sum Gamma_p . omega(x_p), reduced on the device (vpmb200_monitors), appended to `out` and to
    `<save_path>/<run_name><suff>`.  Returns False (a runtime function's "do not stop" flag)."""
    if pfield.np == 0:
        return False
    enstrophy = float(pfield.monitors()["enstrophy"])
    if out is not None:
        out.append(enstrophy)
    if save_path is not None:
        _log_row(save_path, run_name + suff, "nt,t (s),enstrophy (m^3/s^2)", (int(pfield.nt), float(t), enstrophy), pfield.nt == 0)
    return False


def monitor_Cd(pfield, t, dt, save_path=None, run_name="", suff="Chistory.log", vprintln=None, out=None) -> bool:
    """vpm.monitor_Cd(pfield, t, dt; save_path, run_name, vprintln, out) (src/FLOWUnsteady_monitors.jl:697): statistics of
    the SFS model coefficient C_d over the particles where it is non-zero; appends
    [t, ratio of zeros, mean, std, skewness, kurtosis, min, max] to `out` and a CSV row to `<save_path>/<run_name><suff>`."""
    if pfield.np == 0:
        return False
    if getattr(pfield, "_dev_dirty", 0):
        pfield.pull(_E.FM_C)
    stats = cd_statistics(pfield.particles[:pfield.np, C_INDEX][:, 0])
    if out is not None:
        out.append([float(t), *stats])
    if save_path is not None:
        _log_row(save_path, run_name + suff, "nt,t (s),ratio of zeros,mean,std,skewness,kurtosis,min,max",
                 (int(pfield.nt), float(t), *stats), pfield.nt == 0)
    return False


def create_path(save_path: str, prompt: bool = True):
    """vpm.create_path(save_path, prompt) (simulation.jl:318): make `save_path` an empty directory.  With prompt=True an
    existing directory is only replaced after the user confirms on stdin (the reference asks the same question)."""
    import os
    import shutil
    if os.path.isdir(save_path):
        if prompt:
            ans = input(f"\n\nFolder {save_path} already exists. Remove? (y/n) ")
            if ans.strip().lower() != "y":
                return
        shutil.rmtree(save_path)
    os.makedirs(save_path)


def settings_of(pfield) -> dict:
    """The solver settings of a field as plain data (what vpm.save_settings stores)."""
    sfs = pfield.SFS
    d = {"maxparticles": int(pfield.maxparticles), "floattype": "Float32" if pfield.R is np.float32 else "Float64",
         "formulation": {"f": pfield.formulation.f, "g": pfield.formulation.g}, "kernel": pfield.kernel.name,
         "viscous": type(pfield.viscous).__name__, "UJ": getattr(pfield.UJ, "__name__", str(pfield.UJ)),
         "integration": getattr(pfield.integration, "__name__", str(pfield.integration)), "transposed": bool(pfield.transposed),
         "relaxation": {"name": getattr(pfield.relaxation, "name", type(pfield.relaxation).__name__),
                        "rlxf": getattr(pfield.relaxation, "rlxf", None), "nsteps_relax": getattr(pfield.relaxation, "nsteps_relax", None)},
         "fmm": {"p": pfield.fmm.p, "ncrit": pfield.fmm.ncrit, "theta": pfield.fmm.theta, "nonzero_sigma": bool(pfield.fmm.nonzero_sigma)},
         "SFS": {"type": type(sfs).__name__}}
    for k in ("alpha", "rlxf", "minC", "maxC", "Cs"):
        if hasattr(sfs, k):
            d["SFS"][k] = getattr(sfs, k)
    for k in ("clippings", "controls"):
        if hasattr(sfs, k):
            d["SFS"][k] = [getattr(c, "__name__", str(c)) for c in getattr(sfs, k)]
    if hasattr(sfs, "procedure"):
        d["SFS"]["procedure"] = getattr(sfs.procedure, "__name__", str(sfs.procedure))
    if iscorespreading(pfield.viscous):
        v = pfield.viscous
        d["viscous_parameters"] = {"nu": v.nu, "sgm0": v.sgm0, "beta": v.beta, "itmax": v.itmax, "tol": v.tol}
    return d


def save_settings(pfield, file_name: str, path: str = "", suff: str = "_settings") -> str:
    """vpm.save_settings(pfield, file_name; path) (simulation.jl:326).  The reference writes a JLD file; JLD is a Julia
    serialisation no other tool reads, so the mirror writes the same content as `<file_name><suff>.json`."""
    import json
    import os
    fname = os.path.join(path, file_name + suff + ".json")
    with open(fname, "w") as f:
        json.dump(settings_of(pfield), f, indent=1)
    return fname


def initialize_verbose(verbose, save_path, run_name, pfield, dt, nsteps_save, runtime_function=None,
                       static_particles_function=None, v_lvl: int = 0):
    """vpm.initialize_verbose(...) (simulation.jl:333): returns (line1, line2, run_id, file_verbose, vprintln, time_beg);
    `vprintln(str, v_lvl)` prints and mirrors every line into `<save_path>/<run_name>.log`."""
    import datetime
    import os
    line1 = "*" * 73
    line2 = "-" * 73
    run_id = save_path if save_path is not None else run_name
    file_verbose = os.path.join(save_path, run_name + ".log") if save_path is not None else None
    if file_verbose is not None:
        open(file_verbose, "w").close()

    def vprintln(s="", lvl=0):
        text = "\t" * lvl + str(s)
        if verbose:
            print(text)
        if file_verbose is not None:
            with open(file_verbose, "a") as f:
                f.write(text + "\n")

    time_beg = datetime.datetime.now()
    vprintln(line1, v_lvl)
    vprintln(f"START {run_id}\t{time_beg}", v_lvl)
    vprintln(line1, v_lvl)
    vprintln(f"max particles: {pfield.maxparticles}, dt: {dt}, save every {nsteps_save} steps", v_lvl + 1)
    return line1, line2, run_id, file_verbose, vprintln, time_beg


def finalize_verbose(time_beg, line1, vprintln, run_id, v_lvl: int = 0):
    """vpm.finalize_verbose(time_beg, line1, vprintln, run_id, v_lvl) (simulation.jl:450)."""
    import datetime
    time_end = datetime.datetime.now()
    el = time_end - time_beg
    hrs, rem = divmod(int(el.total_seconds()), 3600)
    mins, secs = divmod(rem, 60)
    vprintln(line1, v_lvl)
    vprintln(f"END {run_id}\t{time_end}", v_lvl)
    vprintln(line1, v_lvl)
    vprintln(f"ELAPSED TIME: {hrs} hours {mins} minutes {secs} seconds", v_lvl)


def run_vpm_(pfield, dt: float, nsteps: int, runtime_function=None, static_particles_function=None, nsteps_relax: Optional[int] = None,
             save_path=None, run_name: str = "pfield", nsteps_save: int = 1, verbose: bool = True, v_lvl: int = 0,
             create_savepath: bool = True, prompt: bool = True, save_time: bool = True):
    """vpm.run_vpm!(pfield, dt, nsteps; ...): FLOWVPM's own driver loop (FLOWUnsteady inlines it, simulation.jl:300-447).
    Per step: static particles in -> nextstep -> static particles out -> runtime_function(pfield, t, dt) (True stops the
    run) -> save every `nsteps_save` steps."""
    if save_path is not None:
        if create_savepath:
            create_path(save_path, prompt)
        save_settings(pfield, run_name, path=save_path)
    line1, line2, run_id, file_verbose, vprintln, time_beg = initialize_verbose(
        verbose, save_path, run_name, pfield, dt, nsteps_save, runtime_function, static_particles_function, v_lvl)
    for i in range(nsteps + 1):
        if i % max(1, nsteps // 10) == 0:
            vprintln(f"Time step {i} out of {nsteps}\tParticles: {get_np(pfield)}", v_lvl + 1)
        # relax cadence: the relaxation scheme's own nsteps_relax (simulation.jl:346-348); the keyword overrides it
        every = getattr(pfield.relaxation, "nsteps_relax", 1) if nsteps_relax is None else nsteps_relax
        relax = (getattr(pfield.relaxation, "id", 0) != 0 and every >= 1 and i > 0 and i % every == 0)
        org_np = get_np(pfield)
        if i != 0:
            remove = None
            if static_particles_function is not None:
                remove = static_particles_function(pfield, pfield.t, dt)
            nextstep(pfield, dt, relax=relax)
            if remove is None or remove:                     # simulation.jl:361: statics go unless the function returned false
                for pi in range(get_np(pfield), org_np, -1): # they were appended last (simulation.jl:362-365)
                    remove_particle(pfield, pi - 1)
        breakflag = bool(runtime_function(pfield, pfield.t, dt)) if runtime_function is not None else False
        if save_path is not None and (i % nsteps_save == 0 or i == nsteps or breakflag):
            save(pfield, run_name, path=save_path, add_num=True, overwrite_time=pfield.t if save_time else float(pfield.nt))
        if breakflag:
            break
    finalize_verbose(time_beg, line1, vprintln, run_id, v_lvl)
    return pfield

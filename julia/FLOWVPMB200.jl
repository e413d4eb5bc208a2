#=
FLOWVPMB200.jl — reference-side binding of libvpmb200.so (include/vpmb200.h).

This is the stub a FLOWUnsteady maintainer adds next to `import FLOWVPM; const vpm = FLOWVPM`
(src/FLOWUnsteady.jl:34,42 of the reference).  It defines drop-in callables with FLOWVPM's call shapes

    UJ_b200(pfield; reset=true, reset_sfs=false, sfs=false, optargs...)     # replaces vpm.UJ_direct / vpm.UJ_fmm
    rungekutta3_b200(pfield, dt; relax=false, custom_UJ=nothing)            # replaces vpm.rungekutta3
    euler_b200(pfield, dt; relax=false, custom_UJ=nothing)                  # replaces vpm.euler
    Vvpm_on_Xs_b200(pfield, Xs)                                             # probe fast path for simulation.jl:494-570

selected through the existing keyword arguments of `run_simulation` (src/FLOWUnsteady_simulation.jl:127-137):

    uns.run_simulation(sim, nsteps; vpm_UJ=FLOWVPMB200.UJ_b200, vpm_integration=FLOWVPMB200.rungekutta3_b200, ...)

EXPERIMENTAL — NOT EXECUTED HERE: no julia binary exists in the build image or on the GPU box, so this file is
syntax-reviewed only; run it once under Julia before relying on it.  `VPMB200_NGPUS` > 1 selects the multi-GPU handle
(vpmb200_multi_*: one host thread drives all GPUs of the box; direct path).
Each wrapper is one `ccall`; the particle matrix `pfield.particles` (43 x maxparticles, column-major Float64) is passed
by pointer and is only borrowed for the duration of the call (GC.@preserve).
=#
module FLOWVPMB200

import FLOWVPM
const vpm = FLOWVPM

const LIB = get(ENV, "VPMB200_LIB", "libvpmb200.so")
const NFIELDS = 43
const FM_ALL = UInt32(0x1fff)
const FM_STATE = UInt32(1 << 0 | 1 << 1 | 1 << 2 | 1 << 3 | 1 << 4 | 1 << 10 | 1 << 12)

# mirrors vpmb200_schemes (include/vpmb200.h)
mutable struct Schemes
    kernel::Int32; f::Float64; g::Float64; transposed::Int32
    relaxation::Int32; rlxf::Float64; sfs::Int32; alpha::Float64
    sfs_rlxf::Float64; minC::Float64; maxC::Float64; Cs::Float64
    force_positive::Int32; clippings::Int32; controls::Int32
    viscous::Int32; nu::Float64; integration::Int32
    cs_sgm0::Float64; cs_beta::Float64; cs_itmax::Int32; cs_tol::Float64
    uj::Int32; fmm_p::Int32; fmm_ncrit::Int32; fmm_theta::Float64; fmm_nonzero_sigma::Int32
    Schemes() = new()
end

const NGPUS = parse(Int, get(ENV, "VPMB200_NGPUS", "1"))
# entry points of the handle in use (single-GPU or multi-GPU): constant globals, as ccall requires
const F_DOWNLOAD = NGPUS > 1 ? :vpmb200_multi_download : :vpmb200_download
const F_NEXTSTEP = NGPUS > 1 ? :vpmb200_multi_nextstep : :vpmb200_nextstep
const F_SET_SCHEMES = NGPUS > 1 ? :vpmb200_multi_set_schemes : :vpmb200_set_schemes
const F_SET_TIME = NGPUS > 1 ? :vpmb200_multi_set_time : :vpmb200_set_time
const F_UJ = NGPUS > 1 ? :vpmb200_multi_uj : :vpmb200_uj
const F_UJ_PROBE = NGPUS > 1 ? :vpmb200_multi_uj_probe : :vpmb200_uj_probe
const F_UPLOAD = NGPUS > 1 ? :vpmb200_multi_upload : :vpmb200_upload

# One engine per ParticleField.  The table must not keep the field alive (a strong reference would make the field's finalizer
# — which unregisters the pinned matrix and destroys the engine — unreachable), hence weak keys.
mutable struct _Engine
    h::Ptr{Cvoid}
    pinned::Ptr{Cvoid}
end
const _handles = WeakKeyDict{Any, _Engine}()

function _release(e::_Engine)
    e.pinned != C_NULL && ccall((:vpmb200_host_unregister, LIB), Int32, (Ptr{Cvoid},), e.pinned)
    if e.h != C_NULL
        NGPUS > 1 ? ccall((:vpmb200_multi_destroy, LIB), Int32, (Ptr{Cvoid},), e.h) :
                    ccall((:vpmb200_destroy, LIB), Int32, (Ptr{Cvoid},), e.h)
    end
    e.h = C_NULL; e.pinned = C_NULL
    return nothing
end

function _check(h, rc)
    rc == 0 && return nothing
    msg = NGPUS > 1 ? unsafe_string(ccall((:vpmb200_multi_last_error, LIB), Cstring, (Ptr{Cvoid},), h)) :
                      unsafe_string(ccall((:vpmb200_last_error, LIB), Cstring, (Ptr{Cvoid},), h))
    error("vpmb200 error $rc: $msg")             # reference convention: Julia error() (simulation.jl:201-213)
end

function _handle(pfield)
    e = get!(_handles, pfield) do
        # the engine computes pairs in FP64 or FP32 but its STATE (and this ABI) is Float64: a Float32 particle matrix cannot
        # be passed by pointer — convert the field to Float64 and select the FP32 pair arithmetic with VPMB200_FLOAT_BITS=32
        eltype(pfield.particles) == Float64 ||
            error("FLOWVPMB200 needs a Float64 particle matrix (got $(eltype(pfield.particles))); " *
                  "set ENV[\"VPMB200_FLOAT_BITS\"] = \"32\" for FP32 pair arithmetic on a Float64 field")
        bits = parse(Int, get(ENV, "VPMB200_FLOAT_BITS", "64"))
        h = Ref{Ptr{Cvoid}}(C_NULL)
        rc = NGPUS > 1 ?
            ccall((:vpmb200_multi_create, LIB), Int32, (Int64, Int32, Int32, Int32, Ptr{Int32}, Ref{Ptr{Cvoid}}),
                  pfield.maxparticles, NFIELDS, bits, NGPUS, C_NULL, h) :
            ccall((:vpmb200_create, LIB), Int32, (Int64, Int32, Int32, Int32, Ref{Ptr{Cvoid}}),
                  pfield.maxparticles, NFIELDS, bits, 0, h)
        rc == 0 || error("vpmb200 create failed ($rc): " * (NGPUS > 1 ?
                         unsafe_string(ccall((:vpmb200_multi_last_error, LIB), Cstring, (Ptr{Cvoid},), C_NULL)) :
                         unsafe_string(ccall((:vpmb200_last_error, LIB), Cstring, (Ptr{Cvoid},), C_NULL))))
        # page-lock the particle matrix once so uploads / downloads are full-rate DMA straight from / into it
        P = pfield.particles
        ok = ccall((:vpmb200_host_register, LIB), Int32, (Ptr{Cvoid}, UInt64), pointer(P), sizeof(P))
        eng = _Engine(h[], ok == 0 ? Ptr{Cvoid}(pointer(P)) : C_NULL)
        finalizer(_release, eng)                     # the engine object carries its own finalizer ...
        finalizer(pf -> (haskey(_handles, pf) && _release(_handles[pf])), pfield)   # ... and goes with its field
        eng
    end
    return e.h
end

_kernel_id(k) = k === vpm.gaussianerf ? 0 : k === vpm.winckelmans ? 1 : k === vpm.gaussian ? 2 : 3
_relax_id(r) = r === vpm.pedrizzetti ? 1 : r === vpm.correctedpedrizzetti ? 2 : 0

function _schemes(pfield; uj::Integer=0, integration::Integer=1)
    s = Schemes()
    ccall((:vpmb200_default_schemes, LIB), Int32, (Ref{Schemes},), s)
    s.kernel = _kernel_id(pfield.kernel)
    s.f, s.g = pfield.formulation.f, pfield.formulation.g
    s.transposed = pfield.transposed
    s.relaxation = _relax_id(pfield.relaxation); s.rlxf = pfield.relaxation.rlxf
    sfs = pfield.SFS
    if sfs isa vpm.DynamicSFS
        s.sfs = 2; s.alpha = sfs.alpha; s.sfs_rlxf = sfs.rlxf; s.minC = sfs.minC; s.maxC = sfs.maxC
        s.force_positive = sfs.procedure === vpm.pseudo3level_positive
    elseif sfs isa vpm.ConstantSFS
        s.sfs = 1; s.Cs = sfs.Cs
    end
    if vpm.isSFSenabled(sfs)
        s.clippings = any(c -> c === vpm.clipping_backscatter, sfs.clippings) ? 1 : 0
        s.controls = (any(c -> c === vpm.control_directional, sfs.controls) ? 1 : 0) |
                     (any(c -> c === vpm.control_magnitude, sfs.controls) ? 2 : 0)
    end
    if vpm.iscorespreading(pfield.viscous)
        s.viscous = 1; s.nu = pfield.viscous.nu
        s.cs_sgm0 = pfield.viscous.sgm0; s.cs_beta = pfield.viscous.beta
        s.cs_itmax = pfield.viscous.itmax; s.cs_tol = pfield.viscous.tol
    end
    s.integration = integration
    s.uj = uj
    s.fmm_p, s.fmm_ncrit, s.fmm_theta = pfield.fmm.p, pfield.fmm.ncrit, pfield.fmm.theta
    s.fmm_nonzero_sigma = pfield.fmm.nonzero_sigma
    return s
end

function _push(h, pfield, mask)
    P = pfield.particles
    GC.@preserve P _check(h, ccall((F_UPLOAD, LIB), Int32, (Ptr{Cvoid}, Ptr{Float64}, Int64, Int64, UInt32),
                                   h, P, size(P, 1), pfield.np, mask))
    _check(h, ccall((F_SET_TIME, LIB), Int32, (Ptr{Cvoid}, Float64, Int64), h, pfield.t, pfield.nt))
end

function _pull(h, pfield, mask)
    P = pfield.particles
    GC.@preserve P _check(h, ccall((F_DOWNLOAD, LIB), Int32, (Ptr{Cvoid}, Ptr{Float64}, Int64, Int64, UInt32),
                                   h, P, size(P, 1), pfield.np, mask))
end

"`pfield.UJ(pfield)` on the GPU: U and J (and the SFS term with `sfs=true`) at every particle."
function UJ_b200(pfield; reset=true, reset_sfs=false, sfs=false, fmm=false, optargs...)
    h = _handle(pfield)
    s = _schemes(pfield; uj=(fmm ? 1 : 0))
    _check(h, ccall((F_SET_SCHEMES, LIB), Int32, (Ptr{Cvoid}, Ref{Schemes}), h, s))
    _push(h, pfield, FM_ALL)
    _check(h, ccall((F_UJ, LIB), Int32, (Ptr{Cvoid}, Int32, Int32, Int32), h, reset, reset_sfs, sfs))
    _pull(h, pfield, UInt32(1 << 5 | 1 << 7 | 1 << 8 | 1 << 11))        # U, J, PSE, SFS
    return nothing
end
UJ_fmm_b200(pfield; optargs...) = UJ_b200(pfield; fmm=true, optargs...)

function _nextstep(pfield, dt, integration; relax=false, custom_UJ=nothing)
    custom_UJ === nothing || error("custom_UJ cannot run inside the GPU engine")
    h = _handle(pfield)
    s = _schemes(pfield; uj=(pfield.UJ === UJ_fmm_b200 ? 1 : 0), integration=integration)
    _check(h, ccall((F_SET_SCHEMES, LIB), Int32, (Ptr{Cvoid}, Ref{Schemes}), h, s))
    _push(h, pfield, FM_STATE | UInt32(1 << 9))
    Uinf = Float64.(collect(pfield.Uinf(pfield.t)))                     # evaluated by the host (simulation.jl:238-239)
    _check(h, ccall((F_NEXTSTEP, LIB), Int32, (Ptr{Cvoid}, Float64, Ptr{Float64}, Int32), h, dt, Uinf, relax))
    _pull(h, pfield, FM_ALL)
    return nothing                                                       # vpm.nextstep itself advances pfield.t / nt
end
rungekutta3_b200(pfield, dt; optargs...) = _nextstep(pfield, dt, 1; optargs...)
euler_b200(pfield, dt; optargs...) = _nextstep(pfield, dt, 0; optargs...)

"Velocity induced by the field at probe positions Xs (replaces add_probe + pfield.UJ + get_U, simulation.jl:536-547)."
function Vvpm_on_Xs_b200(pfield, Xs::AbstractVector)
    isempty(Xs) && return [zeros(3) for _ in Xs]
    h = _handle(pfield)
    _check(h, ccall((F_SET_SCHEMES, LIB), Int32, (Ptr{Cvoid}, Ref{Schemes}), h, _schemes(pfield)))
    _push(h, pfield, FM_STATE)
    X = Matrix{Float64}(undef, 3, length(Xs)); for (i, x) in enumerate(Xs); X[:, i] .= x; end
    U = similar(X)
    GC.@preserve X U _check(h, ccall((F_UJ_PROBE, LIB), Int32,
                                     (Ptr{Cvoid}, Ptr{Float64}, Int64, Ptr{Float64}, Ptr{Float64}),
                                     h, X, length(Xs), U, C_NULL))
    return [U[:, i] for i in 1:length(Xs)]
end

"""
Static-particle fast path (simulation.jl:355-365): instead of `static_particles_function(pfield, t, dt)` appending the
embedded particles and the loop removing them after `vpm.nextstep`, hand their columns (43 x n) to the engine; they are parked
behind the field for this step only (single-GPU handle).
"""
function set_statics_b200(pfield, cols::Matrix{Float64})
    NGPUS > 1 && error("the static-particle fast path is wired for the single-GPU handle")
    h = _handle(pfield)
    GC.@preserve cols _check(h, ccall((:vpmb200_set_statics, LIB), Int32, (Ptr{Cvoid}, Ptr{Float64}, Int64, Int64, Int64),
                                      h, cols, size(cols, 1), size(cols, 2), pfield.nt))
end

end # module

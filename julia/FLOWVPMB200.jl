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

NOT EXECUTED HERE: no julia binary exists in the build image or on the GPU box, so this file is syntax-reviewed only.
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

const _handles = IdDict{Any, Ptr{Cvoid}}()        # one engine per ParticleField

function _check(h, rc)
    rc == 0 && return nothing
    msg = unsafe_string(ccall((:vpmb200_last_error, LIB), Cstring, (Ptr{Cvoid},), h))
    error("vpmb200 error $rc: $msg")             # reference convention: Julia error() (simulation.jl:201-213)
end

function _handle(pfield)
    get!(_handles, pfield) do
        h = Ref{Ptr{Cvoid}}(C_NULL)
        bits = eltype(pfield.particles) == Float32 ? 32 : 64
        rc = ccall((:vpmb200_create, LIB), Int32, (Int64, Int32, Int32, Int32, Ref{Ptr{Cvoid}}),
                   pfield.maxparticles, NFIELDS, bits, 0, h)
        rc == 0 || error("vpmb200_create failed ($rc): " *
                         unsafe_string(ccall((:vpmb200_last_error, LIB), Cstring, (Ptr{Cvoid},), C_NULL)))
        # page-lock the particle matrix once so uploads / downloads are full-rate DMA straight from / into it
        P = pfield.particles
        ccall((:vpmb200_host_register, LIB), Int32, (Ptr{Cvoid}, UInt64), pointer(P), sizeof(P))
        finalizer(pfield) do pf
            ccall((:vpmb200_host_unregister, LIB), Int32, (Ptr{Cvoid},), pointer(pf.particles))
            ccall((:vpmb200_destroy, LIB), Int32, (Ptr{Cvoid},), h[])
        end
        h[]
    end
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
    GC.@preserve P _check(h, ccall((:vpmb200_upload, LIB), Int32, (Ptr{Cvoid}, Ptr{Float64}, Int64, Int64, UInt32),
                                   h, P, size(P, 1), pfield.np, mask))
    _check(h, ccall((:vpmb200_set_time, LIB), Int32, (Ptr{Cvoid}, Float64, Int64), h, pfield.t, pfield.nt))
end

function _pull(h, pfield, mask)
    P = pfield.particles
    GC.@preserve P _check(h, ccall((:vpmb200_download, LIB), Int32, (Ptr{Cvoid}, Ptr{Float64}, Int64, Int64, UInt32),
                                   h, P, size(P, 1), pfield.np, mask))
end

"`pfield.UJ(pfield)` on the GPU: U and J (and the SFS term with `sfs=true`) at every particle."
function UJ_b200(pfield; reset=true, reset_sfs=false, sfs=false, fmm=false, optargs...)
    h = _handle(pfield)
    s = _schemes(pfield; uj=(fmm ? 1 : 0))
    _check(h, ccall((:vpmb200_set_schemes, LIB), Int32, (Ptr{Cvoid}, Ref{Schemes}), h, s))
    _push(h, pfield, FM_ALL)
    _check(h, ccall((:vpmb200_uj, LIB), Int32, (Ptr{Cvoid}, Int32, Int32, Int32), h, reset, reset_sfs, sfs))
    _pull(h, pfield, UInt32(1 << 5 | 1 << 7 | 1 << 8 | 1 << 11))        # U, J, PSE, SFS
    return nothing
end
UJ_fmm_b200(pfield; optargs...) = UJ_b200(pfield; fmm=true, optargs...)

function _nextstep(pfield, dt, integration; relax=false, custom_UJ=nothing)
    custom_UJ === nothing || error("custom_UJ cannot run inside the GPU engine")
    h = _handle(pfield)
    s = _schemes(pfield; uj=(pfield.UJ === UJ_fmm_b200 ? 1 : 0), integration=integration)
    _check(h, ccall((:vpmb200_set_schemes, LIB), Int32, (Ptr{Cvoid}, Ref{Schemes}), h, s))
    _push(h, pfield, FM_STATE | UInt32(1 << 9))
    Uinf = Float64.(collect(pfield.Uinf(pfield.t)))                     # evaluated by the host (simulation.jl:238-239)
    _check(h, ccall((:vpmb200_nextstep, LIB), Int32, (Ptr{Cvoid}, Float64, Ptr{Float64}, Int32), h, dt, Uinf, relax))
    _pull(h, pfield, FM_ALL)
    return nothing                                                       # vpm.nextstep itself advances pfield.t / nt
end
rungekutta3_b200(pfield, dt; optargs...) = _nextstep(pfield, dt, 1; optargs...)
euler_b200(pfield, dt; optargs...) = _nextstep(pfield, dt, 0; optargs...)

"Velocity induced by the field at probe positions Xs (replaces add_probe + pfield.UJ + get_U, simulation.jl:536-547)."
function Vvpm_on_Xs_b200(pfield, Xs::AbstractVector)
    isempty(Xs) && return [zeros(3) for _ in Xs]
    h = _handle(pfield)
    _check(h, ccall((:vpmb200_set_schemes, LIB), Int32, (Ptr{Cvoid}, Ref{Schemes}), h, _schemes(pfield)))
    _push(h, pfield, FM_STATE)
    X = Matrix{Float64}(undef, 3, length(Xs)); for (i, x) in enumerate(Xs); X[:, i] .= x; end
    U = similar(X)
    GC.@preserve X U _check(h, ccall((:vpmb200_uj_probe, LIB), Int32,
                                     (Ptr{Cvoid}, Ptr{Float64}, Int64, Ptr{Float64}, Ptr{Float64}),
                                     h, X, length(Xs), U, C_NULL))
    return [U[:, i] for i in 1:length(Xs)]
end

end # module

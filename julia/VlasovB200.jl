# VlasovB200.jl -- Julia glue for libvlasov_b200.so (UNTESTED: no Julia toolchain exists where this repository is
# built and tested; the identical call sequences are exercised through ctypes by vlasovmethods.jl_b200/*.py).
#
# A maintainer of VlasovMethods.jl would `include` this file and forward the hot-path methods to it; the reference
# signatures stay unchanged (see INTEGRATION.md for the mapping, reference file:line per method).
module VlasovB200

const lib = get(ENV, "VLASOV_B200_LIB", joinpath(@__DIR__, "..", "vlasovmethods.jl_b200", "libvlasov_b200.so"))

struct VMError <: Exception
    code::Cint
    msg::String
end

function check(rc::Cint, ctx::Ptr{Cvoid} = C_NULL)
    rc == 0 && return nothing
    throw(VMError(rc, unsafe_string(ccall((:vm_last_error, lib), Cstring, (Ptr{Cvoid},), ctx))))
end

# ------------------------------------------------------------------ context
mutable struct Context
    h::Ptr{Cvoid}
    function Context(device::Integer = parse(Int, get(ENV, "LOCAL_RANK", "0")))
        r = Ref{Ptr{Cvoid}}(C_NULL)
        check(ccall((:vm_ctx_create, lib), Cint, (Cint, Ref{Ptr{Cvoid}}), device, r))
        c = new(r[])
        finalizer(x -> ccall((:vm_ctx_destroy, lib), Cint, (Ptr{Cvoid},), x.h), c)
    end
end
const CTX = Ref{Context}()
ctx() = (isassigned(CTX) || (CTX[] = Context()); CTX[])
sync() = check(ccall((:vm_sync, lib), Cint, (Ptr{Cvoid},), ctx().h), ctx().h)

"One process per GPU: `id` = 128-byte NCCL unique id from rank 0 (vm_comm_unique_id), broadcast by the host (MPI.jl)."
comm_unique_id() = (b = zeros(UInt8, 128); check(ccall((:vm_comm_unique_id, lib), Cint, (Ptr{UInt8},), b)); b)
comm_init!(rank, nranks, id::Vector{UInt8}) =
    check(ccall((:vm_ctx_comm_init, lib), Cint, (Ptr{Cvoid}, Cint, Cint, Ptr{UInt8}), ctx().h, rank, nranks, id), ctx().h)
peer_handle() = (b = zeros(UInt8, 64); check(ccall((:vm_ctx_peer_handle, lib), Cint, (Ptr{Cvoid}, Ptr{UInt8}), ctx().h, b), ctx().h); b)
peer_connect!(handles::Vector{UInt8}) =
    check(ccall((:vm_ctx_peer_connect, lib), Cint, (Ptr{Cvoid}, Ptr{UInt8}), ctx().h, handles), ctx().h)

# ---------------------------------------------------------------- particles
mutable struct DeviceParticles
    h::Ptr{Cvoid}
    n::Int
    function DeviceParticles(n::Integer)
        r = Ref{Ptr{Cvoid}}(C_NULL)
        check(ccall((:vm_particles_create, lib), Cint, (Ptr{Cvoid}, Clong, Ref{Ptr{Cvoid}}), ctx().h, n, r), ctx().h)
        p = new(r[], n)
        finalizer(x -> ccall((:vm_particles_destroy, lib), Cint, (Ptr{Cvoid},), x.h), p)
    end
end
# z3 is the 3 x N column-major matrix [x; v; w] behind ParticleList (src/distributions/particle_distribution.jl:11-18)
upload!(p::DeviceParticles, z3::Matrix{Float64}) =
    check(ccall((:vm_particles_upload_aos, lib), Cint, (Ptr{Cvoid}, Ptr{Float64}), p.h, z3), ctx().h)
download!(z3::Matrix{Float64}, p::DeviceParticles) =
    check(ccall((:vm_particles_download_aos, lib), Cint, (Ptr{Cvoid}, Ptr{Float64}), p.h, z3), ctx().h)
upload!(p::DeviceParticles; x = C_NULL, v = C_NULL, w = C_NULL) =
    check(ccall((:vm_particles_upload_soa, lib), Cint, (Ptr{Cvoid}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}), p.h, x, v, w), ctx().h)
fill!(p::DeviceParticles, kind::Integer, params::Vector{Float64}; seed = 20240601, first = 0, total = p.n) =
    check(ccall((:vm_particles_fill, lib), Cint, (Ptr{Cvoid}, Cint, Ptr{Float64}, Cint, Culonglong, Clong, Clong),
                p.h, kind, params, length(params), seed, first, total), ctx().h)

# -------------------------------------------------------------------- field
mutable struct DeviceField
    h::Ptr{Cvoid}
    n::Int
    function DeviceField(a, b, order::Integer, n_basis::Integer, index_shift::Integer = order ÷ 2 - order + 1)
        r = Ref{Ptr{Cvoid}}(C_NULL)
        check(ccall((:vm_field_create, lib), Cint, (Ptr{Cvoid}, Cdouble, Cdouble, Cint, Cint, Cint, Ref{Ptr{Cvoid}}),
                    ctx().h, a, b, order, n_basis, index_shift, r), ctx().h)
        f = new(r[], n_basis)
        finalizer(x -> ccall((:vm_field_destroy, lib), Cint, (Ptr{Cvoid},), x.h), f)
    end
end
deposit!(f::DeviceField, p::DeviceParticles; mode = 0) =               # projection!(potential, dist)
    check(ccall((:vm_deposit, lib), Cint, (Ptr{Cvoid}, Ptr{Cvoid}, Cint), f.h, p.h, mode), ctx().h)
solve!(f::DeviceField) = check(ccall((:vm_field_solve, lib), Cint, (Ptr{Cvoid},), f.h), ctx().h)   # PoissonSolvers.update!
rhs(f::DeviceField) = (o = zeros(f.n); check(ccall((:vm_field_get_rhs, lib), Cint, (Ptr{Cvoid}, Ptr{Float64}), f.h, o), ctx().h); o)
coefficients(f::DeviceField) = (o = zeros(f.n); check(ccall((:vm_field_get_coefficients, lib), Cint, (Ptr{Cvoid}, Ptr{Float64}), f.h, o), ctx().h); o)
set_coefficients!(f::DeviceField, ϕ::Vector{Float64}) =                # ExternalField update!
    check(ccall((:vm_field_set_coefficients, lib), Cint, (Ptr{Cvoid}, Ptr{Float64}), f.h, ϕ), ctx().h)
energy(f::DeviceField) = (w = Ref{Cdouble}(0); check(ccall((:vm_field_energy, lib), Cint, (Ptr{Cvoid}, Ref{Cdouble}), f.h, w), ctx().h); w[])
function eval_field(f::DeviceField, x::Vector{Float64}; deriv = 1)     # ϕ(x, Derivative(1))
    o = similar(x)
    check(ccall((:vm_field_eval, lib), Cint, (Ptr{Cvoid}, Ptr{Float64}, Clong, Cint, Ptr{Float64}), f.h, x, length(x), deriv, o), ctx().h)
    o
end
drift!(p::DeviceParticles, dt) = check(ccall((:vm_vp_drift, lib), Cint, (Ptr{Cvoid}, Cdouble), p.h, dt), ctx().h)        # s_advection!
kick!(f::DeviceField, p::DeviceParticles, dt; scale = -1.0) =                                                              # s_acceleration! (after deposit!/solve!)
    check(ccall((:vm_vp_kick, lib), Cint, (Ptr{Cvoid}, Ptr{Cvoid}, Cdouble, Cdouble), f.h, p.h, dt, scale), ctx().h)

const RUN_SPLIT_KICK, RUN_FROZEN_FIELD, RUN_ATOMIC_DEPOSIT, RUN_UNFUSED = Cint(1), Cint(2), Cint(4), Cint(8)
"nsteps fused Strang steps; returns the diagnostics rows [W K M Σw] (4 x nrows) when diag_every > 0."
function vp_run!(f::DeviceField, p::DeviceParticles, dt, nsteps::Integer; diag_every = 0, flags = Cint(0), χ = 1.0)
    diag = diag_every > 0 ? zeros(4, nsteps ÷ diag_every + 1) : zeros(4, 0)
    check(ccall((:vm_vp_run, lib), Cint, (Ptr{Cvoid}, Ptr{Cvoid}, Cdouble, Cint, Cint, Cint, Cdouble, Ptr{Float64}),
                f.h, p.h, dt, nsteps, diag_every, flags, χ, diag_every > 0 ? pointer(diag) : C_NULL), ctx().h)
    diag
end

# ------------------------------------------------------------------ vspline
mutable struct DeviceVSpline
    h::Ptr{Cvoid}
    nv::Int
    function DeviceVSpline(vmin, vmax, nknots::Integer, order::Integer; dirichlet = true)
        r = Ref{Ptr{Cvoid}}(C_NULL)
        check(ccall((:vm_vspline_create, lib), Cint, (Ptr{Cvoid}, Cdouble, Cdouble, Cint, Cint, Cint, Ref{Ptr{Cvoid}}),
                    ctx().h, vmin, vmax, nknots, order, dirichlet ? 1 : 0, r), ctx().h)
        s = new(r[], ccall((:vm_vspline_size, lib), Cint, (Ptr{Cvoid},), r[]))
        finalizer(x -> ccall((:vm_vspline_destroy, lib), Cint, (Ptr{Cvoid},), x.h), s)
    end
end
project!(s::DeviceVSpline, p::DeviceParticles) = check(ccall((:vm_vproject, lib), Cint, (Ptr{Cvoid}, Ptr{Cvoid}), s.h, p.h), ctx().h)
coefficients(s::DeviceVSpline) = (o = zeros(s.nv); check(ccall((:vm_vspline_get_coefficients, lib), Cint, (Ptr{Cvoid}, Ptr{Float64}), s.h, o), ctx().h); o)
function moments(s::DeviceVSpline, p::DeviceParticles)                 # compute_f_densities / compute_df_densities / compute_coefficients
    m5 = zeros(5); A = zeros(2)
    check(ccall((:vm_vmoments, lib), Cint, (Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Float64}, Ptr{Float64}), s.h, p.h, m5, A), ctx().h)
    m5, A
end
function lb_rhs!(v̇::Vector{Float64}, s::DeviceVSpline, p::DeviceParticles; ν = 1.0, conservative = false)   # LB_rhs! / CLB_rhs!
    check(ccall((:vm_lb_rhs, lib), Cint, (Ptr{Cvoid}, Ptr{Cvoid}, Cdouble, Cint, Ptr{Float64}), s.h, p.h, ν, conservative, v̇), ctx().h)
    v̇
end
function lb_rk438_run!(s::DeviceVSpline, p::DeviceParticles, dt, nsteps::Integer; ν = 1.0, conservative = false, diag_every = 0)
    diag = diag_every > 0 ? zeros(4, nsteps ÷ diag_every + 1) : zeros(4, 0)
    check(ccall((:vm_lb_rk438_run, lib), Cint, (Ptr{Cvoid}, Ptr{Cvoid}, Cdouble, Cint, Cdouble, Cint, Cint, Ptr{Float64}),
                s.h, p.h, dt, nsteps, ν, conservative, diag_every, diag_every > 0 ? pointer(diag) : C_NULL), ctx().h)
    diag
end

end # module

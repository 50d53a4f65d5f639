# VlasovMethodsB200Ext.jl -- the drop-in: method definitions ON THE REFERENCE'S OWN TYPES whose bodies are ccalls into
# libvlasov_b200.so (include/vlasov_b200.h).  `include` this file after `using VlasovMethods`; the user scripts
# (scripts/vlasov_poisson.jl, lenard_bernstein*.jl) then run unchanged, with the hot path on the GPU.
#
# UNTESTED: no Julia toolchain exists where this repository is built and tested.  The same call sequences are
# exercised through ctypes by vlasovmethods.jl_b200/api.py + legacy.py (the Python mirror the tests drive); the
# signatures below are copied from the reference files cited next to each method.
#
# Device state: the host matrices stay the user-visible state (as in the reference); each reference object gets a
# device twin on first use, kept in a WeakKeyDict, re-uploaded when the host copy was edited (`mark_host_dirty!`).
module VlasovMethodsB200Ext

using VlasovMethods
using VlasovMethods: ParticleDistribution, SplineDistribution, VlasovPoisson, LenardBernstein,
                     ConservativeLenardBernstein, SplittingMethod, GeometricIntegrator
import PoissonSolvers
import PoissonSolvers: Potential
using BSplineKit: PeriodicBSplineBasis, Derivative, order, boundaries, knots
import GeometricEquations
using HDF5

include(joinpath(@__DIR__, "VlasovB200.jl"))
const B = VlasovB200

# ------------------------------------------------------------------------------------------------ device twins --
const PARTICLES = WeakKeyDict{Any, Tuple{B.DeviceParticles, Base.RefValue{Bool}}}()    # dist     -> (handle, current?)
const FIELDS    = WeakKeyDict{Any, B.DeviceField}()                                    # potential -> handle
const VSPLINES  = WeakKeyDict{Any, B.DeviceVSpline}()                                  # sdist    -> handle

"3 x N column-major matrix [x; v; w] behind a ParticleList (src/distributions/particle_distribution.jl:11-18)"
hostmatrix(dist::ParticleDistribution) = parent(dist.particles.z)

function device(dist::ParticleDistribution)
    dev, cur = get!(PARTICLES, dist) do
        (B.DeviceParticles(length(dist.particles)), Ref(false))
    end
    if !cur[]
        B.upload!(dev, hostmatrix(dist))
        cur[] = true
    end
    return dev
end
mark_host_dirty!(dist::ParticleDistribution) = haskey(PARTICLES, dist) && (PARTICLES[dist][2][] = false)
tohost!(dist::ParticleDistribution) = (B.download!(hostmatrix(dist), device(dist)); dist)

# ----------------------------------------------------------------------------------------------- initial loads --
# src/examples/bumpontail.jl:33-75, 90-121: draw!(dist, f_x, params, sampling) on the device.  Fill kinds 7 / 8 of
# vm_particles_fill are the reference's own procedure (2-D Sobol proposals in sequence order, accept-reject in x or
# importance weights, inverse CDF in v); its `rand` calls are Julia's unseeded global RNG there, Philox keyed by
# (seed, proposal index) here.  The host matrix is refreshed so that dist.particles stays the state of record.
function VlasovMethods.draw!(dist::ParticleDistribution{1,1}, fₓ::Base.Callable, params::VlasovMethods.BumpOnTail,
                             sampling::Union{VlasovMethods.AcceptRejectSampling, VlasovMethods.ImportanceSampling}; seed = 20240601)
    dev, cur = get!(PARTICLES, dist) do
        (B.DeviceParticles(length(dist.particles)), Ref(false))
    end
    kind = sampling isa VlasovMethods.ImportanceSampling ? 8 : 7
    B.fill!(dev, kind, Float64[params.ε, params.κ, params.α, params.σ, params.v₀, -1.0]; seed = seed)
    cur[] = true
    tohost!(dist)
    return dist
end

function device(potential::Potential{<:PeriodicBSplineBasis})
    get!(FIELDS, potential) do
        basis = potential.basis
        a, b = boundaries(basis)
        k = order(basis)
        B.DeviceField(a, b, k, length(potential.rhs), k ÷ 2 - k + 1)      # BSplineKit's periodic index rotation
    end
end

function device(sdist::SplineDistribution{1,1})
    get!(VSPLINES, sdist) do
        t = knots(sdist.basis)
        k = order(sdist.basis)
        nknots = length(t) - 2(k - 1)
        B.DeviceVSpline(first(t), last(t), nknots, k; dirichlet = length(sdist.coefficients) == nknots + k - 4)
    end
end

# --------------------------------------------------------------------------- x-space projection and potential --
# src/projections/potential.jl:2-22
function VlasovMethods.projection!(potential::Potential{<:PeriodicBSplineBasis}, distribution::ParticleDistribution)
    f = device(potential)
    B.deposit!(f, device(distribution))
    potential.rhs .= B.rhs(f)
    return potential
end

# src/models/vlasov_poisson.jl:12-15 (PoissonSolvers.update! is replaced by the replicated device solve)
function VlasovMethods.update_potential!(model::VlasovPoisson)
    f = device(model.potential)
    B.deposit!(f, device(model.distribution))
    B.solve!(f)
    model.potential.rhs .= B.rhs(f)
    model.potential.coefficients .= B.coefficients(f)
    return model.potential
end

# Flows of the splitting (src/models/vlasov_poisson.jl:53-67).  GeometricIntegrators calls them with HOST matrices
# z, z̄ (2 x N); the state goes up and comes down once per call -- correct but PCIe-bound; the fused device loop is
# run!(::SplittingMethod) below.  As in the reference, the deposit reads model.distribution.particles (SURVEY F5).
function VlasovMethods.s_advection!(z::AbstractMatrix{Float64}, t, z̄::AbstractMatrix{Float64}, t̄, params)
    p = scratch_particles(size(z̄, 2))
    B.upload!(p; x = vec(z̄[1, :]), v = vec(z̄[2, :]))
    B.drift!(p, t - t̄)
    download_state!(z, p)
end

function VlasovMethods.s_acceleration!(z::AbstractMatrix{Float64}, t, z̄::AbstractMatrix{Float64}, t̄, params)
    VlasovMethods.update_potential!(params.model)
    p = scratch_particles(size(z̄, 2))
    B.upload!(p; x = vec(z̄[1, :]), v = vec(z̄[2, :]))
    B.kick!(device(params.ϕ), p, t - t̄; scale = -1.0)
    download_state!(z, p)
end

# src/models/vlasov_poisson.jl:23-29
function VlasovMethods.lorentz_force!(ż::AbstractMatrix{Float64}, t, z::AbstractMatrix{Float64}, params)
    dist = params.model.distribution
    hostmatrix(dist)[1:2, :] .= z
    mark_host_dirty!(dist)
    n = size(z, 2)
    xdot, vdot = Vector{Float64}(undef, n), Vector{Float64}(undef, n)
    B.check(ccall((:vm_vp_vector_field, B.lib), Cint, (Ptr{Cvoid}, Ptr{Cvoid}, Cint, Ptr{Float64}, Ptr{Float64}),
                  device(params.ϕ).h, device(dist).h, 0, xdot, vdot), B.ctx().h)
    ż[1, :] .= xdot
    ż[2, :] .= vdot
    return ż
end

const SCRATCH = Ref{Union{Nothing, B.DeviceParticles}}(nothing)
scratch_particles(n) = (SCRATCH[] === nothing || SCRATCH[].n != n) ? (SCRATCH[] = B.DeviceParticles(n)) : SCRATCH[]
function download_state!(z::AbstractMatrix{Float64}, p::B.DeviceParticles)
    x, v = Vector{Float64}(undef, p.n), Vector{Float64}(undef, p.n)
    B.check(ccall((:vm_particles_download_soa, B.lib), Cint, (Ptr{Cvoid}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}),
                  p.h, x, v, C_NULL), B.ctx().h)
    z[1, :] .= x
    z[2, :] .= v
    return z
end

# ------------------------------------------------------------------------------------------ run!(SplittingMethod) --
# src/methods/splitting.jl:23-52.  Same datasets ("z", nd x np x nt+1); `save_every` decimates the snapshots (the
# reference writes every step: 1.6 GB per step at 1e8 particles).  field_source = :state is the self-consistent
# reading, :model_ics reproduces the reference as written (field frozen at the model's particles, SURVEY F5).
function VlasovMethods.run!(method::SplittingMethod, h5file; save_every::Int = 1, field_source::Symbol = :state)
    model = method.model
    z₀ = method.equation.ics.q
    nd, np = size(z₀)
    nt = GeometricEquations.ntime(method.equation)
    Δt = GeometricEquations.timestep(method.equation)
    f, p = device(model.potential), device(model.distribution)
    flags = B.RUN_SPLIT_KICK                              # A(Δt/2) B(Δt/2) B(Δt/2) A(Δt/2)
    if field_source == :model_ics
        VlasovMethods.update_potential!(model)
        flags |= B.RUN_FROZEN_FIELD
    end
    nsnap = save_every > 0 ? nt ÷ save_every : 0
    h5 = h5open(h5file, "w")
    try
        h5z = create_dataset(h5, "z", eltype(z₀), ((nd, np, nsnap + 1), (nd, np, -1)), chunk = (nd, np, 1))
        h5z[:, :, 1] = z₀
        z = similar(z₀)
        done = 0
        for s in 1:max(nsnap, 1)
            steps = nsnap > 0 ? save_every : nt
            B.vp_run!(f, p, Δt, steps; flags = flags)
            done += steps
            nsnap > 0 && (download_state!(z, p); h5z[:, :, s + 1] = z)
        end
        done < nt && B.vp_run!(f, p, Δt, nt - done; flags = flags)
    finally
        close(h5)
    end
    tohost!(model.distribution)                            # copy!(model.distribution.particles.z, solstep.q)  (:49)
    return model.distribution
end

# ----------------------------------------------------------------------------- velocity-space projection, LB / CLB --
# src/projections/distribution.jl:35-55 (Float64 only: Dual element types keep the reference's generic method)
function VlasovMethods.projection(velocities::AbstractArray{Float64}, dist::ParticleDistribution, final_dist::SplineDistribution{1,1})
    s = device(final_dist)
    B.check(ccall((:vm_vproject_at, B.lib), Cint, (Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Float64}), s.h, device(dist).h, vec(collect(velocities))), B.ctx().h)
    final_dist.coefficients .= B.coefficients(s)
    return final_dist.spline
end

# src/models/lenard_bernstein_conservative.jl:11-21
function VlasovMethods.compute_coefficients(distribution::SplineDistribution{1,1}, particle_dist::ParticleDistribution, vp::AbstractArray{Float64})
    m5, A = zeros(5), zeros(2)
    B.check(ccall((:vm_vmoments_at, B.lib), Cint, (Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}),
                  device(distribution).h, device(particle_dist).h, vec(collect(vp)), m5, A), B.ctx().h)
    return A[1], A[2]
end

function device_rhs!(v̇, v::AbstractArray{Float64}, params, conservative::Bool)
    sdist = params.model.ent.cache[Float64]
    out = Vector{Float64}(undef, length(v))
    B.check(ccall((:vm_lb_rhs_at, B.lib), Cint, (Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Float64}, Cdouble, Cint, Ptr{Float64}),
                  device(sdist).h, device(params.idist).h, vec(collect(v)), params.ν, conservative, out), B.ctx().h)
    sdist.coefficients .= B.coefficients(device(sdist))     # projection() mutates sdist in the reference (SURVEY 9.6 #6)
    v̇ .= out
end

# src/models/lenard_bernstein.jl:20-34, src/models/lenard_bernstein_conservative.jl:24-50
VlasovMethods.LB_rhs!(v̇, v::AbstractArray{Float64}, params, t) = device_rhs!(v̇, v, params, false)
VlasovMethods.LB_rhs_GI!(v, t, q::AbstractArray{Float64}, params) = device_rhs!(v, q, params, false)
VlasovMethods.CLB_rhs!(v̇, v::AbstractVector{Float64}, params, t) = device_rhs!(v̇, v, params, true)
VlasovMethods.CLB_rhs_GI!(v, t, q::AbstractArray{Float64}, params) = device_rhs!(v, q, params, true)

# ------------------------------------------------------------------------------------- run!(GeometricIntegrator) --
# src/methods/geometric_integrator.jl:12-44: RK438 with the fused stage passes; datasets "z" (np x nt+1) and "t".
function VlasovMethods.run!(method::GeometricIntegrator{<:Union{LenardBernstein{1,1}, ConservativeLenardBernstein{1,1}}}, h5file;
                            save_every::Int = 1, diag_every::Int = 0)
    model = method.model
    t₀ = method.equation.tspan[begin]
    z₀ = method.equation.ics.q
    np = length(z₀)
    nt = GeometricEquations.ntime(method.equation)
    Δt = GeometricEquations.timestep(method.equation)
    s, p = device(model.ent.dist), device(model.dist)
    conservative = model isa ConservativeLenardBernstein
    nsnap = save_every > 0 ? nt ÷ save_every : 0
    h5 = h5open(h5file, "w")
    diags = Matrix{Float64}[]
    try
        h5z = create_dataset(h5, "z", eltype(z₀), ((np, nsnap + 1), (np, -1)), chunk = (np, 1))
        h5t = create_dataset(h5, "t", eltype(t₀), ((nsnap + 1,), (-1,)), chunk = (1,))
        h5z[:, 1] = z₀
        h5t[1] = t₀
        z = zeros(2, np)
        done = 0
        for k in 1:max(nsnap, 1)
            steps = nsnap > 0 ? save_every : nt
            push!(diags, B.lb_rk438_run!(s, p, Δt, steps; ν = model.ν, conservative = conservative, diag_every = diag_every))
            done += steps
            nsnap > 0 && (download_state!(z, p); h5z[:, k + 1] = z[2, :]; h5t[k + 1] = t₀ + done * Δt)
        end
        done < nt && push!(diags, B.lb_rk438_run!(s, p, Δt, nt - done; ν = model.ν, conservative = conservative, diag_every = diag_every))
    finally
        close(h5)
    end
    tohost!(model.dist)                                      # model.dist.particles.v[1,:] .= solstep.q  (:41)
    return model.dist
end

# ------------------------------------------------------------------------------------------------- legacy API --
# src/electric_field.jl + src/vlasov_poisson.jl are present in the reference but not `include`d by the module
# (SURVEY F4).  When a build of VlasovMethods does include them, integrate_vp! (src/vlasov_poisson.jl:70-119) becomes:
if isdefined(VlasovMethods, :integrate_vp!)
    @eval function VlasovMethods.integrate_vp!(P, efield::VlasovMethods.ElectricField, parameters::NamedTuple,
                                               IP::VlasovMethods.VPIntegratorParameters{Float64},
                                               IC::VlasovMethods.VPIntegratorCache{Float64} = VlasovMethods.VPIntegratorCache(IP); save = true)
        inner = efield isa VlasovMethods.ScaledField ? efield.field : efield
        poisson = inner.poisson
        f = get!(() -> B.DeviceField(0.0, poisson.L, poisson.p + 1, IP.nₕ, 0), FIELDS, poisson)
        p = scratch_particles(IP.nₚ)
        IC.x .= P.x[1, :]; IC.v .= P.v[1, :]; IC.w .= P.w[1, :]
        B.upload!(p; x = IC.x, v = IC.v, w = IC.w)
        nsave = save ? div(IP.nₜ, IP.nₛ - 1) : 0
        nrows = nsave > 0 ? IP.nₜ ÷ nsave + 1 : 0
        diag = zeros(4, nrows)
        if inner isa VlasovMethods.ExternalField              # prescribed ϕ(t): src/electric_field.jl:55-77
            coeffs = collect(parent(inner.coeffs))
            B.check(ccall((:vm_vp_run_external, B.lib), Cint,
                          (Ptr{Cvoid}, Ptr{Cvoid}, Cdouble, Cint, Cint, Cdouble, Ptr{Float64}, Cint, Cdouble, Ptr{Float64}),
                          f.h, p.h, IP.dt, IP.nₜ, nsave, parameters.χ, coeffs, size(coeffs, 2), inner.Δt,
                          nrows > 0 ? pointer(diag) : C_NULL), B.ctx().h)
            inner.ts = round(Int, IP.nₜ * IP.dt / inner.Δt)
        else
            diag = B.vp_run!(f, p, IP.dt, IP.nₜ; diag_every = nsave, χ = parameters.χ)
        end
        for ts in 1:min(nrows, IP.nₛ)
            IC.W[ts], IC.K[ts], IC.M[ts] = diag[1, ts], diag[2, ts], diag[3, ts]
        end
        z = zeros(2, IP.nₚ)
        download_state!(z, p)
        IC.x .= z[1, :]; IC.v .= z[2, :]
        IC.ϕ .= B.coefficients(f)
        return IC
    end
end

end # module

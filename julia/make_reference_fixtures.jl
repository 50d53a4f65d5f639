# make_reference_fixtures.jl -- pins the CPU oracle (and through it the CUDA path) to the UNMODIFIED reference.
#
# Run by anyone who has Julia + VlasovMethods.jl v0.2.1 (and its dependencies BSplineKit, PoissonSolvers,
# GeometricIntegrators) installed -- neither exists where this repository is built, which is why parity is
# "unpinned" until the files this script writes are committed:
#
#     julia --project=<env with VlasovMethods> julia/make_reference_fixtures.jl
#
# Inputs : tests/golden/inputs_v1/*.f64   (raw little-endian Float64, written by tests/golden/make_golden.py;
#                                          the same arrays as tests/golden/golden_v1.npz)
# Outputs: tests/golden/reference_v1/*.f64 + manifest.txt (package versions, sizes, detected conventions)
# Consumer: tests/test_reference_fixtures.py (oracle on CPU, CUDA path on the GPU box).
#
# Every call below is a call of the reference's own public API at the call sites the hot path replaces:
#   projection!(potential, dist)                    src/projections/potential.jl:2-22
#   PoissonSolvers.update!(potential), ϕ(x, D(1))   src/models/vlasov_poisson.jl:12-15, 27, 48, 65
#   SplittingMethod(model, tspan, tstep) + integrate!   src/models/vlasov_poisson.jl:73-89, src/methods/splitting.jl:36-43
#   projection(v, dist, sdist)                      src/projections/distribution.jl:35-55
#   compute_coefficients, CLB_rhs_GI!, LB_rhs_GI!   src/models/lenard_bernstein_conservative.jl:11-50, lenard_bernstein.jl:20-34
#   GeometricIntegrator(model, tspan, tstep) (RK438)    src/models/lenard_bernstein_conservative.jl:88-104
using VlasovMethods
using PoissonSolvers
using BSplineKit
using GeometricIntegrators
using LinearAlgebra
using Pkg

const ROOT = normpath(joinpath(@__DIR__, ".."))
const IN   = joinpath(ROOT, "tests", "golden", "inputs_v1")
const OUT  = joinpath(ROOT, "tests", "golden", "reference_v1")
mkpath(OUT)

readf64(name) = (n = filesize(joinpath(IN, name * ".f64")) ÷ 8; v = Vector{Float64}(undef, n); read!(joinpath(IN, name * ".f64"), v); v)
writef64(name, a) = open(io -> write(io, Float64.(vec(collect(a)))), joinpath(OUT, name * ".f64"), "w")

manifest = String[]
note(s) = (push!(manifest, s); println(s))

# ------------------------------------------------------------------ x-space: deposit, solve, gather ----------
# bump-on-tail geometry of scripts/bump_on_tail.jl expressed through the NEW API (Potential on a periodic basis)
x, v, w = readf64("vp_x"), readf64("vp_v"), readf64("vp_w")
npart = length(x)
domain = (0.0, 2π / 0.3)
order, nknot = 4, 16

dist = ParticleDistribution(1, 1, npart)
dist.particles.x .= x'
dist.particles.v .= v'
dist.particles.w .= w'

potential = Potential(PeriodicBasisBSplineKit(domain, order, nknot))
note("vp: length(potential.rhs) = $(length(potential.rhs)) for nknot = $nknot   (n_basis convention, SURVEY 9.1)")
VlasovMethods.projection!(potential, dist)
writef64("vp_rhs", potential.rhs)
PoissonSolvers.update!(potential)
writef64("vp_phi", potential.coefficients)
writef64("vp_dphi", [potential(xi, Derivative(1)) for xi in x])
writef64("vp_phi_at_x", [potential(xi) for xi in x])
note("vp: sum(rhs) - sum(w) = $(sum(potential.rhs) - sum(w))")
note("vp: sum(coefficients) = $(sum(potential.coefficients))   (gauge)")

# one and eight Strang steps of the reference's own SplittingMethod.  NOTE (SURVEY F5): as written the integrator
# advances a COPY of the particle matrix while projection! reads model.distribution.particles, i.e. the field is
# frozen at the initial positions; the self-consistent variant below writes the state back before every step.
model = VlasovPoisson(dist, potential)
tstep = 0.1
function strang_steps(model, nsteps; self_consistent::Bool)
    z0 = copy(model.distribution.particles.z)
    sm = SplittingMethod(model, (0.0, nsteps * tstep), tstep)
    GeometricIntegrators.Integrators.initialize!(sm.integrator)
    for n in 1:nsteps
        GeometricIntegrators.integrate!(sm.integrator)
        self_consistent && copy!(model.distribution.particles.z, sm.integrator.solstep.q)
    end
    q = copy(sm.integrator.solstep.q)
    copy!(model.distribution.particles.z, z0)
    return q
end
for (tag, sc) in (("frozen", false), ("selfconsistent", true))
    q1 = strang_steps(model, 1; self_consistent = sc)
    q8 = strang_steps(model, 8; self_consistent = sc)
    writef64("vp_strang1_$(tag)_x", q1[1, :]); writef64("vp_strang1_$(tag)_v", q1[2, :])
    writef64("vp_strang8_$(tag)_x", q8[1, :]); writef64("vp_strang8_$(tag)_v", q8[2, :])
end

# the script's own default configuration (scripts/vlasov_poisson.jl:6-11): blob on (0,1), quadratic, 16 knots
xs, vs = readf64("st_x"), readf64("st_v")
dist3 = ParticleDistribution(1, 1, length(xs))
dist3.particles.x .= xs'; dist3.particles.v .= vs'; dist3.particles.w .= 1 / length(xs)
pot3 = Potential(PeriodicBasisBSplineKit((0.0, 1.0), 3, 16))
VlasovMethods.projection!(pot3, dist3)
writef64("st_rhs", pot3.rhs)
PoissonSolvers.update!(pot3)
writef64("st_phi", pot3.coefficients)
q5 = strang_steps(VlasovPoisson(dist3, pot3), 5; self_consistent = true)
writef64("st_x5", q5[1, :]); writef64("st_v5", q5[2, :])

# ------------------------------------------------------------------ v-space: projection, LB / CLB, RK438 -----
vv, wv = readf64("lb_v"), readf64("lb_w")
nv = length(vv)
idist = ParticleDistribution(1, 1, nv)
idist.particles.v .= vv'; idist.particles.w .= wv'
sdist = SplineDistribution(1, 1, 41, 4, (-10.0, 10.0), :Dirichlet)
note("lb: length(sdist) = $(length(sdist.coefficients)) for 41 knots, order 4, :Dirichlet")
fs = VlasovMethods.projection(vv, idist, sdist)
writef64("lb_coef", sdist.coefficients)
writef64("lb_mass_matrix", Matrix(sdist.mass_matrix))
dfs = Derivative(1) * fs
writef64("lb_f", fs.(vv)); writef64("lb_df", dfs.(vv))
n, nu, ne = VlasovMethods.compute_f_densities(sdist, vv)
b1, b2 = VlasovMethods.compute_df_densities(sdist, vv)
writef64("lb_m5", [n, nu, ne, b1, b2])
A = VlasovMethods.compute_coefficients(sdist, idist, vv)
writef64("clb_A", [A[1], A[2]])

entropy = CollisionEntropy(sdist)
for (tag, Model, rhs!) in (("lb", LenardBernstein, VlasovMethods.LB_rhs_GI!), ("clb", ConservativeLenardBernstein, VlasovMethods.CLB_rhs_GI!))
    m = Model(idist, entropy)
    params = (ν = m.ν, idist = m.dist, fdist = m.ent.dist, model = m)
    vdot = similar(vv)
    rhs!(vdot, 0.0, vv, params)
    writef64("$(tag)_vdot", vdot)
    gi = GeometricIntegrator(m, (0.0, 3e-2), 1e-2)                  # RK438, three steps
    GeometricIntegrators.Integrators.initialize!(gi.integrator)
    for _ in 1:3
        GeometricIntegrators.integrate!(gi.integrator)
    end
    writef64("$(tag)_v3", gi.integrator.solstep.q)
end

note("julia $(VERSION)")
for (uuid, info) in Pkg.dependencies()
    info.name in ("VlasovMethods", "BSplineKit", "PoissonSolvers", "GeometricIntegrators", "GeometricEquations", "ParticleMethods") &&
        note("$(info.name) $(info.version)")
end
write(joinpath(OUT, "manifest.txt"), join(manifest, "\n") * "\n")
println("wrote ", OUT)

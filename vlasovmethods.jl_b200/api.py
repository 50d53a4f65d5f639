"""Host-side mirror of the VlasovMethods.jl API for the particle hot path.

Same type / function names and argument meaning as the reference (Julia's `f!`
becomes `f_`); every body is a call into the C ABI (libvlasov_b200.so).  The
Julia glue a maintainer would write is shown in INTEGRATION.md; this Python
mirror exists because no Julia toolchain is available where this repo is built
and tested.

Reference files mirrored (paths relative to the reference repo):
  src/distributions/particle_distribution.jl, spline_distribution.jl
  src/entropies/collision_entropy.jl
  src/projections/potential.jl, distribution.jl, density.jl
  src/models/vlasov_poisson.jl, lenard_bernstein.jl, lenard_bernstein_conservative.jl
  src/methods/splitting.jl, geometric_integrator.jl
  src/examples/*.jl, src/sampling/sampling.jl
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field
from typing import Optional, Tuple

import numpy as np

from . import _lib as L
from .core import Context, DeviceField, DeviceParticles, DeviceVSpline, default_context


# =============================================================== sampling tags
class SamplingMethod: ...
class NoSampling(SamplingMethod): ...            # src/sampling/sampling.jl:3-7
class ImportanceSampling(SamplingMethod): ...
class AcceptRejectSampling(SamplingMethod): ...


# ============================================================ ParticleDistribution
class ParticleList:
    """Stand-in for ParticleMethods.ParticleList over a (xdim+vdim+1) x N matrix.

    Host storage `data` is (N, 3) C-ordered, i.e. the same memory as Julia's 3 x N
    column-major matrix; `.x .v .w .z` are row views like the reference's named
    variables (particle_distribution.jl:11-18)."""

    def __init__(self, n: int, host: bool = True):
        self.n = int(n)
        self.data = np.zeros((self.n, 3)) if host else None

    def __len__(self):
        return self.n

    def _need_host(self):
        if self.data is None:
            raise RuntimeError("this ParticleDistribution is device-only; call dist.to_host() first")

    @property
    def x(self): self._need_host(); return self.data.T[0:1, :]
    @property
    def v(self): self._need_host(); return self.data.T[1:2, :]
    @property
    def z(self): self._need_host(); return self.data.T[0:2, :]
    @property
    def w(self): self._need_host(); return self.data.T[2:3, :]


class ParticleDistribution:
    """ParticleDistribution{XD,VD} (src/distributions/particle_distribution.jl:2-24), 1d1v only.

    device_only=True keeps no host copy (needed at 1e8-1e9 particles).
    first_index / total: this process holds particles [first_index, first_index+npart) of `total`
    (one process per GPU; see sharding.py)."""

    def __init__(self, xdim: int, vdim: int, npart: int, *, ctx: Optional[Context] = None, device_only: bool = False,
                 first_index: int = 0, total: Optional[int] = None):
        if (xdim, vdim) != (1, 1):
            raise NotImplementedError("the B200 path implements ParticleDistribution{1,1} only")
        self.xdim, self.vdim = xdim, vdim
        self.particles = ParticleList(npart, host=not device_only)
        self.ctx = ctx
        self.first_index = int(first_index)
        self.total = int(npart if total is None else total)
        self._dev: Optional[DeviceParticles] = None
        self._dev_current = False     # device copy holds the newest state

    def size(self): return (len(self.particles),)
    def __len__(self): return len(self.particles)

    # ---- host <-> device ----
    def device(self) -> DeviceParticles:
        """Device SoA copy, uploaded on first use or after host edits (call mark_host_dirty())."""
        if self._dev is None:
            self._dev = DeviceParticles(self.ctx or default_context(), len(self.particles))
            self._dev_current = self.particles.data is None
        if not self._dev_current:
            self._dev.upload_aos(self.particles.data)
            self._dev_current = True
        return self._dev

    def mark_host_dirty(self):
        self._dev_current = False

    def to_host(self):
        """Copy the device state back into the host matrix (allocating it if device-only)."""
        if self._dev is not None:
            if self.particles.data is None:
                self.particles.data = np.empty((len(self.particles), 3))
            self._dev.download_aos(self.particles.data)
        return self


def xdim(d: ParticleDistribution): return d.xdim
def vdim(d: ParticleDistribution): return d.vdim


# ====================================================================== examples
@dataclass
class NormalDistribution:          # src/examples/normal.jl:2-7
    domain: Tuple[float, float] = (0.0, 1.0)

@dataclass
class BumpOnTail:                  # src/examples/bumpontail.jl:2-12
    ε: float = 0.03
    κ: float = 0.3
    α: float = 0.1
    σ: float = 0.5
    v0: float = 4.5

@dataclass
class DoubleMaxwellian:            # src/examples/doublemaxwellian.jl:1-7
    domain: Tuple[float, float] = (-5.0, 5.0)
    shift: float = 3.0

@dataclass
class UniformDistribution:         # src/examples/uniform.jl:1-7
    xdomain: Tuple[float, float] = (0.0, 1.0)
    vdomain: Tuple[float, float] = (-2.0, 2.0)

@dataclass
class ShiftedNormalV:              # src/examples/shiftednormalv.jl:1-7
    domain: Tuple[float, float] = (-5.0, 5.0)
    shift: float = 2.0

@dataclass
class ShiftedUniformDistribution:  # src/examples/shifteduniform.jl:1-8
    xdomain: Tuple[float, float] = (0.0, 1.0)
    vdomain: Tuple[float, float] = (-2.0, 2.0)
    shift: float = 2.0

@dataclass
class LandauDamping:               # (1 + eps cos(kappa x)) Maxwellian; benchmark load, SURVEY 8d
    ε: float = 0.01
    κ: float = 0.5


def _fill_args(params):
    if isinstance(params, NormalDistribution):
        return L.VM_FILL_NORMAL, [params.domain[0], params.domain[-1]]
    if isinstance(params, BumpOnTail):
        return L.VM_FILL_BUMP_ON_TAIL, [params.ε, params.κ, params.α, params.σ, params.v0]
    if isinstance(params, DoubleMaxwellian):
        return L.VM_FILL_DOUBLE_MAXWELLIAN, [params.domain[0], params.domain[-1], params.shift]
    if isinstance(params, UniformDistribution):
        return L.VM_FILL_UNIFORM, [params.xdomain[0], params.xdomain[-1], params.vdomain[0], params.vdomain[-1]]
    if isinstance(params, ShiftedNormalV):
        return L.VM_FILL_SHIFTED_NORMAL_V, [params.domain[0], params.domain[-1], params.shift]
    if isinstance(params, ShiftedUniformDistribution):
        return L.VM_FILL_SHIFTED_UNIFORM, [params.xdomain[0], params.xdomain[-1], params.vdomain[0], params.vdomain[-1], params.shift]
    if isinstance(params, LandauDamping):
        return L.VM_FILL_LANDAU, [params.ε, params.κ]
    raise TypeError(f"unknown example {type(params).__name__}")


def initialize_(dist: ParticleDistribution, params, sampling: SamplingMethod = None, *, seed: int = 20240601):
    """initialize!(dist, example[, sampling]) (src/examples/*.jl).

    The reference samples from Julia's unseeded global RNG; here the same distributions are
    generated on the device by a counter-based generator keyed by (seed, global particle index)."""
    kind, p = _fill_args(params)
    if isinstance(params, BumpOnTail) and not isinstance(sampling, NoSampling):
        # draw!(dist, f_x, params, ::AcceptRejectSampling) -- the default of initialize!(dist, ::BumpOnTail) -- and
        # ::ImportanceSampling (bumpontail.jl:43-75, 90-121): Sobol proposals as in the reference; NoSampling() selects the
        # inverse-CDF load of the same density (what bench.py fills with)
        kind = L.VM_FILL_BUMP_ON_TAIL_SOBOL_IS if isinstance(sampling, ImportanceSampling) else L.VM_FILL_BUMP_ON_TAIL_SOBOL
        p = list(p) + [-1.0]
    if dist._dev is None:
        dist._dev = DeviceParticles(dist.ctx or default_context(), len(dist.particles))
    dist._dev.fill(kind, p, seed, dist.first_index, dist.total)
    dist._dev_current = True
    if dist.particles.data is not None:
        dist.to_host()
    return dist


# ======================================================================== fields
@dataclass
class PeriodicBasisBSplineKit:
    """PoissonSolvers.PeriodicBasisBSplineKit(domain, order, nknot) (scripts/vlasov_poisson.jl:21).

    n_basis defaults to nknot (unpinned); index_shift to BSplineKit's order//2 - order + 1."""
    domain: Tuple[float, float]
    order: int
    nknot: int
    n_basis: Optional[int] = None
    index_shift: Optional[int] = None

    def __post_init__(self):
        if self.n_basis is None:
            self.n_basis = self.nknot
        if self.index_shift is None:
            self.index_shift = self.order // 2 - self.order + 1


class Potential:
    """PoissonSolvers.Potential(basis): fields .basis .rhs .coefficients, update!(potential)."""

    def __init__(self, basis: PeriodicBasisBSplineKit, *, ctx: Optional[Context] = None):
        self.basis = basis
        self.ctx = ctx or default_context()
        self.field = DeviceField(self.ctx, basis.domain[0], basis.domain[-1], basis.order, basis.n_basis, basis.index_shift)

    @property
    def rhs(self): return self.field.rhs
    @property
    def coefficients(self): return self.field.coefficients
    @coefficients.setter
    def coefficients(self, phi): self.field.coefficients = phi

    def __call__(self, x, derivative: int = 0):
        """phi(x) / phi(x, Derivative(1)) (src/models/vlasov_poisson.jl:27,48,65)."""
        out = self.field.eval(x, derivative)
        return out if np.ndim(x) else float(out[0])

    def mass_matrix(self): return self.field.mass_matrix()
    def stiffness_matrix(self): return self.field.stiffness_matrix()


def update_(potential: Potential):
    """PoissonSolvers.update!(potential): all-reduce + solve for the coefficients."""
    potential.field.solve()
    return potential


def projection_(potential: Potential, distribution: ParticleDistribution, *, mode: int = L.VM_DEPOSIT_DETERMINISTIC):
    """projection!(potential, dist): x-space charge deposit (src/projections/potential.jl:2-22)."""
    potential.field.deposit(distribution.device(), mode)
    return potential


# ========================================================== SplineDistribution etc.
class Spline:
    """View of the projected spline f_s: callable, with a derivative view (Derivative(1) * spline)."""

    def __init__(self, vs: DeviceVSpline, deriv: int = 0):
        self._vs, self._deriv = vs, deriv

    def __call__(self, v):
        f, df = self._vs.eval(v)
        out = df if self._deriv else f
        return out if np.ndim(v) else float(out[0])

    def derivative(self):
        if self._deriv:
            raise NotImplementedError("only first derivatives are available")
        return Spline(self._vs, 1)


class SplineDistribution:
    """SplineDistribution(xdim, vdim, nknots, order, domain, bc) (spline_distribution.jl:23-36)."""

    def __init__(self, xdim: int, vdim: int, nknots: int, order: int, domain: Tuple[float, float], bc: str = "Dirichlet",
                 *, ctx: Optional[Context] = None):
        if (xdim, vdim) != (1, 1):
            raise NotImplementedError("the B200 path implements SplineDistribution{1,1} only")
        self.xdim, self.vdim = xdim, vdim
        self.ctx = ctx or default_context()
        self.vs = DeviceVSpline(self.ctx, domain[0], domain[-1], nknots, order, 1 if bc == "Dirichlet" else 0)
        self.spline = Spline(self.vs)

    def size(self): return (self.vs.nv,)
    def __len__(self): return self.vs.nv
    @property
    def coefficients(self): return self.vs.coefficients
    @coefficients.setter
    def coefficients(self, c): self.vs.coefficients = c
    @property
    def mass_matrix(self): return self.vs.mass_matrix()


class CollisionEntropy:
    """CollisionEntropy(sdist) (src/entropies/collision_entropy.jl:1-10); Float64 only, so the
    per-eltype cache of the reference collapses to the distribution itself."""

    def __init__(self, dist: SplineDistribution):
        self.dist = dist
        self.cache = {np.float64: dist}


def projection(velocities, dist: ParticleDistribution, final_dist: SplineDistribution):
    """projection(v, dist, sdist) (src/projections/distribution.jl:35-55): deposit + mass solve.

    `velocities` is None (use the distribution's device state) or a host vector used instead of it; like the
    reference, a replacement vector only feeds the projection -- dist.particles (host and device) stay untouched."""
    dev = dist.device()
    if velocities is not None:
        final_dist.vs.project_at(dev, np.asarray(velocities, dtype=np.float64).reshape(-1))
    else:
        final_dist.vs.project(dev)
    return final_dist.spline


def _moments(distribution: SplineDistribution, dist: ParticleDistribution, v=None):
    if v is None:
        return distribution.vs.moments(dist.device())
    return distribution.vs.moments_at(dist.device(), np.asarray(v, dtype=np.float64).reshape(-1))


def compute_f_densities(distribution: SplineDistribution, dist: ParticleDistribution, v=None):
    """(n, n u, n eps) = unweighted particle sums of f_s, v f_s, v^2 f_s (density.jl:6-13); v: the `vp` argument
    of the reference (default: the distribution's own velocities)."""
    m5, _ = _moments(distribution, dist, v)
    return m5[0], m5[1], m5[2]


def compute_df_densities(distribution: SplineDistribution, dist: ParticleDistribution, v=None):
    m5, _ = _moments(distribution, dist, v)
    return m5[3], m5[4]


def compute_coefficients(distribution: SplineDistribution, particle_dist: ParticleDistribution, v=None):
    """A1, A2 (lenard_bernstein_conservative.jl:11-21)."""
    _, A = _moments(distribution, particle_dist, v)
    return A[0], A[1]


# ======================================================================== models
class VlasovPoisson:                      # src/models/vlasov_poisson.jl:2-9
    def __init__(self, dist: ParticleDistribution, potential: Potential):
        self.distribution, self.potential = dist, potential


def update_potential_(model: VlasovPoisson):   # :12-15
    projection_(model.potential, model.distribution)
    update_(model.potential)


# Splitting flows and vector fields of src/models/vlasov_poisson.jl:23-67, acting on the model's device state.
# (z, zbar are implicit: the particle state lives on the device; dt = t - tbar.)
def s_advection_(model: VlasovPoisson, dt: float):
    """s_advection! (:53-58): x <- x + dt * v."""
    model.distribution.device().drift(dt)


def s_acceleration_(model: VlasovPoisson, dt: float):
    """s_acceleration! (:61-67): update_potential! then v <- v - dt * phi'(x)."""
    update_potential_(model)
    model.potential.field.kick(model.distribution.device(), dt, -1.0)


def lorentz_force_(model: VlasovPoisson, *, to_host: bool = True):
    """lorentz_force! (:23-29): (xdot, vdot) = (v, -phi'(x)) with the potential refreshed from the state.
    One C-ABI call (vm_vp_vector_field); to_host=False keeps both on the device (vdot in the handle's work array,
    xdot is the v array) and returns (None, None) -- at 1e8 particles the host round trip is 1.6 GB per call."""
    return model.potential.field.vector_field(model.distribution.device(), False, to_host)


def v_advection_(model: VlasovPoisson):
    """v_advection! (:36-41): (xdot, vdot) = (v, 0)."""
    v = model.distribution.device().download(x=False, w=False)[1]
    return v, np.zeros_like(v)


def v_acceleration_(model: VlasovPoisson):
    """v_acceleration! (:44-50): (xdot, vdot) = (0, -phi'(x))."""
    xdot, vdot = lorentz_force_(model)
    return np.zeros_like(xdot), vdot


class LenardBernstein:                    # src/models/lenard_bernstein.jl:1-9
    conservative = False

    def __init__(self, dist: ParticleDistribution, ent: CollisionEntropy, ν: float = 1.0):
        self.dist, self.ent, self.ν = dist, ent, ν


class ConservativeLenardBernstein(LenardBernstein):   # lenard_bernstein_conservative.jl:1-9
    conservative = True


def LB_rhs_(model: LenardBernstein, v=None):
    """LB_rhs! / CLB_rhs! (lenard_bernstein.jl:20-30, lenard_bernstein_conservative.jl:24-36):
    vdot for the model's particles (optionally with replacement velocities v)."""
    dev = model.dist.device()
    if v is not None:      # stage values of a user-side integrator: the particle state is not modified
        return model.ent.dist.vs.lb_rhs_at(dev, np.asarray(v, dtype=np.float64).reshape(-1), model.ν, model.conservative)
    return model.ent.dist.vs.lb_rhs(dev, model.ν, model.conservative)


CLB_rhs_ = LB_rhs_


# ======================================================================= methods
def _ntime(tspan, tstep):
    return int(round((tspan[1] - tspan[0]) / tstep))


class SplittingMethod:
    """SplittingMethod(model::VlasovPoisson, tspan, tstep): Strang splitting
    (src/models/vlasov_poisson.jl:73-89, src/methods/splitting.jl:2-52).

    field_source = "state"     self-consistent field from the advancing particles (default)
                 = "model_ics" field frozen at the model's initial particles: bug-compatible with the
                               reference as written (the integrator advances a copy)."""

    def __init__(self, model: VlasovPoisson, tspan, tstep, *, field_source: str = "state"):
        assert field_source in ("state", "model_ics")
        self.model, self.tspan, self.tstep, self.field_source = model, tuple(tspan), float(tstep), field_source


class GeometricIntegrator:
    """GeometricIntegrator(model::(Conservative)LenardBernstein, tspan, tstep): explicit RK438
    (lenard_bernstein.jl:68-84, lenard_bernstein_conservative.jl:88-104, geometric_integrator.jl:1-44)."""

    def __init__(self, model: LenardBernstein, tspan, tstep):
        self.model, self.tspan, self.tstep = model, tuple(tspan), float(tstep)


class DiffEqIntegrator:
    """TRBDF2 + ForwardDiff driver of the reference (src/methods/diffeq_integrator.jl): needs dual-number
    eltypes and a dense N x N Jacobian -- out of scope for the Float64 GPU path."""

    def __init__(self, *a, **k):
        raise NotImplementedError("DiffEqIntegrator (implicit TRBDF2 with AD Jacobian) is not part of the B200 hot path")


def run_(method, h5file: Optional[str] = None, *, save_every: int = 0, diag_every: int = 0):
    """run!(method, h5file).

    The reference writes the full state to HDF5 after EVERY step (splitting.jl:42); at 1e8
    particles that is 1.6 GB per step, so snapshots are decimated: every `save_every`-th step
    (0 = final state only) is stored.  Without an HDF5 library the container is a .npz with the
    reference's dataset names: z[nd, np, nt] (and t[nt]).
    Returns the model's distribution with the final state copied back (splitting.jl:49)."""
    nt = _ntime(method.tspan, method.tstep)
    if diag_every > 0 and save_every > 0 and save_every % diag_every != 0:
        # every chunk of save_every steps then starts on a diagnostics step, so the cadence continues across chunks
        raise ValueError("save_every must be a multiple of diag_every")
    if isinstance(method, SplittingMethod):
        return _run_splitting(method, nt, h5file, save_every, diag_every)
    if isinstance(method, GeometricIntegrator):
        return _run_rk438(method, nt, h5file, save_every, diag_every)
    raise TypeError(type(method))


def _chunks(nt, save_every):
    if save_every <= 0:
        return [nt] if nt > 0 else []
    out = [save_every] * (nt // save_every)
    if nt % save_every:
        out.append(nt % save_every)
    return out


class _SnapshotWriter:
    """Decimated, asynchronous snapshots: the device->host copy of snapshot i (pinned double buffers, second
    stream) overlaps with the steps of chunk i+1; replaces the per-step HDF5 write of splitting.jl:42."""

    def __init__(self, dev: DeviceParticles, nsnaps: int, with_x: bool):
        from .core import PinnedArray
        self.dev, self.with_x = dev, with_x
        self.bufs = [(PinnedArray(dev.n) if with_x else None, PinnedArray(dev.n)) for _ in range(2)]
        self.z = np.empty((2, dev.n, nsnaps)) if with_x else np.empty((dev.n, nsnaps))
        self.i = 0            # snapshots started
        self.done = 0         # snapshots stored

    def begin(self):
        bx, bv = self.bufs[self.i % 2]
        self.dev.snapshot_begin(bx.array if bx is not None else None, bv.array)
        self.i += 1

    def finish(self):
        if self.done == self.i:
            return
        self.dev.snapshot_wait()
        bx, bv = self.bufs[self.done % 2]
        if self.with_x:
            self.z[0, :, self.done] = bx.array
            self.z[1, :, self.done] = bv.array
        else:
            self.z[:, self.done] = bv.array
        self.done += 1


def _run_splitting(method: SplittingMethod, nt, h5file, save_every, diag_every):
    model = method.model
    dist, pot = model.distribution, model.potential
    dev = dist.device()
    flags = L.VM_RUN_SPLIT_KICK
    if method.field_source == "model_ics":
        update_potential_(model)          # phi of the model's (initial) particles, never refreshed
        flags |= L.VM_RUN_FROZEN_FIELD
    chunks = _chunks(nt, save_every)
    keep = h5file is not None or save_every > 0
    snaps = _SnapshotWriter(dev, len(chunks) + 1, True) if keep else None
    times, diags = [method.tspan[0]], []
    if keep:
        snaps.begin()
    done = 0
    dtimes = []
    for n in chunks:
        d = pot.field.run(dev, method.tstep, n, diag_every, flags, 1.0)     # enqueued; overlaps the pending snapshot copy
        if d is not None:                      # rows at steps done, done + diag_every, ... (row 0 repeats the previous chunk's last)
            tt = method.tspan[0] + (done + diag_every * np.arange(d.shape[0])) * method.tstep
            diags.append(d if not diags else d[1:])
            dtimes.append(tt if not dtimes else tt[1:])
        done += n
        times.append(method.tspan[0] + done * method.tstep)
        if keep:
            snaps.finish()
            snaps.begin()
    if keep:
        snaps.finish()
    method.diagnostics = np.concatenate(diags) if diags else None            # rows [W, K, M, sum_w]
    method.diagnostics_t = np.concatenate(dtimes) if dtimes else None
    if h5file is not None:
        np.savez(h5file, z=snaps.z, t=np.asarray(times))
    if dist.particles.data is not None:
        dist.to_host()                    # copy!(model.distribution.particles.z, solstep.q)
    return dist


def _run_rk438(method: GeometricIntegrator, nt, h5file, save_every, diag_every):
    model = method.model
    dev = model.dist.device()
    vs = model.ent.dist.vs
    chunks = _chunks(nt, save_every)
    keep = h5file is not None or save_every > 0
    snaps = _SnapshotWriter(dev, len(chunks) + 1, False) if keep else None
    times, diags = [method.tspan[0]], []
    if keep:
        snaps.begin()
    done = 0
    for n in chunks:
        d = vs.rk438_run(dev, method.tstep, n, model.ν, model.conservative, diag_every)
        if d is not None:
            d = d.copy(); d[:, 0] += method.tspan[0] + done * method.tstep
            diags.append(d if not diags else d[1:])
        done += n
        times.append(method.tspan[0] + done * method.tstep)
        if keep:
            snaps.finish()
            snaps.begin()
    if keep:
        snaps.finish()
    method.diagnostics = np.concatenate(diags) if diags else None
    if h5file is not None:
        np.savez(h5file, z=snaps.z, t=np.asarray(times))
    if model.dist.particles.data is not None:
        model.dist.to_host()              # model.dist.particles.v[1,:] .= solstep.q
    return model.dist

"""Legacy API of the reference (files present but not `include`d in VlasovMethods.jl v0.2.1):
src/electric_field.jl (field functors) and src/vlasov_poisson.jl (integrate_vp!), as used by
scripts/bump_on_tail.jl and test/electric_field_tests.jl.  Bodies are C-ABI calls."""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Optional

import numpy as np

from . import _lib as L
from .core import Context, DeviceField, DeviceParticles, default_context


class PoissonSolverPBSplines:
    """PoissonSolvers.PoissonSolverPBSplines(p, nh, L): degree-p periodic B-splines, nh elements on [0, L)
    (scripts/bump_on_tail.jl:38).  Fields .ϕ (coefficients) and .S (stiffness matrix) as used by
    src/electric_field.jl:47-49."""

    def __init__(self, p: int, nh: int, L_: float, *, ctx: Optional[Context] = None, index_shift: int = 0):
        self.p, self.nh, self.L = int(p), int(nh), float(L_)
        self.ctx = ctx or default_context()
        self.field = DeviceField(self.ctx, 0.0, self.L, self.p + 1, self.nh, index_shift)
        self._scratch: Optional[DeviceParticles] = None

    @property
    def ϕ(self): return self.field.coefficients
    @ϕ.setter
    def ϕ(self, c): self.field.coefficients = c
    phi = ϕ
    @property
    def S(self): return self.field.stiffness_matrix()
    @property
    def rhs(self): return self.field.rhs

    def _particles(self, n: int) -> DeviceParticles:
        if self._scratch is None or self._scratch.n != n:
            self._scratch = DeviceParticles(self.ctx, n)
        return self._scratch


def solve_(poisson: PoissonSolverPBSplines, x, w):
    """PoissonSolvers.solve!(poisson, x, w): deposit + solve from host arrays."""
    x = np.ascontiguousarray(x, dtype=np.float64).reshape(-1)
    p = poisson._particles(x.size)
    p.upload(x=x, w=np.ascontiguousarray(w, dtype=np.float64).reshape(-1))
    poisson.field.deposit(p)
    poisson.field.solve()


def eval_field_(e, poisson: PoissonSolverPBSplines, x):
    """PoissonSolvers.eval_field!(e, poisson, x): e = -phi'(x)."""
    e[...] = -poisson.field.eval(np.asarray(x, dtype=np.float64).reshape(-1), 1).reshape(np.shape(e))
    return e


class ElectricField:                      # src/electric_field.jl:2-17
    def __call__(self, *args):
        if len(args) == 4:                # f(e, x, w, t)
            e, x, w, t = args
            update_(self, x, w, t)
            return efield_(self, e, x)
        if len(args) == 2:                # f(e, x)
            return efield_(self, *args)
        if len(args) == 3:                # f(x, w, t) -> e
            x, w, t = args
            e = np.zeros_like(np.asarray(x, dtype=np.float64))
            return self(e, x, w, t)
        raise TypeError("ElectricField functor takes (e,x,w,t), (e,x) or (x,w,t)")


class PoissonField(ElectricField):        # :39-51
    def __init__(self, poisson: PoissonSolverPBSplines):
        self.poisson = poisson


class ExternalField(ElectricField):       # :55-77
    def __init__(self, poisson: PoissonSolverPBSplines, coeffs, Δt: float):
        self.poisson, self.coeffs, self.Δt, self.ts = poisson, np.asarray(coeffs, dtype=np.float64), float(Δt), 0


class ScaledField(ElectricField):         # :21-35
    def __init__(self, field: ElectricField, χ: float):
        self.field, self.χ = field, float(χ)


def ScaledPoissonField(poisson, χ): return ScaledField(PoissonField(poisson), χ)
def ScaledExternalField(poisson, coeffs, Δt, χ): return ScaledField(ExternalField(poisson, coeffs, Δt), χ)


def update_(f: ElectricField, x, w, t):
    if isinstance(f, ScaledField):
        return update_(f.field, x, w, t)
    if isinstance(f, PoissonField):
        return solve_(f.poisson, x, w)
    if isinstance(f, ExternalField):      # :66-69
        f.ts = int(round(t / f.Δt))
        f.poisson.ϕ = f.coeffs[:, f.ts]
        return None
    raise TypeError(type(f))


def efield_(f: ElectricField, e, x):
    if isinstance(f, ScaledField):        # :26-29
        efield_(f.field, e, x)
        e /= f.χ ** 2
        return e
    return eval_field_(e, f.poisson, x)


def energy(f: ElectricField) -> float:
    if isinstance(f, ScaledField):
        return energy(f.field) / f.χ ** 2
    return f.poisson.field.energy()       # dot(phi, S, phi) / 2  (:47, :75)


def coefficients(f: ElectricField):
    return coefficients(f.field) if isinstance(f, ScaledField) else f.poisson.ϕ


def _poisson_of(f: ElectricField) -> PoissonSolverPBSplines:
    return _poisson_of(f.field) if isinstance(f, ScaledField) else f.poisson


class VPIntegratorParameters:             # src/vlasov_poisson.jl:5-18
    def __init__(self, dt: float, nt: int, ns: int, nh: int, np_: int):
        self.dt, self.nₜ, self.nₛ, self.nₕ, self.nₚ = float(dt), int(nt), int(ns), int(nh), int(np_)
        self.t = np.linspace(0.0, dt * nt, ns)


class VPIntegratorCache:                  # :21-56
    """x, v, w final state; Φ coefficient history; W, K, M histories.  The N x ns trajectory
    histories X, V, A of the reference are only kept when trajectories=True."""

    def __init__(self, IP: VPIntegratorParameters, trajectories: bool = False):
        n, ns, nh = IP.nₚ, IP.nₛ, IP.nₕ
        self.x, self.v, self.a, self.w = np.zeros(n), np.zeros(n), np.zeros(n), np.zeros(n)
        self.ϕ = np.zeros(nh)
        self.Φ = np.zeros((nh, ns))
        self.W, self.K, self.M = np.zeros(ns), np.zeros(ns), np.zeros(ns)
        self.X = np.zeros((n, ns)) if trajectories else None
        self.V = np.zeros((n, ns)) if trajectories else None
        self.A = np.zeros((n, ns)) if trajectories else None


def _run_external_from(fld, dev, dt, steps, inner, steps_done, diag_every, χ):
    """Continue an ExternalField run after `steps_done` steps: vm_vp_run_external counts time from 0, so the
    coefficient history is handed over re-indexed (column j of the slice = round((steps_done + it) dt / Δt))."""
    cols = [int(round((steps_done + it) * dt / inner.Δt)) for it in range(steps + 1)]      # update!: ts = round(t / Δt)
    return fld.run_external(dev, dt, steps, inner.coeffs[:, cols], dt, diag_every, χ)


def integrate_vp_(P, efield: ElectricField, parameters, IP: VPIntegratorParameters,
                  IC: Optional[VPIntegratorCache] = None, *, save: bool = True):
    """integrate_vp!(P, efield, parameters, IP, IC) (src/vlasov_poisson.jl:70-119).

    P: object with .x .v .w (length-N arrays or 1xN views).  parameters: mapping with key "χ"."""
    IC = IC or VPIntegratorCache(IP)
    χ = float(parameters["χ"] if isinstance(parameters, dict) else getattr(parameters, "χ"))
    inner = efield.field if isinstance(efield, ScaledField) else efield
    external = isinstance(inner, ExternalField)
    poisson = _poisson_of(efield)
    fld = poisson.field
    n = IP.nₚ
    dev = poisson._particles(n)
    dev.upload(np.asarray(P.x).reshape(-1), np.asarray(P.v).reshape(-1), np.asarray(P.w).reshape(-1))
    IC.w[:] = np.asarray(P.w).reshape(-1)
    nsave = IP.nₜ // (IP.nₛ - 1) if IP.nₛ > 1 else 0               # :77
    need_hist = save and (IC.X is not None)
    t_done = [0]                                                   # steps taken so far (ExternalField indexes by time)

    def advance(steps, diag_every):
        if external:      # prescribed phi(t): no deposit / solve, the coefficient history stays on the device
            d = _run_external_from(fld, dev, IP.dt, steps, inner, t_done[0], diag_every, χ)
            inner.ts = int(round((t_done[0] + steps) * IP.dt / inner.Δt))
        else:
            d = fld.run(dev, IP.dt, steps, diag_every, 0, χ)
        t_done[0] += steps
        return d

    if not need_hist:
        diag = advance(IP.nₜ, nsave if save else 0)
        if save and diag is not None:
            m = min(diag.shape[0], IP.nₛ)
            IC.W[:m], IC.K[:m], IC.M[:m] = diag[:m, 0], diag[:m, 1], diag[:m, 2]
    else:
        def snapshot(ts):
            x, v, _ = dev.download(w=False)
            if external:      # update!(efield, x, w, t) only selects the column; K, M on the host (trajectory mode is small-N)
                update_(inner, None, None, t_done[0] * IP.dt)
                d = [energy(efield), 0.5 * float(np.dot(IC.w * v, v)), float(np.dot(IC.w, v))]
            else:
                d = fld.diagnostics(dev, χ)
            IC.W[ts], IC.K[ts], IC.M[ts] = d[0], d[1], d[2]
            IC.Φ[:, ts] = fld.coefficients
            IC.X[:, ts], IC.V[:, ts] = x, v
            IC.A[:, ts] = fld.gather_E(dev, 1.0 / χ ** 2)
        snapshot(0)
        done, ts = 0, 0
        while done < IP.nₜ:
            step = min(nsave, IP.nₜ - done) if nsave > 0 else IP.nₜ - done
            advance(step, 0)
            done += step
            if nsave > 0 and done % nsave == 0 and ts + 1 < IP.nₛ:
                ts += 1
                snapshot(ts)
    x, v, _ = dev.download(w=False)
    IC.x[:], IC.v[:] = x, v
    IC.ϕ[:] = fld.coefficients
    return IC

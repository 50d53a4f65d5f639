"""GPU parity tests added in round 2 (all through the C ABI, checker = the CPU oracle):

  * fused cross-CTA finish on meshes of 200 .. 1024 cells (two-level reduction inside the pass kernel)
  * ExternalField time loop (vm_vp_run_external), device-resident lorentz_force! and its RK4 driver
  * declared uniform weight, replacement velocities that leave the particle state alone
  * the reference's own projections test (test/projections_tests.jl:6-34) through the CUDA path
  * v-space quantities at the north star's 1e-12 with the measured errors printed
  * full-length runs of the script configurations: histories [W, K, M] / [sum v, sum v^2] against the oracle

Tolerances: relative <= 1e-12 on deposited moments and fields (north star).  Where a quantity is conditioned
worse than that (mass-matrix solve, long chaotic trajectories) the measured error is printed and the assert sits
within a factor ~10-30 of the errors measured on B200 (recorded next to each assert).
"""
import math

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

RTOL = 1e-12


def relmax(a, b):
    return np.max(np.abs(np.asarray(a) - np.asarray(b))) / max(np.max(np.abs(b)), 1e-300)


@pytest.fixture(scope="module")
def ctx(vm):
    c = vm.Context(0)
    yield c
    c.close()


def report(name, **errs):
    print(f"[measured] {name}: " + ", ".join(f"{k}={v:.3e}" for k, v in errs.items()))


# ------------------------------------------------------- large-mesh finish ---
@pytest.mark.parametrize("n", [130, 200, 256, 512, 1024])
@pytest.mark.parametrize("k", [3, 4])
def test_two_level_finish_matches_oracle_and_unfused(vm, oracle, rng, n, k):
    """n_h > 128: the per-CTA rows are reduced inside the pass kernel (group tickets, two levels) and the solve is
    the multi-CTA kernel; must agree with the oracle and with the separate reduce kernel (tuning no_fuse)."""
    a, b = 0.0, 2 * math.pi / 0.3
    npart = 250_001
    x = rng.uniform(a - (b - a), b + (b - a), npart); v = rng.standard_normal(npart)
    w = rng.uniform(0.5, 1.5, npart) * (b - a) / npart
    S = oracle.periodic_stiffness(a, b, n, k, 0)
    rhs_ref = oracle.deposit_periodic(x, w, a, b, n, k, 0)
    xo, vo = x.copy(), v.copy()
    dref, phiref = oracle.integrate_vp(xo, vo, w, 0.1, 1.0, 4, 2, a, b, n, k, 0, S, want_phi=True)
    got = {}
    for no_fuse in (0, 1):
        c = vm.Context(0)
        c.set_tuning("no_fuse", no_fuse)
        fld = vm.DeviceField(c, a, b, k, n, 0)
        p = vm.DeviceParticles(c, npart)
        p.upload(x, v, w)
        fld.deposit(p, 0)
        assert relmax(fld.rhs, rhs_ref) <= RTOL, (n, k, no_fuse)
        r1 = fld.rhs.tobytes()
        fld.deposit(p, 0)
        assert fld.rhs.tobytes() == r1                      # run-to-run bitwise
        diag = fld.run(p, 0.1, 4, 2, 0, 1.0)
        xg, vg, _ = p.download(w=False)
        assert np.max(np.abs(xg - xo)) <= 1e-11 and np.max(np.abs(vg - vo)) <= 1e-11
        assert np.allclose(diag[:, :3], dref, rtol=1e-10, atol=1e-13)
        assert relmax(fld.coefficients, phiref[-1]) <= 1e-10
        got[no_fuse] = (xg, vg)
        fld.close(); p.close(); c.close()
    assert np.max(np.abs(got[0][0] - got[1][0])) <= 1e-12


# ----------------------------------------------------------- ExternalField ---
@pytest.mark.parametrize("chi", [1.0, 0.7])
def test_external_field_run_matches_oracle(vm, oracle, ctx, rng, chi):
    """integrate_vp! with an ExternalField (src/electric_field.jl:55-77): prescribed phi(t), gather only."""
    a, b, n, k = 0.0, 2 * math.pi / 0.3, 16, 4
    npart, dt, nt, nsave = 20001, 0.1, 24, 4
    x = rng.uniform(a, b, npart); v = rng.standard_normal(npart); w = rng.uniform(0.5, 1.5, npart) * (b - a) / npart
    dt_c = 0.25                                             # coarser than the step: several steps share a column
    ncols = int(round(nt * dt / dt_c)) + 1
    coeffs = rng.standard_normal((n, ncols)); coeffs -= coeffs.mean(axis=0)
    S = oracle.periodic_stiffness(a, b, n, k, 0)
    xo, vo = x.copy(), v.copy()
    dref = oracle.integrate_vp_external(xo, vo, w, dt, chi, nt, nsave, a, b, n, k, 0, S, coeffs, dt_c)
    fld = vm.DeviceField(ctx, a, b, k, n, 0)
    p = vm.DeviceParticles(ctx, npart)
    p.upload(x, v, w)
    diag = fld.run_external(p, dt, nt, coeffs, dt_c, nsave, chi)
    xg, vg, _ = p.download(w=False)
    report("external run", dx=np.max(np.abs(xg - xo)), dv=np.max(np.abs(vg - vo)))
    assert np.max(np.abs(xg - xo)) <= 1e-11 and np.max(np.abs(vg - vo)) <= 1e-11
    assert diag.shape == (nt // nsave + 1, 4)
    assert np.allclose(diag[:, :3], dref, rtol=1e-11, atol=1e-13)
    last = int(np.rint(nt * dt / dt_c))
    assert np.array_equal(fld.coefficients, coeffs[:, last])          # poisson.phi holds the last column used
    with pytest.raises(vm.VMError):
        fld.run_external(p, dt, nt + 50, coeffs, dt_c, 0, chi)         # time index outside the history


def test_integrate_vp_mirror_with_external_and_scaled_fields(vm, oracle, ctx, rng):
    """The legacy mirror no longer refuses ExternalField: a recorded self-consistent run replayed as an
    ExternalField reproduces the trajectory (x positions at the kick are the same, so is the field)."""
    nh, pdeg, L_, npart, dt, nt = 16, 3, 2 * math.pi / 0.3, 5000, 0.1, 10
    x = rng.uniform(0, L_, npart); v = rng.standard_normal(npart); w = np.full(npart, L_ / npart)

    class P: pass
    P.x, P.v, P.w = x, v, w
    chi = 0.8
    poisson = vm.PoissonSolverPBSplines(pdeg, nh, L_, ctx=ctx)
    IP = vm.VPIntegratorParameters(dt, nt, nt + 1, nh, npart)
    S = oracle.periodic_stiffness(0.0, L_, nh, pdeg + 1, 0)
    # phi at the half-drift positions of every step = what the kick of step `it` used
    xo, vo = x.copy(), v.copy()
    cols = [np.zeros(nh)]
    for it in range(nt):
        xo += 0.5 * dt * chi * vo
        phi = oracle.poisson_solve(S, oracle.deposit_periodic(xo, w, 0.0, L_, nh, pdeg + 1, 0))
        cols.append(phi)
        vo += dt * chi * (-oracle.eval_dphi(xo, 0.0, L_, nh, pdeg + 1, 0, phi) / chi ** 2)
        xo += 0.5 * dt * chi * vo
    coeffs = np.column_stack(cols)
    ic_self = vm.integrate_vp_(P, vm.ScaledPoissonField(poisson, chi), {"χ": chi}, IP, save=False)
    ext = vm.ScaledExternalField(poisson, coeffs, dt, chi)
    ic_ext = vm.integrate_vp_(P, ext, {"χ": chi}, IP, save=True)
    assert np.max(np.abs(ic_self.x - xo)) <= 1e-11 and np.max(np.abs(ic_self.v - vo)) <= 1e-11
    assert np.max(np.abs(ic_ext.x - xo)) <= 1e-11 and np.max(np.abs(ic_ext.v - vo)) <= 1e-11
    assert ext.field.ts == nt
    Kref = 0.5 * np.dot(w * vo, vo)
    assert abs(ic_ext.K[nt] - Kref) <= 1e-12 * Kref
    assert abs(ic_ext.W[nt] - oracle.field_energy(S, coeffs[:, nt]) / chi ** 2) <= 1e-11 * abs(ic_ext.W[nt])


# ------------------------------------------------- lorentz_force! and RK4 ----
def test_vector_field_device_resident_and_rk4(vm, oracle, ctx, rng):
    a, b, n, k = 0.0, 1.0, 16, 3
    npart = 30001
    shift = oracle.bspline_shift_bsplinekit(k)
    x = rng.uniform(a, b, npart); v = rng.standard_normal(npart); w = np.full(npart, 1.0 / npart)
    S = oracle.periodic_stiffness(a, b, n, k, shift)
    fld = vm.DeviceField(ctx, a, b, k, n, shift)
    p = vm.DeviceParticles(ctx, npart)
    p.upload(x, v, w)
    xd, vd = fld.vector_field(p)
    xr, vr = oracle.lorentz_force(x, v, w, a, b, n, k, shift, S)
    assert np.array_equal(xd, xr) and relmax(vd, vr) <= RTOL
    assert fld.vector_field(p, to_host=False) == (None, None)            # nothing crosses PCIe
    xs, vs, _ = p.download(w=False)
    assert np.array_equal(xs, x) and np.array_equal(vs, v)               # the state is untouched
    assert relmax(fld.gather_E(p, 1.0), vr) <= RTOL                      # ... and phi is the refreshed one
    fld.coefficients = np.zeros(n)
    _, v0 = fld.vector_field(p, keep_potential=True)
    assert np.all(v0 == 0)
    xo, vo = x.copy(), v.copy()
    for _ in range(6):
        oracle.vp_rk4_step(xo, vo, w, 0.05, a, b, n, k, shift, S)
    fld.rk4_run(p, 0.05, 6)
    xg, vg, _ = p.download(w=False)
    report("rk4 run", dx=np.max(np.abs(xg - xo)), dv=np.max(np.abs(vg - vo)))
    assert np.max(np.abs(xg - xo)) <= 1e-12 and np.max(np.abs(vg - vo)) <= 1e-12
    # mirror
    vm.set_default_context(ctx)
    dist = vm.ParticleDistribution(1, 1, npart)
    dist.particles.data[:, 0], dist.particles.data[:, 1], dist.particles.data[:, 2] = x, v, w
    model = vm.VlasovPoisson(dist, vm.Potential(vm.PeriodicBasisBSplineKit((a, b), k, n)))
    xd2, vd2 = vm.lorentz_force_(model)
    assert np.array_equal(xd2, xr) and relmax(vd2, vr) <= RTOL
    assert vm.lorentz_force_(model, to_host=False) == (None, None)
    vm.set_default_context(None)


# --------------------------------------------- order-independent deposit -----
@pytest.mark.parametrize("n", [16, 24, 64, 256, 1024])
def test_fixed_point_deposit_is_bitwise_independent_of_geometry(vm, oracle, rng, n):
    """VM_DEPOSIT_FIXED (SURVEY 8c KAT 10): contributions are rounded once to 64-bit fixed point and summed as
    integers, so the deposited vector -- and a whole run -- has the same bits for every CTA shape, replica layout
    (lane-private / bank-sorted queues) and pipeline depth, and stays within 1e-12 of the oracle."""
    k = 4
    a, b = 0.0, 2 * math.pi / 0.3
    npart = 300_001
    x = rng.uniform(a - (b - a), b + (b - a), npart); v = rng.standard_normal(npart)
    tunings = [{}, {"bankq": 1}, {"ctas_per_sm": 1, "threads_per_cta": 256, "replicas": 32}, {"ctas_per_sm": 1, "threads_per_cta": 128, "replicas": 32},
               {"no_uniform_w": 1}]
    for w in (np.full(npart, (b - a) / npart), rng.uniform(0.5, 1.5, npart) * (b - a) / npart):
        ref = oracle.deposit_periodic(x, w, a, b, n, k, 0)
        outs = []
        for tune in tunings:
            c = vm.Context(0)
            for key, val in tune.items():
                c.set_tuning(key, val)
            fld = vm.DeviceField(c, a, b, k, n, 0)
            p = vm.DeviceParticles(c, npart)
            p.upload(x, v, w)
            try:
                fld.deposit(p, vm._lib.VM_DEPOSIT_FIXED)
            except vm.VMError as e:          # a hand-tuned shape whose replica grids do not fit this mesh
                assert "do not fit" in str(e) or "no fixed-point" in str(e), e
                fld.close(); p.close(); c.close()
                continue
            rhs = fld.rhs
            assert relmax(rhs, ref) <= RTOL, (n, tune)
            fld.run(p, 0.1, 5, 0, vm._lib.VM_RUN_FIXED_DEPOSIT, 1.0)
            xs, vs_, _ = p.download(w=False)
            outs.append((tune, rhs.tobytes(), xs.tobytes(), vs_.tobytes(), fld.coefficients.tobytes()))
            fld.close(); p.close(); c.close()
        assert len(outs) >= 3
        for o in outs[1:]:
            assert o[1:] == outs[0][1:], (n, o[0])
    # and the run agrees with the fp64-accumulating default to rounding
    c = vm.Context(0)
    fld = vm.DeviceField(c, a, b, k, n, 0)
    p = vm.DeviceParticles(c, npart)
    p.upload(x, v, w)
    fld.run(p, 0.1, 5, 0, 0, 1.0)
    xd = p.download(w=False)[0]
    assert np.max(np.abs(xd - np.frombuffer(outs[0][2]))) <= 1e-11
    c.close()


@pytest.mark.parametrize("n,k", [(20, 4), (33, 2), (64, 3), (64, 4), (200, 5), (256, 4), (512, 6), (1024, 4)])
def test_limb_atomic_deposit_matches_oracle_and_is_order_independent(vm, oracle, rng, n, k):
    """The default layout from 20 cells on: 64-bit fixed-point rows as two 32-bit limbs, native shared-
    memory atomics with an exact carry, bank-steered replicas shared by the CTA.  Mixed-sign per-particle weights (the
    two's-complement carry path), positions far outside the domain; against the oracle at 1e-12, and the SAME BITS for
    every replica count / CTA shape, for the explicit VM_DEPOSIT_FIXED mode, and for the bank-sorted fixed-point layout."""
    a, b = 0.0, 2 * math.pi / 0.3
    npart = 200_001
    x = rng.uniform(a - 3 * (b - a), b + 3 * (b - a), npart); v = rng.standard_normal(npart)
    w = rng.standard_normal(npart) * (b - a) / npart                      # both signs
    ref = oracle.deposit_periodic(x, w, a, b, n, k, 0)
    scale = np.sum(np.abs(w)) / n                                         # (cancellation: compare against the mean |deposit|)
    S = oracle.periodic_stiffness(a, b, n, k, 0)
    xo, vo = x.copy(), v.copy()
    oracle.integrate_vp(xo, vo, w, 0.1, 1.0, 4, 0, a, b, n, k, 0, S)
    outs = []
    tunings = [({}, 0), ({}, 2), ({"af_replicas": 1}, 0), ({"af_replicas": 8}, 0), ({"af_replicas": 16}, 0), ({"af_ctas": 2}, 0),
               ({"af_ctas": 4}, 0), ({"no_repg": 1}, 0), ({"no_uniform_w": 1}, 0), ({"no_fuse": 1}, 0), ({"bankq": 1}, 2)]
    for tune, mode in tunings:
        c = vm.Context(0)
        if n < 88 and "bankq" not in tune:
            tune = dict(tune, af=1)          # below 88 cells the deposit-only CALL is lane-private by default (the fused step is not)
        for key, val in tune.items():
            c.set_tuning(key, val)
        plan = vm._lib.pass_plan(n, k, 1)
        assert plan.variant == 5                                          # (the default plan of this mesh: limb atomics)
        fld = vm.DeviceField(c, a, b, k, n, 0)
        p = vm.DeviceParticles(c, npart)
        p.upload(x, v, w)
        fld.deposit(p, mode)
        rhs = fld.rhs
        err = np.max(np.abs(rhs - ref)) / scale
        assert err <= RTOL, (n, k, tune, mode, err)
        p.upload(x, v, w)
        fld.run(p, 0.1, 4, 0, vm._lib.VM_RUN_FIXED_DEPOSIT if mode == 2 else 0, 1.0)
        xs, vs_, _ = p.download(w=False)
        assert np.max(np.abs(xs - xo)) <= 1e-11 * (b - a) and np.max(np.abs(vs_ - vo)) <= 1e-11
        if "no_fuse" in tune:            # separate fp64 reduce of converted CTA rows: equal to rounding only
            assert relmax(rhs, np.frombuffer(outs[0][1])) <= 1e-13
        else:
            outs.append((tune, rhs.tobytes(), xs.tobytes(), vs_.tobytes()))
        fld.close(); p.close(); c.close()
    report(f"limb-atomic deposit n={n} k={k}", rhs=np.max(np.abs(np.frombuffer(outs[0][1]) - ref)) / scale)
    for o in outs[1:]:
        assert o[1:] == outs[0][1:], (n, k, o[0])
    # the fp64-accumulating layouts it replaced stay selectable (tuning af = -1: lane-private tiers / bank-sorted queues)
    c = vm.Context(0)
    c.set_tuning("af", -1)
    fld = vm.DeviceField(c, a, b, k, n, 0)
    p = vm.DeviceParticles(c, npart)
    p.upload(x, v, w)
    fld.deposit(p, 0)
    assert np.max(np.abs(fld.rhs - ref)) / scale <= RTOL, (n, k, "af=-1")
    fld.run(p, 0.1, 4, 0, 0, 1.0)
    xs, vs_, _ = p.download(w=False)
    assert np.max(np.abs(xs - xo)) <= 1e-11 * (b - a) and np.max(np.abs(vs_ - vo)) <= 1e-11
    fld.close(); p.close(); c.close()


def test_limb_atomic_edge_cases(vm, oracle, rng):
    """Limb-atomic layout at its edges: no particle, one particle on knots / domain ends / far outside, odd counts (the
    tail particle), every particle in one cell (32-way same-address atomics), a mesh above the fused finish (per-CTA
    conversion + separate reduce kernel), and short fused runs of 1 and 3 particles."""
    a, b, k = 0.0, 1.0, 4
    c = vm.Context(0)
    c.set_tuning("af", 1)                       # also for the deposit-only call below 88 cells
    for n in (24, 64, 2048):
        fld = vm.DeviceField(c, a, b, k, n, 0)
        p0 = vm.DeviceParticles(c, 0)
        fld.deposit(p0, 0)
        assert np.all(fld.rhs == 0), n
        for xv in [0.0, 1.0, 0.5, 1.0 - 1e-17, -1e-17, 1.0 / n, 7.0, -7.0, 100.25]:
            p1 = vm.DeviceParticles(c, 1)
            p1.upload(np.array([xv]), np.array([0.0]), np.array([-2.5]))
            fld.deposit(p1, 0)
            ref = oracle.deposit_periodic(np.array([xv]), np.array([-2.5]), a, b, n, k, 0)
            assert np.max(np.abs(fld.rhs - ref)) <= 1e-12 * 2.5, (n, xv)
            p1.close()
        for npart in (3, 65, 100_001):
            x = rng.uniform(0.5, 0.5 + 1.0 / n, npart)                       # all in one cell
            w = np.full(npart, 1.0 / npart)
            p = vm.DeviceParticles(c, npart)
            p.upload(x, np.zeros(npart), w)
            fld.deposit(p, 0)
            assert relmax(fld.rhs, oracle.deposit_periodic(x, w, a, b, n, k, 0)) <= RTOL, (n, npart)
            p.close()
        fld.close(); p0.close()
    # fused runs with fewer particles than threads
    n = 64
    S = oracle.periodic_stiffness(a, b, n, k, 0)
    fld = vm.DeviceField(c, a, b, k, n, 0)
    for npart in (1, 3):
        x = rng.uniform(a, b, npart); v = rng.standard_normal(npart); w = np.full(npart, (b - a) / npart)
        p = vm.DeviceParticles(c, npart)
        p.upload(x, v, w)
        d = fld.run(p, 0.05, 6, 2, 0, 1.0)
        xo, vo = x.copy(), v.copy()
        dref = oracle.integrate_vp(xo, vo, w, 0.05, 1.0, 6, 2, a, b, n, k, 0, S)
        xs, vs_, _ = p.download(w=False)
        assert np.max(np.abs(xs - xo)) <= 1e-12 and np.max(np.abs(vs_ - vo)) <= 1e-12, npart
        assert np.allclose(d[:, :3], dref, rtol=1e-10, atol=1e-13), npart
        p.close()
    fld.close(); c.close()


def test_fixed_point_scale_edge_cases(vm, oracle, ctx):
    """One particle, huge / tiny / mixed-sign weights, zero weights: the scale keeps every contribution exact enough."""
    a, b, n, k = 0.0, 1.0, 16, 4
    fld = vm.DeviceField(ctx, a, b, k, n, 0)
    cases = [(np.array([0.37]), np.array([2.5])),
             (np.linspace(0.01, 0.99, 1000), np.full(1000, 1e-30)),
             (np.linspace(0.01, 0.99, 1000), np.full(1000, 1e30)),
             (np.linspace(0.01, 0.99, 1001), np.where(np.arange(1001) % 2 == 0, 1.0, -0.75) / 1001),
             (np.linspace(0.01, 0.99, 100), np.zeros(100))]
    for x, w in cases:
        p = vm.DeviceParticles(ctx, x.size)
        p.upload(x, np.zeros(x.size), w)
        fld.deposit(p, vm._lib.VM_DEPOSIT_FIXED)
        ref = oracle.deposit_periodic(x, w, a, b, n, k, 0)
        scale = max(np.max(np.abs(ref)), np.sum(np.abs(w)) * 1e-3, 1e-300)
        assert np.max(np.abs(fld.rhs - ref)) <= 1e-12 * scale, (x.size, w[0])
        p.close()


# ------------------------------------------------------ weights / scratch ----
def test_declared_uniform_weight_is_bitwise_the_uploaded_one(vm, ctx, rng):
    a, b, n, k = 0.0, 2 * math.pi / 0.3, 16, 4
    npart = 100_001
    x = rng.uniform(a, b, npart); v = rng.standard_normal(npart); w0 = (b - a) / npart
    fld = vm.DeviceField(ctx, a, b, k, n, 0)
    outs = []
    for declared in (False, True):
        p = vm.DeviceParticles(ctx, npart)
        if declared:
            p.upload(x, v, None)
            p.set_uniform_weight(w0)
        else:
            p.upload(x, v, np.full(npart, w0))
        d = fld.run(p, 0.1, 5, 5, 0, 1.0)
        xs, vs, ws = p.download()
        assert np.all(ws == w0)
        outs.append((xs.tobytes(), vs.tobytes(), d.tobytes()))
        p.close()
    assert outs[0] == outs[1]


def test_replacement_velocities_do_not_touch_the_particle_state(vm, oracle, ctx, rng):
    """projection(v, dist, sdist) / CLB_rhs!(vdot, v, ...) with stage values v: dist.particles stays as it was
    (the reference only reads dist.particles.w there, src/projections/distribution.jl:47)."""
    a, b, nknots, k = -10.0, 10.0, 41, 4
    npart = 20001
    v = rng.standard_normal(npart) * 1.5; w = np.full(npart, 1.0 / npart)
    vstage = v + 0.1 * rng.standard_normal(npart)
    M = oracle.dirichlet_mass(a, b, nknots, k)
    vm.set_default_context(ctx)
    dist = vm.ParticleDistribution(1, 1, npart)
    dist.particles.data[:, 1], dist.particles.data[:, 2] = v, w
    sdist = vm.SplineDistribution(1, 1, nknots, k, (a, b), "Dirichlet")
    model = vm.ConservativeLenardBernstein(dist, vm.CollisionEntropy(sdist))
    vdot = vm.CLB_rhs_(model, vstage)
    ref, coef, A = oracle.lb_rhs(vstage, w, a, b, nknots, k, M, 1.0, True)
    assert relmax(vdot, ref) <= 1e-11
    assert np.array_equal(dist.device().download()[1], v)                 # device state untouched
    vm.projection(vstage, dist, sdist)
    assert relmax(sdist.coefficients, coef) <= 1e-11
    A1, A2 = vm.compute_coefficients(sdist, dist, vstage)
    assert np.allclose([A1, A2], A, rtol=1e-9)
    assert np.array_equal(dist.device().download()[1], v)
    vdot0 = vm.CLB_rhs_(model)                                            # own velocities
    ref0, _, _ = oracle.lb_rhs(v, w, a, b, nknots, k, M, 1.0, True)
    assert relmax(vdot0, ref0) <= 1e-11
    vm.set_default_context(None)


# ------------------------------------ the reference's own test, CUDA path ----
def test_reference_projections_test_through_cuda(vm, oracle, ctx):
    """test/projections_tests.jl:6-34 restated: N = 1e6 samples of f on (0,1), periodic order 5, 32 functions,
    projection!(potential, dist), rho = M \\ rhs compared with f at x = 0:0.1:1 minus two points each end,
    atol 5e-2 (the reference's own tolerance).  Sampler: truncated-normal inverse CDF (f is a Gaussian in
    4 pi x) instead of adaptive rejection sampling -- same distribution."""
    from scipy.special import ndtri, ndtr
    npart, nknot, order, sigma = 1_000_000, 32, 5, 2.0
    f = lambda xx: np.exp(-0.5 * (4 * np.pi * xx - 2 * np.pi) ** 2 / sigma ** 2) * np.sqrt(np.pi * sigma ** 2) / np.sqrt(2)
    rng = np.random.default_rng(12345)
    s = sigma / (4 * np.pi)                                  # f ~ N(0.5, s) truncated to (0, 1)
    lo, hi = ndtr((0 - 0.5) / s), ndtr((1 - 0.5) / s)
    x = 0.5 + s * ndtri(lo + (hi - lo) * rng.uniform(size=npart))
    w = np.full(npart, 1.0 / npart)
    shift = oracle.bspline_shift_bsplinekit(order)
    fld = vm.DeviceField(ctx, 0.0, 1.0, order, nknot, shift)
    p = vm.DeviceParticles(ctx, npart)
    p.upload(x, np.zeros(npart), w)
    fld.deposit(p, 0)                                        # projection!(potential, dist)
    rhs = fld.rhs
    rho = np.linalg.solve(fld.mass_matrix(), rhs)            # potential.solver.Mfac \ potential.rhs
    xs = np.arange(0.0, 1.0 + 1e-12, 0.1)
    vals = []
    for xx in xs:                                            # rho(x) = sum_j rho_{c+j+shift} B_j(x)
        c, N, _ = oracle.periodic_eval(0.0, 1.0, nknot, order, xx)
        vals.append(sum(rho[(c + j + shift) % nknot] * N[j] for j in range(order)))
    vals = np.array(vals)
    cutoff = 2
    err = np.max(np.abs(f(xs)[cutoff:-cutoff] - vals[cutoff:-cutoff]))
    report("projections_tests.jl restated (CUDA deposit)", max_abs_err=err)
    assert err <= 5e-2
    assert relmax(rhs, oracle.deposit_periodic(x, w, 0.0, 1.0, nknot, order, shift)) <= RTOL


# --------------------------------------------- v-space at the stated 1e-12 ---
@pytest.mark.parametrize("nknots,k", [(41, 4), (41, 3), (17, 5), (129, 4)])
def test_vspace_quantities_at_north_star_tolerance(vm, oracle, ctx, rng, nknots, k):
    a, b = -10.0, 10.0
    npart = 60001
    v = np.concatenate([rng.standard_normal(npart // 2) + 2.0, rng.standard_normal(npart - npart // 2) - 2.0])
    w = np.full(npart, 1.0 / npart)
    M = oracle.dirichlet_mass(a, b, nknots, k)
    vs = vm.DeviceVSpline(ctx, a, b, nknots, k, 1)
    p = vm.DeviceParticles(ctx, npart)
    p.upload(np.zeros(npart), v, w)
    vs.project(p)
    coef, rhs = oracle.vproject(v, w, a, b, nknots, k, M)
    f, df = vs.eval(v)
    fr, dfr = oracle.vspline_eval(v, a, b, nknots, k, coef)
    m5, A = vs.moments(p)
    m5r = oracle.vmoments(v, a, b, nknots, k, coef)
    e = dict(rhs=relmax(vs.rhs, rhs), coef=relmax(vs.coefficients, coef), f=relmax(f, fr), df=relmax(df, dfr),
             m5=float(np.max(np.abs(m5 - m5r) / np.max(np.abs(m5r)))))
    for cons in (False, True):
        vdot = vs.lb_rhs(p, 1.3, cons)
        ref, _, Aref = oracle.lb_rhs(v, w, a, b, nknots, k, M, 1.3, cons)
        e["vdot_clb" if cons else "vdot_lb"] = relmax(vdot, ref)
        if cons:
            e["A"] = float(np.max(np.abs(np.asarray(A) - np.asarray(Aref)) / np.max(np.abs(Aref))))
    report(f"v-space nknots={nknots} order={k}", **e)
    # measured on B200 (round 2, gpurun_out r02a): every entry between 1.2e-16 and 2.1e-14 for these bases
    # (worst: f' and the CLB right-hand side at 129 knots); asserted at the north star's 1e-12
    assert e["rhs"] <= RTOL and e["coef"] <= RTOL and e["f"] <= RTOL
    assert e["df"] <= 1e-12 and e["m5"] <= 1e-12
    assert e["vdot_lb"] <= 1e-12 and e["vdot_clb"] <= 1e-12
    assert e["A"] <= 1e-12


# ------------------------------------------- full-length script histories ----
def test_config1_vlasov_poisson_script_full_length(vm, oracle, ctx):
    """scripts/vlasov_poisson.jl:6-11 at its own length: N = 1e4, 16 knots, order 3, dt = 0.1, 200 steps of the
    new-API Strang composition (self-consistent field); energy / momentum history every 10 steps."""
    npart, nknot, order, tstep, nt = 10_000, 16, 3, 0.1, 200
    rng = np.random.default_rng(42)
    z = rng.standard_normal(npart)
    X = math.ceil(np.max(np.abs(z)))
    x = (z + X) / (2 * X); v = rng.standard_normal(npart); w = np.full(npart, 1.0 / npart)
    shift = oracle.bspline_shift_bsplinekit(order)
    S = oracle.periodic_stiffness(0.0, 1.0, nknot, order, shift)
    fld = vm.DeviceField(ctx, 0.0, 1.0, order, nknot, shift)
    p = vm.DeviceParticles(ctx, npart)
    p.upload(x, v, w)
    xo, vo = x.copy(), v.copy()
    hist_g, hist_o = [], []
    for blk in range(nt // 10):
        fld.run(p, tstep, 10, 0, vm._lib.VM_RUN_SPLIT_KICK, 1.0)
        for _ in range(10):
            oracle.vp_strang_step(xo, vo, w, tstep, 0.0, 1.0, nknot, order, shift, S)
        d = fld.diagnostics(p, 1.0)
        phi = oracle.poisson_solve(S, oracle.deposit_periodic(xo, w, 0.0, 1.0, nknot, order, shift))
        hist_g.append(d[:3])
        hist_o.append([oracle.field_energy(S, phi), 0.5 * np.dot(w * vo, vo), np.dot(w, vo)])
    hist_g, hist_o = np.array(hist_g), np.array(hist_o)
    xg, vg, _ = p.download(w=False)
    ex, ev = np.max(np.abs(xg - xo)), np.max(np.abs(vg - vo))
    eh = np.max(np.abs(hist_g - hist_o), axis=0) / np.max(np.abs(hist_o), axis=0)
    report("config 1, 200 Strang steps", dx=ex, dv=ev, W=eh[0], K=eh[1], M=eh[2])
    assert ex <= 2e-12 and ev <= 1e-12          # measured 7.1e-14 / 9.8e-15 (rounding differences grow along trajectories)
    assert np.all(eh <= 1e-12)                  # measured W 1.9e-14, K 4e-16, M 9e-16


def test_config2_bump_on_tail_script_full_length(vm, oracle, ctx):
    """scripts/bump_on_tail.jl:14-32 at its own size: N = 5e4, n_h = 16, p = 3, dt = 0.1, 500 steps of the legacy
    leapfrog loop, [W, K, M] rows every 5 steps vs the oracle's integrate_vp!."""
    npart, nh, pdeg, dt, nt, nsave = 50_000, 16, 3, 0.1, 500, 5
    eps, kappa, alpha, sigma, v0 = 0.03, 0.3, 0.1, 0.5, 4.5
    L_ = 2 * math.pi / kappa
    p = vm.DeviceParticles(ctx, npart)
    p.fill(vm._lib.VM_FILL_BUMP_ON_TAIL, [eps, kappa, alpha, sigma, v0], 7)
    x, v, w = p.download()
    S = oracle.periodic_stiffness(0.0, L_, nh, pdeg + 1, 0)
    xo, vo = x.copy(), v.copy()
    dref = oracle.integrate_vp(xo, vo, w, dt, 1.0, nt, nsave, 0.0, L_, nh, pdeg + 1, 0, S)
    fld = vm.DeviceField(ctx, 0.0, L_, pdeg + 1, nh, 0)
    diag = fld.run(p, dt, nt, nsave, 0, 1.0)
    xg, vg, _ = p.download(w=False)
    ex, ev = np.max(np.abs(xg - xo)), np.max(np.abs(vg - vo))
    eh = np.max(np.abs(diag[:, :3] - dref), axis=0) / np.max(np.abs(dref), axis=0)
    report("config 2, 500 leapfrog steps", dx=ex, dv=ev, W=eh[0], K=eh[1], M=eh[2])
    assert diag.shape == (nt // nsave + 1, 4)
    assert ex <= 1e-8 and ev <= 3e-9            # measured 2.5e-10 / 6.1e-11 (t = 50: trajectories are weakly chaotic,
    #                                             the bump-on-tail instability amplifies rounding differences)
    assert eh[0] <= 3e-12 and eh[1] <= 1e-12 and eh[2] <= 1e-12      # measured W 7.1e-14, K 3.2e-15, M 3.7e-16
    etot = diag[:, 0] + diag[:, 1]
    assert abs(etot[-1] - etot[0]) / etot[0] <= 1e-3


def test_config4_conservative_lb_long_run(vm, oracle, ctx):
    """scripts/lenard_bernstein_conservative.jl:10-18: N = 1e3, 41 knots, order 4, dt = 1e-2, DoubleMaxwellian(+-2),
    RK438; 2 000 of its 50 000 steps, [sum v, sum v^2] rows every 100 steps vs the oracle."""
    npart, dt, nt, every = 1000, 1e-2, 2000, 100
    a, b, nknots, k = -10.0, 10.0, 41, 4
    rng = np.random.default_rng(3)
    v = np.concatenate([rng.standard_normal(npart // 2) + 2.0, rng.standard_normal(npart - npart // 2) - 2.0])
    w = np.full(npart, 1.0 / npart)
    M = oracle.dirichlet_mass(a, b, nknots, k)
    vs = vm.DeviceVSpline(ctx, a, b, nknots, k, 1)
    p = vm.DeviceParticles(ctx, npart)
    p.upload(np.zeros(npart), v, w)
    diag = vs.rk438_run(p, dt, nt, 1.0, True, every)
    vg = p.download(x=False, w=False)[1]
    vo = v.copy()
    rows = [[vo.sum(), (vo ** 2).sum()]]
    for s in range(1, nt + 1):
        oracle.lb_rk438_step(vo, w, dt, a, b, nknots, k, M, 1.0, True)
        if s % every == 0:
            rows.append([vo.sum(), (vo ** 2).sum()])
    rows = np.array(rows)
    ev = np.max(np.abs(vg - vo))
    e1 = np.max(np.abs(diag[:, 1] - rows[:, 0])) / npart
    e2 = np.max(np.abs(diag[:, 2] - rows[:, 1])) / np.max(rows[:, 1])
    report("config 4, 2000 RK438 steps", dv=ev, sum_v=e1, sum_v2=e2)
    assert diag.shape == (nt // every + 1, 4) and np.allclose(diag[:, 0], dt * every * np.arange(nt // every + 1))
    assert ev <= 1e-12                          # measured 4.0e-15
    assert e1 <= 1e-13 and e2 <= 1e-13          # measured 3.4e-16 / 1.8e-16
    # what the script prints (:64): relative momentum / energy drift of the conservative operator
    assert abs(diag[-1, 1] - diag[0, 1]) / npart <= 1e-6 and abs(diag[-1, 2] - diag[0, 2]) / diag[0, 2] <= 1e-6


def test_run_diagnostics_cadence_across_chunks(vm, ctx):
    """run_(method, save_every, diag_every): rows and their times continue across snapshot chunks (incl. the tail)."""
    vm.set_default_context(ctx)
    npart = 2000
    dist = vm.initialize_(vm.ParticleDistribution(1, 1, npart), vm.NormalDistribution(), seed=5)
    model = vm.VlasovPoisson(dist, vm.Potential(vm.PeriodicBasisBSplineKit((0.0, 1.0), 3, 16)))
    integ = vm.SplittingMethod(model, (0.0, 1.4), 0.1)       # 14 steps = 3 chunks of 4 + a tail of 2
    vm.run_(integ, None, save_every=4, diag_every=2)
    assert integ.diagnostics.shape == (8, 4)
    assert np.allclose(integ.diagnostics_t, 0.2 * np.arange(8))
    with pytest.raises(ValueError):
        vm.run_(integ, None, save_every=3, diag_every=2)
    vm.set_default_context(None)


# ------------------------------------------------------- in-pass solve -------
@pytest.mark.parametrize("n", [130, 256, 1000, 1024])
def test_in_pass_solve_is_bitwise_the_solve_kernel(vm, oracle, rng, n):
    """Meshes above 128 cells: between fused passes the Poisson solve runs in the prologue of the next pass
    (pass_presolve) instead of k_poisson_solve.  Same tile routine: trajectories, diagnostics and the final potential
    have the same bits as with the separate kernel (tuning no_presolve), through runs with and without diagnostic
    steps in between, and match the oracle."""
    a, b, k = 0.0, 2 * math.pi / 0.3, 4
    npart = 60_001
    x = rng.uniform(a, b, npart); v = rng.standard_normal(npart); w = np.full(npart, (b - a) / npart)
    S = oracle.periodic_stiffness(a, b, n, k, 0)
    xo, vo = x.copy(), v.copy()
    dref = oracle.integrate_vp(xo, vo, w, 0.1, 1.0, 12, 4, a, b, n, k, 0, S)
    outs = []
    for no_presolve in (0, 1):
        c = vm.Context(0)
        c.set_tuning("no_presolve", no_presolve)
        fld = vm.DeviceField(c, a, b, k, n, 0)
        p = vm.DeviceParticles(c, npart)
        p.upload(x, v, w)
        d = fld.run(p, 0.1, 12, 4, 0, 1.0)              # fused stretches of 3 steps between diagnostic steps
        fld.run(p, 0.1, 7, 0, 0, 1.0)                   # and a stretch without any
        xs, vs_, _ = p.download(w=False)
        outs.append((d.tobytes(), xs.tobytes(), vs_.tobytes(), fld.coefficients.tobytes()))
        assert np.allclose(d[:, :3], dref, rtol=1e-10, atol=1e-13), (n, no_presolve)
        fld.close(); p.close(); c.close()
    assert outs[0] == outs[1], n


# ------------------------------------------------------------- Sobol loads ---
def test_sobol_bump_on_tail_loads(vm, ctx):
    """draw!(dist, f_x, ::BumpOnTail, ::ImportanceSampling / ::AcceptRejectSampling) (bumpontail.jl:90-121, 43-75) on the
    device: proposals are the points of the unscrambled 2-D Sobol sequence in Gray-code order (checked against scipy's
    generator, point for point), kept in order; accept-reject thins them against f_x / (1 + eps)."""
    from scipy.stats import qmc
    from scipy.special import ndtri
    eps, kappa, alpha, sigma, v0 = 0.03, 0.3, 0.1, 0.5, 4.5
    L = 2 * math.pi / kappa
    n = 300_001
    skip = 1
    while 2 * skip <= 2 * n + 1:
        skip *= 2                                   # Sobol.jl: skip(s, 2N) skips the largest power of two <= 2N + 1
    nprop = int(n * (1 + eps) * 1.02) + 65536
    sob = qmc.Sobol(2, scramble=False, bits=32)
    sob.fast_forward(skip + 1)                      # (scipy's point 0 is the origin, which Sobol.jl leaves out)
    y = sob.random(nprop)
    p = vm.DeviceParticles(ctx, n)
    # importance sampling: particle i IS proposal i
    p.fill(vm._lib.VM_FILL_BUMP_ON_TAIL_SOBOL_IS, [eps, kappa, alpha, sigma, v0, -1.0], 7)
    x, v, w = p.download()
    assert np.max(np.abs(x - y[:n, 0] * L)) <= 1e-12 * L
    assert np.max(np.abs(w - (1 - eps * np.cos(kappa * x)) * L / n)) <= 1e-15
    bulk = ndtri(y[:n, 1])
    is_bulk = np.abs(v - bulk) <= 1e-9 * (1 + np.abs(bulk))
    is_tail = np.abs(v - (bulk * sigma + v0)) <= 1e-9 * (1 + np.abs(bulk))
    assert np.all(is_bulk | is_tail) and abs(is_tail.mean() - alpha) < 5e-3
    assert abs(w.sum() - L) < 2e-4 * L                                   # f_x integrates to L (quasi-Monte-Carlo error)
    # accept-reject: an ordered subsequence of the same proposals, density 1 - eps cos(kappa x), w = L / N
    p.fill(vm._lib.VM_FILL_BUMP_ON_TAIL_SOBOL, [eps, kappa, alpha, sigma, v0, -1.0], 7)
    xa, va, wa = p.download()
    assert np.all(wa == L / n)
    prop = y[:, 0] * L
    order = np.argsort(prop, kind="stable")
    where = np.searchsorted(prop[order], xa)
    where = np.clip(where, 0, nprop - 1)
    cand = order[where]
    assert np.max(np.abs(prop[cand] - xa)) <= 1e-12 * L                  # every particle is one of the proposals ...
    assert np.all(np.diff(cand) > 0)                                     # ... taken in sequence order
    assert abs(n / (cand[-1] + 1) - 1 / (1 + eps)) < 5e-3                # acceptance rate 1 / (1 + eps)
    assert abs(np.mean(np.cos(kappa * xa)) - (-eps / 2)) < 2e-3
    assert abs((va > 3.0).mean() - alpha) < 5e-3
    # sharding independence: two half shards == one full load
    q = vm.DeviceParticles(ctx, n // 2)
    q.fill(vm._lib.VM_FILL_BUMP_ON_TAIL_SOBOL, [eps, kappa, alpha, sigma, v0, -1.0], 7, 0, n)
    x1, v1, _ = q.download()
    q2 = vm.DeviceParticles(ctx, n - n // 2)
    q2.fill(vm._lib.VM_FILL_BUMP_ON_TAIL_SOBOL, [eps, kappa, alpha, sigma, v0, -1.0], 7, n // 2, n)
    x2, v2, _ = q2.download()
    assert np.array_equal(np.concatenate([x1, x2]), xa) and np.array_equal(np.concatenate([v1, v2]), va)
    # the mirror: initialize!(dist, BumpOnTail()) defaults to AcceptRejectSampling (bumpontail.jl:33-35)
    vm.set_default_context(ctx)
    dist = vm.initialize_(vm.ParticleDistribution(1, 1, 20_000), vm.BumpOnTail(), seed=7)
    d2 = vm.initialize_(vm.ParticleDistribution(1, 1, 20_000), vm.BumpOnTail(), vm.ImportanceSampling(), seed=7)
    assert np.all(dist.particles.w == L / 20_000) and d2.particles.w.std() > 0
    vm.set_default_context(None)
    p.close(); q.close(); q2.close()

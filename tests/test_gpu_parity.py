"""GPU parity tests: CUDA path (through the C ABI) vs the CPU oracle on the same seeded inputs.

Tolerance (north star): relative <= 1e-12 on deposited moments and fields, measured against the
max-norm of the reference vector; trajectories after several steps <= 1e-11 absolute (O(1) values).
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


def make_particles(rng, npart, a, b, spread=3.0):
    L = b - a
    x = rng.uniform(a - spread * L, b + spread * L, npart)
    v = rng.standard_normal(npart)
    w = rng.uniform(0.5, 1.5, npart) * L / npart
    return x, v, w


# ------------------------------------------------------------------ deposit --
@pytest.mark.parametrize("k", [2, 3, 4, 5, 6])
@pytest.mark.parametrize("n", [16, 33, 64, 200, 1024])
def test_deposit_matches_oracle(vm, oracle, ctx, rng, k, n):
    a, b = -1.0, 2 * math.pi / 0.3 - 1.0
    npart = 50001
    x, v, w = make_particles(rng, npart, a, b)
    shift = oracle.bspline_shift_bsplinekit(k)
    ref = oracle.deposit_periodic(x, w, a, b, n, k, shift)
    fld = vm.DeviceField(ctx, a, b, k, n, shift)
    p = vm.DeviceParticles(ctx, npart)
    p.upload(x, v, w)
    for mode in (0, 1):
        fld.deposit(p, mode)
        got = fld.rhs
        assert relmax(got, ref) <= RTOL, (k, n, mode)
        assert abs(got.sum() - w.sum()) <= 1e-13 * w.sum()


@pytest.mark.parametrize("k,n", [(4, 4), (4, 5), (3, 3), (5, 6), (2, 2), (6, 7)])
def test_tiny_grids_where_the_stencil_wraps(vm, oracle, ctx, rng, k, n):
    """n_basis barely above the spline order: a particle's K basis functions wrap around the period and the
    circulant stiffness stencil aliases onto itself."""
    a, b = 0.0, 1.0
    npart = 5001
    x, v, w = make_particles(rng, npart, a, b, spread=2.0)
    fld = vm.DeviceField(ctx, a, b, k, n, 0)
    p = vm.DeviceParticles(ctx, npart)
    p.upload(x, v, w)
    fld.deposit(p, 0); fld.solve()
    S = oracle.periodic_stiffness(a, b, n, k, 0)
    rhs = oracle.deposit_periodic(x, w, a, b, n, k, 0)
    phi = oracle.poisson_solve(S, rhs)
    assert relmax(fld.rhs, rhs) <= RTOL
    assert relmax(fld.stiffness_matrix(), S) <= 1e-12
    assert relmax(fld.coefficients, phi) <= 1e-11
    assert relmax(fld.gather_E(p, 1.0), -oracle.eval_dphi(x, a, b, n, k, 0, phi)) <= 1e-11
    assert abs(fld.energy() - oracle.field_energy(S, phi)) <= 1e-10 * abs(oracle.field_energy(S, phi))
    xo, vo = x.copy(), v.copy()
    dref = oracle.integrate_vp(xo, vo, w, 0.05, 1.0, 4, 2, a, b, n, k, 0, S)
    diag = fld.run(p, 0.05, 4, 2, 0, 1.0)
    xg, vg, _ = p.download(w=False)
    assert np.max(np.abs(xg - xo)) <= 1e-10 and np.max(np.abs(vg - vo)) <= 1e-10
    assert np.allclose(diag[:, :3], dref, rtol=1e-9, atol=1e-12)


def test_deposit_deterministic_bitwise(vm, ctx, rng):
    a, b, n, k = 0.0, 1.0, 128, 4
    npart = 400003
    x, v, w = make_particles(rng, npart, a, b)
    fld = vm.DeviceField(ctx, a, b, k, n, 0)
    p = vm.DeviceParticles(ctx, npart)
    p.upload(x, v, w)
    runs = []
    for _ in range(4):
        fld.deposit(p, 0)
        runs.append(fld.rhs.tobytes())
    assert all(r == runs[0] for r in runs)
    # lane-private variant (n small)
    fld2 = vm.DeviceField(ctx, a, b, k, 16, 0)
    runs = []
    for _ in range(4):
        fld2.deposit(p, 0)
        runs.append(fld2.rhs.tobytes())
    assert all(r == runs[0] for r in runs)


@pytest.mark.parametrize("replicas", [1, 2, 4, 8, 16, 32])
def test_deposit_all_replica_variants(vm, oracle, rng, replicas):
    c = vm.Context(0)
    c.set_tuning("replicas", replicas)
    a, b, n, k = 0.0, 3.0, 24, 4
    x, v, w = make_particles(rng, 30000, a, b)
    ref = oracle.deposit_periodic(x, w, a, b, n, k, 0)
    fld = vm.DeviceField(c, a, b, k, n, 0)
    p = vm.DeviceParticles(c, x.size)
    p.upload(x, v, w)
    fld.deposit(p, 0)
    assert relmax(fld.rhs, ref) <= RTOL
    fld.close(); p.close(); c.close()

@pytest.mark.parametrize("n,k", [(24, 4), (32, 4), (64, 4), (100, 4), (128, 4), (64, 3), (40, 5), (128, 6), (64, 2), (200, 4)])
@pytest.mark.parametrize("tune", [{}, {"pairs": 1, "no_repg": 1, "priv_min_warps": 12}, {"pairs": 4}, {"pairs": 8, "no_repg": 1}],
                         ids=["auto", "round1", "pairs4", "pairs8-plain-table"])
def test_pipeline_depths_match_oracle(vm, oracle, rng, n, k, tune):
    """Mid-size meshes run the lane-private passes with few warps: every software-pipeline depth (`pairs`), the
    two-phase batch form of the deep tiers and the 16-fold conflict-free gather table (`no_repg` switches it
    off) against the oracle: deposit, uniform and per-particle weights, fused steps.  Enough particles for
    several loop iterations per thread, odd count."""
    c = vm.Context(0)
    for key, val in tune.items():
        c.set_tuning(key, val)
    kappa = 0.3
    a, b = 0.0, 2 * math.pi / kappa
    npart = 700_001
    x = rng.uniform(a, b, npart); v = rng.standard_normal(npart)
    fld = vm.DeviceField(c, a, b, k, n, 0)
    p = vm.DeviceParticles(c, npart)
    S = oracle.periodic_stiffness(a, b, n, k, 0)
    for w in (np.full(npart, (b - a) / npart), rng.uniform(0.5, 1.5, npart) * (b - a) / npart):
        p.upload(x, v, w)
        fld.deposit(p, 0)
        assert relmax(fld.rhs, oracle.deposit_periodic(x, w, a, b, n, k, 0)) <= RTOL
        xo, vo = x.copy(), v.copy()
        dref = oracle.integrate_vp(xo, vo, w, 0.1, 1.0, 3, 3, a, b, n, k, 0, S)
        diag = fld.run(p, 0.1, 3, 3, 0, 1.0)
        xg, vg, _ = p.download(w=False)
        assert np.max(np.abs(xg - xo)) <= 1e-11 and np.max(np.abs(vg - vo)) <= 1e-11
        assert np.allclose(diag[:, :3], dref, rtol=1e-10, atol=1e-13)
    fld.close(); p.close(); c.close()


def test_deposit_edge_cases(vm, oracle, ctx, rng):
    a, b, n, k = 0.0, 1.0, 16, 3
    fld = vm.DeviceField(ctx, a, b, k, n, 0)
    # empty
    p0 = vm.DeviceParticles(ctx, 0)
    fld.deposit(p0, 0)
    assert np.all(fld.rhs == 0)
    # single particle, exactly on knots and domain ends
    for xv in [0.0, 1.0, 0.5, 1.0 - 1e-17, -1e-17, 0.0625, 7.0, -7.0]:
        p1 = vm.DeviceParticles(ctx, 1)
        p1.upload(np.array([xv]), np.array([0.0]), np.array([2.5]))
        fld.deposit(p1, 0)
        ref = oracle.deposit_periodic(np.array([xv]), np.array([2.5]), a, b, n, k, 0)
        assert np.max(np.abs(fld.rhs - ref)) <= 1e-14, xv
    # all particles in one cell (maximum collision pressure)
    x = rng.uniform(0.5, 0.5 + 1.0 / n, 100000)
    w = np.full(x.size, 1.0 / x.size)
    p = vm.DeviceParticles(ctx, x.size)
    p.upload(x, np.zeros_like(x), w)
    for nn in (16, 256):
        f2 = vm.DeviceField(ctx, a, b, 4, nn, 0)
        for mode in (0, 1):
            f2.deposit(p, mode)
            assert relmax(f2.rhs, oracle.deposit_periodic(x, w, a, b, nn, 4, 0)) <= RTOL


# ------------------------------------------------------- solve / gather / W --
@pytest.mark.parametrize("k,n", [(3, 16), (4, 16), (4, 100), (5, 64), (4, 512)])
def test_solve_gather_energy(vm, oracle, ctx, rng, k, n):
    a, b = 0.0, 2 * math.pi / 0.5
    npart = 40000
    L = b - a
    u = rng.uniform(0, 1, npart)
    x = u * L + 0.05 / 0.5 * np.sin(0.5 * u * L)       # mildly perturbed
    v = rng.standard_normal(npart)
    w = np.full(npart, L / npart)
    fld = vm.DeviceField(ctx, a, b, k, n, 0)
    p = vm.DeviceParticles(ctx, npart)
    p.upload(x, v, w)
    fld.deposit(p, 0)
    fld.solve()
    S = oracle.periodic_stiffness(a, b, n, k, 0)
    M = oracle.periodic_mass(a, b, n, k, 0)
    assert relmax(fld.stiffness_matrix(), S) <= 1e-13
    assert relmax(fld.mass_matrix(), M) <= 1e-13
    rhs = oracle.deposit_periodic(x, w, a, b, n, k, 0)
    phi = oracle.poisson_solve(S, rhs)
    assert relmax(fld.coefficients, phi) <= RTOL
    assert abs(fld.coefficients.sum()) <= 1e-12 * np.abs(phi).sum()
    dphi = oracle.eval_dphi(x, a, b, n, k, 0, phi)
    E = fld.gather_E(p, 1.0)
    assert relmax(E, -dphi) <= RTOL
    E2 = fld.gather_E(p, 0.25)
    assert relmax(E2, -0.25 * dphi) <= RTOL
    xs = rng.uniform(a - L, b + L, 333)
    assert relmax(fld.eval(xs, 1), oracle.eval_dphi(xs, a, b, n, k, 0, phi)) <= RTOL
    W = fld.energy()
    Wref = oracle.field_energy(S, phi)
    assert abs(W - Wref) <= 1e-11 * abs(Wref)


def test_gather_is_pure_function_of_phi(vm, ctx, rng):
    """Restated test/electric_field_tests.jl:37,46: PoissonField == ExternalField fed the same phi, bitwise."""
    nh, pdeg, L = 16, 3, 2 * math.pi
    x = rng.uniform(0, L, 100); w = np.full(100, L / 100)
    poisson = vm.PoissonSolverPBSplines(pdeg, nh, L, ctx=ctx)
    pf = vm.PoissonField(poisson)
    ep = pf(x, w, 0.0)
    phi = np.column_stack([poisson.ϕ, rng.uniform(size=(nh, 10))])
    e2 = vm.ExternalField(poisson, phi, 0.1)
    ef2 = e2(x, w, 0.0)
    assert np.array_equal(ep, ef2)
    s1 = vm.ScaledField(pf, 1.0)
    assert np.array_equal(ep, s1(x, w, 0.0))
    s2 = vm.ScaledField(e2, 1.0)
    assert np.array_equal(ep, s2(x, w, 0.0))
    assert np.any(ep != 0)


# -------------------------------------------------------------- time steps ---
@pytest.mark.parametrize("k,n,chi", [(4, 16, 1.0), (3, 16, 1.0), (5, 40, 1.0), (4, 128, 0.7)])
def test_vp_run_matches_legacy_loop(vm, oracle, ctx, rng, k, n, chi):
    kappa = 0.3
    a, b = 0.0, 2 * math.pi / kappa
    npart = 30001
    x = rng.uniform(a, b, npart); v = rng.standard_normal(npart); w = np.full(npart, (b - a) / npart)
    S = oracle.periodic_stiffness(a, b, n, k, 0)
    nt, nsave, dt = 12, 3, 0.1
    xo, vo = x.copy(), v.copy()
    dref, phiref = oracle.integrate_vp(xo, vo, w, dt, chi, nt, nsave, a, b, n, k, 0, S, want_phi=True)
    fld = vm.DeviceField(ctx, a, b, k, n, 0)
    p = vm.DeviceParticles(ctx, npart)
    for flags in (0, 8, 4):      # fused, unfused, fused with atomic deposit
        p.upload(x, v, w)
        diag = fld.run(p, dt, nt, nsave, flags, chi)
        xg, vg, _ = p.download(w=False)
        assert np.max(np.abs(xg - xo)) <= 1e-11, flags
        assert np.max(np.abs(vg - vo)) <= 1e-11, flags
        assert diag.shape == (nt // nsave + 1, 4)
        assert np.allclose(diag[:, :3], dref, rtol=1e-10, atol=1e-13), flags
        assert np.allclose(diag[:, 3], w.sum(), rtol=1e-13)
        assert relmax(fld.coefficients, phiref[-1]) <= 1e-10


def test_vp_run_matches_strang_new_api(vm, oracle, ctx, rng):
    """New-API Strang A(dt/2) B(dt/2) B(dt/2) A(dt/2), self-consistent and frozen-field."""
    a, b, n, k = 0.0, 1.0, 16, 3
    npart = 10000
    z = rng.standard_normal(npart)
    X = math.ceil(np.max(np.abs(z)))
    x = (z + X) / (2 * X); v = rng.standard_normal(npart); w = np.full(npart, 1.0 / npart)
    shift = oracle.bspline_shift_bsplinekit(k)
    S = oracle.periodic_stiffness(a, b, n, k, shift)
    fld = vm.DeviceField(ctx, a, b, k, n, shift)
    p = vm.DeviceParticles(ctx, npart)
    # self-consistent
    xo, vo = x.copy(), v.copy()
    for _ in range(7):
        oracle.vp_strang_step(xo, vo, w, 0.1, a, b, n, k, shift, S)
    p.upload(x, v, w)
    fld.run(p, 0.1, 7, 0, 1, 1.0)
    xg, vg, _ = p.download(w=False)
    assert np.max(np.abs(xg - xo)) <= 1e-11 and np.max(np.abs(vg - vo)) <= 1e-11
    # frozen at the initial particles
    xo, vo = x.copy(), v.copy()
    for _ in range(7):
        oracle.vp_strang_step(xo, vo, w, 0.1, a, b, n, k, shift, S, x_src=x)
    p.upload(x, v, w)
    fld.deposit(p, 0); fld.solve()
    fld.run(p, 0.1, 7, 0, 1 | 2, 1.0)
    xg, vg, _ = p.download(w=False)
    assert np.max(np.abs(xg - xo)) <= 1e-11 and np.max(np.abs(vg - vo)) <= 1e-11


def test_mirror_splitting_method(vm, oracle, ctx, rng, tmp_path):
    """scripts/vlasov_poisson.jl through the mirrored API (default config: N=1e4, 16 knots, order 3)."""
    vm.set_default_context(ctx)
    npart, nknot, order, tstep = 10000, 16, 3, 0.1
    dist = vm.initialize_(vm.ParticleDistribution(1, 1, npart), vm.NormalDistribution(), seed=3)
    x0 = dist.particles.x[0].copy(); v0 = dist.particles.v[0].copy(); w0 = dist.particles.w[0].copy()
    assert x0.min() >= 0 and x0.max() <= 1 and abs(w0.sum() - 1) < 1e-12 and abs(v0.std() - 1) < 0.05
    potential = vm.Potential(vm.PeriodicBasisBSplineKit((0.0, 1.0), order, nknot))
    model = vm.VlasovPoisson(dist, potential)
    integ = vm.SplittingMethod(model, (0.0, 1.0), tstep)
    out = tmp_path / "vp.npz"
    vm.run_(integ, str(out), save_every=5)
    z = np.load(out)["z"]
    assert z.shape == (2, npart, 3)
    shift = potential.basis.index_shift
    S = oracle.periodic_stiffness(0.0, 1.0, nknot, order, shift)
    xo, vo = x0.copy(), v0.copy()
    for _ in range(10):
        oracle.vp_strang_step(xo, vo, w0, tstep, 0.0, 1.0, nknot, order, shift, S)
    assert np.max(np.abs(dist.particles.x[0] - xo)) <= 1e-11
    assert np.max(np.abs(dist.particles.v[0] - vo)) <= 1e-11
    assert np.array_equal(z[0, :, -1], dist.particles.x[0])
    vm.set_default_context(None)


def test_uniform_weight_fast_path_is_bitwise_identical(vm, rng):
    """When every particle carries the same weight the library passes w0 as a kernel parameter instead of
    streaming w: same arithmetic, so deposits, trajectories and LB right-hand sides must agree bit for bit
    with the general path (tuning no_uniform_w = 1)."""
    a, b, n, k = 0.0, 2 * math.pi / 0.3, 16, 4
    npart = 150001
    x = rng.uniform(a, b, npart); v = rng.standard_normal(npart); w = np.full(npart, (b - a) / npart)
    res = []
    for general in (0, 1):
        c = vm.Context(0)
        c.set_tuning("no_uniform_w", general)
        fld = vm.DeviceField(c, a, b, k, n, 0)
        p = vm.DeviceParticles(c, npart)
        p.upload(x, v, w)
        fld.deposit(p, 0)
        rhs = fld.rhs.tobytes()
        d = fld.run(p, 0.1, 6, 3, 0, 1.0)
        xs, vs_, ws = p.download()
        vsp = vm.DeviceVSpline(c, -10.0, 10.0, 41, 4, 1)
        p.upload(v=v, w=np.full(npart, 1.0 / npart))
        vdot = vsp.lb_rhs(p, 1.0, True)
        vsp.rk438_run(p, 1e-2, 2, 1.0, True, 0)
        vend = p.download(x=False, w=False)[1]
        res.append((rhs, d.tobytes(), xs.tobytes(), vs_.tobytes(), vdot.tobytes(), vend.tobytes()))
        assert np.array_equal(ws, w)
        # non-uniform weights after a uniform phase must be picked up again
        w2 = w.copy(); w2[npart // 2] *= 2.0
        p.upload(x, v, w2)
        fld.deposit(p, 0)
        assert abs(fld.rhs.sum() - w2.sum()) <= 1e-13 * w2.sum()
        vsp.close(); p.close(); fld.close(); c.close()
    assert res[0] == res[1]


def test_async_snapshot(vm, ctx, rng):
    """Snapshot taken asynchronously while stepping continues == state at the time of the call."""
    npart = 300001
    a, b = 0.0, 2 * math.pi / 0.3
    x = rng.uniform(a, b, npart); v = rng.standard_normal(npart); w = np.full(npart, (b - a) / npart)
    fld = vm.DeviceField(ctx, a, b, 4, 16, 0)
    p = vm.DeviceParticles(ctx, npart)
    p.upload(x, v, w)
    fld.run(p, 0.1, 3, 0, 0, 1.0)
    x3, v3, _ = p.download(w=False)
    bx, bv = vm.PinnedArray(npart), vm.PinnedArray(npart)
    p.snapshot_begin(bx.array, bv.array)
    fld.run(p, 0.1, 4, 0, 0, 1.0)          # keeps the GPU busy while the copy drains
    p.snapshot_wait()
    assert np.array_equal(bx.array, x3) and np.array_equal(bv.array, v3)
    x7 = p.download(w=False)[0]
    assert not np.array_equal(x7, x3)
    p.snapshot_begin(None, bv.array)       # v only, twice in a row (staging buffer reuse)
    p.snapshot_begin(bx.array, None)
    p.snapshot_wait()
    assert np.array_equal(bx.array, x7)
    vm.DeviceParticles(ctx, 0).snapshot_begin(None, None)


def test_particles_aos_roundtrip_and_kick_drift(vm, oracle, ctx, rng):
    npart = 12347
    z = rng.standard_normal((npart, 3))
    p = vm.DeviceParticles(ctx, npart)
    p.upload_aos(z)
    x, v, w = p.download()
    assert np.array_equal(x, z[:, 0]) and np.array_equal(v, z[:, 1]) and np.array_equal(w, z[:, 2])
    assert np.array_equal(p.download_aos(), z)
    p.drift(0.25)
    assert np.array_equal(p.download()[0], z[:, 0] + 0.25 * z[:, 1])
    fld = vm.DeviceField(ctx, 0.0, 1.0, 4, 16, 0)
    phi = rng.standard_normal(16); phi -= phi.mean()
    fld.coefficients = phi
    xcur = p.download()[0]
    fld.kick(p, 0.1, -1.0)
    vref = z[:, 1] - 0.1 * oracle.eval_dphi(xcur, 0.0, 1.0, 16, 4, 0, phi)
    assert np.max(np.abs(p.download()[1] - vref)) <= 1e-12 * np.max(np.abs(vref))


# ---------------------------------------------------------------- v-space ----
@pytest.mark.parametrize("nknots,k", [(41, 4), (41, 3), (17, 5), (129, 4), (9, 2), (300, 6)])
def test_vspline_projection_and_rhs(vm, oracle, ctx, rng, nknots, k):
    a, b = -10.0, 10.0
    npart = 60001
    v = np.concatenate([rng.standard_normal(npart - 7) * 2.0 + 0.3, [-9.99, 9.99, 10.0, -10.0, 10.5, -11.0, 0.0]])
    w = rng.uniform(0.5, 1.5, npart) / npart
    M = oracle.dirichlet_mass(a, b, nknots, k)
    vs = vm.DeviceVSpline(ctx, a, b, nknots, k, 1)
    assert vs.nv == nknots + k - 4
    assert relmax(vs.mass_matrix(), M) <= 1e-13
    p = vm.DeviceParticles(ctx, npart)
    p.upload(np.zeros(npart), v, w)
    vs.project(p)
    coef, rhs = oracle.vproject(v, w, a, b, nknots, k, M)
    assert relmax(vs.rhs, rhs) <= RTOL
    # round 2: asserted at the north star's 1e-12 (measured 1e-16 .. 2e-14, tests/test_gpu_round2.py prints them);
    # the 300-knot order-6 basis has a mass matrix of condition ~1e3 and gets 1e-11
    tol = 1e-11 if nknots >= 300 else 1e-12
    assert relmax(vs.coefficients, coef) <= tol
    pts = np.concatenate([rng.uniform(a, b, 500), [a, b, a - 1, b + 1]])
    f, df = vs.eval(pts)
    fr, dfr = oracle.vspline_eval(pts, a, b, nknots, k, coef)
    assert relmax(f, fr) <= tol and relmax(df, dfr) <= tol
    m5, A = vs.moments(p)
    m5r = oracle.vmoments(v, a, b, nknots, k, coef)
    assert np.max(np.abs(m5 - m5r)) <= tol * np.max(np.abs(m5r))
    for cons in (False, True):
        vdot = vs.lb_rhs(p, 1.3, cons)
        ref, _, Aref = oracle.lb_rhs(v, w, a, b, nknots, k, M, 1.3, cons)
        assert relmax(vdot, ref) <= tol, (cons,)
        if cons and k >= 3:
            assert np.allclose(A, Aref, rtol=10 * tol)


def test_clb_rk438_matches_oracle_and_conserves(vm, oracle, ctx, rng):
    """scripts/lenard_bernstein_conservative.jl: DoubleMaxwellian(+-2), 41 knots, order 4, RK438."""
    a, b, nknots, k = -10.0, 10.0, 41, 4
    npart, dt, nsteps = 20000, 1e-2, 5
    v = np.concatenate([rng.standard_normal(npart // 2) + 2.0, rng.standard_normal(npart - npart // 2) - 2.0])
    w = np.full(npart, 1.0 / npart)
    M = oracle.dirichlet_mass(a, b, nknots, k)
    vs = vm.DeviceVSpline(ctx, a, b, nknots, k, 1)
    p = vm.DeviceParticles(ctx, npart)
    for cons in (True, False):
        p.upload(np.zeros(npart), v, w)
        diag = vs.rk438_run(p, dt, nsteps, 1.0, cons, 1)
        vg = p.download(x=False, w=False)[1]
        vo = v.copy()
        for _ in range(nsteps):
            oracle.lb_rk438_step(vo, w, dt, a, b, nknots, k, M, 1.0, cons)
        assert np.max(np.abs(vg - vo)) <= 1e-11, cons
        assert diag.shape == (nsteps + 1, 4)
        assert np.allclose(diag[:, 0], dt * np.arange(nsteps + 1))
        assert abs(diag[-1, 1] - vo.sum()) <= 1e-9 * npart and abs(diag[-1, 2] - (vo ** 2).sum()) <= 1e-9 * npart
        if cons:     # momentum and energy drift as printed by the script (:64)
            assert abs(diag[-1, 1] - diag[0, 1]) <= 1e-6 * npart
            assert abs(diag[-1, 2] - diag[0, 2]) / diag[0, 2] <= 1e-6


def test_rk438_fused_equals_unfused(vm, rng):
    """The fused stage pass (RHS + stage assembly + next deposit) must reproduce the kernel-per-operation
    driver: same per-particle arithmetic; only the summation order of the deposit differs (the two
    kernels walk the particles with different strides), so agreement is to rounding, and each driver
    is bit-reproducible on its own."""
    a, b, nknots, k = -10.0, 10.0, 41, 4
    npart = 500001           # several iterations per thread: the software-pipelined stage loop gets exercised
    v = np.concatenate([rng.standard_normal(npart // 2) + 2.0, rng.standard_normal(npart - npart // 2) - 2.0])
    w = np.full(npart, 1.0 / npart)
    res = {}
    for no_fuse in (0, 1):
        c = vm.Context(0)
        c.set_tuning("no_fuse", no_fuse)
        vs = vm.DeviceVSpline(c, a, b, nknots, k, 1)
        p = vm.DeviceParticles(c, npart)
        for cons in (True, False):
            p.upload(np.zeros(npart), v, w)
            vs.rk438_run(p, 1e-2, 3, 1.0, cons, 0)
            res[(no_fuse, cons)] = p.download(x=False, w=False)[1].copy()
        vs.close(); p.close(); c.close()
    for cons in (True, False):
        assert np.max(np.abs(res[(0, cons)] - res[(1, cons)])) <= 1e-12, cons
    c = vm.Context(0)
    vs = vm.DeviceVSpline(c, a, b, nknots, k, 1)
    p = vm.DeviceParticles(c, npart)
    p.upload(np.zeros(npart), v, w)
    vs.rk438_run(p, 1e-2, 3, 1.0, True, 0)
    assert np.array_equal(p.download(x=False, w=False)[1], res[(0, True)])      # run-to-run bitwise


def test_vp_run_bitwise_reproducible(vm, ctx, rng):
    """Deterministic deposit => the whole fused time loop is bit-reproducible run to run."""
    a, b, n, k = 0.0, 2 * math.pi / 0.3, 16, 4
    npart = 200001
    x = rng.uniform(a, b, npart); v = rng.standard_normal(npart); w = np.full(npart, (b - a) / npart)
    fld = vm.DeviceField(ctx, a, b, k, n, 0)
    p = vm.DeviceParticles(ctx, npart)
    outs = []
    for _ in range(3):
        p.upload(x, v, w)
        d = fld.run(p, 0.1, 10, 5, 0, 1.0)
        xs, vs_, _ = p.download(w=False)
        outs.append((xs.tobytes(), vs_.tobytes(), d.tobytes(), fld.coefficients.tobytes()))
    assert outs[0] == outs[1] == outs[2]


def test_mirror_lenard_bernstein(vm, oracle, ctx, rng, tmp_path):
    vm.set_default_context(ctx)
    npart = 1000
    dist = vm.initialize_(vm.ParticleDistribution(1, 1, npart), vm.DoubleMaxwellian((-10.0, 10.0), 2.0), seed=11)
    v0 = dist.particles.v[0].copy(); w0 = dist.particles.w[0].copy()
    assert abs(v0[: npart // 2].mean() - 2) < 0.2 and abs(v0[npart // 2:].mean() + 2) < 0.2
    sdist = vm.SplineDistribution(1, 1, 41, 4, (-10.0, 10.0), "Dirichlet")
    model = vm.ConservativeLenardBernstein(dist, vm.CollisionEntropy(sdist))
    integ = vm.GeometricIntegrator(model, (0.0, 0.05), 1e-2)
    vm.run_(integ, str(tmp_path / "lb.npz"), diag_every=1, save_every=2)
    zz = np.load(tmp_path / "lb.npz")
    assert zz["z"].shape == (npart, 4) and np.array_equal(zz["z"][:, 0], v0) and np.allclose(zz["t"], [0, 0.02, 0.04, 0.05])
    assert np.array_equal(zz["z"][:, -1], dist.particles.v[0])
    M = oracle.dirichlet_mass(-10.0, 10.0, 41, 4)
    vo = v0.copy()
    for _ in range(5):
        oracle.lb_rk438_step(vo, w0, 1e-2, -10.0, 10.0, 41, 4, M, 1.0, True)
    assert np.max(np.abs(dist.particles.v[0] - vo)) <= 1e-11
    assert integ.diagnostics.shape == (6, 4)
    fs = vm.projection(None, dist, sdist)
    assert abs(fs(0.0) - oracle.vspline_eval(np.array([0.0]), -10.0, 10.0, 41, 4, sdist.coefficients)[0][0]) < 1e-12
    vm.set_default_context(None)


# ----------------------------------------------------------------- loads -----
def test_device_fills(vm, ctx):
    n = 400000
    p = vm.DeviceParticles(ctx, n)
    p.fill(1, [0.03, 0.3, 0.1, 0.5, 4.5], 123)          # bump on tail
    x, v, w = p.download()
    L = 2 * math.pi / 0.3
    assert x.min() >= 0 and x.max() < L and abs(w.sum() - L) < 1e-9
    assert abs(np.mean(np.cos(0.3 * x)) - (-0.03 / 2)) < 5e-3            # density 1 - eps cos(kx)
    assert abs((v > 3.0).mean() - 0.1) < 5e-3
    p.fill(6, [0.05, 0.5], 5)                            # Landau
    x, v, w = p.download()
    assert abs(np.mean(np.cos(0.5 * x)) - 0.05 / 2) < 5e-3 and abs(v.std() - 1) < 5e-3
    # sharding independence: two half shards == one full load
    q = vm.DeviceParticles(ctx, n // 2)
    q.fill(6, [0.05, 0.5], 5, 0, n)
    xa = q.download()[0]
    q.fill(6, [0.05, 0.5], 5, n // 2, n)
    xb = q.download()[0]
    assert np.array_equal(np.concatenate([xa, xb]), x)
    p.fill(2, [-5.0, 5.0, 3.0], 9)
    x, v, w = p.download()
    assert abs(v[: n // 2].mean() - 3) < 1e-2 and abs(v[n // 2:].mean() + 3) < 1e-2 and abs(w.sum() - 1) < 1e-9


def test_error_reporting(vm, ctx):
    with pytest.raises(vm.VMError):
        vm.DeviceField(ctx, 0.0, 1.0, 9, 16, 0)
    with pytest.raises(vm.VMError):
        vm.DeviceField(ctx, 1.0, 0.0, 4, 16, 0)
    with pytest.raises(vm.VMError):
        vm.DeviceField(ctx, 0.0, 1.0, 4, 2, 0)
    with pytest.raises(vm.VMError):
        vm.DeviceVSpline(ctx, 0.0, 1.0, 1, 4, 1)
    with pytest.raises(vm.VMError):
        vm.DeviceParticles(ctx, 4).fill(1, [0.1], 0)
    c2 = vm.Context(0)
    f = vm.DeviceField(c2, 0.0, 1.0, 4, 16, 0)
    with pytest.raises(vm.VMError):
        f.deposit(vm.DeviceParticles(ctx, 4), 0)       # different contexts


def test_interpreter_exit_without_close(tmp_path):
    """A script that never calls close() must exit cleanly: shutdown finalises the wrappers in no particular
    order, and a child handle destroyed after its context used to be a use-after-free (segfault at exit)."""
    import subprocess
    import sys
    from pathlib import Path
    root = Path(__file__).resolve().parent.parent
    script = tmp_path / "noclose.py"
    script.write_text(
        "import sys; sys.path.insert(0, %r)\n"
        "from __graft_entry__ import load_package\n"
        "vm = load_package()\n"
        "ctx = vm.Context(0)\n"
        "p = vm.DeviceParticles(ctx, 1000); f = vm.DeviceField(ctx, 0.0, 1.0, 4, 16, 0)\n"
        "vs = vm.DeviceVSpline(ctx, -10.0, 10.0, 41, 4, 1)\n"
        "p.fill(vm._lib.VM_FILL_UNIFORM, [0.0, 1.0, -1.0, 1.0], 1); f.run(p, 0.1, 2, 0, 0, 1.0)\n"
        "print('done')\n" % str(root))
    r = subprocess.run([sys.executable, str(script)], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0 and "done" in r.stdout, (r.returncode, r.stderr[-2000:])


def test_handles_outliving_their_context(vm, oracle, rng):
    """Julia and Python finalise objects in no particular order: a context may be destroyed before its children.
    The library must then free the orphans without touching the context, refuse to compute with them, accept a
    second destroy of the context, and leave the device usable."""
    import ctypes as C
    L = vm._lib
    c = vm.Context(0)
    p = vm.DeviceParticles(c, 1000)
    f = vm.DeviceField(c, 0.0, 1.0, 4, 16, 0)
    vs = vm.DeviceVSpline(c, -10.0, 10.0, 41, 4, 1)
    h = C.c_void_p(c._h.value)
    assert L.lib().vm_ctx_destroy(h) == 0            # the context goes first
    assert L.lib().vm_ctx_destroy(h) == 0            # twice: no-op
    with pytest.raises(vm.VMError) as ei:
        f.deposit(p, 0)
    assert ei.value.code == L.VM_ERR_INVALID and "already been destroyed" in str(ei.value)
    with pytest.raises(vm.VMError):
        p.upload(np.zeros(1000), np.zeros(1000), np.ones(1000))
    with pytest.raises(vm.VMError):
        vs.lb_rhs(p, 1.0, False)
    assert L.lib().vm_launch_count(h) == 0
    p.close(); f.close(); vs.close()                 # orphans: freed without the context
    c._h = C.c_void_p()
    c2 = vm.Context(0)                               # the device is still usable
    a, b, n, k = 0.0, 3.0, 16, 4
    x, v, w = make_particles(rng, 5000, a, b)
    f2 = vm.DeviceField(c2, a, b, k, n, 0)
    p2 = vm.DeviceParticles(c2, x.size)
    p2.upload(x, v, w)
    f2.deposit(p2, 0)
    assert relmax(f2.rhs, oracle.deposit_periodic(x, w, a, b, n, k, 0)) <= RTOL
    c2.close()


@pytest.mark.parametrize("n", [64, 128])
def test_split_kick_on_midsize_meshes(vm, oracle, rng, n):
    """The two-half-kick form of the new-API Strang step (VM_RUN_SPLIT_KICK) in the deep-pipeline tiers of the
    fused pass (n_h = 64: 4 pairs in flight, n_h = 128: 8 pairs) against the oracle's Strang step."""
    c = vm.Context(0)
    a, b, k = 0.0, 2 * math.pi / 0.3, 4
    npart = 300_001
    x = rng.uniform(a, b, npart); v = rng.standard_normal(npart); w = np.full(npart, (b - a) / npart)
    shift = oracle.bspline_shift_bsplinekit(k)
    S = oracle.periodic_stiffness(a, b, n, k, shift)
    fld = vm.DeviceField(c, a, b, k, n, shift)
    p = vm.DeviceParticles(c, npart)
    xo, vo = x.copy(), v.copy()
    for _ in range(4):
        oracle.vp_strang_step(xo, vo, w, 0.1, a, b, n, k, shift, S)
    p.upload(x, v, w)
    fld.run(p, 0.1, 4, 0, 1, 1.0)
    xg, vg, _ = p.download(w=False)
    assert np.max(np.abs(xg - xo)) <= 1e-11 and np.max(np.abs(vg - vo)) <= 1e-11
    fld.close(); p.close(); c.close()

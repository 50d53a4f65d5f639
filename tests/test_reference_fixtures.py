"""Reference-pinned parity: outputs of the UNMODIFIED VlasovMethods.jl (written by julia/make_reference_fixtures.jl
on a machine that has Julia -- none exists where this repository is built) against the CPU oracle (CPU test) and the
CUDA path through the C ABI (GPU test), on the inputs of tests/golden/inputs_v1.

While tests/golden/reference_v1/ is absent these tests SKIP with the reason "PARITY UNPINNED": the oracle is then
pinned only by its known-answer tests (tests/test_oracle_kat.py), as DESIGN.md section 4 states.  When the directory
is present the conventions no reference test pins (periodic index rotation, n_basis vs nknot, sign/gauge of phi) are
detected from the reference's own vectors and reported, then everything is compared at the north star's tolerance.
"""
import math
import warnings
from pathlib import Path

import numpy as np
import pytest

HERE = Path(__file__).resolve().parent
IN, REF = HERE / "golden" / "inputs_v1", HERE / "golden" / "reference_v1"
UNPINNED = ("PARITY UNPINNED: tests/golden/reference_v1/ is absent -- run julia/make_reference_fixtures.jl with "
            "Julia + VlasovMethods.jl v0.2.1 and commit its output to pin the oracle to the reference")


def load(d, name):
    return np.fromfile(d / f"{name}.f64", dtype="<f8")


def need_reference():
    if not (REF / "manifest.txt").exists():
        warnings.warn(UNPINNED)
        pytest.skip(UNPINNED)


def relmax(a, b):
    return np.max(np.abs(np.asarray(a) - np.asarray(b))) / max(np.max(np.abs(b)), 1e-300)


def test_inputs_are_the_golden_arrays():
    """The raw files the Julia script reads are bit for bit the arrays of golden_v1.npz."""
    G = np.load(HERE / "golden" / "golden_v1.npz")
    for key in ("vp_x", "vp_v", "vp_w", "st_x", "st_v", "lb_v", "lb_w"):
        assert np.array_equal(load(IN, key), G[key]), key


def detect_conventions(oracle, x, w, a, b, order, rhs_ref):
    """n_basis = length of the reference's rhs; index rotation = the shift whose oracle deposit matches it."""
    n = rhs_ref.size
    best = min(range(-order, order + 1), key=lambda s: relmax(oracle.deposit_periodic(x, w, a, b, n, order, s), rhs_ref))
    return n, best


def align_phi(phi, phi_ref):
    """phi is defined up to the gauge constant and (were the reference to solve the other sign) a sign."""
    sign = 1.0 if np.dot(phi - phi.mean(), phi_ref - phi_ref.mean()) >= 0 else -1.0
    return sign, phi_ref.mean() - sign * phi.mean()


def test_oracle_matches_reference_fixtures(oracle):
    need_reference()
    a, b, k = 0.0, 2 * math.pi / 0.3, 4
    x, v, w = load(IN, "vp_x"), load(IN, "vp_v"), load(IN, "vp_w")
    rhs_ref, phi_ref, dphi_ref = load(REF, "vp_rhs"), load(REF, "vp_phi"), load(REF, "vp_dphi")
    n, shift = detect_conventions(oracle, x, w, a, b, k, rhs_ref)
    print(f"[reference conventions] n_basis = {n} (nknot = 16), index_shift = {shift} "
          f"(oracle default {oracle.bspline_shift_bsplinekit(k)})")
    assert relmax(oracle.deposit_periodic(x, w, a, b, n, k, shift), rhs_ref) <= 1e-12
    S = oracle.periodic_stiffness(a, b, n, k, shift)
    phi = oracle.poisson_solve(S, rhs_ref)
    sign, gauge = align_phi(phi, phi_ref)
    print(f"[reference conventions] phi sign = {sign:+.0f}, gauge offset = {gauge:.3e}")
    assert sign > 0, "the reference solves the opposite sign of the Poisson equation: kick scale must flip"
    assert relmax(phi + gauge, phi_ref) <= 1e-11
    assert relmax(oracle.eval_dphi(x, a, b, n, k, shift, phi), dphi_ref) <= 1e-11
    for nsteps in (1, 8):
        for tag, src in (("selfconsistent", None), ("frozen", x)):
            xo, vo = x.copy(), v.copy()
            for _ in range(nsteps):
                oracle.vp_strang_step(xo, vo, w, 0.1, a, b, n, k, shift, S, x_src=src)
            assert np.max(np.abs(xo - load(REF, f"vp_strang{nsteps}_{tag}_x"))) <= 1e-11, (nsteps, tag)
            assert np.max(np.abs(vo - load(REF, f"vp_strang{nsteps}_{tag}_v"))) <= 1e-11, (nsteps, tag)
    # v-space
    vv, wv = load(IN, "lb_v"), load(IN, "lb_w")
    M = oracle.dirichlet_mass(-10.0, 10.0, 41, 4)
    assert relmax(M, load(REF, "lb_mass_matrix").reshape(M.shape)) <= 1e-13
    coef, _ = oracle.vproject(vv, wv, -10.0, 10.0, 41, 4, M)
    assert relmax(coef, load(REF, "lb_coef")) <= 1e-11
    f, df = oracle.vspline_eval(vv, -10.0, 10.0, 41, 4, coef)
    assert relmax(f, load(REF, "lb_f")) <= 1e-11 and relmax(df, load(REF, "lb_df")) <= 1e-10
    assert np.allclose(oracle.vmoments(vv, -10.0, 10.0, 41, 4, coef), load(REF, "lb_m5"), rtol=1e-10)
    for tag, cons in (("lb", False), ("clb", True)):
        vdot, _, A = oracle.lb_rhs(vv, wv, -10.0, 10.0, 41, 4, M, 1.0, cons)
        assert relmax(vdot, load(REF, f"{tag}_vdot")) <= 1e-10
        vend = vv.copy()
        for _ in range(3):
            oracle.lb_rk438_step(vend, wv, 1e-2, -10.0, 10.0, 41, 4, M, 1.0, cons)
        assert np.max(np.abs(vend - load(REF, f"{tag}_v3"))) <= 1e-11
    assert np.allclose(A, load(REF, "clb_A"), rtol=1e-9)


@pytest.mark.gpu
def test_cuda_matches_reference_fixtures(vm, oracle):
    need_reference()
    a, b, k = 0.0, 2 * math.pi / 0.3, 4
    x, v, w = load(IN, "vp_x"), load(IN, "vp_v"), load(IN, "vp_w")
    rhs_ref, phi_ref, dphi_ref = load(REF, "vp_rhs"), load(REF, "vp_phi"), load(REF, "vp_dphi")
    n, shift = detect_conventions(oracle, x, w, a, b, k, rhs_ref)
    ctx = vm.Context(0)
    fld = vm.DeviceField(ctx, a, b, k, n, shift)
    p = vm.DeviceParticles(ctx, x.size)
    p.upload(x, v, w)
    fld.deposit(p, 0); fld.solve()
    assert relmax(fld.rhs, rhs_ref) <= 1e-12
    phi = fld.coefficients
    sign, gauge = align_phi(phi, phi_ref)
    assert sign > 0 and relmax(phi + gauge, phi_ref) <= 1e-11
    assert relmax(fld.gather_E(p, 1.0), -dphi_ref) <= 1e-11
    for nsteps in (1, 8):
        p.upload(x, v, w)
        fld.run(p, 0.1, nsteps, 0, vm._lib.VM_RUN_SPLIT_KICK, 1.0)
        xg, vg, _ = p.download(w=False)
        assert np.max(np.abs(xg - load(REF, f"vp_strang{nsteps}_selfconsistent_x"))) <= 1e-11
        assert np.max(np.abs(vg - load(REF, f"vp_strang{nsteps}_selfconsistent_v"))) <= 1e-11
        p.upload(x, v, w)
        fld.deposit(p, 0); fld.solve()
        fld.run(p, 0.1, nsteps, 0, vm._lib.VM_RUN_SPLIT_KICK | vm._lib.VM_RUN_FROZEN_FIELD, 1.0)
        xg, vg, _ = p.download(w=False)
        assert np.max(np.abs(xg - load(REF, f"vp_strang{nsteps}_frozen_x"))) <= 1e-11
        assert np.max(np.abs(vg - load(REF, f"vp_strang{nsteps}_frozen_v"))) <= 1e-11
    vv, wv = load(IN, "lb_v"), load(IN, "lb_w")
    vs = vm.DeviceVSpline(ctx, -10.0, 10.0, 41, 4, 1)
    q = vm.DeviceParticles(ctx, vv.size)
    q.upload(np.zeros(vv.size), vv, wv)
    vs.project(q)
    assert relmax(vs.coefficients, load(REF, "lb_coef")) <= 1e-11
    f, df = vs.eval(vv)
    assert relmax(f, load(REF, "lb_f")) <= 1e-11 and relmax(df, load(REF, "lb_df")) <= 1e-10
    for tag, cons in (("lb", False), ("clb", True)):
        assert relmax(vs.lb_rhs(q, 1.0, cons), load(REF, f"{tag}_vdot")) <= 1e-10
        q.upload(v=vv)
        vs.rk438_run(q, 1e-2, 3, 1.0, cons, 0)
        assert np.max(np.abs(q.download(x=False, w=False)[1] - load(REF, f"{tag}_v3"))) <= 1e-11
        q.upload(v=vv)
    ctx.close()

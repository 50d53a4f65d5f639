#!/usr/bin/env python
"""Generate tests/golden/golden_v1.npz from the CPU oracle (oracle/vm_oracle.c).

The reference (Julia) cannot run where this repository is built and ships no golden vectors for the hot
path, so these fixtures do NOT pin the oracle to the reference -- the known-answer tests in
tests/test_oracle_kat.py do that job.  They freeze the oracle's own outputs on small seeded inputs so that
(a) later changes to the oracle are detected, and (b) the CUDA path is compared against numbers that are
committed next to the code.  Regenerate with:  python tests/golden/make_golden.py
"""
import math
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))
from oracle import vm_oracle as orc  # noqa: E402


def build():
    rng = np.random.default_rng(20240601)
    out = {}
    # --- x-space: bump-on-tail-like load, cubic, 16 basis functions (scripts/bump_on_tail.jl geometry)
    n, k, npart, dt, chi = 16, 4, 4001, 0.1, 1.0
    a, b = 0.0, 2 * math.pi / 0.3
    x = rng.uniform(a - 2 * (b - a), b + 2 * (b - a), npart)
    v = rng.standard_normal(npart)
    w = rng.uniform(0.5, 1.5, npart) * (b - a) / npart
    S = orc.periodic_stiffness(a, b, n, k, 0)
    rhs = orc.deposit_periodic(x, w, a, b, n, k, 0)
    phi = orc.poisson_solve(S, rhs)
    out.update(vp_x=x, vp_v=v, vp_w=w, vp_rhs=rhs, vp_phi=phi, vp_dphi=orc.eval_dphi(x, a, b, n, k, 0, phi),
               vp_W=np.array([orc.field_energy(S, phi)]), vp_stiffness_row=S[0].copy(),
               vp_mass_row=orc.periodic_mass(a, b, n, k, 0)[0].copy())
    xo, vo = x.copy(), v.copy()
    diag = orc.integrate_vp(xo, vo, w, dt, chi, 8, 2, a, b, n, k, 0, S)
    out.update(vp_x8=xo, vp_v8=vo, vp_diag=diag)
    # --- new-API default: N(0,1) blob on (0,1), quadratic, 16 functions, BSplineKit index rotation
    k3, sh = 3, orc.bspline_shift_bsplinekit(3)
    z = rng.standard_normal(2000)
    X = math.ceil(np.max(np.abs(z)))
    xs = (z + X) / (2 * X); vs = rng.standard_normal(2000); ws = np.full(2000, 1.0 / 2000)
    S3 = orc.periodic_stiffness(0.0, 1.0, 16, k3, sh)
    xo, vo = xs.copy(), vs.copy()
    for _ in range(5):
        orc.vp_strang_step(xo, vo, ws, 0.1, 0.0, 1.0, 16, k3, sh, S3)
    out.update(st_x=xs, st_v=vs, st_x5=xo, st_v5=vo)
    # --- v-space: 41 knots, order 4, Dirichlet on (-10, 10) (scripts/lenard_bernstein*.jl)
    nk, kv, av, bv = 41, 4, -10.0, 10.0
    vv = np.concatenate([rng.standard_normal(1500) + 2.0, rng.standard_normal(1500) - 2.0, [-9.9, 9.9, 10.0, -10.0, 10.5]])
    wv = rng.uniform(0.5, 1.5, vv.size) / vv.size
    M = orc.dirichlet_mass(av, bv, nk, kv)
    coef, vrhs = orc.vproject(vv, wv, av, bv, nk, kv, M)
    f, df = orc.vspline_eval(vv, av, bv, nk, kv, coef)
    lb, _, _ = orc.lb_rhs(vv, wv, av, bv, nk, kv, M, 1.0, False)
    clb, _, A = orc.lb_rhs(vv, wv, av, bv, nk, kv, M, 1.0, True)
    vend = vv.copy()
    for _ in range(3):
        orc.lb_rk438_step(vend, wv, 1e-2, av, bv, nk, kv, M, 1.0, True)
    out.update(lb_v=vv, lb_w=wv, lb_rhs=vrhs, lb_coef=coef, lb_f=f, lb_df=df, lb_vdot=lb, clb_vdot=clb, clb_A=A,
               lb_m5=orc.vmoments(vv, av, bv, nk, kv, coef), clb_v3=vend, lb_mass_diag=np.diag(M).copy())
    return out


INPUT_KEYS = ("vp_x", "vp_v", "vp_w", "st_x", "st_v", "lb_v", "lb_w")


def write_inputs(out):
    """The seeded INPUT arrays as raw little-endian Float64 files (tests/golden/inputs_v1/<key>.f64): what
    julia/make_reference_fixtures.jl reads to run the unmodified reference on the very same particles (Julia's
    standard library reads them with read!; no .npz reader is needed on that side)."""
    d = Path(__file__).with_name("inputs_v1")
    d.mkdir(exist_ok=True)
    for key in INPUT_KEYS:
        np.ascontiguousarray(out[key], dtype="<f8").tofile(d / f"{key}.f64")
    return d


if __name__ == "__main__":
    out = build()
    np.savez_compressed(Path(__file__).with_name("golden_v1.npz"), **out)
    print("wrote", Path(__file__).with_name("golden_v1.npz"))
    print("wrote", write_inputs(out))

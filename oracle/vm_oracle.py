"""ctypes front-end of the CPU oracle (oracle/vm_oracle.c).

TEST INFRASTRUCTURE ONLY -- see oracle/vm_oracle.h.  Only tests/,
__graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this module.  Parity status: "parity unpinned" for the third-party
arithmetic (BSplineKit / PoissonSolvers / GeometricIntegrators are absent and
the reference ships no golden vectors); pinned by the known-answer tests in
tests/test_oracle_kat.py.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from pathlib import Path

import numpy as np

_HERE = Path(__file__).resolve().parent
_LIB = None

_d = C.c_double
_i = C.c_int
_l = C.c_long
_p = C.POINTER(C.c_double)


def build(native_out: str | None = None) -> Path:
    """Compile the oracle.  Default: portable oracle/libvm_oracle.so.
    native_out: path for a -march=native copy (timed CPU baseline)."""
    if native_out is not None:
        subprocess.check_call(["make", "-s", "-C", str(_HERE), "native", f"OUT={native_out}"])
        return Path(native_out)
    subprocess.check_call(["make", "-s", "-C", str(_HERE)])
    return _HERE / "libvm_oracle.so"


def _sig(lib):
    lib.vmo_eval_all.argtypes = [_p, _i, _i, _d, _p]
    lib.vmo_eval_all_deriv.argtypes = [_p, _i, _i, _d, _p]
    lib.vmo_periodic_eval.argtypes = [_d, _d, _i, _i, _d, _p, _p]
    lib.vmo_periodic_eval.restype = _i
    lib.vmo_clamped_eval.argtypes = [_d, _d, _i, _i, _d, _p, _p]
    lib.vmo_clamped_eval.restype = _i
    lib.vmo_periodic_mass.argtypes = [_d, _d, _i, _i, _i, _p]
    lib.vmo_periodic_stiffness.argtypes = [_d, _d, _i, _i, _i, _p]
    lib.vmo_dirichlet_mass.argtypes = [_d, _d, _i, _i, _p]
    lib.vmo_deposit_periodic.argtypes = [_p, _p, _l, _d, _d, _i, _i, _i, _p]
    lib.vmo_poisson_solve.argtypes = [_p, _i, _p, _p]
    lib.vmo_eval_dphi.argtypes = [_p, _l, _d, _d, _i, _i, _i, _p, _p]
    lib.vmo_field_energy.argtypes = [_p, _i, _p]
    lib.vmo_field_energy.restype = _d
    lib.vmo_s_advection.argtypes = [_p, _p, _l, _d]
    lib.vmo_s_acceleration.argtypes = [_p, _p, _p, _l, _d, _d, _d, _i, _i, _i, _p, _p]
    lib.vmo_vp_strang_step.argtypes = [_p, _p, _p, _l, _d, _d, _d, _i, _i, _i, _p, _p]
    lib.vmo_integrate_vp.argtypes = [_p, _p, _p, _l, _d, _d, _i, _i, _d, _d, _i, _i, _i, _p, _p, _p]
    lib.vmo_integrate_vp_external.argtypes = [_p, _p, _p, _l, _d, _d, _i, _i, _d, _d, _i, _i, _i, _p, _p, _i, _d, _p]
    lib.vmo_lorentz_force.argtypes = [_p, _p, _p, _l, _d, _d, _i, _i, _i, _p, _p, _p, _p]
    lib.vmo_vp_rk4_step.argtypes = [_p, _p, _p, _l, _d, _d, _d, _i, _i, _i, _p]
    lib.vmo_vproject.argtypes = [_p, _p, _l, _d, _d, _i, _i, _p, _p, _p]
    lib.vmo_vspline_eval.argtypes = [_p, _l, _d, _d, _i, _i, _p, _p, _p]
    lib.vmo_vmoments.argtypes = [_p, _l, _d, _d, _i, _i, _p, _p]
    lib.vmo_clb_coefficients.argtypes = [_p, _p, _p]
    lib.vmo_lb_rhs.argtypes = [_p, _p, _l, _d, _d, _i, _i, _p, _d, _i, _p, _p, _p]
    lib.vmo_lb_rk438_step.argtypes = [_p, _p, _l, _d, _d, _d, _i, _i, _p, _d, _i]
    lib.vmo_baseline_vp_steps.argtypes = [_p, _p, _p, _l, _d, _i, _d, _d, _i, _i, _i, _i]
    lib.vmo_baseline_lb_rhs.argtypes = [_p, _p, _l, _d, _d, _i, _i, _d, _i, _p, _i]
    lib.vmo_max_threads.restype = _i
    for name in ("vmo_eval_all", "vmo_eval_all_deriv", "vmo_periodic_mass", "vmo_periodic_stiffness",
                 "vmo_dirichlet_mass", "vmo_deposit_periodic", "vmo_poisson_solve", "vmo_eval_dphi",
                 "vmo_s_advection", "vmo_s_acceleration", "vmo_vp_strang_step", "vmo_integrate_vp",
                 "vmo_vproject", "vmo_vspline_eval", "vmo_vmoments", "vmo_clb_coefficients",
                 "vmo_lb_rhs", "vmo_lb_rk438_step", "vmo_baseline_vp_steps", "vmo_baseline_lb_rhs"):
        getattr(lib, name).restype = None
    return lib


def lib(path: str | None = None):
    """Load (building if needed) the oracle library."""
    global _LIB
    if path is not None:
        return _sig(C.CDLL(path))
    if _LIB is None:
        so = _HERE / "libvm_oracle.so"
        src = _HERE / "vm_oracle.c"
        if (not so.exists()) or so.stat().st_mtime < src.stat().st_mtime:
            build()
        _LIB = _sig(C.CDLL(str(so)))
    return _LIB


def _a(x):
    return np.ascontiguousarray(x, dtype=np.float64)


def _ptr(a):
    return a.ctypes.data_as(_p) if a is not None else None


# ------------------------------------------------------------------ basis ---
def bspline_shift_bsplinekit(order: int) -> int:
    """Index rotation of BSplineKit's periodic basis (knots shifted by k//2)."""
    return order // 2 - order + 1


def periodic_eval(a, b, n, k, x):
    N = np.zeros(k)
    dN = np.zeros(k)
    c = lib().vmo_periodic_eval(a, b, n, k, float(x), _ptr(N), _ptr(dN))
    return c, N, dN


def clamped_eval(a, b, nknots, k, x):
    N = np.zeros(k)
    dN = np.zeros(k)
    c = lib().vmo_clamped_eval(a, b, nknots, k, float(x), _ptr(N), _ptr(dN))
    return c, N, dN


def periodic_mass(a, b, n, k, shift=0):
    M = np.zeros((n, n))
    lib().vmo_periodic_mass(a, b, n, k, shift, _ptr(M))
    return M


def periodic_stiffness(a, b, n, k, shift=0):
    S = np.zeros((n, n))
    lib().vmo_periodic_stiffness(a, b, n, k, shift, _ptr(S))
    return S


def dirichlet_mass(a, b, nknots, k):
    nv = nknots + k - 4
    M = np.zeros((nv, nv))
    lib().vmo_dirichlet_mass(a, b, nknots, k, _ptr(M))
    return M


# ---------------------------------------------------------------- x-space ---
def deposit_periodic(x, w, a, b, n, k, shift=0):
    x = _a(x); w = _a(w)
    rhs = np.zeros(n)
    lib().vmo_deposit_periodic(_ptr(x), _ptr(w), x.size, a, b, n, k, shift, _ptr(rhs))
    return rhs


def poisson_solve(S, rhs):
    S = _a(S); rhs = _a(rhs)
    phi = np.zeros(rhs.size)
    lib().vmo_poisson_solve(_ptr(S), rhs.size, _ptr(rhs), _ptr(phi))
    return phi


def eval_dphi(x, a, b, n, k, shift, phi):
    x = _a(x); phi = _a(phi)
    out = np.zeros(x.size)
    lib().vmo_eval_dphi(_ptr(x), x.size, a, b, n, k, shift, _ptr(phi), _ptr(out))
    return out


def field_energy(S, phi):
    S = _a(S); phi = _a(phi)
    return lib().vmo_field_energy(_ptr(S), phi.size, _ptr(phi))


def s_advection(x, v, dt):
    lib().vmo_s_advection(_ptr(x), _ptr(v), x.size, dt)


def s_acceleration(x, v, w, dt, a, b, n, k, shift, S, x_src=None):
    lib().vmo_s_acceleration(_ptr(x), _ptr(v), _ptr(w), x.size, dt, a, b, n, k, shift, _ptr(S),
                             _ptr(x_src) if x_src is not None else None)


def vp_strang_step(x, v, w, dt, a, b, n, k, shift, S, x_src=None):
    """In-place on contiguous float64 x, v."""
    lib().vmo_vp_strang_step(_ptr(x), _ptr(v), _ptr(w), x.size, dt, a, b, n, k, shift, _ptr(S),
                             _ptr(x_src) if x_src is not None else None)


def integrate_vp(x, v, w, dt, chi, nt, nsave, a, b, n, k, shift, S, want_phi=False):
    nrec = (nt // nsave + 1) if nsave > 0 else 1
    diag = np.zeros((nrec, 3))
    phi_hist = np.zeros((nrec, n)) if want_phi else None
    lib().vmo_integrate_vp(_ptr(x), _ptr(v), _ptr(w), x.size, dt, chi, nt, nsave, a, b, n, k, shift,
                           _ptr(S), _ptr(diag), _ptr(phi_hist) if want_phi else None)
    return (diag, phi_hist) if want_phi else diag


def integrate_vp_external(x, v, w, dt, chi, nt, nsave, a, b, n, k, shift, S, coeffs, dt_c):
    """Legacy loop in a prescribed field; coeffs is (n, ncols), column ts = phi at time ts * dt_c.  In-place on x, v."""
    cm = np.asfortranarray(np.asarray(coeffs, dtype=np.float64))
    nrec = (nt // nsave + 1) if nsave > 0 else 0
    diag = np.zeros((max(nrec, 1), 3))
    lib().vmo_integrate_vp_external(_ptr(x), _ptr(v), _ptr(w), x.size, dt, chi, nt, nsave, a, b, n, k, shift,
                                    _ptr(S), cm.ctypes.data_as(_p), cm.shape[1], dt_c, _ptr(diag))
    return diag[:nrec]


def lorentz_force(x, v, w, a, b, n, k, shift, S, x_src=None):
    xdot, vdot = np.empty_like(x), np.empty_like(x)
    lib().vmo_lorentz_force(_ptr(x), _ptr(v), _ptr(w), x.size, a, b, n, k, shift, _ptr(S),
                            _ptr(x_src) if x_src is not None else None, _ptr(xdot), _ptr(vdot))
    return xdot, vdot


def vp_rk4_step(x, v, w, dt, a, b, n, k, shift, S):
    lib().vmo_vp_rk4_step(_ptr(x), _ptr(v), _ptr(w), x.size, dt, a, b, n, k, shift, _ptr(S))


# ---------------------------------------------------------------- v-space ---
def vproject(v, w, a, b, nknots, k, M):
    v = _a(v); w = _a(w); M = _a(M)
    nv = nknots + k - 4
    coef = np.zeros(nv); rhs = np.zeros(nv)
    lib().vmo_vproject(_ptr(v), _ptr(w), v.size, a, b, nknots, k, _ptr(M), _ptr(coef), _ptr(rhs))
    return coef, rhs


def vspline_eval(v, a, b, nknots, k, coef):
    v = _a(v); coef = _a(coef)
    f = np.zeros(v.size); df = np.zeros(v.size)
    lib().vmo_vspline_eval(_ptr(v), v.size, a, b, nknots, k, _ptr(coef), _ptr(f), _ptr(df))
    return f, df


def vmoments(v, a, b, nknots, k, coef):
    v = _a(v); coef = _a(coef)
    out = np.zeros(5)
    lib().vmo_vmoments(_ptr(v), v.size, a, b, nknots, k, _ptr(coef), _ptr(out))
    return out


def clb_coefficients(m5):
    m5 = _a(m5)
    A1 = C.c_double(); A2 = C.c_double()
    lib().vmo_clb_coefficients(_ptr(m5), C.byref(A1), C.byref(A2))
    return A1.value, A2.value


def lb_rhs(v, w, a, b, nknots, k, M, nu=1.0, conservative=False):
    v = _a(v); w = _a(w); M = _a(M)
    nv = nknots + k - 4
    vdot = np.zeros(v.size); coef = np.zeros(nv); A = np.zeros(2)
    lib().vmo_lb_rhs(_ptr(v), _ptr(w), v.size, a, b, nknots, k, _ptr(M), nu, int(conservative),
                     _ptr(vdot), _ptr(coef), _ptr(A))
    return vdot, coef, A


def lb_rk438_step(v, w, dt, a, b, nknots, k, M, nu=1.0, conservative=False):
    """In-place on contiguous float64 v."""
    lib().vmo_lb_rk438_step(_ptr(v), _ptr(w), v.size, dt, a, b, nknots, k, _ptr(M), nu, int(conservative))


# ------------------------------------------------------- timed CPU baseline -
def max_threads(native_lib=None):
    return (native_lib or lib()).vmo_max_threads()


def baseline_vp_steps(x, v, w, dt, nsteps, a, b, n, k, shift, nthreads, native_lib=None):
    (native_lib or lib()).vmo_baseline_vp_steps(_ptr(x), _ptr(v), _ptr(w), x.size, dt, nsteps, a, b, n, k,
                                                shift, nthreads)


def baseline_lb_rhs(v, w, a, b, nknots, k, nu, conservative, nthreads, native_lib=None):
    vdot = np.zeros(v.size)
    (native_lib or lib()).vmo_baseline_lb_rhs(_ptr(v), _ptr(w), v.size, a, b, nknots, k, nu, int(conservative),
                                              _ptr(vdot), nthreads)
    return vdot

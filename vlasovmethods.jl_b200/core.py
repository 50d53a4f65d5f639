"""Thin object wrappers over the C-ABI handles (context, particles, field, vspline).

All device memory is owned by libvlasov_b200.so; these classes only hold the
opaque handles and convert numpy arrays at the boundary.
"""
from __future__ import annotations

import atexit
import ctypes as C
import weakref

import numpy as np

from . import _lib as L


_live_contexts = weakref.WeakSet()


def _close_all_contexts():
    """Interpreter shutdown finalises objects in no particular order; a child handle destroyed after its context
    is a use-after-free inside the library.  Close every context (children first) while the world is intact."""
    for c in list(_live_contexts):
        try:
            c.close()
        except Exception:
            pass


atexit.register(_close_all_contexts)


def _f64(a, shape=None):
    a = np.ascontiguousarray(a, dtype=np.float64)
    if shape is not None and a.shape != shape:
        raise ValueError(f"expected shape {shape}, got {a.shape}")
    return a


class Context:
    """vm_ctx: one CUDA device, one stream, optional NCCL communicator (one process per GPU)."""

    def __init__(self, device: int = 0):
        self._h = C.c_void_p()
        self._children = weakref.WeakSet()      # particles / fields / splines created on this context
        L.check(L.lib().vm_ctx_create(int(device), C.byref(self._h)))
        self.device = int(device)
        _live_contexts.add(self)

    # -- lifetime ---------------------------------------------------------
    def close(self):
        """Destroy the context.  The C ABI requires every handle created on a context to be destroyed
        before the context itself; the wrappers register themselves here so that order is kept."""
        if getattr(self, "_h", None) is not None and self._h.value:
            for child in list(self._children):
                child.close()
            L.lib().vm_ctx_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def sync(self):
        L.check(L.lib().vm_sync(self._h), self._h)

    def device_info(self):
        sm, ma, mi = C.c_int(), C.c_int(), C.c_int()
        fr, to = C.c_size_t(), C.c_size_t()
        L.check(L.lib().vm_ctx_device_info(self._h, C.byref(sm), C.byref(ma), C.byref(mi), C.byref(fr), C.byref(to)), self._h)
        return {"sm_count": sm.value, "cc": (ma.value, mi.value), "free_bytes": fr.value, "total_bytes": to.value}

    def set_tuning(self, key: str, value: int):
        L.check(L.lib().vm_ctx_set_tuning(self._h, key.encode(), int(value)), self._h)

    # -- multi-GPU --------------------------------------------------------
    @staticmethod
    def unique_id() -> bytes:
        buf = C.create_string_buffer(128)
        L.check(L.lib().vm_comm_unique_id(buf))
        return buf.raw

    def comm_init(self, rank: int, nranks: int, uid: bytes | None):
        buf = C.create_string_buffer(uid, 128) if uid is not None else None
        L.check(L.lib().vm_ctx_comm_init(self._h, int(rank), int(nranks), buf), self._h)

    def peer_handle(self) -> bytes:
        buf = C.create_string_buffer(64)
        L.check(L.lib().vm_ctx_peer_handle(self._h, buf), self._h)
        return buf.raw

    def peer_connect(self, handles: bytes):
        """handles: nranks x 64 bytes in rank order (all-gathered peer_handle() results)."""
        buf = C.create_string_buffer(handles, len(handles))
        L.check(L.lib().vm_ctx_peer_connect(self._h, buf), self._h)
        self._peer_connected = True

    def peer_connected(self) -> bool:
        return bool(getattr(self, "_peer_connected", False))

    def comm_info(self):
        r, n = C.c_int(), C.c_int()
        L.check(L.lib().vm_ctx_comm_info(self._h, C.byref(r), C.byref(n)), self._h)
        return r.value, n.value

    # -- timing -----------------------------------------------------------
    def event_record(self, slot: int):
        L.check(L.lib().vm_event_record(self._h, int(slot)), self._h)

    def event_elapsed_ms(self, a: int, b: int) -> float:
        ms = C.c_double()
        L.check(L.lib().vm_event_elapsed_ms(self._h, int(a), int(b), C.byref(ms)), self._h)
        return ms.value

    def launch_count(self) -> int:
        return int(L.lib().vm_launch_count(self._h))

    def profile_read(self):
        """(launches, total_ms) of the event-bracketed dominant kernel since the last read."""
        n, ms = C.c_long(), C.c_double()
        L.check(L.lib().vm_profile_read(self._h, C.byref(n), C.byref(ms)), self._h)
        return n.value, ms.value


_default_ctx = None


def default_context() -> Context:
    """Process-wide context on device LOCAL_RANK (or 0)."""
    global _default_ctx
    if _default_ctx is None:
        import os
        _default_ctx = Context(int(os.environ.get("LOCAL_RANK", "0")))
    return _default_ctx


def set_default_context(ctx: Context | None):
    global _default_ctx
    _default_ctx = ctx


class PinnedArray:
    """float64 numpy view of page-locked host memory from vm_host_alloc (freed with the object)."""

    def __init__(self, n: int):
        self._ptr = C.c_void_p()
        L.check(L.lib().vm_host_alloc(C.c_size_t(8 * max(int(n), 1)), C.byref(self._ptr)))
        buf = (C.c_double * int(n)).from_address(self._ptr.value)
        self.array = np.frombuffer(buf, dtype=np.float64, count=int(n))

    def __del__(self):
        try:
            if getattr(self, "_ptr", None) is not None and self._ptr.value:
                self.array = None
                L.lib().vm_host_free(self._ptr)
                self._ptr = C.c_void_p()
        except Exception:
            pass


class DeviceParticles:
    """vm_particles: device SoA x[N], v[N], w[N]."""

    def __init__(self, ctx: Context, n: int):
        self.ctx = ctx
        self.n = int(n)
        self._h = C.c_void_p()
        L.check(L.lib().vm_particles_create(ctx._h, self.n, C.byref(self._h)), ctx._h)
        ctx._children.add(self)

    def close(self):
        if getattr(self, "_h", None) is not None and self._h.value:
            L.lib().vm_particles_destroy(self._h)          # safe in any order: the library tolerates orphaned handles
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __len__(self):
        return self.n

    def upload_aos(self, z):
        """z: (N, 3) C-ordered == Julia's 3xN column-major ParticleList matrix [x; v; w]."""
        z = _f64(z, (self.n, 3))
        L.check(L.lib().vm_particles_upload_aos(self._h, L.dptr(z)), self.ctx._h)

    def download_aos(self, out=None):
        out = np.empty((self.n, 3)) if out is None else out
        assert out.flags.c_contiguous and out.dtype == np.float64 and out.shape == (self.n, 3)
        L.check(L.lib().vm_particles_download_aos(self._h, L.dptr(out)), self.ctx._h)
        return out

    def upload(self, x=None, v=None, w=None):
        arrs = [None if a is None else _f64(a, (self.n,)) for a in (x, v, w)]
        L.check(L.lib().vm_particles_upload_soa(self._h, *[L.dptr(a) for a in arrs]), self.ctx._h)

    def download(self, x=True, v=True, w=True, out=None):
        """Returns (x, v, w) host arrays (None where not requested).  `out` = preallocated triple."""
        arrs = list(out) if out is not None else [np.empty(self.n) if f else None for f in (x, v, w)]
        L.check(L.lib().vm_particles_download_soa(self._h, *[L.dptr(a) for a in arrs]), self.ctx._h)
        return tuple(arrs)

    def snapshot_begin(self, x_host=None, v_host=None):
        """Start an asynchronous copy of x and/or v into (ideally pinned) host arrays; returns at once."""
        for a in (x_host, v_host):
            assert a is None or (a.dtype == np.float64 and a.flags.c_contiguous and a.size == self.n)
        L.check(L.lib().vm_particles_snapshot_begin(self._h, L.dptr(x_host), L.dptr(v_host)), self.ctx._h)

    def snapshot_wait(self):
        L.check(L.lib().vm_particles_snapshot_wait(self._h), self.ctx._h)

    def set_uniform_weight(self, w0: float):
        """Declare w_p = w0 for every particle (what every sampler of the reference produces) instead of uploading w."""
        L.check(L.lib().vm_particles_set_uniform_weight(self._h, float(w0)), self.ctx._h)

    def copy_from(self, other: "DeviceParticles"):
        L.check(L.lib().vm_particles_copy(self._h, other._h), self.ctx._h)

    def fill(self, kind: int, params, seed: int, first_index: int = 0, total_n: int | None = None):
        p = _f64(params)
        L.check(L.lib().vm_particles_fill(self._h, int(kind), L.dptr(p), p.size, C.c_ulonglong(seed),
                                          int(first_index), int(self.n if total_n is None else total_n)), self.ctx._h)

    def drift(self, dt: float):
        L.check(L.lib().vm_vp_drift(self._h, float(dt)), self.ctx._h)


class DeviceField:
    """vm_field: periodic B-spline potential (deposit target, Poisson solve, E gather)."""

    def __init__(self, ctx: Context, a: float, b: float, order: int, n_basis: int, index_shift: int = 0):
        self.ctx = ctx
        self.a, self.b, self.order, self.n, self.shift = float(a), float(b), int(order), int(n_basis), int(index_shift)
        self._h = C.c_void_p()
        L.check(L.lib().vm_field_create(ctx._h, self.a, self.b, self.order, self.n, self.shift, C.byref(self._h)), ctx._h)
        ctx._children.add(self)

    def close(self):
        if getattr(self, "_h", None) is not None and self._h.value:
            L.lib().vm_field_destroy(self._h)          # safe in any order: the library tolerates orphaned handles
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    @property
    def rhs(self):
        out = np.empty(self.n)
        L.check(L.lib().vm_field_get_rhs(self._h, L.dptr(out)), self.ctx._h)
        return out

    @property
    def coefficients(self):
        out = np.empty(self.n)
        L.check(L.lib().vm_field_get_coefficients(self._h, L.dptr(out)), self.ctx._h)
        return out

    @coefficients.setter
    def coefficients(self, phi):
        phi = _f64(phi, (self.n,))
        L.check(L.lib().vm_field_set_coefficients(self._h, L.dptr(phi)), self.ctx._h)

    def stencils(self):
        m, s = np.empty(self.order), np.empty(self.order)
        L.check(L.lib().vm_field_get_stencils(self._h, L.dptr(m), L.dptr(s)), self.ctx._h)
        return m, s

    def _circulant(self, st):
        n = self.n
        M = np.zeros((n, n))
        for i in range(n):
            M[i, i] += st[0]
            for d in range(1, self.order):
                M[i, (i + d) % n] += st[d]
                M[i, (i - d) % n] += st[d]
        return M

    def stiffness_matrix(self):
        return self._circulant(self.stencils()[1])

    def mass_matrix(self):
        return self._circulant(self.stencils()[0])

    def deposit(self, p: DeviceParticles, mode: int = L.VM_DEPOSIT_DETERMINISTIC):
        L.check(L.lib().vm_deposit(self._h, p._h, int(mode)), self.ctx._h)

    def solve(self):
        L.check(L.lib().vm_field_solve(self._h), self.ctx._h)

    def energy(self) -> float:
        W = C.c_double()
        L.check(L.lib().vm_field_energy(self._h, C.byref(W)), self.ctx._h)
        return W.value

    def gather_E(self, p: DeviceParticles, inv_chi2: float = 1.0, to_host: bool = True):
        e = np.empty(p.n) if to_host else None
        L.check(L.lib().vm_gather_E(self._h, p._h, L.dptr(e), float(inv_chi2)), self.ctx._h)
        return e

    def eval(self, x, deriv: int = 0):
        x = _f64(np.atleast_1d(x))
        out = np.empty(x.size)
        L.check(L.lib().vm_field_eval(self._h, L.dptr(x), x.size, int(deriv), L.dptr(out)), self.ctx._h)
        return out

    def kick(self, p: DeviceParticles, dt: float, scale: float = -1.0):
        L.check(L.lib().vm_vp_kick(self._h, p._h, float(dt), float(scale)), self.ctx._h)

    def run(self, p: DeviceParticles, dt: float, nsteps: int, diag_every: int = 0, flags: int = 0, chi: float = 1.0):
        """nsteps fused Strang steps; returns diag rows [W, K, M, sum_w] (or None)."""
        diag = np.zeros((nsteps // diag_every + 1, 4)) if diag_every > 0 else None
        L.check(L.lib().vm_vp_run(self._h, p._h, float(dt), int(nsteps), int(diag_every), int(flags), float(chi),
                                  L.dptr(diag)), self.ctx._h)
        return diag

    def run_external(self, p: DeviceParticles, dt: float, nsteps: int, coeffs, coeff_dt: float, diag_every: int = 0,
                     chi: float = 1.0):
        """nsteps leapfrog steps in a prescribed field: coeffs is (n_basis, ncols), column ts = phi at time ts*coeff_dt
        (ExternalField, src/electric_field.jl:55-77).  Returns diag rows [W, K, M, sum_w] (or None)."""
        coeffs = np.asarray(coeffs, dtype=np.float64)
        if coeffs.ndim != 2 or coeffs.shape[0] != self.n:
            raise ValueError(f"coeffs must be ({self.n}, ncols)")
        cm = np.asfortranarray(coeffs)                  # column-major: the memory of the Julia matrix
        diag = np.zeros((nsteps // diag_every + 1, 4)) if diag_every > 0 else None
        L.check(L.lib().vm_vp_run_external(self._h, p._h, float(dt), int(nsteps), int(diag_every), float(chi),
                                           cm.ctypes.data_as(L._dp), int(cm.shape[1]), float(coeff_dt), L.dptr(diag)), self.ctx._h)
        return diag

    def vector_field(self, p: DeviceParticles, keep_potential: bool = False, to_host: bool = True):
        """lorentz_force!: (xdot, vdot) = (v, -phi'(x)) after update_potential!; to_host=False leaves them on the device."""
        xdot = np.empty(p.n) if to_host else None
        vdot = np.empty(p.n) if to_host else None
        L.check(L.lib().vm_vp_vector_field(self._h, p._h, L.VM_VF_KEEP_POTENTIAL if keep_potential else 0,
                                           L.dptr(xdot), L.dptr(vdot)), self.ctx._h)
        return xdot, vdot

    def rk4_run(self, p: DeviceParticles, dt: float, nsteps: int):
        """nsteps classical RK4 steps of the unsplit vector field lorentz_force!, device-resident."""
        L.check(L.lib().vm_vp_rk4_run(self._h, p._h, float(dt), int(nsteps)), self.ctx._h)

    def diagnostics(self, p: DeviceParticles, chi: float = 1.0):
        out = np.empty(4)
        L.check(L.lib().vm_diagnostics(self._h, p._h, float(chi), L.dptr(out)), self.ctx._h)
        return out


class DeviceVSpline:
    """vm_vspline: clamped (Dirichlet) velocity-space spline f_s(v) with its mass-matrix solve."""

    def __init__(self, ctx: Context, vmin: float, vmax: float, nknots: int, order: int, bc: int = 1):
        self.ctx = ctx
        self.a, self.b, self.nknots, self.order, self.bc = float(vmin), float(vmax), int(nknots), int(order), int(bc)
        self._h = C.c_void_p()
        L.check(L.lib().vm_vspline_create(ctx._h, self.a, self.b, self.nknots, self.order, self.bc, C.byref(self._h)), ctx._h)
        ctx._children.add(self)
        self.nv = L.lib().vm_vspline_size(self._h)

    def close(self):
        if getattr(self, "_h", None) is not None and self._h.value:
            L.lib().vm_vspline_destroy(self._h)          # safe in any order: the library tolerates orphaned handles
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __len__(self):
        return self.nv

    @property
    def coefficients(self):
        out = np.empty(self.nv)
        L.check(L.lib().vm_vspline_get_coefficients(self._h, L.dptr(out)), self.ctx._h)
        return out

    @coefficients.setter
    def coefficients(self, c):
        c = _f64(c, (self.nv,))
        L.check(L.lib().vm_vspline_set_coefficients(self._h, L.dptr(c)), self.ctx._h)

    @property
    def rhs(self):
        out = np.empty(self.nv)
        L.check(L.lib().vm_vspline_get_rhs(self._h, L.dptr(out)), self.ctx._h)
        return out

    def mass_matrix(self):
        out = np.empty((self.nv, self.nv))
        L.check(L.lib().vm_vspline_get_mass_matrix(self._h, L.dptr(out)), self.ctx._h)
        return out

    def project(self, p: DeviceParticles):
        L.check(L.lib().vm_vproject(self._h, p._h), self.ctx._h)

    def project_at(self, p: DeviceParticles, v):
        """projection with replacement velocities v (the particle state on the device is not modified)."""
        v = _f64(v, (p.n,))
        L.check(L.lib().vm_vproject_at(self._h, p._h, L.dptr(v)), self.ctx._h)

    def moments_at(self, p: DeviceParticles, v):
        v = _f64(v, (p.n,))
        m5, A = np.empty(5), np.empty(2)
        L.check(L.lib().vm_vmoments_at(self._h, p._h, L.dptr(v), L.dptr(m5), L.dptr(A)), self.ctx._h)
        return m5, A

    def lb_rhs_at(self, p: DeviceParticles, v, nu: float = 1.0, conservative: bool = False):
        v = _f64(v, (p.n,))
        vdot = np.empty(p.n)
        L.check(L.lib().vm_lb_rhs_at(self._h, p._h, L.dptr(v), float(nu), int(conservative), L.dptr(vdot)), self.ctx._h)
        return vdot

    def eval(self, v):
        v = _f64(np.atleast_1d(v))
        f, df = np.empty(v.size), np.empty(v.size)
        L.check(L.lib().vm_vspline_eval(self._h, L.dptr(v), v.size, L.dptr(f), L.dptr(df)), self.ctx._h)
        return f, df

    def moments(self, p: DeviceParticles):
        m5, A = np.empty(5), np.empty(2)
        L.check(L.lib().vm_vmoments(self._h, p._h, L.dptr(m5), L.dptr(A)), self.ctx._h)
        return m5, A

    def lb_rhs(self, p: DeviceParticles, nu: float = 1.0, conservative: bool = False, to_host: bool = True):
        vdot = np.empty(p.n) if to_host else None
        L.check(L.lib().vm_lb_rhs(self._h, p._h, float(nu), int(conservative), L.dptr(vdot)), self.ctx._h)
        return vdot

    def rk438_run(self, p: DeviceParticles, dt: float, nsteps: int, nu: float = 1.0, conservative: bool = False,
                  diag_every: int = 0):
        diag = np.zeros((nsteps // diag_every + 1, 4)) if diag_every > 0 else None
        L.check(L.lib().vm_lb_rk438_run(self._h, p._h, float(dt), int(nsteps), float(nu), int(conservative),
                                        int(diag_every), L.dptr(diag)), self.ctx._h)
        return diag

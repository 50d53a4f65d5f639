"""ctypes binding of libvlasov_b200.so (include/vlasov_b200.h).

There is deliberately NO fallback: if the shared library is missing or no CUDA
device is visible the calls raise.  Nothing here imports the CPU oracle.
"""
from __future__ import annotations

import ctypes as C
import subprocess
from pathlib import Path

import os

_HERE = Path(__file__).resolve().parent
# VLASOV_B200_LIB selects another build of the same ABI (A/B experiments between kernel revisions)
LIB_PATH = Path(os.environ.get("VLASOV_B200_LIB", _HERE / "libvlasov_b200.so"))

_dp = C.POINTER(C.c_double)
_vp = C.c_void_p
_i = C.c_int
_l = C.c_long
_d = C.c_double

# name -> (restype, argtypes); mirrors include/vlasov_b200.h one to one
SIGNATURES = {
    "vm_abi_version": (_i, []),
    "vm_ctx_create": (_i, [_i, C.POINTER(_vp)]),
    "vm_ctx_destroy": (_i, [_vp]),
    "vm_last_error": (C.c_char_p, [_vp]),
    "vm_sync": (_i, [_vp]),
    "vm_ctx_device_info": (_i, [_vp, C.POINTER(_i), C.POINTER(_i), C.POINTER(_i), C.POINTER(C.c_size_t), C.POINTER(C.c_size_t)]),
    "vm_ctx_set_tuning": (_i, [_vp, C.c_char_p, _i]),
    "vm_comm_unique_id": (_i, [_vp]),
    "vm_ctx_comm_init": (_i, [_vp, _i, _i, _vp]),
    "vm_ctx_comm_info": (_i, [_vp, C.POINTER(_i), C.POINTER(_i)]),
    "vm_ctx_peer_handle": (_i, [_vp, _vp]),
    "vm_ctx_peer_connect": (_i, [_vp, _vp]),
    "vm_event_record": (_i, [_vp, _i]),
    "vm_event_elapsed_ms": (_i, [_vp, _i, _i, C.POINTER(_d)]),
    "vm_launch_count": (C.c_ulonglong, [_vp]),
    "vm_profile_read": (_i, [_vp, C.POINTER(_l), C.POINTER(_d)]),
    "vm_particles_create": (_i, [_vp, _l, C.POINTER(_vp)]),
    "vm_particles_destroy": (_i, [_vp]),
    "vm_particles_size": (_l, [_vp]),
    "vm_particles_upload_aos": (_i, [_vp, _dp]),
    "vm_particles_download_aos": (_i, [_vp, _dp]),
    "vm_particles_upload_soa": (_i, [_vp, _dp, _dp, _dp]),
    "vm_particles_download_soa": (_i, [_vp, _dp, _dp, _dp]),
    "vm_particles_copy": (_i, [_vp, _vp]),
    "vm_particles_set_uniform_weight": (_i, [_vp, _d]),
    "vm_particles_snapshot_begin": (_i, [_vp, _dp, _dp]),
    "vm_particles_snapshot_wait": (_i, [_vp]),
    "vm_host_alloc": (_i, [C.c_size_t, C.POINTER(_vp)]),
    "vm_host_free": (_i, [_vp]),
    "vm_particles_fill": (_i, [_vp, _i, _dp, _i, C.c_ulonglong, _l, _l]),
    "vm_field_create": (_i, [_vp, _d, _d, _i, _i, _i, C.POINTER(_vp)]),
    "vm_field_destroy": (_i, [_vp]),
    "vm_field_get_rhs": (_i, [_vp, _dp]),
    "vm_field_get_coefficients": (_i, [_vp, _dp]),
    "vm_field_set_coefficients": (_i, [_vp, _dp]),
    "vm_field_get_stencils": (_i, [_vp, _dp, _dp]),
    "vm_pass_plan_query": (_i, [_i, C.c_size_t, _i, _i, _i, _i, _vp]),
    "vm_deposit": (_i, [_vp, _vp, _i]),
    "vm_field_solve": (_i, [_vp]),
    "vm_field_energy": (_i, [_vp, C.POINTER(_d)]),
    "vm_gather_E": (_i, [_vp, _vp, _dp, _d]),
    "vm_field_eval": (_i, [_vp, _dp, _l, _i, _dp]),
    "vm_vp_drift": (_i, [_vp, _d]),
    "vm_vp_kick": (_i, [_vp, _vp, _d, _d]),
    "vm_vp_run": (_i, [_vp, _vp, _d, _i, _i, _i, _d, _dp]),
    "vm_vp_run_external": (_i, [_vp, _vp, _d, _i, _i, _d, _dp, _i, _d, _dp]),
    "vm_vp_vector_field": (_i, [_vp, _vp, _i, _dp, _dp]),
    "vm_vp_rk4_run": (_i, [_vp, _vp, _d, _i]),
    "vm_diagnostics": (_i, [_vp, _vp, _d, _dp]),
    "vm_vspline_create": (_i, [_vp, _d, _d, _i, _i, _i, C.POINTER(_vp)]),
    "vm_vspline_destroy": (_i, [_vp]),
    "vm_vspline_size": (_i, [_vp]),
    "vm_vspline_get_coefficients": (_i, [_vp, _dp]),
    "vm_vspline_set_coefficients": (_i, [_vp, _dp]),
    "vm_vspline_get_rhs": (_i, [_vp, _dp]),
    "vm_vspline_get_mass_matrix": (_i, [_vp, _dp]),
    "vm_vproject": (_i, [_vp, _vp]),
    "vm_vproject_at": (_i, [_vp, _vp, _dp]),
    "vm_vmoments_at": (_i, [_vp, _vp, _dp, _dp, _dp]),
    "vm_lb_rhs_at": (_i, [_vp, _vp, _dp, _d, _i, _dp]),
    "vm_vspline_eval": (_i, [_vp, _dp, _l, _dp, _dp]),
    "vm_vmoments": (_i, [_vp, _vp, _dp, _dp]),
    "vm_lb_rhs": (_i, [_vp, _vp, _d, _i, _dp]),
    "vm_lb_rk438_run": (_i, [_vp, _vp, _d, _i, _d, _i, _i, _dp]),
}

# enums of the header
VM_DEPOSIT_DETERMINISTIC, VM_DEPOSIT_ATOMIC, VM_DEPOSIT_FIXED = 0, 1, 2
VM_RUN_SPLIT_KICK, VM_RUN_FROZEN_FIELD, VM_RUN_ATOMIC_DEPOSIT, VM_RUN_UNFUSED, VM_RUN_FIXED_DEPOSIT = 1, 2, 4, 8, 16
VM_VF_KEEP_POTENTIAL = 1
(VM_FILL_NORMAL, VM_FILL_BUMP_ON_TAIL, VM_FILL_DOUBLE_MAXWELLIAN, VM_FILL_UNIFORM,
 VM_FILL_SHIFTED_NORMAL_V, VM_FILL_SHIFTED_UNIFORM, VM_FILL_LANDAU, VM_FILL_BUMP_ON_TAIL_SOBOL,
 VM_FILL_BUMP_ON_TAIL_SOBOL_IS) = range(9)
VM_OK, VM_ERR_INVALID, VM_ERR_CUDA, VM_ERR_NOMEM, VM_ERR_NCCL, VM_ERR_UNSUPPORTED, VM_ERR_NO_DEVICE = range(7)   # vm_status

_lib = None


class VMError(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__(f"libvlasov_b200 error {code}: {msg}")
        self.code = code


class PassPlan(C.Structure):
    """vm_pass_plan (include/vlasov_b200.h)."""
    _fields_ = [("variant", _i), ("replicas", _i), ("grid", _i), ("threads", _i), ("pairs", _i), ("max_threads", _i),
                ("gather_copies", _i), ("smem_bytes", C.c_size_t)]


def pass_plan(n_basis: int, order: int, pass_: int, deposit_mode: int = 0, sm_count: int = 148,
              smem_optin_bytes: int = 232448) -> PassPlan:
    """Launch plan for a particle pass (pure host logic: works without a GPU)."""
    out = PassPlan()
    check(lib().vm_pass_plan_query(int(sm_count), C.c_size_t(int(smem_optin_bytes)), int(n_basis), int(order), int(pass_),
                                   int(deposit_mode), C.byref(out)))
    return out


def build() -> Path:
    """Compile the CUDA library in-tree for sm_100a (nvcc cross-compiles without a GPU)."""
    subprocess.check_call(["make", "-s", "-j8", "-C", str(_HERE / "csrc")])
    return LIB_PATH


def lib():
    """Load the shared library; raise loudly if it has not been built."""
    global _lib
    if _lib is None:
        if not LIB_PATH.exists():
            raise FileNotFoundError(
                f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                "(there is no CPU fallback for this package)")
        handle = C.CDLL(str(LIB_PATH))
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(handle, name)       # AttributeError if the ABI lost a symbol
            fn.restype = res
            fn.argtypes = args
        if handle.vm_abi_version() != 2:
            raise RuntimeError("libvlasov_b200.so ABI version mismatch")
        _lib = handle
    return _lib


def check(rc: int, ctx=None):
    if rc != 0:
        msg = lib().vm_last_error(ctx)
        raise VMError(rc, msg.decode() if msg else "unknown error")


def dptr(a):
    """numpy float64 contiguous array -> double* (None -> NULL)."""
    return a.ctypes.data_as(_dp) if a is not None else None

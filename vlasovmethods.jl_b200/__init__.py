"""vlasovmethods.jl_b200 -- B200-native (sm_100a) drop-in for the particle hot path of
JuliaPlasma/VlasovMethods.jl.  The directory name contains a dot, so load it with
`__graft_entry__.load_package()` (registers it as module `vlasovmethods_jl_b200`).

Layout: csrc/ (hand-written CUDA kernels + C ABI, builds libvlasov_b200.so), _lib.py (ctypes
binding), core.py (handle wrappers), api.py (mirror of the reference's new API), legacy.py
(mirror of src/electric_field.jl + src/vlasov_poisson.jl), sharding.py (one process per GPU)."""
from . import _lib
from ._lib import VMError, build, lib
from .core import (Context, DeviceField, DeviceParticles, DeviceVSpline, PinnedArray, default_context,
                   set_default_context)
from .api import *          # noqa: F401,F403
from .api import (initialize_, projection_, projection, update_ as update_potential_solver_, run_, LB_rhs_, CLB_rhs_,
                  update_potential_, compute_coefficients, compute_f_densities, compute_df_densities,
                  s_advection_, s_acceleration_, lorentz_force_, v_advection_, v_acceleration_)
from . import legacy
from .legacy import (PoissonSolverPBSplines, PoissonField, ExternalField, ScaledField, ScaledPoissonField,
                     ScaledExternalField, VPIntegratorParameters, VPIntegratorCache, integrate_vp_, energy,
                     coefficients, solve_, eval_field_, efield_)
from .sharding import shard_bounds, init_distributed_context

/*
 * vlasov_b200.h -- C ABI of libvlasov_b200.so
 *
 * B200-native (sm_100a, fp64) implementation of the particle hot path of
 * JuliaPlasma/VlasovMethods.jl: one step of the 1d1v spline particle-in-cell
 * loop (Vlasov-Poisson) and the spline velocity-space projection + right-hand
 * side of the Lenard-Bernstein collision operators.
 *
 * The reference has no FFI: its boundary is Julia multiple dispatch.  Each entry
 * point below names the reference call site it replaces (paths relative to the
 * reference repo root); INTEGRATION.md shows the `ccall` bodies a maintainer
 * would add behind the unchanged Julia signatures.
 *
 * Conventions
 *  - every function returns a vm_status (0 = OK); the message of the last
 *    failure is returned by vm_last_error(ctx) (ctx may be NULL for failures of
 *    vm_ctx_create).  No exception crosses this boundary.
 *  - the library owns all device memory behind opaque handles; host pointers
 *    are borrowed for the duration of the call only.
 *  - a vm_ctx is bound to ONE device and ONE host thread at a time (the
 *    reference is single-threaded).  Calls enqueue work on the context's stream
 *    and return; any call that hands data back to the host synchronises.
 *  - multi-GPU: one process (one vm_ctx) per GPU; particles are sharded, the
 *    deposited grid is all-reduced over NCCL inside vm_field_solve /
 *    vm_vproject, the small solves are replicated (bit-identical on all ranks).
 *  - Float64 only, 1d1v only.  There is NO CPU fallback: without a CUDA device
 *    vm_ctx_create fails with VM_ERR_NO_DEVICE.
 */
#ifndef VLASOV_B200_H
#define VLASOV_B200_H

#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

#define VM_ABI_VERSION 2

typedef struct vm_ctx vm_ctx;
typedef struct vm_particles vm_particles;
typedef struct vm_field vm_field;
typedef struct vm_vspline vm_vspline;

typedef enum vm_status {
    VM_OK = 0,
    VM_ERR_INVALID = 1,      /* bad argument */
    VM_ERR_CUDA = 2,         /* CUDA runtime failure (message has the detail) */
    VM_ERR_NOMEM = 3,
    VM_ERR_NCCL = 4,
    VM_ERR_UNSUPPORTED = 5,  /* e.g. spline order outside 2..6 */
    VM_ERR_NO_DEVICE = 6     /* no CUDA device: the library has no CPU path */
} vm_status;

/* ---------------------------------------------------------------- context */
int vm_abi_version(void);

/* Create a context on CUDA device `device` (ordinal). */
int vm_ctx_create(int device, vm_ctx** out);
/* Handles created on a context should be destroyed before it.  Host languages that finalise objects in no
 * particular order (Julia, Python at exit) are tolerated: destroying a child handle after its context frees the
 * child's device memory without touching the context, any other call on it returns VM_ERR_INVALID, and destroying
 * a context twice is a no-op. */
int vm_ctx_destroy(vm_ctx* ctx);
const char* vm_last_error(vm_ctx* ctx);
/* Block until all work enqueued on the context has finished. */
int vm_sync(vm_ctx* ctx);
int vm_ctx_device_info(vm_ctx* ctx, int* sm_count, int* cc_major, int* cc_minor,
                       size_t* free_bytes, size_t* total_bytes);
/* Tuning knobs (defaults are chosen from the device and problem size; everything below exists for A/B
 * measurements and tests -- results are the same to rounding for every setting).
 * key: "ctas_per_sm", "threads_per_cta", "replicas" (0 = auto): CTA shape / replica grids per warp of the passes;
 *      "pairs" (0 = auto, 1, 2, 4, 8): pairs of particles in flight per thread in the lane-private passes;
 *      "priv_min_warps" (0 = auto): fewest warps per SM for which the lane-private deposit is still chosen;
 *      "no_repg" (1: single field table in the fused pass instead of 16 bank-conflict-free copies);
 *      "af" (limb-atomic fixed-point pass, the default layout of larger meshes: 0 = auto -- fused step from 20 cells,
 *      deposit-only pass from 88 --, 1 = always, -1 = never), "af_ctas" (its CTAs per SM: 0 = auto, 1, 2, 4),
 *      "af_replicas" (its bank-steered replicas per CTA: 0 = as many as fit, else a power of two <= 32);
 *      "bankq" (bank-sorted pass, the layout "af" replaced: 0 = auto -- from 88 cells when af = -1 --, 1 = always, -1 = never);
 *      "force_match" (1: MATCH.ANY grouping instead of xor-shuffle rounds), "no_uniform_w" (1: always stream the
 *      weight array), "no_pdl" (1: no programmatic dependent launch), "no_fuse" (1: separate reduce / solve kernels),
 *      "no_presolve" (1: meshes above 128 cells launch the solve kernel between fused passes instead of solving in the
 *      next pass's prologue);
 *      "profile" (see vm_profile_read). */
int vm_ctx_set_tuning(vm_ctx* ctx, const char* key, int value);

/* Multi-GPU, one process per GPU.  Rank 0 obtains a 128-byte NCCL unique id
 * and distributes it by any host channel (MPI.jl, torch.distributed, a file). */
int vm_comm_unique_id(void* out128);
int vm_ctx_comm_init(vm_ctx* ctx, int rank, int nranks, const void* id128);
int vm_ctx_comm_info(vm_ctx* ctx, int* rank, int* nranks);
/* Optional: fused deposit + exchange + solve over NVLink peer memory (<= 8 ranks on one node).
 * Every rank exports a 64-byte CUDA IPC handle of its inbox (vm_ctx_peer_handle), the handles are
 * all-gathered over the host channel and passed in rank order to vm_ctx_peer_connect.  From then on
 * the last CTA of the deposit pass writes this rank's partial grid into every peer's inbox, waits
 * for the peers' grids, sums them in rank order (bit-identical on all ranks) and solves -- one
 * kernel per step, no NCCL call on the hot path.  Without it the all-reduce uses NCCL. */
int vm_ctx_peer_handle(vm_ctx* ctx, void* out64);
int vm_ctx_peer_connect(vm_ctx* ctx, const void* handles_nranks_x_64);

/* Device timing on the context's own stream (CUDA events), slots 0..15. */
int vm_event_record(vm_ctx* ctx, int slot);
int vm_event_elapsed_ms(vm_ctx* ctx, int slot_start, int slot_stop, double* ms);
/* Number of kernels this context has launched so far. */
unsigned long long vm_launch_count(vm_ctx* ctx);
/* Per-launch device timing of the dominant kernel (the fused gather+kick+drift+deposit pass of
 * vm_vp_run, or the deposit pass of vm_deposit / vm_lb_rhs): when tuning key "profile" is 1 every
 * such launch is bracketed by CUDA events on the context's stream.  vm_profile_read synchronises,
 * returns the number of bracketed launches and their summed duration since the last read, and
 * resets the counters. */
int vm_profile_read(vm_ctx* ctx, long* launches, double* total_ms);

/* ------------------------------------------------------------- particles --
 * Replaces ParticleDistribution{1,1} (src/distributions/particle_distribution.jl:2-24),
 * a ParticleMethods.ParticleList over a 3xN Float64 matrix [x;v;w] (AoS).
 * Device layout: three SoA arrays x[N], v[N], w[N]. */
int vm_particles_create(vm_ctx* ctx, long n, vm_particles** out);
int vm_particles_destroy(vm_particles* p);
long vm_particles_size(vm_particles* p);
/* z3xn: column-major 3xN matrix = the memory of ParticleList.z with the weight row. */
int vm_particles_upload_aos(vm_particles* p, const double* z3xn);
int vm_particles_download_aos(vm_particles* p, double* z3xn);
/* Any of x, v, w may be NULL (left untouched). */
int vm_particles_upload_soa(vm_particles* p, const double* x, const double* v, const double* w);
int vm_particles_download_soa(vm_particles* p, double* x, double* v, double* w);
int vm_particles_copy(vm_particles* dst, vm_particles* src);
/* Every sampler of the reference gives all particles ONE weight (w = 1/N: src/examples/normal.jl:33, w = L/N:
 * bumpontail.jl:70).  A host that knows this declares it instead of uploading the weight row: the device array is
 * filled at HBM rate (no 8 B/particle over PCIe) and the passes take w0 as a kernel parameter. */
int vm_particles_set_uniform_weight(vm_particles* p, double w0);

/* Asynchronous, decimated snapshots for the driver loops (run!(::SplittingMethod) writes the full state
 * after every step, src/methods/splitting.jl:42 -- 1.6 GB per step at 1e8 particles).
 * vm_particles_snapshot_begin copies x and v (either may be NULL) into a device staging buffer on the
 * context's stream and starts the device->host copy into the caller's buffers on a second stream; the
 * call returns at once and stepping may continue.  vm_particles_snapshot_wait blocks until the data has
 * landed.  Host buffers should be page-locked (vm_host_alloc) for the copy to overlap with compute. */
int vm_particles_snapshot_begin(vm_particles* p, double* x_host, double* v_host);
int vm_particles_snapshot_wait(vm_particles* p);
int vm_host_alloc(size_t bytes, void** out);   /* page-locked host memory */
int vm_host_free(void* ptr);

/* Device-side synthetic loads reproducing the *distributions* of
 * the example files under src/examples (the reference draws from Julia's unseeded global RNG, so
 * streams cannot match).  Counter-based Philox4x32-10 keyed by
 * (seed, global particle index): the load is independent of the sharding.
 * first_index/total_n: this shard holds particles first_index .. first_index+n-1
 * of a global population of total_n. */
typedef enum vm_fill_kind {
    VM_FILL_NORMAL = 0,          /* normal.jl:10-36    params: xlo, xhi                     */
    VM_FILL_BUMP_ON_TAIL = 1,    /* bumpontail.jl:43-75 params: eps, kappa, alpha, sigma, v0 */
    VM_FILL_DOUBLE_MAXWELLIAN = 2,/* doublemaxwellian.jl:9-39 params: xlo, xhi, shift         */
    VM_FILL_UNIFORM = 3,         /* uniform.jl:10-33   params: xlo, xhi, vlo, vhi           */
    VM_FILL_SHIFTED_NORMAL_V = 4,/* shiftednormalv.jl  params: xlo, xhi, shift              */
    VM_FILL_SHIFTED_UNIFORM = 5, /* shifteduniform.jl  params: xlo, xhi, vlo, vhi, shift    */
    VM_FILL_LANDAU = 6,          /* (1+eps cos kx) Maxwellian  params: eps, kappa           */
    VM_FILL_BUMP_ON_TAIL_SOBOL = 7,    /* bumpontail.jl:43-75 as written: proposals from a 2-D Sobol sequence (Gray-code
                                          order, `skip` points skipped), accept-reject in x against f_x / (1 + eps),
                                          inverse CDF in v, w = L/N.  params: eps, kappa, alpha, sigma, v0, skip
                                          (skip < 0: Sobol.jl's skip(s, 2N) = the largest power of two <= 2N + 1) */
    VM_FILL_BUMP_ON_TAIL_SOBOL_IS = 8  /* bumpontail.jl:90-121 (ImportanceSampling): every Sobol proposal kept,
                                          w = f_x(x) L/N.  params as kind 7 */
} vm_fill_kind;
int vm_particles_fill(vm_particles* p, int kind, const double* params, int nparams,
                      unsigned long long seed, long first_index, long total_n);

/* ----------------------------------------------------------------- field --
 * Replaces Potential(PeriodicBasisBSplineKit(domain, order, nknot))
 * (scripts/vlasov_poisson.jl:21) and the legacy PoissonSolverPBSplines(p, nh, L)
 * (scripts/bump_on_tail.jl:38).  n_basis periodic B-splines of order `order`
 * (= degree + 1) on [a,b).  index_shift: a particle in cell c touches basis
 * indices (c + j + index_shift) mod n_basis, j = 0..order-1 (a pure rotation of
 * the coefficient vectors; BSplineKit's periodic basis uses order/2 - order + 1). */
int vm_field_create(vm_ctx* ctx, double a, double b, int order, int n_basis, int index_shift,
                    vm_field** out);
int vm_field_destroy(vm_field* f);
int vm_field_get_rhs(vm_field* f, double* host_n);            /* potential.rhs            */
int vm_field_get_coefficients(vm_field* f, double* host_n);   /* potential.coefficients   */
/* ExternalField (src/electric_field.jl:66-69): prescribe phi, no deposit/solve. */
int vm_field_set_coefficients(vm_field* f, const double* host_n);
/* Circulant stencils (centre, +1, ..., +order-1) of the mass and stiffness matrices. */
int vm_field_get_stencils(vm_field* f, double* mass_k, double* stiff_k);

typedef enum vm_deposit_mode {
    VM_DEPOSIT_DETERMINISTIC = 0, /* bit-reproducible run to run.  Small meshes: lane-private replica grids, fixed-order
                                     tree across warps / CTAs / ranks.  Larger meshes (fused step from 20 cells,
                                     deposit-only pass from 88): the fixed-point sum of VM_DEPOSIT_FIXED accumulated
                                     with native 32-bit shared-memory atomics (two limbs, exact carry) -- an integer
                                     sum, so the bits do not depend on the order the atomics land in */
    VM_DEPOSIT_ATOMIC = 1,        /* warp-aggregated shared-memory atomics + global fp64 RED */
    VM_DEPOSIT_FIXED = 2          /* order-independent: every contribution w B_j(x) is rounded once to a 64-bit
                                     fixed-point integer (scale 2^S from an exact bound on sum |w|) and all sums --
                                     replicas, CTAs, GPUs -- are integer sums: identical bits for every launch
                                     geometry, deposit layout and number of GPUs (SURVEY 8c KAT 10, 8e).  Needs the
                                     fused finish (n_basis <= 1024) and, across ranks, the peer exchange. */
} vm_deposit_mode;

/* projection!(potential, dist): src/projections/potential.jl:2-22.
 * rhs_i = sum_p w_p B_i(x_p) over this rank's particles (local partial). */
int vm_deposit(vm_field* f, vm_particles* p, int mode);

/* Introspection (no reference counterpart; no device needed): the launch plan the library would choose for a
 * particle pass over a periodic mesh of n_basis cells on a device with sm_count SMs and smem_optin_bytes of
 * opt-in shared memory per CTA (B200: 148, 232448).  pass: 0 = deposit only (vm_deposit), 1 = fused
 * kick+drift+deposit step (vm_vp_run), 2 = drift+deposit prologue. */
typedef struct vm_pass_plan {
    int variant;        /* 0 lane-private replicas, 1 MATCH.ANY grouping, 2 shared atomics, 3 xor-shuffle, 4 bank-sorted queues,
                           5 limb atomics (64-bit fixed point as two 32-bit words, native shared-memory adds with exact carry) */
    int replicas;       /* replica grids per warp (per CTA for variants 2 and 5) */
    int grid, threads;  /* CTAs, threads per CTA */
    int pairs;          /* pairs of particles in flight per thread */
    int max_threads;    /* launch bound of the kernel instantiation (registers per thread = 65536 / max_threads) */
    int gather_copies;  /* copies of the field table in shared memory (16: bank-conflict-free gather) */
    size_t smem_bytes;  /* dynamic shared memory per CTA */
} vm_pass_plan;
int vm_pass_plan_query(int sm_count, size_t smem_optin_bytes, int n_basis, int order, int pass, int deposit_mode,
                       vm_pass_plan* out);
/* PoissonSolvers.update!(potential) (src/models/vlasov_poisson.jl:14),
 * update!(::PoissonField, x, w, t) (src/electric_field.jl:45): all-reduce rhs over
 * the ranks, then solve S phi = rhs - mean(rhs), sum(phi) = 0. */
int vm_field_solve(vm_field* f);
/* energy(::PoissonField) = 1/2 phi' S phi (src/electric_field.jl:47). */
int vm_field_energy(vm_field* f, double* W);
/* efield!(f, e, x) (src/electric_field.jl:43) with ScaledField's 1/chi^2 (:26-29):
 * e_host[p] = -phi'(x_p) * inv_chi2.   e_host may be NULL (result stays on device). */
int vm_gather_E(vm_field* f, vm_particles* p, double* e_host, double inv_chi2);
/* phi'(x) at arbitrary host points (Potential functor phi(x, Derivative(1)),
 * src/models/vlasov_poisson.jl:27,48,65); deriv = 0 evaluates phi itself. */
int vm_field_eval(vm_field* f, const double* x_host, long n, int deriv, double* out_host);

/* s_advection! (src/models/vlasov_poisson.jl:53-58): x += dt * v. */
int vm_vp_drift(vm_particles* p, double dt);
/* s_acceleration! without the potential update (:63-66): v += dt * scale * phi'(x);
 * the reference uses scale = -1 (new API) or -1/chi^2 (legacy). */
int vm_vp_kick(vm_field* f, vm_particles* p, double dt, double scale);

typedef enum vm_run_flags {
    VM_RUN_SPLIT_KICK = 1,    /* two half kicks B(dt/2)B(dt/2) like GeometricIntegrators' Strang
                                 composition (new API); default is one B(dt) (legacy loop)       */
    VM_RUN_FROZEN_FIELD = 2,  /* field_source = model_ics: never re-deposit; reproduces the
                                 frozen-field quirk of the new API                   */
    VM_RUN_ATOMIC_DEPOSIT = 4,/* use VM_DEPOSIT_ATOMIC inside the fused step                     */
    VM_RUN_UNFUSED = 8,       /* separate drift / deposit / solve / kick passes (for A/B tests)  */
    VM_RUN_FIXED_DEPOSIT = 16 /* use VM_DEPOSIT_FIXED inside the step: the whole run is bit-identical for
                                 every launch geometry and GPU count                              */
} vm_run_flags;

/* nsteps Strang steps A(dt/2) B(dt) A(dt/2) with self-consistent field:
 * integrate!(SplittingMethod) (src/methods/splitting.jl:40-43) and the body of
 * integrate_vp! (src/vlasov_poisson.jl:94-115) incl. ScaledField: effective step
 * dt*chi, E/chi^2.  Consecutive steps are fused into one particle pass per step
 * (gather E -> kick -> drift -> deposit).  State is at integer time on return.
 * diag_every > 0: every diag_every-th step (and step 0) append a row
 * [W, K, M, sum_w] to diag_host (save_timestep!, src/vlasov_poisson.jl:58-67);
 * diag_host must hold nsteps/diag_every + 1 rows.  diag_every = 0: no diagnostics. */
int vm_vp_run(vm_field* f, vm_particles* p, double dt, int nsteps, int diag_every, int flags,
              double chi, double* diag_host);

/* integrate_vp! (src/vlasov_poisson.jl:94-115) driven by an ExternalField (src/electric_field.jl:55-77): no
 * deposit, no solve -- at step `it` the coefficients are column round(it*dt / coeff_dt) of the prescribed history
 * coeffs_host (column-major n_basis x ncols: the memory of the Julia matrix `coeffs`, time index 0 first), and
 * x += dt*chi/2 v ; v += dt*chi E(x)/chi^2 ; x += dt*chi/2 v.  The history is uploaded once and stays on the
 * device for the call; on return the field holds the last column used (poisson.phi of the reference).
 * Diagnostics rows [W, K, M, sum_w] as vm_vp_run, W = energy(::ExternalField) of the column in use (:75). */
int vm_vp_run_external(vm_field* f, vm_particles* p, double dt, int nsteps, int diag_every, double chi,
                       const double* coeffs_host, int ncols, double coeff_dt, double* diag_host);

typedef enum vm_vf_flags {
    VM_VF_KEEP_POTENTIAL = 1  /* do not re-deposit / re-solve: evaluate with the field's current coefficients */
} vm_vf_flags;
/* lorentz_force!(zdot, t, z, params) (src/models/vlasov_poisson.jl:23-29), and with VM_VF_KEEP_POTENTIAL the
 * second halves of v_advection! / v_acceleration! (:36-50): update_potential!(model) from the particle state,
 * then xdot = v, vdot = -phi'(x).  The result stays on the device (vdot in the handle's work array, read by
 * vm_vp_rk_run; xdot IS the v array); either host pointer may be NULL -- nothing is then copied. */
int vm_vp_vector_field(vm_field* f, vm_particles* p, int flags, double* xdot_host, double* vdot_host);

/* nsteps classical RK4 steps of zdot = lorentz_force(z) with everything on the device: the unsplit vector field
 * (src/models/vlasov_poisson.jl:23-29) driven the way a generic explicit Runge-Kutta integrator of
 * GeometricIntegrators would drive it -- four update_potential! + gather evaluations per step. */
int vm_vp_rk4_run(vm_field* f, vm_particles* p, double dt, int nsteps);

/* [W, K, M, sum_w] for the current state (deposit + solve + reductions). */
int vm_diagnostics(vm_field* f, vm_particles* p, double chi, double* out4);

/* --------------------------------------------------------------- vspline --
 * Replaces SplineDistribution(1, 1, nknots, order, domain, :Dirichlet)
 * (src/distributions/spline_distribution.jl:23-36): clamped B-splines on
 * LinRange(vmin, vmax, nknots), bc = 1: homogeneous Dirichlet recombination
 * (nknots + order - 4 functions), bc = 0: parent basis (nknots + order - 2). */
int vm_vspline_create(vm_ctx* ctx, double vmin, double vmax, int nknots, int order, int bc,
                      vm_vspline** out);
int vm_vspline_destroy(vm_vspline* s);
int vm_vspline_size(vm_vspline* s);                                  /* length(basis)  */
int vm_vspline_get_coefficients(vm_vspline* s, double* host_nv);      /* sdist.coefficients */
int vm_vspline_set_coefficients(vm_vspline* s, const double* host_nv);
int vm_vspline_get_rhs(vm_vspline* s, double* host_nv);
int vm_vspline_get_mass_matrix(vm_vspline* s, double* host_nv_x_nv);  /* galerkin_matrix   */
/* projection(v, dist, sdist): src/projections/distribution.jl:35-55.
 * Uses the particles' v and w arrays. */
int vm_vproject(vm_vspline* s, vm_particles* p);
/* The same with replacement velocities `v_host` (length N) instead of the particles' own: the `velocities`
 * argument of projection(velocities, dist, sdist) when a user-side integrator passes stage values.  The device
 * particle state is NOT modified (the values live in a scratch array for the call). */
int vm_vproject_at(vm_vspline* s, vm_particles* p, const double* v_host);
int vm_vmoments_at(vm_vspline* s, vm_particles* p, const double* v_host, double* out5, double* A2);
int vm_lb_rhs_at(vm_vspline* s, vm_particles* p, const double* v_host, double nu, int conservative, double* vdot_host);
/* spline.(v), (Derivative(1)*spline).(v) at host points (either output may be NULL). */
int vm_vspline_eval(vm_vspline* s, const double* v_host, long n, double* f_host, double* df_host);
/* compute_f_densities / compute_df_densities (src/projections/density.jl:6-20):
 * out5 = [sum f, sum v f, sum v^2 f, sum f', sum v f'] (unweighted particle sums,
 * all ranks) and A = [A1, A2] of compute_coefficients
 * (src/models/lenard_bernstein_conservative.jl:11-21). */
int vm_vmoments(vm_vspline* s, vm_particles* p, double* out5, double* A2);
/* LB_rhs! (src/models/lenard_bernstein.jl:20-30), conservative = 0, and
 * CLB_rhs! (src/models/lenard_bernstein_conservative.jl:24-36), conservative = 1:
 * projection + (moments) + vdot_p = -nu (f' + (A1 + A2 v) f).
 * vdot_host may be NULL; the result is kept on the device either way. */
int vm_lb_rhs(vm_vspline* s, vm_particles* p, double nu, int conservative, double* vdot_host);
/* nsteps RK438 (3/8-rule) steps of the velocity ODE: the loop of
 * run!(::GeometricIntegrator) (src/methods/geometric_integrator.jl:31-35).
 * diag_every > 0: rows [t, sum_p v_p, sum_p v_p^2] (scripts/
 * lenard_bernstein_conservative.jl:49-50) at step 0 and every diag_every-th step. */
int vm_lb_rk438_run(vm_vspline* s, vm_particles* p, double dt, int nsteps, double nu,
                    int conservative, int diag_every, double* diag_host);

#ifdef __cplusplus
}
#endif
#endif /* VLASOV_B200_H */

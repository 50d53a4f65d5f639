// vm_push.cu -- x-space particle passes of the Vlasov-Poisson step (sm_100a, fp64).
//
//   k_vp_pass<K, VAR, MODE>   one streaming pass over the particle SoA:
//        MODE_DEPOSIT        rhs_i += w_p B_i(x_p)                (projection!, potential.jl:2-22)
//        MODE_DRIFT_DEPOSIT  x += d1 v ; deposit                  (prologue of the fused loop)
//        MODE_PUSH_DEPOSIT   E(x) gather ; v += kick E ; x += d1 v ; x += d2 v ; deposit
//                            (s_acceleration! + 2 x s_advection! + projection! in ONE pass: 40 B/particle)
//   k_vp_push<K>              x += d0 v ; gather ; kick ; x += d1 v ; optional K/M/sum_w sums (no deposit)
//   k_gather<K>               e_p = scale * phi'(x_p)  (or phi(x_p))
//
// Deposition never uses shared-memory atomics on the hot variants (fp64 shared atomics are CAS
// loops): every warp owns private replica grids in shared memory and resolves intra-warp
// collisions by grouping lanes by cell (__match_any_sync) and reducing each group in lane order,
// so the result is bit-reproducible for a fixed launch geometry:
//   VAR_PRIV    32 replicas per warp (one per lane): no collisions possible, plain RMW
//   VAR_MATCH   R < 32 replicas per warp: sort-by-cell segmented reduce inside the warp, leader RMW
//   VAR_ATOMIC  warp-aggregated atomicAdd on a per-CTA grid + global RED flush (A/B reference)
#include "vm_pass.cuh"
#include "vm_pass_bq.cuh"

// ------------------------------------------- kick + drift without deposit ---
template <int K>
__device__ __forceinline__ void push_one(double& xp, double& vp, const PassParams& P, const double* __restrict__ dsh)
{
    if (P.drift0 != 0.0) xp = __dadd_rn(xp, __dmul_rn(P.drift0, vp));
    if (P.kick != 0.0) {
        int b0;
        double xi;
        cell_of<false>(P.map, xp, b0, xi);
        const double dphi = gather_dphi<K>(dsh, b0, xi);
        vp = __dadd_rn(vp, __dmul_rn(P.kick, dphi));
        if (P.kick2 != 0.0) vp = __dadd_rn(vp, __dmul_rn(P.kick2, dphi));
    }
    if (P.drift1 != 0.0) xp = __dadd_rn(xp, __dmul_rn(P.drift1, vp));
}

template <int K>
__global__ void __launch_bounds__(512, 2)
k_vp_push(double* __restrict__ x, double* __restrict__ v, const double* __restrict__ w,
          const double* __restrict__ dcoef, double* __restrict__ out, const PassParams P)
{
    extern __shared__ double smem[];
    const int n = P.map.n;
    double* dsh = smem;
    load_dcoef_ext(dsh, dcoef, n, K - 2 > 0 ? K - 2 : 0);
    __syncthreads();
    constexpr int U = 2;
    const long npairs = P.n >> 1;
    const long stride = (long)gridDim.x * blockDim.x;
    const long gtid = (long)blockIdx.x * blockDim.x + threadIdx.x;
    double s2 = 0.0, s1 = 0.0, s0 = 0.0;
    for (long base = gtid; base < npairs; base += U * stride) {
        double2 cx[U], cv[U], cw[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const long q = base + u * stride;
            if (q < npairs) {
                cx[u] = ld_stream2(x + 2 * q);
                cv[u] = ld_stream2(v + 2 * q);
                if (P.diag) cw[u] = ld_stream2(w + 2 * q);
            }
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const long q = base + u * stride;
            if (q < npairs) {
                push_one<K>(cx[u].x, cv[u].x, P, dsh);
                push_one<K>(cx[u].y, cv[u].y, P, dsh);
                st_stream2(x + 2 * q, cx[u]);
                st_stream2(v + 2 * q, cv[u]);
                if (P.diag) {
                    const double a = cw[u].x * cv[u].x, b = cw[u].y * cv[u].y;
                    s2 = fma(a, cv[u].x, s2); s2 = fma(b, cv[u].y, s2);
                    s1 += a; s1 += b;
                    s0 += cw[u].x; s0 += cw[u].y;
                }
            }
        }
    }
    if ((P.n & 1) && gtid == 0) {
        double xp = x[P.n - 1], vp = v[P.n - 1];
        push_one<K>(xp, vp, P, dsh);
        x[P.n - 1] = xp;
        v[P.n - 1] = vp;
        if (P.diag) {
            const double wp = w[P.n - 1];
            s2 = fma(wp * vp, vp, s2); s1 += wp * vp; s0 += wp;
        }
    }
    if (P.diag) {
        __shared__ double red[3][32];
        const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
        s2 = warp_sum(s2); s1 = warp_sum(s1); s0 = warp_sum(s0);
        if (lane == 0) { red[0][warp] = s2; red[1][warp] = s1; red[2][warp] = s0; }
        __syncthreads();
        if (threadIdx.x < 3) {
            double s = 0.0;
            for (int q = 0; q < nwarps; ++q) s += red[threadIdx.x][q];
            out[(size_t)blockIdx.x * VM_DIAG_COLS + threadIdx.x] = s;
        }
        if (threadIdx.x == 3) out[(size_t)blockIdx.x * VM_DIAG_COLS + 3] = 0.0;
    }
}

// sums sum w v^2, sum w v, sum w of the current state (no push)
__global__ void __launch_bounds__(512, 2)
k_wv_moments(const double* __restrict__ v, const double* __restrict__ w, long n, double* __restrict__ out)
{
    const long stride = (long)gridDim.x * blockDim.x;
    double s2 = 0.0, s1 = 0.0, s0 = 0.0;
    for (long p = (long)blockIdx.x * blockDim.x + threadIdx.x; p < n; p += stride) {
        const double vp = ld_stream(v + p), wp = ld_stream(w + p);
        const double wv = wp * vp;
        s2 = fma(wv, vp, s2);
        s1 += wv;
        s0 += wp;
    }
    __shared__ double red[3][32];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
    s2 = warp_sum(s2); s1 = warp_sum(s1); s0 = warp_sum(s0);
    if (lane == 0) { red[0][warp] = s2; red[1][warp] = s1; red[2][warp] = s0; }
    __syncthreads();
    if (threadIdx.x < 3) {
        double s = 0.0;
        for (int q = 0; q < nwarps; ++q) s += red[threadIdx.x][q];
        out[(size_t)blockIdx.x * VM_DIAG_COLS + threadIdx.x] = s;
    }
    if (threadIdx.x == 3) out[(size_t)blockIdx.x * VM_DIAG_COLS + 3] = 0.0;
}

// ------------------------------------------------------------- gather E -----
// e_p = scale * d^deriv/dx^deriv phi (x_p), deriv in {0, 1}
template <int K>
__global__ void __launch_bounds__(512, 2)
k_gather(const double* __restrict__ x, long np, const double* __restrict__ coef /* phi or dcoef */,
         double* __restrict__ e, CellMap map, double scale, int deriv)
{
    extern __shared__ double smem[];
    const int n = map.n;
    for (int i = threadIdx.x; i < n + K; i += blockDim.x) smem[i] = coef[i < n ? i : i - n];   // periodic extension
    __syncthreads();
    const long stride = (long)gridDim.x * blockDim.x;
    for (long p = (long)blockIdx.x * blockDim.x + threadIdx.x; p < np; p += stride) {
        int b0;
        double xi;
        cell_of<true>(map, x[p], b0, xi);
        double s;
        if (deriv) {
            s = gather_dphi<K>(smem, b0, xi);
        } else {
            double N[K];
            bspline_uniform<K>(xi, N);
            s = 0.0;
#pragma unroll
            for (int j = 0; j < K; ++j) s = fma(N[j], smem[b0 + j], s);
        }
        e[p] = scale * s;
    }
}

__global__ void __launch_bounds__(512, 2) k_drift(double* __restrict__ x, const double* __restrict__ v, long n, double dt)
{
    const long stride = (long)gridDim.x * blockDim.x;
    for (long p = (long)blockIdx.x * blockDim.x + threadIdx.x; p < n; p += stride)
        st_stream(x + p, __dadd_rn(ld_stream(x + p), __dmul_rn(dt, ld_stream(v + p))));
}

// One stage of the classical RK4 scheme for zdot = lorentz_force(z) with low storage: accumulators (ax, av) and
// the next stage state (xs, vs); kx = stage velocity, kv = a (the gathered -phi').
__global__ void __launch_bounds__(512, 2)
k_rk4_stage(double* __restrict__ x, double* __restrict__ v, const double* __restrict__ a, double* __restrict__ xs,
            double* __restrict__ vs, double* __restrict__ ax, double* __restrict__ av, long n, double cdt_next,
            double bdt, int stage)
{
    const long stride = (long)gridDim.x * blockDim.x;
    for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        const double kx = stage == 0 ? v[i] : vs[i], kv = a[i];
        const double sx = (stage == 0 ? 0.0 : ax[i]) + bdt * kx, sv = (stage == 0 ? 0.0 : av[i]) + bdt * kv;
        if (stage < 3) {
            ax[i] = sx; av[i] = sv;
            xs[i] = x[i] + cdt_next * kx;
            vs[i] = v[i] + cdt_next * kv;
        } else {
            x[i] += sx;
            v[i] += sv;
        }
    }
}

// ================================================================ host ======
// one translation unit per spline order (vm_pass_order.cu)
#define VM_DECL_PASS(k)                                                                                                   \
    void vm_launch_pass_k##k(vm_ctx*, int, const DepositPlan&, double*, double*, const double*, const double*, double*, \
                             const PassParams&, const FinishParams&);
VM_DECL_PASS(2) VM_DECL_PASS(3) VM_DECL_PASS(4) VM_DECL_PASS(5) VM_DECL_PASS(6)
#undef VM_DECL_PASS
#define VM_DECL_PASS_BQ(k)                                                                                           \
    void vm_launch_pass_bq_k##k(vm_ctx*, int, const BqPlan&, double*, double*, const double*, const double*, double*, \
                                const PassParams&, const FinishParams&);
VM_DECL_PASS_BQ(2) VM_DECL_PASS_BQ(3) VM_DECL_PASS_BQ(4) VM_DECL_PASS_BQ(5) VM_DECL_PASS_BQ(6)
#undef VM_DECL_PASS_BQ

// Meshes from this size on run the bank-sorted pass (vm_pass_bq.cuh); below it lane-private replicas fit with enough
// warps.  Tuning key "bankq": 0 = this rule (the first size without a lane-private plan), 1 = always (n >= 8), -1 = never (A/B: the round-1 variants).
#define VM_BQ_MIN_N 88

namespace {

void launch_pass_bq(vm_ctx* ctx, int mode, int order, const BqPlan& bp, double* x, double* v, const double* w,
                    const double* dcoef, double* out, const PassParams& P, const FinishParams& F)
{
    switch (order) {
        case 2: vm_launch_pass_bq_k2(ctx, mode, bp, x, v, w, dcoef, out, P, F); break;
        case 3: vm_launch_pass_bq_k3(ctx, mode, bp, x, v, w, dcoef, out, P, F); break;
        case 4: vm_launch_pass_bq_k4(ctx, mode, bp, x, v, w, dcoef, out, P, F); break;
        case 5: vm_launch_pass_bq_k5(ctx, mode, bp, x, v, w, dcoef, out, P, F); break;
        case 6: vm_launch_pass_bq_k6(ctx, mode, bp, x, v, w, dcoef, out, P, F); break;
        default: throw vm_error(VM_ERR_UNSUPPORTED, "spline order must be in 2..6");
    }
}

bool want_bankq(vm_ctx* ctx, int n, int deposit_mode)
{
    if (deposit_mode == VM_DEPOSIT_FIXED) return n >= 8;      // (the caller prefers a shallow lane-private plan when one exists)
    if (deposit_mode != VM_DEPOSIT_DETERMINISTIC || ctx->bankq < 0) return false;
    if (ctx->ctas_per_sm > 0 || ctx->threads_per_cta > 0 || ctx->replicas > 0 || ctx->force_match) return false;   // hand-tuned round-1 variants
    return ctx->bankq > 0 ? n >= 8 : n >= VM_BQ_MIN_N;
}

// Meshes from this size on run the limb-atomic fixed-point pass (VAR_AF, vm_deposit.cuh) in the default deposit mode.
// Its cost does not depend on the mesh size while 32 bank-steered replicas fit (up to 512 cells); the lane-private
// per-warp replicas are only cheaper while two CTAs of 16 warps fit next to them.  Measured, interleaved A/B on one box
// (profiles/r02c_af_variants2_thresholds.jsonl), fused step, fraction of the HBM roofline lane-private vs limb-atomic:
// 16 cells 0.877 vs 0.861, 24 cells 0.803 vs 0.886, 32 cells 0.765 vs 0.878, 40 cells 0.708 vs 0.871.  The deposit-only
// pass (8-16 B/particle: bound by instruction issue in this layout, by the shared-memory data pipe in the lane-private
// one, 0.25-0.29 ms vs 0.23-0.26 ms per 1e8 particles) switches only where no lane-private plan exists.
// Tuning key "af": 0 = this rule, 1 = always (n >= 8), -1 = never (the bank-sorted / round-1 layouts).
#ifndef VM_AF_MIN_N
#define VM_AF_MIN_N 20
#endif
#define VM_AF_MIN_N_DEPOSIT 88
bool want_af(vm_ctx* ctx, int n, int deposit_mode, int pass_mode)
{
    if (ctx->af < 0 || ctx->bankq > 0) return false;
    if (deposit_mode == VM_DEPOSIT_FIXED) return n >= 8;      // (the caller prefers a shallow lane-private plan when one exists)
    if (deposit_mode != VM_DEPOSIT_DETERMINISTIC) return false;
    if (ctx->ctas_per_sm > 0 || ctx->threads_per_cta > 0 || ctx->replicas > 0 || ctx->force_match) return false;   // hand-tuned round-1 variants
    if (ctx->af > 0) return n >= 8;
    return n >= (pass_mode == MODE_DEPOSIT ? VM_AF_MIN_N_DEPOSIT : VM_AF_MIN_N);
}

void launch_pass(vm_ctx* ctx, int mode, int order, const DepositPlan& pl, double* x, double* v, const double* w,
                 const double* dcoef, double* out, const PassParams& P, const FinishParams& F)
{
    switch (order) {
        case 2: vm_launch_pass_k2(ctx, mode, pl, x, v, w, dcoef, out, P, F); break;
        case 3: vm_launch_pass_k3(ctx, mode, pl, x, v, w, dcoef, out, P, F); break;
        case 4: vm_launch_pass_k4(ctx, mode, pl, x, v, w, dcoef, out, P, F); break;
        case 5: vm_launch_pass_k5(ctx, mode, pl, x, v, w, dcoef, out, P, F); break;
        case 6: vm_launch_pass_k6(ctx, mode, pl, x, v, w, dcoef, out, P, F); break;
        default: throw vm_error(VM_ERR_UNSUPPORTED, "spline order must be in 2..6");
    }
}

template <int K>
void launch_push_inst(vm_ctx* ctx, vm_field* f, vm_particles* p, double* out, const PassParams& P, const double* dcoef)
{
    int grid, threads;
    vm_launch_geometry(ctx, &grid, &threads);
    if (threads > 512) threads = 512;
    k_vp_push<K><<<grid, threads, (size_t)(f->n + K) * sizeof(double), ctx->stream>>>(p->x, p->v, p->w, dcoef, out, P);
    VM_LAUNCHED(ctx);
}

// dcoef: derivative-spline coefficients to gather from (default: the field's current ones)
void launch_push(vm_ctx* ctx, vm_field* f, vm_particles* p, double* out, const PassParams& P, const double* dcoef = nullptr)
{
    if (!dcoef) dcoef = f->dcoef;
    switch (f->order) {
        case 2: launch_push_inst<2>(ctx, f, p, out, P, dcoef); break;
        case 3: launch_push_inst<3>(ctx, f, p, out, P, dcoef); break;
        case 4: launch_push_inst<4>(ctx, f, p, out, P, dcoef); break;
        case 5: launch_push_inst<5>(ctx, f, p, out, P, dcoef); break;
        case 6: launch_push_inst<6>(ctx, f, p, out, P, dcoef); break;
        default: throw vm_error(VM_ERR_UNSUPPORTED, "spline order must be in 2..6");
    }
}

template <int K>
void launch_gather_inst(vm_ctx* ctx, vm_field* f, const double* x, long np, double* e, double scale, int deriv)
{
    int grid, threads;
    vm_launch_geometry(ctx, &grid, &threads);
    if (threads > 512) threads = 512;
    long need = (np + threads - 1) / threads;
    if (need < grid) grid = (int)(need > 0 ? need : 1);
    k_gather<K><<<grid, threads, (size_t)(f->n + K) * sizeof(double), ctx->stream>>>(x, np, deriv ? f->dcoef : f->phi, e,
                                                                             f->map, scale, deriv);
    VM_LAUNCHED(ctx);
}

}  // namespace

// internal entry points shared with vm_field.cu ------------------------------------------------
void vm_field_reduce_rows(vm_field* f, const double* rows, int nrows, int ncols, double* out);   // vm_field.cu
void vm_field_solve_local(vm_field* f, bool allreduce);                                            // vm_field.cu
void vm_field_energy_dev(vm_field* f, const double* phi = nullptr);                                // vm_field.cu
void vm_field_ext_upload(vm_field* f, const double* coeffs_host, int ncols);                       // vm_field.cu
void vm_field_ext_select(vm_field* f, int col);                                                    // vm_field.cu
void vm_field_store_diag(vm_field* f, int row, double chi);                                        // vm_field.cu
double* vm_field_wv(vm_field* f);                                                                  // vm_field.cu
double* vm_field_diag_rows(vm_field* f, int rows);                                                 // vm_field.cu

void vm_gather_dev(vm_field* f, const double* x_dev, long np, double* e_dev, double scale, int deriv)
{
    vm_ctx* ctx = f->ctx;
    switch (f->order) {
        case 2: launch_gather_inst<2>(ctx, f, x_dev, np, e_dev, scale, deriv); break;
        case 3: launch_gather_inst<3>(ctx, f, x_dev, np, e_dev, scale, deriv); break;
        case 4: launch_gather_inst<4>(ctx, f, x_dev, np, e_dev, scale, deriv); break;
        case 5: launch_gather_inst<5>(ctx, f, x_dev, np, e_dev, scale, deriv); break;
        case 6: launch_gather_inst<6>(ctx, f, x_dev, np, e_dev, scale, deriv); break;
        default: throw vm_error(VM_ERR_UNSUPPORTED, "spline order must be in 2..6");
    }
}

// One particle pass with deposition; leaves the LOCAL (this rank's) deposit in f->rhs[0..n).
// want_solve: also produce phi/dcoef (all-reduce + replicated solve); on a single GPU with a small
// grid both the reduction and the solve are fused into the pass kernel's last CTA.
// defer_solve: the caller's next operation on this field is another fused pass (vm_vp_run); on meshes above 128 cells
// that pass can then do the solve in its own prologue (pass_presolve) instead of k_poisson_solve here.
static void pass_with_deposit(vm_field* f, vm_particles* p, int pass_mode, int deposit_mode, PassParams P,
                              bool want_solve, bool prof_deposit = false, double* xsrc = nullptr /* deposit-only: positions to use */,
                              bool defer_solve = false)
{
    vm_ctx* ctx = f->ctx;
    const int n = f->n;
    const int ncols = n;   // partial rows hold the grid only (the K/M sums live in their own rows)
    P.map = f->map;
    P.n = p->n;
    P.ncols = ncols;
    P.uw = vm_particles_uniform_weight(p, &P.w0) ? 1 : 0;
    BqPlan bp{};
    DepositPlan pl{};
    bool bq = false, af = false;
    const bool fixed = deposit_mode == VM_DEPOSIT_FIXED;
    PassPlan afp{};
    if (fixed) {
        // order-independent fixed-point accumulation: the shallow lane-private pass where it is the plan (small meshes),
        // limb atomics (or, tuning af = -1 / bankq = 1, the bank-sorted pass) otherwise; same bits in every layout (and
        // for any launch geometry / GPU count)
        const int S = vm_particles_fixed_scale(p);
        P.fixscale = ldexp(1.0, S);
        PassPlan pp{};
        bool priv = false;
        if (ctx->bankq <= 0 && ctx->af <= 0) {
            try {
                pp = plan_pass(ctx, n, f->order, pass_mode, VM_DEPOSIT_DETERMINISTIC);
                const PassTier t = vm_pass_tier(pass_mode, pp.pl.var, pp.pl.threads * (pp.pl.grid / ctx->sm_count), ctx->pairs);
                priv = pp.pl.var == VAR_PRIV && t.max_threads == 1024 && t.pairs == (pass_mode == MODE_DEPOSIT ? 2 : 1);
            } catch (const vm_error&) { priv = false; }
        }
        if (priv) { pl = pp.pl; P.repg = pp.repg ? 1 : 0; }
        else {
            af = want_af(ctx, n, deposit_mode, pass_mode) && plan_af(ctx, n, f->order, pass_mode, &afp);
            if (!af) bq = n >= 8 && plan_bq(ctx, n, f->order, pass_mode, P.uw != 0, &bp);
            if (!af && !bq) throw vm_error(VM_ERR_UNSUPPORTED, "VM_DEPOSIT_FIXED: no fixed-point deposit layout for this mesh size / tuning");
        }
    } else {
        if (want_af(ctx, n, deposit_mode, pass_mode) && plan_af(ctx, n, f->order, pass_mode, &afp)) {
            // the default mode on larger meshes IS the fixed-point sum (bit-reproducible for any geometry, <= 2^-S absolute
            // per contribution); weights without a finite scale keep the fp64 layouts
            const int S = vm_particles_fixed_scale(p);
            if (p->fix_ok) { af = true; P.fixscale = ldexp(1.0, S); }
        }
        if (!af) bq = want_bankq(ctx, n, deposit_mode) && plan_bq(ctx, n, f->order, pass_mode, P.uw != 0, &bp);
    }
    if (af) {          // limb atomics: bank-steered two-limb grids shared by the CTA
        pl = afp.pl;
        P.repg = afp.repg ? 1 : 0;
    } else if (bq) {   // bank-sorted pass: one CTA per SM, one replica grid per warp
        pl.var = VAR_MATCH; pl.rep_log2 = 0; pl.grid = ctx->sm_count; pl.threads = bp.warps * 32; pl.smem = bp.smem;
    } else if (!fixed) {
        const PassPlan pp = plan_pass(ctx, n, f->order, pass_mode, deposit_mode);
        pl = pp.pl;
        P.repg = pp.repg ? 1 : 0;
    }
    P.rep_log2 = pl.rep_log2;
    if (f->solve_pending) {
        // phi / dcoef are one deposit behind: this pass solves in its prologue when it is the limb-atomic fused pass with room
        // for the scratch (its grids) and a CTA per tile; any other pass gets the separate kernel first
        const int tiles = (n + 31) / 32;
        const bool can = af && pass_mode == MODE_PUSH_DEPOSIT && pl.threads >= 256 && pl.grid >= tiles && !ctx->no_fuse &&
                         vm_af_core_doubles(n, f->order, pl.rep_log2) >= (size_t)VM_SOLVE_SCRATCH_DOUBLES(n);
        if (can) {
            f->solve_target += (unsigned)tiles;
            P.ps_tiles = tiles; P.ps_target = f->solve_target; P.ps_count = f->solve_count; P.ps_err = f->solve_count + 1;
            P.ps_rhs = f->rhs; P.ps_G = f->G; P.ps_phi = f->phi; P.ps_dcoef = f->dcoef;
            f->solve_pending = false;
        } else {
            vm_field_solve_local(f, false);
        }
    }
    FinishParams F{};
    F.mode = FINISH_NONE;
    double* out;
    if (pl.var == VAR_ATOMIC) {
        VM_CUDA(cudaMemsetAsync(f->rhs, 0, (size_t)ncols * sizeof(double), ctx->stream));
        out = f->rhs;
    } else {
        const int ngroups = (pl.grid + VM_GROUP_CTAS - 1) / VM_GROUP_CTAS;
        out = vm_partials(ctx, (size_t)(pl.grid + ngroups) * ncols);
        const size_t gdoubles = af ? vm_af_core_doubles(n, f->order, pl.rep_log2)
                                   : ((size_t)(n + f->order - 1) << pl.rep_log2) * (size_t)(pl.threads / 32);
        const bool xchg = want_solve && ctx->nranks > 1 && ctx->peers_connected && n <= VM_X_MAX_N;
        F.ticket = ctx->ticket;
        F.rhs = f->rhs; F.G = f->G; F.phi = f->phi; F.dcoef = f->dcoef; F.inv_h = f->map.inv_h;
        // the one-level finish works with one thread per basis function (hand-tuned CTA shapes may be smaller)
        if (n <= VM_FUSE_MAX_N && !ctx->no_fuse && gdoubles >= (size_t)3 * n + 1 && pl.threads >= n) {
            // single GPU, or all ranks connected through peer memory: reduce + (exchange) + solve in the pass kernel
            F.mode = (want_solve && (ctx->nranks == 1 || xchg)) ? FINISH_REDUCE_SOLVE : FINISH_REDUCE;
        } else if (n <= VM_X_MAX_N && !ctx->no_fuse && gdoubles >= (size_t)n && ngroups <= VM_MAX_GROUPS) {
            F.mode = FINISH_REDUCE;               // two-level reduce (+ exchange); the solve is its own multi-CTA kernel
            F.two_level = 1;
            F.grows = out + (size_t)pl.grid * ncols;
        }
        if (F.mode != FINISH_NONE && xchg) vm_xchg_setup(ctx, F);   // deposit + all-gather over NVLink (+ solve) in one kernel
        if (fixed) {
            // the integer sums are converted to fp64 by the in-kernel finish, after the LAST addition -- across ranks that
            // is the peer exchange (an fp64 NCCL all-reduce of converted partial sums would depend on the rank count)
            if (F.mode == FINISH_NONE) throw vm_error(VM_ERR_UNSUPPORTED, "VM_DEPOSIT_FIXED needs the fused finish (n_basis <= 1024, tuning no_fuse = 0)");
            if (ctx->nranks > 1 && want_solve && !F.xchg) throw vm_error(VM_ERR_UNSUPPORTED, "VM_DEPOSIT_FIXED across ranks needs the peer-memory exchange (vm_ctx_peer_connect)");
            F.fixed = 1;
            F.inv_scale = 1.0 / P.fixscale;
        } else if (af) {
            // default mode on the limb-atomic layout: integer sums as far as the fused finish reaches (CTA rows, and the
            // ranks when the peer exchange runs), one conversion to fp64 at its end; without a fused finish every CTA row
            // is converted by the pass itself
            F.fixed = F.mode != FINISH_NONE ? 1 : 0;
            F.inv_scale = 1.0 / P.fixscale;
        }
    }
    // the dominant kernel of its caller: the fused pass inside vm_vp_run, the deposit pass elsewhere
    const bool prof = (pass_mode == MODE_PUSH_DEPOSIT) || (pass_mode == MODE_DEPOSIT && prof_deposit);
    if (prof) vm_prof_mark(ctx);
    if (bq) launch_pass_bq(ctx, pass_mode, f->order, bp, xsrc ? xsrc : p->x, p->v, p->w, f->dcoef, out, P, F);
    else launch_pass(ctx, pass_mode, f->order, pl, xsrc ? xsrc : p->x, p->v, p->w, f->dcoef, out, P, F);
    if (prof) vm_prof_mark(ctx);
    if (pl.var != VAR_ATOMIC && F.mode == FINISH_NONE) vm_field_reduce_rows(f, out, pl.grid, ncols, f->rhs);
    f->rhs_global = (ctx->nranks == 1) || F.xchg;
    if (want_solve && F.mode != FINISH_REDUCE_SOLVE) {
        // (deferred only when rhs already holds the sum over the ranks and the fused finish produced it: the next pass's
        // prologue then reads it straight away)
        if (defer_solve && f->rhs_global && F.mode == FINISH_REDUCE && af && !ctx->no_presolve) f->solve_pending = true;
        else vm_field_solve_local(f, true);
    }
}

static void wv_moments(vm_field* f, vm_particles* p)
{
    vm_ctx* ctx = f->ctx;
    int grid, threads;
    vm_launch_geometry(ctx, &grid, &threads);
    if (threads > 512) threads = 512;
    double* out = vm_partials(ctx, (size_t)grid * VM_DIAG_COLS);
    k_wv_moments<<<grid, threads, 0, ctx->stream>>>(p->v, p->w, p->n, out);
    VM_LAUNCHED(ctx);
    vm_field_reduce_rows(f, out, grid, VM_DIAG_COLS, vm_field_wv(f));
    vm_allreduce_sum(ctx, vm_field_wv(f), VM_DIAG_COLS);
}

static void check_pair(vm_field* f, vm_particles* p, const char* who)
{
    if (!f || !p) throw vm_error(VM_ERR_INVALID, std::string(who) + ": NULL handle");
    if (f->ctx != p->ctx) throw vm_error(VM_ERR_INVALID, std::string(who) + ": field and particles belong to different contexts");
}

extern "C" {

int vm_pass_plan_query(int sm_count, size_t smem_optin_bytes, int n_basis, int order, int pass, int deposit_mode,
                       vm_pass_plan* out)
{
    vm_ctx* ctx__ = nullptr;
    try {
        VM_REQUIRE(out != nullptr, "vm_pass_plan_query: out is NULL");
        VM_REQUIRE(sm_count >= 1 && smem_optin_bytes >= 16 * 1024, "vm_pass_plan_query: bad device description");
        VM_REQUIRE(n_basis >= 1 && n_basis <= VM_MAX_NBASIS, "vm_pass_plan_query: n_basis out of range");
        VM_REQUIRE(order >= VM_MIN_ORDER && order <= VM_MAX_ORDER, "vm_pass_plan_query: spline order must be in 2..6");
        VM_REQUIRE(pass >= 0 && pass <= 2, "vm_pass_plan_query: pass must be 0 (deposit), 1 (fused step) or 2 (drift + deposit)");
        VM_REQUIRE(deposit_mode == VM_DEPOSIT_DETERMINISTIC || deposit_mode == VM_DEPOSIT_ATOMIC, "vm_pass_plan_query: unknown mode (the fixed-point mode picks between the plans of mode 0 at run time)");
        vm_ctx dev;                      // host-side description only: no CUDA call is made
        dev.sm_count = sm_count;
        dev.smem_optin = smem_optin_bytes;
        const int mode = pass == 0 ? MODE_DEPOSIT : (pass == 1 ? MODE_PUSH_DEPOSIT : MODE_DRIFT_DEPOSIT);
        BqPlan bp{};
        PassPlan afp{};
        if (want_af(&dev, n_basis, deposit_mode, mode) && plan_af(&dev, n_basis, order, mode, &afp)) {
            out->variant = VAR_AF;
            out->replicas = 1 << afp.pl.rep_log2;
            out->grid = afp.pl.grid;
            out->threads = afp.pl.threads;
            out->pairs = mode == MODE_DEPOSIT ? 2 : 1;
            out->max_threads = 1024;
            out->gather_copies = afp.repg ? VM_GATHER_COPIES : 1;
            out->smem_bytes = afp.pl.smem;
            return VM_OK;
        }
        if (want_bankq(&dev, n_basis, deposit_mode) && plan_bq(&dev, n_basis, order, mode, true, &bp)) {
            out->variant = 4;
            out->replicas = 1;
            out->grid = sm_count;
            out->threads = bp.warps * 32;
            out->pairs = 1;
            out->max_threads = bp.warps <= 16 ? 512 : 1024;
            out->gather_copies = mode == MODE_PUSH_DEPOSIT ? (1 << bp.gshift) : 1;
            out->smem_bytes = bp.smem;
            return VM_OK;
        }
        const PassPlan pp = plan_pass(&dev, n_basis, order, mode, deposit_mode);
        const int per_sm = pp.pl.threads * (pp.pl.grid / sm_count);
        const PassTier t = vm_pass_tier(mode, pp.pl.var, per_sm, 0);
        out->variant = pp.pl.var;
        out->replicas = 1 << pp.pl.rep_log2;
        out->grid = pp.pl.grid;
        out->threads = pp.pl.threads;
        out->pairs = t.pairs;
        out->max_threads = t.max_threads;
        out->gather_copies = pp.repg ? VM_GATHER_COPIES : 1;
        out->smem_bytes = pp.pl.smem;
    }
    catch (const vm_error& e) { vm_set_error(ctx__, e.what()); return e.code; }
    return VM_OK;
}

int vm_deposit(vm_field* f, vm_particles* p, int mode)
{
    VM_API_BEGIN(f ? f->ctx : nullptr)
    check_pair(f, p, "vm_deposit");
    VM_REQUIRE(mode == VM_DEPOSIT_DETERMINISTIC || mode == VM_DEPOSIT_ATOMIC || mode == VM_DEPOSIT_FIXED, "vm_deposit: unknown mode");
    PassParams P{};
    pass_with_deposit(f, p, MODE_DEPOSIT, mode, P, false, true);
    VM_API_END
}

int vm_vp_drift(vm_particles* p, double dt)
{
    VM_API_BEGIN(p ? p->ctx : nullptr)
    VM_REQUIRE(p != nullptr, "vm_vp_drift: NULL handle");
    if (p->n > 0) {
        int grid, threads;
        vm_launch_geometry(p->ctx, &grid, &threads);
        if (threads > 512) threads = 512;
        k_drift<<<grid, threads, 0, p->ctx->stream>>>(p->x, p->v, p->n, dt);
        VM_LAUNCHED(p->ctx);
    }
    VM_API_END
}

int vm_vp_kick(vm_field* f, vm_particles* p, double dt, double scale)
{
    VM_API_BEGIN(f ? f->ctx : nullptr)
    check_pair(f, p, "vm_vp_kick");
    PassParams P{};
    P.map = f->map;
    P.n = p->n;
    P.kick = dt * scale;
    launch_push(f->ctx, f, p, nullptr, P);
    VM_API_END
}

int vm_gather_E(vm_field* f, vm_particles* p, double* e_host, double inv_chi2)
{
    VM_API_BEGIN(f ? f->ctx : nullptr)
    check_pair(f, p, "vm_gather_E");
    vm_ctx* ctx = f->ctx;
    if (!p->a) VM_CUDA(cudaMalloc(&p->a, (size_t)(p->n > 0 ? p->n : 1) * sizeof(double)));
    if (p->n > 0) vm_gather_dev(f, p->x, p->n, p->a, -inv_chi2, 1);
    if (e_host) {
        if (p->n > 0) VM_CUDA(cudaMemcpyAsync(e_host, p->a, (size_t)p->n * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
        VM_CUDA(cudaStreamSynchronize(ctx->stream));
    }
    VM_API_END
}

int vm_field_eval(vm_field* f, const double* x_host, long n, int deriv, double* out_host)
{
    VM_API_BEGIN(f ? f->ctx : nullptr)
    VM_REQUIRE(f != nullptr && (n == 0 || (x_host && out_host)), "vm_field_eval: NULL argument");
    VM_REQUIRE(deriv == 0 || deriv == 1, "vm_field_eval: deriv must be 0 or 1");
    if (n > 0) {
        vm_ctx* ctx = f->ctx;
        double* buf = nullptr;
        VM_CUDA(cudaMalloc(&buf, (size_t)2 * n * sizeof(double)));
        try {
            VM_CUDA(cudaMemcpyAsync(buf, x_host, (size_t)n * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
            vm_gather_dev(f, buf, n, buf + n, 1.0, deriv);
            VM_CUDA(cudaMemcpyAsync(out_host, buf + n, (size_t)n * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
            VM_CUDA(cudaStreamSynchronize(ctx->stream));
        } catch (...) { cudaFree(buf); throw; }
        VM_CUDA(cudaFree(buf));
    }
    VM_API_END
}

int vm_diagnostics(vm_field* f, vm_particles* p, double chi, double* out4)
{
    VM_API_BEGIN(f ? f->ctx : nullptr)
    check_pair(f, p, "vm_diagnostics");
    VM_REQUIRE(out4 != nullptr && chi != 0.0, "vm_diagnostics: bad argument");
    vm_ctx* ctx = f->ctx;
    PassParams P{};
    pass_with_deposit(f, p, MODE_DEPOSIT, VM_DEPOSIT_DETERMINISTIC, P, true);
    vm_field_energy_dev(f);
    wv_moments(f, p);
    double* rows = vm_field_diag_rows(f, 1);
    vm_field_store_diag(f, 0, chi);
    double* host = vm_pinned(ctx, 4);
    VM_CUDA(cudaMemcpyAsync(host, rows, 4 * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    VM_CUDA(cudaStreamSynchronize(ctx->stream));
    for (int i = 0; i < 4; ++i) out4[i] = host[i];
    VM_API_END
}

int vm_vp_run(vm_field* f, vm_particles* p, double dt, int nsteps, int diag_every, int flags, double chi,
              double* diag_host)
{
    VM_API_BEGIN(f ? f->ctx : nullptr)
    check_pair(f, p, "vm_vp_run");
    VM_REQUIRE(nsteps >= 0 && diag_every >= 0 && chi != 0.0, "vm_vp_run: bad argument");
    VM_REQUIRE(diag_every == 0 || diag_host != nullptr, "vm_vp_run: diag_host is NULL");
    vm_ctx* ctx = f->ctx;
    const double dte = dt * chi;                 // effective step (src/vlasov_poisson.jl:80)
    const double kick_full = dte * (-1.0 / (chi * chi));   // v += dte * E / chi^2, E = -phi'
    const bool split = (flags & VM_RUN_SPLIT_KICK) != 0;
    const bool frozen = (flags & VM_RUN_FROZEN_FIELD) != 0;
    const bool unfused = (flags & VM_RUN_UNFUSED) != 0;
    const int dmode = (flags & VM_RUN_ATOMIC_DEPOSIT) ? VM_DEPOSIT_ATOMIC
                      : ((flags & VM_RUN_FIXED_DEPOSIT) ? VM_DEPOSIT_FIXED : VM_DEPOSIT_DETERMINISTIC);
    const double k1 = split ? 0.5 * kick_full : kick_full, k2 = split ? 0.5 * kick_full : 0.0;
    const double hd = 0.5 * dte;
    const int nrows = diag_every > 0 ? nsteps / diag_every + 1 : 0;
    double* rows = nrows ? vm_field_diag_rows(f, nrows) : nullptr;
    int grid, threads;
    vm_launch_geometry(ctx, &grid, &threads);
    if (threads > 512) threads = 512;
    int row = 0;

    auto record_diag = [&](bool have_wv) {
        // W needs phi at the current (integer-time) positions: update!(efield, x, w, t) + save_timestep!
        if (!frozen) {
            PassParams D{};
            pass_with_deposit(f, p, MODE_DEPOSIT, dmode, D, true);
        }
        vm_field_energy_dev(f);
        if (!have_wv) wv_moments(f, p);
        vm_field_store_diag(f, row++, chi);
    };

    // a deposit whose solve was left to the next fused pass (solve_pending) must not outlive this call
    struct SolveGuard {
        vm_field* f;
        ~SolveGuard() { try { if (f->solve_pending) vm_field_solve_local(f, false); } catch (...) { f->solve_pending = false; } }
    } solve_guard{f};
    auto fused_step = [&](int s) { return s <= nsteps && !frozen && !unfused && !(diag_every > 0 && (s % diag_every == 0)) && s != nsteps; };
    // a rank with an empty shard still runs every pass: it takes part in the exchanges / all-reduces of the others
    if (diag_every > 0) record_diag(false);
    bool staggered = false;   // true: x holds x^n + dt/2 v^n and f holds phi of those positions
    for (int s = 1; s <= nsteps; ++s) {
        const bool is_diag = diag_every > 0 && (s % diag_every == 0);
        const bool last = (s == nsteps);
        if (frozen) {
            PassParams P{};
            P.map = f->map; P.n = p->n;
            P.drift0 = hd; P.kick = k1; P.kick2 = k2; P.drift1 = hd; P.diag = is_diag;
            double* out = is_diag ? vm_partials(ctx, (size_t)grid * VM_DIAG_COLS) : nullptr;
            launch_push(ctx, f, p, out, P);
            if (is_diag) {
                vm_field_reduce_rows(f, out, grid, VM_DIAG_COLS, vm_field_wv(f));
                vm_allreduce_sum(ctx, vm_field_wv(f), VM_DIAG_COLS);
                record_diag(true);
            }
            continue;
        }
        if (!staggered) {
            if (unfused) {
                k_drift<<<grid, threads, 0, ctx->stream>>>(p->x, p->v, p->n, hd);
                VM_LAUNCHED(ctx);
                PassParams D{};
                pass_with_deposit(f, p, MODE_DEPOSIT, dmode, D, true);
            } else {
                PassParams P{};
                P.drift1 = hd;
                pass_with_deposit(f, p, MODE_DRIFT_DEPOSIT, dmode, P, true, false, nullptr, fused_step(s));
            }
            staggered = true;
        }
        if (is_diag || last || unfused) {
            PassParams P{};
            P.map = f->map; P.n = p->n;
            P.kick = k1; P.kick2 = k2; P.drift1 = hd; P.diag = is_diag;
            double* out = is_diag ? vm_partials(ctx, (size_t)grid * VM_DIAG_COLS) : nullptr;
            launch_push(ctx, f, p, out, P);
            staggered = false;
            if (is_diag) {
                vm_field_reduce_rows(f, out, grid, VM_DIAG_COLS, vm_field_wv(f));
                vm_allreduce_sum(ctx, vm_field_wv(f), VM_DIAG_COLS);
                record_diag(true);
            }
        } else {
            PassParams P{};
            P.kick = k1; P.kick2 = k2; P.drift1 = hd; P.drift2 = hd;
            pass_with_deposit(f, p, MODE_PUSH_DEPOSIT, dmode, P, true, false, nullptr, fused_step(s + 1));
        }
    }
    if (nrows > 0) {
        for (int i = 0; i < nrows * 4; ++i) diag_host[i] = 0.0;
        if (row > 0) {
            double* host = vm_pinned(ctx, (size_t)row * 4);
            VM_CUDA(cudaMemcpyAsync(host, rows, (size_t)row * 4 * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
            VM_CUDA(cudaStreamSynchronize(ctx->stream));
            for (int i = 0; i < row * 4; ++i) diag_host[i] = host[i];
        }
    }
    VM_API_END
}

int vm_vp_run_external(vm_field* f, vm_particles* p, double dt, int nsteps, int diag_every, double chi,
                       const double* coeffs_host, int ncols, double coeff_dt, double* diag_host)
{
    VM_API_BEGIN(f ? f->ctx : nullptr)
    check_pair(f, p, "vm_vp_run_external");
    VM_REQUIRE(nsteps >= 0 && diag_every >= 0 && chi != 0.0, "vm_vp_run_external: bad argument");
    VM_REQUIRE(coeffs_host != nullptr && ncols >= 1 && coeff_dt > 0.0, "vm_vp_run_external: bad coefficient history");
    VM_REQUIRE(diag_every == 0 || diag_host != nullptr, "vm_vp_run_external: diag_host is NULL");
    vm_ctx* ctx = f->ctx;
    // update!(::ExternalField, x, w, t): ts = round(t / dt_coeffs), phi = coeffs[:, ts]  (src/electric_field.jl:66-69;
    // Julia's round: ties to even, as nearbyint in the default rounding mode)
    auto col_of = [&](int it) {
        const double ts = nearbyint((it * dt) / coeff_dt);
        VM_REQUIRE(ts >= 0.0 && ts < (double)ncols, "vm_vp_run_external: time index outside the coefficient history");
        return (int)ts;
    };
    for (int it = 0; it <= nsteps; ++it) (void)col_of(it);          // validate before touching the state
    vm_field_ext_upload(f, coeffs_host, ncols);
    const double dte = dt * chi;                           // effective step (src/vlasov_poisson.jl:80)
    const double kick_full = dte * (-1.0 / (chi * chi));   // v += dte * E / chi^2, E = -phi'
    const double hd = 0.5 * dte;
    const int nrows = diag_every > 0 ? nsteps / diag_every + 1 : 0;
    double* rows = nrows ? vm_field_diag_rows(f, nrows) : nullptr;
    int grid, threads;
    vm_launch_geometry(ctx, &grid, &threads);
    if (threads > 512) threads = 512;
    int row = 0;
    if (diag_every > 0) {                                  // update!(efield, x, w, 0.0); save_timestep!(IC, efield, 1)
        vm_field_energy_dev(f, f->ext_phi + (size_t)col_of(0) * f->n);
        wv_moments(f, p);
        vm_field_store_diag(f, row++, chi);
    }
    int last_col = col_of(0);
    for (int it = 1; it <= nsteps; ++it) {
        const bool is_diag = diag_every > 0 && (it % diag_every == 0);
        const int col = col_of(it);
        last_col = col;
        PassParams P{};
        P.map = f->map; P.n = p->n;
        P.drift0 = hd; P.kick = kick_full; P.drift1 = hd; P.diag = is_diag;
        double* out = is_diag ? vm_partials(ctx, (size_t)grid * VM_DIAG_COLS) : nullptr;
        launch_push(ctx, f, p, out, P, f->ext_dcoef + (size_t)col * f->n);
        if (is_diag) {
            vm_field_reduce_rows(f, out, grid, VM_DIAG_COLS, vm_field_wv(f));
            vm_allreduce_sum(ctx, vm_field_wv(f), VM_DIAG_COLS);
            vm_field_energy_dev(f, f->ext_phi + (size_t)col * f->n);
            vm_field_store_diag(f, row++, chi);
        }
    }
    vm_field_ext_select(f, last_col);                      // poisson.phi of the reference holds the last prescribed column
    if (nrows > 0) {
        double* host = vm_pinned(ctx, (size_t)nrows * 4);
        VM_CUDA(cudaMemcpyAsync(host, rows, (size_t)row * 4 * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
        VM_CUDA(cudaStreamSynchronize(ctx->stream));
        for (int i = 0; i < nrows * 4; ++i) diag_host[i] = i < row * 4 ? host[i] : 0.0;
    }
    VM_API_END
}

int vm_vp_vector_field(vm_field* f, vm_particles* p, int flags, double* xdot_host, double* vdot_host)
{
    VM_API_BEGIN(f ? f->ctx : nullptr)
    check_pair(f, p, "vm_vp_vector_field");
    vm_ctx* ctx = f->ctx;
    if (!(flags & VM_VF_KEEP_POTENTIAL)) {                 // update_potential!(model): projection! + update!
        PassParams D{};
        pass_with_deposit(f, p, MODE_DEPOSIT, VM_DEPOSIT_DETERMINISTIC, D, true);
    }
    if (!p->a) VM_CUDA(cudaMalloc(&p->a, (size_t)(p->n > 0 ? p->n : 1) * sizeof(double)));
    if (p->n > 0) vm_gather_dev(f, p->x, p->n, p->a, -1.0, 1);      // vdot = -phi'(x); xdot is the v array itself
    if (xdot_host && p->n > 0) VM_CUDA(cudaMemcpyAsync(xdot_host, p->v, (size_t)p->n * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    if (vdot_host && p->n > 0) VM_CUDA(cudaMemcpyAsync(vdot_host, p->a, (size_t)p->n * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    if (xdot_host || vdot_host) VM_CUDA(cudaStreamSynchronize(ctx->stream));
    VM_API_END
}

int vm_vp_rk4_run(vm_field* f, vm_particles* p, double dt, int nsteps)
{
    VM_API_BEGIN(f ? f->ctx : nullptr)
    check_pair(f, p, "vm_vp_rk4_run");
    VM_REQUIRE(nsteps >= 0, "vm_vp_rk4_run: bad argument");
    vm_ctx* ctx = f->ctx;
    const size_t bytes = (size_t)(p->n > 0 ? p->n : 1) * sizeof(double);
    for (int i = 0; i < 4; ++i) if (!p->work[i]) VM_CUDA(cudaMalloc(&p->work[i], bytes));
    if (!p->a) VM_CUDA(cudaMalloc(&p->a, bytes));
    double *xs = p->work[0], *vs = p->work[1], *ax = p->work[2], *av = p->work[3];
    int grid, threads;
    vm_launch_geometry(ctx, &grid, &threads);
    if (threads > 512) threads = 512;
    const double c_next[4] = {0.5, 0.5, 1.0, 0.0}, bw[4] = {1.0 / 6.0, 2.0 / 6.0, 2.0 / 6.0, 1.0 / 6.0};
    for (int s = 0; s < nsteps; ++s) {
        for (int st = 0; st < 4; ++st) {
            double* xq = st == 0 ? p->x : xs;                      // lorentz_force! at the stage state:
            PassParams D{};
            pass_with_deposit(f, p, MODE_DEPOSIT, VM_DEPOSIT_DETERMINISTIC, D, true, false, xq);   // update_potential!
            if (p->n > 0) vm_gather_dev(f, xq, p->n, p->a, -1.0, 1);                                // vdot = -phi'(x)
            k_rk4_stage<<<grid, threads, 0, ctx->stream>>>(p->x, p->v, p->a, xs, vs, ax, av, p->n, c_next[st] * dt, bw[st] * dt, st);
            VM_LAUNCHED(ctx);
        }
    }
    VM_API_END
}

}  // extern "C"

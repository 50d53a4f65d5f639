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
#include "vm_internal.cuh"
#include "vm_deposit.cuh"

enum { MODE_DEPOSIT = 0, MODE_PUSH_DEPOSIT = 1, MODE_DRIFT_DEPOSIT = 2 };

struct PassParams {
    CellMap map;
    double kick, kick2;       // v += kick * phi'(x) ; v += kick2 * phi'(x)   (kick2 == 0: skipped)
    double drift0;            // k_vp_push only: x += drift0 * v before the gather
    double drift1, drift2;    // x += drift1 * v ; x += drift2 * v            (drift2 == 0: skipped)
    long n;                   // particles
    int rep_log2;             // log2(replicas per warp (PRIV/MATCH) or per CTA (ATOMIC))
    int ncols;                // row length of the per-CTA partial output (n_basis + VM_DIAG_COLS)
    int diag;                 // k_vp_push: accumulate sum w v^2, sum w v, sum w
    int uw;                   // all particles carry the weight w0: the weight array is not read
    double w0;
};

// ---------------------------------------------------------------- gather ----
// dsh is the derivative-coefficient vector D extended periodically by K-2 entries (dsh[n+i] = D[i]),
// so the K-1 reads at b0 .. b0+K-2 need no index wrap.
template <int K>
__device__ __forceinline__ double gather_dphi(const double* __restrict__ dsh, int b0, double xi)
{
    // phi'(x) = sum_{j<K-1} N^{K-1}_j(xi) * D[(b0 + j) mod n],  D_m = (phi_{m+1} - phi_m) / h
    double Nd[K - 1 > 0 ? K - 1 : 1];
    bspline_uniform<(K - 1 > 0 ? K - 1 : 1)>(xi, Nd);
    const double* d = dsh + b0;
    double s = 0.0;
#pragma unroll
    for (int j = 0; j < K - 1; ++j) s = fma(Nd[j], d[j], s);
    return s;
}

__device__ __forceinline__ void load_dcoef_ext(double* __restrict__ dsh, const double* __restrict__ dcoef, int n, int ext)
{
    for (int i = threadIdx.x; i < n + ext; i += blockDim.x) dsh[i] = dcoef[i < n ? i : i - n];
}

// --------------------------------------------------------- the fused pass ---
// SPLIT: two half kicks (new-API Strang) instead of one.  The fused mode always applies the two
// separately rounded half drifts of consecutive Strang steps (drift1 then drift2).
template <int K, int VAR, int MODE, bool SPLIT, bool POW2>
__device__ __forceinline__ void process(double& xp, double& vp, double wp, bool active, const PassParams& P,
                                        const double* __restrict__ dsh, double* __restrict__ wg, int rep, int lane)
{
    int b0;
    double xi;
    constexpr bool CONV = (MODE == MODE_DEPOSIT);      // see cell_of: measured per mode
    if (MODE == MODE_PUSH_DEPOSIT) {
        cell_of<CONV, POW2>(P.map, xp, b0, xi);
        const double dphi = gather_dphi<K>(dsh, b0, xi);
        // literal (unfused) update order of s_acceleration!: v = v - dt * phi'
        vp = __dadd_rn(vp, __dmul_rn(P.kick, dphi));
        if (SPLIT) vp = __dadd_rn(vp, __dmul_rn(P.kick2, dphi));
    }
    if (MODE != MODE_DEPOSIT) xp = __dadd_rn(xp, __dmul_rn(P.drift1, vp));
    if (MODE == MODE_PUSH_DEPOSIT) xp = __dadd_rn(xp, __dmul_rn(P.drift2, vp));
    cell_of<CONV, POW2>(P.map, xp, b0, xi);
    double val[K];
    bspline_uniform_w<K>(xi, wp, val);                 // inactive lanes carry wp == 0
    scatter<K, VAR>(wg, P.rep_log2, rep, lane, b0, val, active);
}

template <int MODE>
struct PairBuf {
    double2 x, v, w;
};

// U = pairs of particles each thread keeps in flight per (half-)iteration, loads one iteration ahead.
// Pair indices are 32-bit (N < 2^32 particles per GPU).
template <int K, int VAR, int MODE, int U, bool SPLIT, bool POW2>
__global__ void __launch_bounds__(1024, 1)
k_vp_pass(double* __restrict__ x, double* __restrict__ v, const double* __restrict__ w,
          const double* __restrict__ dcoef, double* __restrict__ out, const PassParams P, const FinishParams F)
{
    extern __shared__ double smem[];
    const int n = P.map.n;
    constexpr int GHOST = K - 1;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
    double* dsh = smem;
    double* grid = smem + (MODE == MODE_PUSH_DEPOSIT ? n + K : 0);
    const int gsz = (n + GHOST) << P.rep_log2;
    const int gtotal = (VAR == VAR_ATOMIC) ? gsz : gsz * nwarps;
    double* scratch = grid + gtotal;
    for (int i = threadIdx.x; i < gtotal; i += blockDim.x) grid[i] = 0.0;
    // Programmatic dependent launch: everything above overlaps the previous kernel's tail (its last
    // CTA is still reducing / exchanging / solving); nothing it wrote is read before this point.
    asm volatile("griddepcontrol.wait;" ::: "memory");
    if (MODE == MODE_PUSH_DEPOSIT) load_dcoef_ext(dsh, dcoef, n, K - 2 > 0 ? K - 2 : 0);
    __syncthreads();
    double* wg = (VAR == VAR_ATOMIC) ? grid : grid + warp * gsz;
    const int rep = ((VAR == VAR_ATOMIC) ? warp : lane) & ((1 << P.rep_log2) - 1);

    const unsigned npairs = (unsigned)(P.n >> 1);
    const unsigned stride = gridDim.x * blockDim.x;
    const unsigned gtid = blockIdx.x * blockDim.x + threadIdx.x;
    const unsigned chunk = U * stride;
    const unsigned iters = (npairs + chunk - 1) / chunk;   // uniform trip count: the scatter is warp-collective

    PairBuf<MODE> A[U], B[U];
    auto load = [&](PairBuf<MODE> (&buf)[U], unsigned q0) {   // q0 = first pair of this thread's slice
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const unsigned q = q0 + u * stride;
            buf[u].w = make_double2(0., 0.);            // out-of-range pairs deposit nothing
            if (q < npairs) {
                buf[u].x = ld_stream2(x + 2 * (size_t)q);
                if (MODE != MODE_DEPOSIT) buf[u].v = ld_stream2(v + 2 * (size_t)q);
                buf[u].w = P.uw ? make_double2(P.w0, P.w0) : ld_stream2(w + 2 * (size_t)q);
            }
        }
    };
    auto work = [&](PairBuf<MODE> (&buf)[U], unsigned q0) {
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const unsigned q = q0 + u * stride;
            const bool active = q < npairs;
            process<K, VAR, MODE, SPLIT, POW2>(buf[u].x.x, buf[u].v.x, buf[u].w.x, active, P, dsh, wg, rep, lane);
            process<K, VAR, MODE, SPLIT, POW2>(buf[u].x.y, buf[u].v.y, buf[u].w.y, active, P, dsh, wg, rep, lane);
            if (active && MODE != MODE_DEPOSIT) {
                st_stream2(x + 2 * (size_t)q, buf[u].x);
                if (MODE == MODE_PUSH_DEPOSIT) st_stream2(v + 2 * (size_t)q, buf[u].v);
            }
        }
    };
#pragma unroll
    for (int u = 0; u < U; ++u) A[u].x = A[u].v = B[u].x = B[u].v = make_double2(0., 0.);
    // q advances by `chunk` per iteration; npairs + 3*chunk < 2^32 is guaranteed by vm_particles_create
    unsigned q = gtid;
    load(A, q);
    if (MODE == MODE_DEPOSIT) {
        // 16 B/particle pass, issue-bound: unrolled twice over the two buffer sets (no register moves)
        for (unsigned it = 0; it < iters; it += 2, q += 2 * chunk) {
            load(B, q + chunk);
            work(A, q);
            load(A, q + 2 * chunk);
            work(B, q + chunk);    // all-inactive when iters is odd (costs one idle half-iteration)
        }
    } else {
        // 32-40 B/particle passes, HBM-bound: keeping the next pair's loads at the very top of the
        // iteration measured 5 % faster than the unrolled form (ptxas sinks the loads otherwise)
        for (unsigned it = 0; it < iters; ++it, q += chunk) {
            load(B, q + chunk);
            work(A, q);
#pragma unroll
            for (int u = 0; u < U; ++u) A[u] = B[u];
        }
    }
    // let the next kernel of the stream start its prologue while this grid drains and finishes
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    if ((P.n & 1) && blockIdx.x == 0 && warp == 0) {   // odd particle count: last particle, lane 0 of one warp
        const bool active = (lane == 0);
        double xp = 0., vp = 0., wp = 0.;
        if (active) {
            xp = x[P.n - 1];
            if (MODE != MODE_DEPOSIT) vp = v[P.n - 1];
            wp = P.uw ? P.w0 : w[P.n - 1];
        }
        process<K, VAR, MODE, SPLIT, POW2>(xp, vp, wp, active, P, dsh, wg, rep, lane);
        if (active && MODE != MODE_DEPOSIT) {
            x[P.n - 1] = xp;
            if (MODE == MODE_PUSH_DEPOSIT) v[P.n - 1] = vp;
        }
    }
    flush_grid<VAR>(grid, scratch, out, n, GHOST, P.rep_log2, nwarps, P.ncols);
    if (VAR != VAR_ATOMIC && F.mode != FINISH_NONE) finish_last_cta(F, out, gridDim.x, n, grid, scratch);
}

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

// ================================================================ host ======
namespace {

template <int K, int VAR, int MODE, bool SPLIT, bool POW2>
void launch_pass_inst(vm_ctx* ctx, const DepositPlan& pl, double* x, double* v, const double* w,
                      const double* dcoef, double* out, const PassParams& P, const FinishParams& F)
{
    constexpr int U = (MODE == MODE_DEPOSIT) ? 2 : 1;
    static size_t configured[64] = {};   // per device: max dynamic smem already opted into for this instantiation
    size_t& conf = configured[ctx->device & 63];
    if (pl.smem > conf) {
        VM_CUDA(cudaFuncSetAttribute(k_vp_pass<K, VAR, MODE, U, SPLIT, POW2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)pl.smem));
        conf = pl.smem;
    }
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(pl.grid);
    cfg.blockDim = dim3(pl.threads);
    cfg.dynamicSmemBytes = pl.smem;
    cfg.stream = ctx->stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;   // PDL: see griddepcontrol.wait in the kernel
    attr[0].val.programmaticStreamSerializationAllowed = ctx->no_pdl ? 0 : 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    VM_CUDA(cudaLaunchKernelEx(&cfg, k_vp_pass<K, VAR, MODE, U, SPLIT, POW2>, x, v, w, dcoef, out, P, F));
    ++ctx->launches;
}

template <int K, int MODE>
void launch_pass_var(vm_ctx* ctx, const DepositPlan& pl, double* x, double* v, const double* w,
                     const double* dcoef, double* out, const PassParams& P, const FinishParams& F)
{
    const bool split = (MODE == MODE_PUSH_DEPOSIT) && P.kick2 != 0.0;
    // the mask form of the periodic wrap is only specialised for the lane-private variant (small grids)
    const bool pow2 = P.map.mask >= 0;
#define VM_PASS_VAR(V, PW)                                                                               \
    if (MODE == MODE_PUSH_DEPOSIT && split) launch_pass_inst<K, V, MODE, (MODE == MODE_PUSH_DEPOSIT), PW>(ctx, pl, x, v, w, dcoef, out, P, F); \
    else launch_pass_inst<K, V, MODE, false, PW>(ctx, pl, x, v, w, dcoef, out, P, F)
    switch (pl.var) {
        case VAR_PRIV:
            if (pow2) { VM_PASS_VAR(VAR_PRIV, true); } else { VM_PASS_VAR(VAR_PRIV, false); }
            break;
        case VAR_MATCH: VM_PASS_VAR(VAR_MATCH, false); break;
        case VAR_XOR: VM_PASS_VAR(VAR_XOR, false); break;
        default: VM_PASS_VAR(VAR_ATOMIC, false); break;
    }
#undef VM_PASS_VAR
}

template <int MODE>
void launch_pass(vm_ctx* ctx, int order, const DepositPlan& pl, double* x, double* v, const double* w,
                 const double* dcoef, double* out, const PassParams& P, const FinishParams& F)
{
    switch (order) {
        case 2: launch_pass_var<2, MODE>(ctx, pl, x, v, w, dcoef, out, P, F); break;
        case 3: launch_pass_var<3, MODE>(ctx, pl, x, v, w, dcoef, out, P, F); break;
        case 4: launch_pass_var<4, MODE>(ctx, pl, x, v, w, dcoef, out, P, F); break;
        case 5: launch_pass_var<5, MODE>(ctx, pl, x, v, w, dcoef, out, P, F); break;
        case 6: launch_pass_var<6, MODE>(ctx, pl, x, v, w, dcoef, out, P, F); break;
        default: throw vm_error(VM_ERR_UNSUPPORTED, "spline order must be in 2..6");
    }
}

template <int K>
void launch_push_inst(vm_ctx* ctx, vm_field* f, vm_particles* p, double* out, const PassParams& P)
{
    int grid, threads;
    vm_launch_geometry(ctx, &grid, &threads);
    if (threads > 512) threads = 512;
    k_vp_push<K><<<grid, threads, (size_t)(f->n + K) * sizeof(double), ctx->stream>>>(p->x, p->v, p->w, f->dcoef, out, P);
    VM_LAUNCHED(ctx);
}

void launch_push(vm_ctx* ctx, vm_field* f, vm_particles* p, double* out, const PassParams& P)
{
    switch (f->order) {
        case 2: launch_push_inst<2>(ctx, f, p, out, P); break;
        case 3: launch_push_inst<3>(ctx, f, p, out, P); break;
        case 4: launch_push_inst<4>(ctx, f, p, out, P); break;
        case 5: launch_push_inst<5>(ctx, f, p, out, P); break;
        case 6: launch_push_inst<6>(ctx, f, p, out, P); break;
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
void vm_field_energy_dev(vm_field* f);                                                             // vm_field.cu
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
static void pass_with_deposit(vm_field* f, vm_particles* p, int pass_mode, int deposit_mode, PassParams P,
                              bool want_solve, bool prof_deposit = false)
{
    vm_ctx* ctx = f->ctx;
    const int n = f->n;
    const int ncols = n;   // partial rows hold the grid only (the K/M sums live in their own rows)
    // the 32-40 B/particle passes need more resident warps than the deposit-only pass (measured)
    DepositPlan pl = plan_deposit(ctx, n, f->order - 1, pass_mode == MODE_PUSH_DEPOSIT ? n + f->order : 0, deposit_mode,
                                  pass_mode == MODE_DEPOSIT ? VM_PRIV_MIN_WARPS : 12);
    P.map = f->map;
    P.n = p->n;
    P.rep_log2 = pl.rep_log2;
    P.ncols = ncols;
    P.uw = vm_particles_uniform_weight(p, &P.w0) ? 1 : 0;
    FinishParams F{};
    F.mode = FINISH_NONE;
    double* out;
    if (pl.var == VAR_ATOMIC) {
        VM_CUDA(cudaMemsetAsync(f->rhs, 0, (size_t)ncols * sizeof(double), ctx->stream));
        out = f->rhs;
    } else {
        out = vm_partials(ctx, (size_t)pl.grid * ncols);
        const size_t gdoubles = ((size_t)(n + f->order - 1) << pl.rep_log2) * (size_t)(pl.var == VAR_ATOMIC ? 1 : pl.threads / 32);
        if (n <= VM_FUSE_MAX_N && !ctx->no_fuse && gdoubles >= (size_t)3 * n + 1) {
            F.mode = (want_solve && ctx->nranks == 1) ? FINISH_REDUCE_SOLVE : FINISH_REDUCE;
            F.ticket = ctx->ticket;
            F.rhs = f->rhs; F.G = f->G; F.phi = f->phi; F.dcoef = f->dcoef; F.inv_h = f->map.inv_h;
            if (want_solve && ctx->nranks > 1 && ctx->peers_connected && n <= VM_XSLOT) {
                F.mode = FINISH_EXCHANGE_SOLVE;       // deposit + all-gather over NVLink + solve in one kernel
                F.nranks = ctx->nranks; F.rank = ctx->rank; F.seq = ++ctx->xseq;
                F.inbox = ctx->inbox; F.err = ctx->xerr;
                for (int r = 0; r < ctx->nranks; ++r) F.peer[r] = ctx->peer_inbox[r];
            }
        }
    }
    // the dominant kernel of its caller: the fused pass inside vm_vp_run, the deposit pass elsewhere
    const bool prof = (pass_mode == MODE_PUSH_DEPOSIT) || (pass_mode == MODE_DEPOSIT && prof_deposit);
    if (prof) vm_prof_mark(ctx);
    switch (pass_mode) {
        case MODE_DEPOSIT: launch_pass<MODE_DEPOSIT>(ctx, f->order, pl, p->x, p->v, p->w, f->dcoef, out, P, F); break;
        case MODE_PUSH_DEPOSIT: launch_pass<MODE_PUSH_DEPOSIT>(ctx, f->order, pl, p->x, p->v, p->w, f->dcoef, out, P, F); break;
        default: launch_pass<MODE_DRIFT_DEPOSIT>(ctx, f->order, pl, p->x, p->v, p->w, f->dcoef, out, P, F); break;
    }
    if (prof) vm_prof_mark(ctx);
    if (pl.var != VAR_ATOMIC && F.mode == FINISH_NONE) vm_field_reduce_rows(f, out, pl.grid, ncols, f->rhs);
    if (want_solve && F.mode != FINISH_REDUCE_SOLVE && F.mode != FINISH_EXCHANGE_SOLVE) vm_field_solve_local(f, true);
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

int vm_deposit(vm_field* f, vm_particles* p, int mode)
{
    VM_API_BEGIN(f ? f->ctx : nullptr)
    check_pair(f, p, "vm_deposit");
    VM_REQUIRE(mode == VM_DEPOSIT_DETERMINISTIC || mode == VM_DEPOSIT_ATOMIC, "vm_deposit: unknown mode");
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
    const int dmode = (flags & VM_RUN_ATOMIC_DEPOSIT) ? VM_DEPOSIT_ATOMIC : VM_DEPOSIT_DETERMINISTIC;
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

    if (p->n == 0) nsteps = 0;
    if (diag_every > 0 && p->n > 0) record_diag(false);
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
                pass_with_deposit(f, p, MODE_DRIFT_DEPOSIT, dmode, P, true);
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
            pass_with_deposit(f, p, MODE_PUSH_DEPOSIT, dmode, P, true);
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

}  // extern "C"

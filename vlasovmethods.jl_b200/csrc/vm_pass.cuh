// vm_pass.cuh -- the fused x-space particle pass (kernel template + launch machinery).
//
// Included by vm_pass_order.cu, which is compiled once per spline order (-DVM_PASS_ORDER=k: the
// instantiations of one order per translation unit, so the orders build in parallel), and by vm_push.cu
// for the pieces the other kernels share (PassParams, the field gather).
#pragma once
#include "vm_internal.cuh"
#include "vm_deposit.cuh"

enum { MODE_DEPOSIT = 0, MODE_PUSH_DEPOSIT = 1, MODE_DRIFT_DEPOSIT = 2 };
#ifndef VM_PASS_EARLY_LOAD
#define VM_PASS_EARLY_LOAD 0     // 1: first batch of particle loads before griddepcontrol.wait -- measured neutral (8.49 vs 8.43 us fixed cost, profiles/r02_early_load_ab.txt), off
#endif

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
    int repg;                 // fused pass: the gather table is stored VM_GATHER_COPIES times (conflict-free reads)
    double w0;
    double fixscale;          // VM_DEPOSIT_FIXED: 2^S (contributions are accumulated as 64-bit integers); 0: fp64 accumulation
    // presolve (fused limb-atomic pass on meshes above 128 cells): the Poisson solve of the PREVIOUS pass's deposit is
    // done by the first ps_tiles CTAs of this grid before the particle loop (0: dcoef is already current)
    int ps_tiles;
    unsigned ps_target;       // value of *ps_count once every tile is done
    unsigned* ps_count;
    unsigned* ps_err;
    const double *ps_rhs, *ps_G;
    double *ps_phi, *ps_dcoef;
};

// ---------------------------------------------------------------- gather ----
// dsh is the derivative-coefficient vector D extended periodically by K-2 entries (dsh[n+i] = D[i]),
// so the K-1 reads at b0 .. b0+K-2 need no index wrap.
//
// REPG: the table is stored 16 times, entry m of copy c at dsh[m*16 + c], and lane l reads copy l & 15.  On
// meshes with more than 16 cells two lanes of a half-warp often sit in cells whose table entries share a bank
// (ncu, n_h = 32: 25 % of all shared-memory wavefronts of the fused pass were such conflicts, and the pass is
// co-limited by the shared-memory data pipe); with one copy per bank pair every gather is conflict-free.
#define VM_GATHER_COPIES 16
#define VM_GATHER_TABLE_MAX_BYTES (45 * 1024)   // beyond this the copies cost more replica-grid warps than they save
template <int K, bool REPG = false>
__device__ __forceinline__ double gather_dphi(const double* __restrict__ dsh, int b0, double xi)
{
    // phi'(x) = sum_{j<K-1} N^{K-1}_j(xi) * D[(b0 + j) mod n],  D_m = (phi_{m+1} - phi_m) / h
    double Nd[K - 1 > 0 ? K - 1 : 1];
    bspline_uniform<(K - 1 > 0 ? K - 1 : 1)>(xi, Nd);
    constexpr int ST = REPG ? VM_GATHER_COPIES : 1;
    const double* d = REPG ? dsh + b0 * VM_GATHER_COPIES + (threadIdx.x & (VM_GATHER_COPIES - 1)) : dsh + b0;
    double s = 0.0;
#pragma unroll
    for (int j = 0; j < K - 1; ++j) s = fma(Nd[j], d[j * ST], s);
    return s;
}

template <bool REPG = false>
__device__ __forceinline__ void load_dcoef_ext(double* __restrict__ dsh, const double* __restrict__ dcoef, int n, int ext)
{
    if (REPG) {
        for (int i = threadIdx.x; i < (n + ext) * VM_GATHER_COPIES; i += blockDim.x) {
            const int m = i / VM_GATHER_COPIES;
            dsh[i] = dcoef[m < n ? m : m - n];
        }
    } else {
        for (int i = threadIdx.x; i < n + ext; i += blockDim.x) dsh[i] = dcoef[i < n ? i : i - n];
    }
}

// doubles of dynamic shared memory the gather table of the fused pass takes
inline size_t vm_gather_table_doubles(int n, int order, bool repg) { return (size_t)(n + order) * (repg ? VM_GATHER_COPIES : 1); }

// --------------------------------------------------------- the fused pass ---
// SPLIT: two half kicks (new-API Strang) instead of one.  The fused mode always applies the two
// separately rounded half drifts of consecutive Strang steps (drift1 then drift2).
// Phase A of a particle: everything up to the deposit weights (reads the read-only dcoef table only).
// Limb-atomic pass: with conflict-free atomics it is bound by instruction issue at the sustained (power-capped) SM clock,
// so it takes the cheaper forms the HBM-bound lane-private pass does not need.  One box, 64 / 256 / 1024 cells, fraction
// of the HBM roofline (profiles/r02c_af_variants.txt): as first written 0.763 / 0.756 / 0.733 (255 instructions per pair
// of particles); canonical xi without the fmin 0.770 / 0.765 / 0.739 (236); + FRND/F2I cell lookup 0.787 / 0.780 / 0.748
// (212); + main loop unrolled over the two buffer sets (no register moves) 0.851 / 0.839 / 0.78 (196).
#ifndef VM_AF_CONV
#define VM_AF_CONV 1
#endif
#ifndef VM_AF_UNROLL
#define VM_AF_UNROLL 1
#endif
#ifndef VM_PRIV_CONV
#define VM_PRIV_CONV 0          // 1: FRND/F2I cell lookup in the fused lane-private pass too (A/B)
#endif
#ifndef VM_PRIV_UNROLL
#define VM_PRIV_UNROLL 0        // 1: unrolled main loop in the shallow fused lane-private pass too (A/B)
#endif
#ifndef VM_AF_THREADS
#define VM_AF_THREADS 1024      // resident threads per SM of the limb-atomic pass (A/B: 768 = 85 registers per thread)
#endif
#ifndef VM_AF_PAIRS
#define VM_AF_PAIRS 1           // pairs of particles in flight per thread in its fused / drift modes (A/B)
#endif
// The canonical xi of the fixed-point layouts is (xi + 1) - 1.  (The bank-sorted pass clamps xi below 1 first because it
// stores the mantissa of xi + 1; for the value computed here the clamp changes nothing: it only bites at xi == 1, where
// both forms give 1.)
#ifndef VM_FIX_FMIN
#define VM_FIX_FMIN 0
#endif
template <int K, int MODE, bool SPLIT, bool POW2, bool REPG, bool FIXED = false, bool AF = false>
__device__ __forceinline__ void prepare(double& xp, double& vp, double wp, const PassParams& P,
                                        const double* __restrict__ dsh, int& b0, double (&val)[K])
{
    double xi;
    constexpr bool CONV = (MODE == MODE_DEPOSIT) || (VM_AF_CONV && AF) || VM_PRIV_CONV;      // see cell_of: measured per mode
    if (MODE == MODE_PUSH_DEPOSIT) {
        cell_of<CONV, POW2>(P.map, xp, b0, xi);
        const double dphi = gather_dphi<K, REPG>(dsh, b0, xi);
        // literal (unfused) update order of s_acceleration!: v = v - dt * phi'
        vp = __dadd_rn(vp, __dmul_rn(P.kick, dphi));
        if (SPLIT) vp = __dadd_rn(vp, __dmul_rn(P.kick2, dphi));
    }
    if (MODE != MODE_DEPOSIT) xp = __dadd_rn(xp, __dmul_rn(P.drift1, vp));
    if (MODE == MODE_PUSH_DEPOSIT) xp = __dadd_rn(xp, __dmul_rn(P.drift2, vp));
    cell_of<CONV, POW2>(P.map, xp, b0, xi);
    if (FIXED) {
        // the fixed-point deposit must round the SAME per-particle numbers in every layout: the bank-sorted pass
        // carries xi as the 52-bit mantissa of xi + 1 (vm_pass_bq.cuh: bq_pack), so that is the canonical xi
#if VM_FIX_FMIN
        xi = (fmin(xi, 0x1.fffffffffffffp-1) + 1.0) - 1.0;
#else
        xi = (xi + 1.0) - 1.0;
#endif
    }
    bspline_uniform_w<K>(xi, wp, val);                 // inactive lanes carry wp == 0
}

template <int K, int VAR, int MODE, bool SPLIT, bool POW2, bool REPG, bool FIXED = false>
__device__ __forceinline__ void process(double& xp, double& vp, double wp, bool active, const PassParams& P,
                                        const double* __restrict__ dsh, double* __restrict__ wg, int rep, int lane)
{
    int b0;
    double val[K];
    prepare<K, MODE, SPLIT, POW2, REPG, FIXED, VAR == VAR_AF>(xp, vp, wp, P, dsh, b0, val);
    scatter<K, VAR, FIXED>(wg, P.rep_log2, rep, lane, b0, val, active, P.fixscale);
}

// The solve between two fused passes without a kernel of its own (meshes above 128 cells; up to 128 the last CTA of the
// depositing pass solves).  All CTAs of this grid are resident (one per SM), so the first ps_tiles of them each solve one
// tile of 32 outputs -- the arithmetic of k_poisson_solve, same bits -- publish phi / dcoef, and bump a counter that every
// CTA waits for before it loads its gather table (with ld.cg: the table was written during this kernel).  scr: shared
// memory scratch of VM_SOLVE_SCRATCH_DOUBLES(n) doubles, zero on entry and zeroed again on exit (the replica grids).
// Removes two launch boundaries per step: 14.4 -> see profiles/r02c_fixed_cost_per_step.jsonl.
__device__ __forceinline__ void pass_presolve(const PassParams& P, double* __restrict__ scr)
{
    const int n = P.map.n;
    if ((int)blockIdx.x < P.ps_tiles) {
        poisson_solve_tile((int)blockIdx.x, P.ps_rhs, P.ps_G, n, P.map.inv_h, P.ps_phi, P.ps_dcoef, scr, scr + n, scr + n + 2 * 8 * 33);
        __threadfence();
        __syncthreads();
        if (threadIdx.x == 0) atomicAdd(P.ps_count, 1u);
        for (int i = threadIdx.x; i < VM_SOLVE_SCRATCH_DOUBLES(n); i += blockDim.x) scr[i] = 0.0;
    }
    if (threadIdx.x == 0) {
        const long long t0 = clock64();
        unsigned c;
        do {
            asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(c) : "l"(P.ps_count) : "memory");
            if ((int)(c - P.ps_target) < 0 && clock64() - t0 > (1ll << 33)) { atomicExch(P.ps_err, 1u); break; }   // ~4 s: give up, host reports
        } while ((int)(c - P.ps_target) < 0);
    }
    __syncthreads();
}

template <int MODE>
struct PairBuf {
    double2 x, v, w;
};

// U = pairs of particles each thread keeps in flight per (half-)iteration, loads one iteration ahead.
// MAXT = largest CTA this instantiation is launched with: the lane-private replica grids of mid-size
// meshes leave room for few warps per SM (12 at n_h = 64), so those launches get the register budget of
// the absent warps (65536 / MAXT per thread) and spend it on more pairs in flight -- bytes in flight per
// SM, not resident warps, is what hides the HBM latency.
// Pair indices are 32-bit (N < 2^32 particles per GPU).
template <int K, int VAR, int MODE, int U, bool SPLIT, bool POW2, int MAXT, bool REPG, bool FIXED = false>
__global__ void __launch_bounds__(MAXT, 1)
k_vp_pass(double* __restrict__ x, double* __restrict__ v, const double* __restrict__ w,
          const double* __restrict__ dcoef, double* __restrict__ out, const PassParams P, const FinishParams F)
{
    extern __shared__ double smem[];
    const int n = P.map.n;
    constexpr int GHOST = K - 1;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
    double* dsh = smem;
    double* grid = smem + (MODE == MODE_PUSH_DEPOSIT ? (n + K) * (REPG ? VM_GATHER_COPIES : 1) : 0);
    // VAR_AF: R = 2^rep_log2 bank-steered replicas of (pitch) lo words, then as many hi words: R * pitch doubles in all
    const int pitch = vm_af_pitch(n + GHOST, P.rep_log2);
    const int gsz = (VAR == VAR_AF) ? (pitch << P.rep_log2) : ((n + GHOST) << P.rep_log2);
    const int gtotal = (VAR == VAR_ATOMIC || VAR == VAR_AF) ? gsz : gsz * nwarps;
    // (VAR_AF: the finish needs 3n + 1 doubles of work area where the limb arrays were, see vm_af_core_doubles)
    double* scratch = grid + ((VAR == VAR_AF && 3 * n + 2 > gtotal) ? 3 * n + 2 : gtotal);
    for (int i = threadIdx.x; i < gtotal; i += blockDim.x) grid[i] = 0.0;
    double* wg = (VAR == VAR_ATOMIC || VAR == VAR_AF) ? grid : grid + warp * gsz;
    const int rep = (VAR == VAR_AF) ? pitch : (((VAR == VAR_ATOMIC) ? warp : lane) & ((1 << P.rep_log2) - 1));

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
    // Deep tiers (U >= 4, few resident warps) run two phases per batch of 2U particles: the compiler must assume
    // that a replica-grid store may alias the next particle's read of the field table (both live in the dynamic
    // shared array), so the particle-after-particle form serialises the dependency chains of consecutive
    // particles -- which only many resident warps can hide.  Phase A (gather, push, cell, weights: reads only)
    // is free to interleave all 2U particles; phase B applies the read-modify-writes in particle order, so the
    // bits are those of the sequential form.  Measured: -6 % at n_h = 64, -10 % at n_h = 128; with 64 registers
    // (full CTAs) and in the warp-collective collision variants the batch form is slower and not used.
    auto work = [&](PairBuf<MODE> (&buf)[U], unsigned q0) {
        if constexpr (VAR == VAR_PRIV && U >= 4) {
            int b0[2 * U];
            double val[2 * U][K];
#pragma unroll
            for (int u = 0; u < U; ++u) {
                prepare<K, MODE, SPLIT, POW2, REPG, FIXED>(buf[u].x.x, buf[u].v.x, buf[u].w.x, P, dsh, b0[2 * u], val[2 * u]);
                prepare<K, MODE, SPLIT, POW2, REPG, FIXED>(buf[u].x.y, buf[u].v.y, buf[u].w.y, P, dsh, b0[2 * u + 1], val[2 * u + 1]);
            }
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const bool active = (q0 + u * stride) < npairs;
                scatter<K, VAR, FIXED>(wg, P.rep_log2, rep, lane, b0[2 * u], val[2 * u], active, P.fixscale);
                scatter<K, VAR, FIXED>(wg, P.rep_log2, rep, lane, b0[2 * u + 1], val[2 * u + 1], active, P.fixscale);
            }
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const unsigned q = q0 + u * stride;
                if (q < npairs && MODE != MODE_DEPOSIT) {
                    st_stream2(x + 2 * (size_t)q, buf[u].x);
                    if (MODE == MODE_PUSH_DEPOSIT) st_stream2(v + 2 * (size_t)q, buf[u].v);
                }
            }
        } else {
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const unsigned q = q0 + u * stride;
                bool active = q < npairs;
                if constexpr (VAR == VAR_AF) {          // nothing warp-collective in this layout: idle threads skip the pair
                    if (!active) continue;
                    active = true;
                }
                process<K, VAR, MODE, SPLIT, POW2, REPG, FIXED>(buf[u].x.x, buf[u].v.x, buf[u].w.x, active, P, dsh, wg, rep, lane);
                process<K, VAR, MODE, SPLIT, POW2, REPG, FIXED>(buf[u].x.y, buf[u].v.y, buf[u].w.y, active, P, dsh, wg, rep, lane);
                if (active && MODE != MODE_DEPOSIT) {
                    st_stream2(x + 2 * (size_t)q, buf[u].x);
                    if (MODE == MODE_PUSH_DEPOSIT) st_stream2(v + 2 * (size_t)q, buf[u].v);
                }
            }
        }
    };
#pragma unroll
    for (int u = 0; u < U; ++u) A[u].x = A[u].v = B[u].x = B[u].v = make_double2(0., 0.);
    // q advances by `chunk` per iteration; npairs + 3*chunk < 2^32 is guaranteed by vm_particles_create
    unsigned q = gtid;
    // Programmatic dependent launch: the replica zero-fill above and the first batch of particle loads below overlap
    // the previous kernel's tail (its last CTA is still reducing / exchanging / solving).  The particle arrays are
    // safe to read early: a pass kernel fences its stores before griddepcontrol.launch_dependents, and this grid only
    // starts once every CTA of the previous one has passed that point (any other kind of predecessor runs to
    // completion first).  What the finish of the previous kernel writes (dcoef) is read after griddepcontrol.wait.
    if (MODE == MODE_PUSH_DEPOSIT && VM_PASS_EARLY_LOAD) load(A, q);
    asm volatile("griddepcontrol.wait;" ::: "memory");
    if (MODE == MODE_PUSH_DEPOSIT) {
        if (VAR == VAR_AF && P.ps_tiles > 0) {
            __syncthreads();                                  // (the zero-fill of the grids, which double as scratch here)
            pass_presolve(P, grid);
            constexpr int ext = K - 2 > 0 ? K - 2 : 0, copies = REPG ? VM_GATHER_COPIES : 1;
            for (int i = threadIdx.x; i < (n + ext) * copies; i += blockDim.x) {
                const int m = i / copies;
                dsh[i] = __ldcg(dcoef + (m < n ? m : m - n));
            }
        } else {
            load_dcoef_ext<REPG>(dsh, dcoef, n, K - 2 > 0 ? K - 2 : 0);
        }
    }
    __syncthreads();
    if (!(MODE == MODE_PUSH_DEPOSIT && VM_PASS_EARLY_LOAD)) load(A, q);
    if (MODE == MODE_DEPOSIT || (VM_AF_UNROLL && VAR == VAR_AF) || (VM_PRIV_UNROLL && U == 1)) {
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
    // let the next kernel of the stream start its prologue while this grid drains and finishes; the fence makes this
    // thread's particle stores visible before the trigger (the next kernel loads its first batch before its wait)
    if (MODE != MODE_DEPOSIT && VM_PASS_EARLY_LOAD) __threadfence();
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    if ((P.n & 1) && blockIdx.x == 0 && warp == 0) {   // odd particle count: last particle, lane 0 of one warp
        const bool active = (lane == 0);
        double xp = 0., vp = 0., wp = 0.;
        if (active) {
            xp = x[P.n - 1];
            if (MODE != MODE_DEPOSIT) vp = v[P.n - 1];
            wp = P.uw ? P.w0 : w[P.n - 1];
        }
        process<K, VAR, MODE, SPLIT, POW2, REPG, FIXED>(xp, vp, wp, active, P, dsh, wg, rep, lane);
        if (active && MODE != MODE_DEPOSIT) {
            x[P.n - 1] = xp;
            if (MODE == MODE_PUSH_DEPOSIT) v[P.n - 1] = vp;
        }
    }
    if (VAR == VAR_AF) flush_limbs((const unsigned*)grid, out, n, GHOST, P.rep_log2, pitch, P.ncols, F.mode != FINISH_NONE && F.fixed, F.inv_scale);
    else flush_grid<VAR, FIXED>(grid, scratch, out, n, GHOST, P.rep_log2, nwarps, P.ncols);
    if (VAR != VAR_ATOMIC && F.mode != FINISH_NONE) finish_grid(F, out, gridDim.x, n, grid, scratch);
}

// ================================================================ host ======
// Plan of one particle pass: deposit variant + CTA shape, and whether the fused pass stores its gather table
// VM_GATHER_COPIES times.  The replicated table only exists in the lane-private variant: plan with it first and
// fall back to the plain table when the plan picks another variant (or nothing fits).
struct PassPlan {
    DepositPlan pl;
    bool repg;
};
inline PassPlan plan_pass(vm_ctx* ctx, int n, int order, int pass_mode, int deposit_mode)
{
    const int pmw = ctx->priv_min_warps > 0 ? ctx->priv_min_warps
                                            : (pass_mode == MODE_DEPOSIT ? VM_PRIV_MIN_WARPS : VM_PRIV_MIN_WARPS_PUSH);
    PassPlan pp{};
    pp.repg = pass_mode == MODE_PUSH_DEPOSIT && n > 16 && !ctx->no_repg && deposit_mode != VM_DEPOSIT_ATOMIC &&
              vm_gather_table_doubles(n, order, true) * sizeof(double) <= VM_GATHER_TABLE_MAX_BYTES;
    if (pp.repg) {
        try { pp.pl = plan_deposit(ctx, n, order - 1, (int)vm_gather_table_doubles(n, order, true), deposit_mode, pmw); }
        catch (const vm_error&) { pp.pl.var = -1; }
        if (pp.pl.var != VAR_PRIV) pp.repg = false;
    }
    if (!pp.repg)
        pp.pl = plan_deposit(ctx, n, order - 1, pass_mode == MODE_PUSH_DEPOSIT ? (int)vm_gather_table_doubles(n, order, false) : 0,
                             deposit_mode, pmw);
    return pp;
}

// Limb-atomic pass (VAR_AF): R bank-steered replicas of a grid of (n + K - 1) two-limb rows per CTA, shared by all its
// warps (vm_deposit.cuh) -- 8 B * R per row and CTA, so the CTA shape is one full CTA per SM (1024 threads; tuning key
// af_ctas: 2 or 4 smaller ones), the gather table is stored 16-fold whenever it fits, and R is the largest power of
// two <= 32 that still fits (32 up to 512 cells with the 16-fold table; tuning key "replicas" overrides it here too).
inline size_t vm_af_core_doubles(int n, int order, int rep_log2)
{
    const size_t limbs = (size_t)vm_af_pitch(n + order - 1, rep_log2) << rep_log2;
    return limbs > (size_t)(3 * n + 2) ? limbs : (size_t)(3 * n + 2);
}
inline bool plan_af(vm_ctx* ctx, int n, int order, int pass_mode, PassPlan* out)
{
    // Replicas first, the 16-fold gather table second: only 32 (one wavefront per atomic) and 16 replicas (two) pay --
    // ncu, 1024 cells: 8 replicas cost 3.9 wavefronts per atomic, as many as a single grid -- and at 1024 cells 16 replicas
    // with the plain table beat 8 replicas with the 16-fold one by 10 % (profiles/r02c_priv_variants_and_1024_replicas.jsonl).
    const size_t sm_total = 227 * 1024;
    int rl_max = 5;
    if (ctx->af_replicas > 0) { rl_max = 0; while ((1 << rl_max) < ctx->af_replicas) ++rl_max; }
    // Shared memory is carved out of the 256 KB L1: a 198 KB plan (512 cells, 32 replicas + 16-fold table) left too little
    // L1 for the 64 KB of streaming loads the CTA keeps in flight and ran 10 % slower than either 140 KB alternative
    // (profiles/r02c_af_smem_footprint_ab.jsonl), so plans stay below VM_AF_SMEM_CAP while one exists.
    const int ctas = ctx->af_ctas > 0 ? ctx->af_ctas : 1;    // (1 x 1024 measured >= 2 x 512 from 96 cells on, equal below)
    const int threads = VM_AF_THREADS / ctas;
    if (threads < 128) return false;
    for (int capped = 1; capped >= 0; --capped) {
        for (int rl = rl_max; rl >= 0; --rl) {
            for (int rg = 1; rg >= 0; --rg) {
                if (rg && (pass_mode != MODE_PUSH_DEPOSIT || n <= 16 || ctx->no_repg)) continue;
                const size_t table = pass_mode == MODE_PUSH_DEPOSIT ? vm_gather_table_doubles(n, order, rg != 0) : 0;
                const size_t smem = (table + vm_af_core_doubles(n, order, rl) + (size_t)threads) * sizeof(double);
                if (capped && (size_t)ctas * smem > VM_AF_SMEM_CAP) continue;
                if (smem <= ctx->smem_optin && (size_t)ctas * (smem + 1024) <= sm_total) {
                    out->pl = DepositPlan{VAR_AF, rl, ctx->sm_count * ctas, threads, smem};
                    out->repg = rg != 0;
                    return true;
                }
            }
        }
    }
    return false;
}

template <int K, int VAR, int MODE, bool SPLIT, bool POW2, int U, int MAXT, bool REPG, bool FIXED = false>
void launch_pass_depth(vm_ctx* ctx, const DepositPlan& pl, double* x, double* v, const double* w,
                       const double* dcoef, double* out, const PassParams& P, const FinishParams& F)
{
    static size_t configured[64] = {};   // per device: max dynamic smem already opted into for this instantiation
    size_t& conf = configured[ctx->device & 63];
    if (pl.smem > conf) {
        VM_CUDA(cudaFuncSetAttribute(k_vp_pass<K, VAR, MODE, U, SPLIT, POW2, MAXT, REPG, FIXED>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)pl.smem));
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
    VM_CUDA(cudaLaunchKernelEx(&cfg, k_vp_pass<K, VAR, MODE, U, SPLIT, POW2, MAXT, REPG, FIXED>, x, v, w, dcoef, out, P, F));
    ++ctx->launches;
}

// Pairs in flight per thread and the launch bound of the instantiation that runs them.  Full CTAs (1024 threads
// per SM, 64 registers each) run the shallow loop; the lane-private variant with fewer warps trades its spare
// registers for depth (vm_auto_pairs; `pairs` tuning key: A/B override, rounded down to a tier the CTA fits).
struct PassTier { int pairs, max_threads; };
inline PassTier vm_pass_tier(int mode, int var, int per_sm, int pairs_req)
{
    const PassTier base{mode == MODE_DEPOSIT ? 2 : 1, 1024};
    if (var != VAR_PRIV) return base;
    const int u = pairs_req > 0 ? pairs_req : vm_auto_pairs(mode == MODE_DEPOSIT, per_sm);
    if (mode == MODE_DEPOSIT) {
        if (u >= 8 && per_sm <= 256) return PassTier{8, 256};
        if (u >= 4 && per_sm <= 512) return PassTier{4, 512};
    } else {
        if (u >= 8 && per_sm <= 192) return PassTier{8, 192};
        if (u >= 4 && per_sm <= 448) return PassTier{4, 448};
    }
    return base;
}

// P.repg (lane-private and limb-atomic fused passes on meshes with more than 16 cells): replicated gather table.
template <int K, int VAR, int MODE, bool SPLIT, bool POW2>
void launch_pass_inst(vm_ctx* ctx, const DepositPlan& pl, double* x, double* v, const double* w,
                      const double* dcoef, double* out, const PassParams& P, const FinishParams& F)
{
    constexpr int U0 = (MODE == MODE_DEPOSIT) ? 2 : 1;
    const int per_sm = pl.threads * (pl.grid / ctx->sm_count);      // resident threads per SM
    const PassTier t = vm_pass_tier(MODE, VAR, per_sm, ctx->pairs);
    if (pl.threads > t.max_threads) throw vm_error(VM_ERR_UNSUPPORTED, "internal: CTA larger than the launch bound of its tier");
    if constexpr (VAR == VAR_AF) {     // limb atomics: always fixed-point, shallow tier
        if (P.fixscale == 0.0) throw vm_error(VM_ERR_UNSUPPORTED, "internal: the limb-atomic deposit needs a fixed-point scale");
        constexpr int UA = (MODE == MODE_DEPOSIT) ? 2 : VM_AF_PAIRS;
        if (P.repg) return launch_pass_depth<K, VAR, MODE, SPLIT, POW2, UA, VM_AF_THREADS, (MODE == MODE_PUSH_DEPOSIT), true>(ctx, pl, x, v, w, dcoef, out, P, F);
        return launch_pass_depth<K, VAR, MODE, SPLIT, POW2, UA, VM_AF_THREADS, false, true>(ctx, pl, x, v, w, dcoef, out, P, F);
    } else {
        if (P.fixscale != 0.0) {      // fixed-point accumulation: lane-private layout, shallow tier (the planner sends everything else to the limb-atomic or bank-sorted pass)
            if constexpr (VAR == VAR_PRIV) {
                if (P.repg) return launch_pass_depth<K, VAR, MODE, SPLIT, POW2, U0, 1024, (MODE == MODE_PUSH_DEPOSIT), true>(ctx, pl, x, v, w, dcoef, out, P, F);
                return launch_pass_depth<K, VAR, MODE, SPLIT, POW2, U0, 1024, false, true>(ctx, pl, x, v, w, dcoef, out, P, F);
            }
            throw vm_error(VM_ERR_UNSUPPORTED, "internal: fixed-point deposit requested for a layout without it");
        }
        if constexpr (VAR == VAR_PRIV) {
            if constexpr (MODE == MODE_DEPOSIT) {
                if (t.pairs == 8) return launch_pass_depth<K, VAR, MODE, SPLIT, POW2, 8, 256, false>(ctx, pl, x, v, w, dcoef, out, P, F);
                if (t.pairs == 4) return launch_pass_depth<K, VAR, MODE, SPLIT, POW2, 4, 512, false>(ctx, pl, x, v, w, dcoef, out, P, F);
            } else if constexpr (MODE == MODE_DRIFT_DEPOSIT) {
                if (t.pairs == 8) return launch_pass_depth<K, VAR, MODE, SPLIT, POW2, 8, 192, false>(ctx, pl, x, v, w, dcoef, out, P, F);
                if (t.pairs == 4) return launch_pass_depth<K, VAR, MODE, SPLIT, POW2, 4, 448, false>(ctx, pl, x, v, w, dcoef, out, P, F);
            } else if (P.repg) {
                if (t.pairs == 8) return launch_pass_depth<K, VAR, MODE, SPLIT, POW2, 8, 192, true>(ctx, pl, x, v, w, dcoef, out, P, F);
                if (t.pairs == 4) return launch_pass_depth<K, VAR, MODE, SPLIT, POW2, 4, 448, true>(ctx, pl, x, v, w, dcoef, out, P, F);
                return launch_pass_depth<K, VAR, MODE, SPLIT, POW2, 1, 1024, true>(ctx, pl, x, v, w, dcoef, out, P, F);
            }
        }
        if (P.repg) throw vm_error(VM_ERR_UNSUPPORTED, "internal: replicated gather table requested for a deposit variant without it");
        launch_pass_depth<K, VAR, MODE, SPLIT, POW2, U0, 1024, false>(ctx, pl, x, v, w, dcoef, out, P, F);
    }
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
        case VAR_AF:
            if (pow2) { VM_PASS_VAR(VAR_AF, true); } else { VM_PASS_VAR(VAR_AF, false); }
            break;
        default: VM_PASS_VAR(VAR_ATOMIC, false); break;
    }
#undef VM_PASS_VAR
}

// vm_vspline.cu -- velocity-space spline projection and the Lenard-Bernstein right-hand sides.
//
// Replaces (reference paths):
//   SplineDistribution(1,1,nknots,order,domain,:Dirichlet)   src/distributions/spline_distribution.jl:1-36
//   projection(v, dist, sdist)                               src/projections/distribution.jl:35-55
//   compute_f_densities / compute_df_densities               src/projections/density.jl:6-52
//   compute_coefficients, CLB_rhs!, CLB_rhs_GI!              src/models/lenard_bernstein_conservative.jl:11-50
//   LB_rhs!, LB_rhs_GI!                                      src/models/lenard_bernstein.jl:20-34
//   run!(::GeometricIntegrator) with RK438                   src/methods/geometric_integrator.jl:12-44
//
// Data layout: the clamped basis is kept in PARENT indexing (npar = nknots + order - 2 functions;
// the Dirichlet recombination just drops parent 0 and npar-1).  Every cell c carries the exact
// polynomial pieces A[c][j][m] of its `order` nonzero B-splines (general knots at the clamped
// ends, uniform in the interior), so evaluation of f_s and f_s' at a particle is one Horner sweep
// over poly[c][0..order) -- no knot search, no boundary special cases.
#include <cstring>

#include "vm_deposit.cuh"
#include "vm_internal.cuh"
#include "vm_spline_host.hpp"

struct VCell {
    double inv_h, off;   // t = v * inv_h + off  in cell units
    double ncell_d;      // (double)ncell
    double slack;        // rounding slack of the affine map at the domain ends
    int ncell, k;
};

// Cell index and local coordinate of velocity v.  Always returns an in-range cell (so callers can stay
// branch-free); the boolean says whether v lies inside [vmin, vmax] -- outside (or NaN) nothing is deposited
// and the spline evaluates to zero.  The end points themselves may land a few ulp outside [0, ncell] after
// the affine map, hence the slack.
__device__ __forceinline__ bool vcell_of(const VCell& m, double v, int& c, double& xi)
{
    const double t = fma(v, m.inv_h, m.off);
    const bool ok = (t >= -m.slack) && (t <= m.ncell_d + m.slack);      // false for NaN
    // the conversion saturates and maps NaN to 0, so the integer clamp alone keeps c in range (the fp64
    // fmin/fmax clamp of t it replaces cost ~15 instructions of selects and moves per lookup)
    c = max(0, min(__double2int_rd(t), m.ncell - 1));
    xi = ok ? t - (double)c : 0.0;                                      // end points: xi within a few ulp of [0, 1]
    return ok;
}

// ------------------------------------------------------------ v deposit -----
// STRAIGHT = true: interior values always computed, the rare clamped-end cells override them (best in the
// stand-alone deposit pass); false: if/else (best inside the register-tight RK stage pass) -- A/B measured.
// Cell and deposit weights of one particle (no shared-memory access: callers batch this for several particles
// before the read-modify-writes, see k_vp_pass).  Returns whether the particle deposits at all.
template <int K, bool STRAIGHT>
__device__ __forceinline__ bool vdeposit_prepare(double vp, double wp, bool active, const VCell& m,
                                                 const double* __restrict__ cellpoly, int& c, double (&val)[K])
{
    double xi = 0.0;
    c = 0;
    active = active && vcell_of(m, vp, c, xi);
    if (!active) wp = 0.0;
    const bool interior = (c >= K - 1 && c <= m.ncell - K);
    if (STRAIGHT || interior) bspline_uniform_w<K>(xi, wp, val);   // interior cells: uniform cardinal splines (x weight)
    if (!interior) {                                      // rare: the 2(K-1) clamped end cells, exact pieces
        const double* A = cellpoly + (size_t)c * K * K;
#pragma unroll
        for (int j = 0; j < K; ++j) {
            double s = __ldg(A + j * K + K - 1);
#pragma unroll
            for (int q = K - 2; q >= 0; --q) s = fma(s, xi, __ldg(A + j * K + q));
            val[j] = s * wp;
        }
    }
    return active;
}

template <int K, int VAR, bool STRAIGHT = true>
__device__ __forceinline__ void vdeposit_one(double vp, double wp, bool active, const VCell& m,
                                             const double* __restrict__ cellpoly, double* __restrict__ wg,
                                             int npar, int rep_log2, int rep, int lane)
{
    int c;
    double val[K];
    active = vdeposit_prepare<K, STRAIGHT>(vp, wp, active, m, cellpoly, c, val);
    scatter<K, VAR>(wg, rep_log2, rep, lane, c, val, active);
}

// U = pairs of particles per thread and half-iteration, MAXT = launch bound (as k_vp_pass): the lane-private replica
// grids of the default 41-knot basis leave room for 20 warps, which get 96 registers and, per half-iteration, the
// cell/weight computation of the whole batch ahead of its read-modify-writes (ncu, round 2: the 2-pair / 64-register
// form ran at 62 % issue-active, 64 % of the shared-memory wavefront rate and 31 % occupancy -- latency-bound).
template <int K, int VAR, int U, int MAXT>
__global__ void __launch_bounds__(MAXT, 1)
k_v_deposit(const double* __restrict__ v, const double* __restrict__ w, long np, VCell m,
            const double* __restrict__ cellpoly, int npar, int rep_log2, double* __restrict__ out,
            const FinishParams F, int uw, double w0)
{
    extern __shared__ double smem[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
    const int gsz = npar << rep_log2;
    const int gtotal = (VAR == VAR_ATOMIC) ? gsz : gsz * nwarps;
    double* grid = smem;
    double* scratch = grid + gtotal;
    for (int i = threadIdx.x; i < gtotal; i += blockDim.x) grid[i] = 0.0;
    __syncthreads();
    double* wg = (VAR == VAR_ATOMIC) ? grid : grid + warp * gsz;
    const int rep = ((VAR == VAR_ATOMIC) ? warp : lane) & ((1 << rep_log2) - 1);

    // 8-16 B/particle pass: same loop shape as the x-space deposit-only pass (unrolled twice over two register
    // buffer sets, 32-bit pair indices)
    const unsigned npairs = (unsigned)(np >> 1);
    const unsigned stride = gridDim.x * blockDim.x;
    const unsigned gtid = blockIdx.x * blockDim.x + threadIdx.x;
    const unsigned chunk = U * stride;
    const unsigned iters = (npairs + chunk - 1) / chunk;
    double2 Av[U], Aw[U], Bv[U], Bw[U];
    auto load = [&](double2 (&bv)[U], double2 (&bw)[U], unsigned q0) {
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const unsigned q = q0 + u * stride;
            bw[u] = make_double2(0., 0.);
            if (q < npairs) {
                bv[u] = ld_stream2(v + 2 * (size_t)q);
                bw[u] = uw ? make_double2(w0, w0) : ld_stream2(w + 2 * (size_t)q);
            }
        }
    };
    auto work = [&](double2 (&bv)[U], double2 (&bw)[U], unsigned q0) {
        if constexpr (VAR == VAR_PRIV && MAXT < 1024) {
            int c[2 * U];
            double val[2 * U][K];
#pragma unroll
            for (int u = 0; u < U; ++u) {      // (inactive or out-of-domain particles come back with zero weights)
                const bool active = (q0 + u * stride) < npairs;
                vdeposit_prepare<K, true>(bv[u].x, bw[u].x, active, m, cellpoly, c[2 * u], val[2 * u]);
                vdeposit_prepare<K, true>(bv[u].y, bw[u].y, active, m, cellpoly, c[2 * u + 1], val[2 * u + 1]);
            }
#pragma unroll
            for (int u = 0; u < 2 * U; ++u) scatter<K, VAR>(wg, rep_log2, rep, lane, c[u], val[u], true);
        } else {
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const bool active = (q0 + u * stride) < npairs;
                vdeposit_one<K, VAR>(bv[u].x, bw[u].x, active, m, cellpoly, wg, npar, rep_log2, rep, lane);
                vdeposit_one<K, VAR>(bv[u].y, bw[u].y, active, m, cellpoly, wg, npar, rep_log2, rep, lane);
            }
        }
    };
#pragma unroll
    for (int u = 0; u < U; ++u) Av[u] = Bv[u] = make_double2(0., 0.);
    unsigned q = gtid;
    load(Av, Aw, q);
    for (unsigned it = 0; it < iters; it += 2, q += 2 * chunk) {
        load(Bv, Bw, q + chunk);
        work(Av, Aw, q);
        load(Av, Aw, q + 2 * chunk);
        work(Bv, Bw, q + chunk);
    }
    if ((np & 1) && blockIdx.x == 0 && warp == 0) {
        const bool active = (lane == 0);
        vdeposit_one<K, VAR>(active ? v[np - 1] : 0.0, active ? (uw ? w0 : w[np - 1]) : 0.0, active, m, cellpoly, wg, npar,
                             rep_log2, rep, lane);
    }
    flush_grid<VAR>(grid, scratch, out, npar, 0, rep_log2, nwarps, npar);
    if (VAR != VAR_ATOMIC && F.mode != FINISH_NONE) finish_grid(F, out, gridDim.x, npar, grid, scratch);
}

// ------------------------------------------------------------- small solves -
// coef_par[off + i] = sum_j Minv[i][j] rhs_par[off + j]: one warp per row, fixed-order tree.
__global__ void __launch_bounds__(256) k_v_solve(const double* __restrict__ minv, const double* __restrict__ rhs_par,
                                                 int nv, int off, int npar, double* __restrict__ coef_par)
{
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int i = blockIdx.x * 8 + warp;
    if (i < nv) {
        double s = 0.0;
        for (int j = lane; j < nv; j += 32) s = fma(minv[(size_t)i * nv + j], rhs_par[off + j], s);
        s = warp_sum(s);
        if (lane == 0) coef_par[off + i] = s;
    }
    if (blockIdx.x == 0 && threadIdx.x == 0 && off == 1) { coef_par[0] = 0.0; coef_par[npar - 1] = 0.0; }
}

// poly[c][m] = sum_j coef_par[c + j] * A[c][j][m]
__global__ void k_v_poly(const double* __restrict__ coef_par, const double* __restrict__ cellpoly, int ncell, int k,
                         double* __restrict__ poly)
{
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t < ncell * k) {
        const int c = t / k, m = t % k;
        double s = 0.0;
        for (int j = 0; j < k; ++j) s = fma(coef_par[c + j], cellpoly[((size_t)c * k + j) * k + m], s);
        poly[t] = s;
    }
}

// The per-cell polynomial table sits in shared memory with an ODD row stride (K | 1 doubles): rows of
// 16 consecutive cells then start in 16 distinct bank pairs, so a warp whose lanes sit in different cells
// reads its coefficients without bank conflicts (28 % of the shared wavefronts were conflicts at stride K).
template <int K>
struct PolyRow { static constexpr int stride = K | 1; };

template <int K>
__device__ __forceinline__ void load_poly(double* __restrict__ psh, const double* __restrict__ poly, int ncell)
{
    for (int i = threadIdx.x; i < ncell * K; i += blockDim.x) psh[(i / K) * PolyRow<K>::stride + (i % K)] = poly[i];
}

template <int K>
__device__ __forceinline__ void eval_f_df(const double* __restrict__ psh, const VCell& m, double v, double& f, double& df)
{
    int c;
    double xi;
    const bool ok = vcell_of(m, v, c, xi);          // branch-free: out-of-range values are zeroed by selects
    const double* q = psh + c * PolyRow<K>::stride;
    double sf = q[K - 1], sd = (double)(K - 1) * q[K - 1];
#pragma unroll
    for (int j = K - 2; j >= 0; --j) {
        sf = fma(sf, xi, q[j]);
        if (j >= 1) sd = fma(sd, xi, (double)j * q[j]);
    }
    f = ok ? sf : 0.0;
    df = (ok && K > 1) ? sd * m.inv_h : 0.0;
}

// ------------------------------------------------ fused RK438 stage pass ----
// One pass per Runge-Kutta stage of the Lenard-Bernstein velocity ODE
//   q_s    = v + dt (a1 k1 + a2 k2 + a3 k3)                      (stage state, recomputed, never re-read)
//   k_s    = -nu (f_s'(q_s) + (A1 + A2 q_s) f_s(q_s))             (LB_rhs! / CLB_rhs!)
//   q_next = v + dt (c1 k1 + c2 k2 + c3 k3 + cs k_s)              (next stage state, or v_new on the last stage)
//   deposit w * phi_i(q_next)                                      (projection for the next right-hand side)
// i.e. RHS evaluation, stage assembly and the next projection's deposit in ONE sweep over the particles
// (LB: 144 B/particle per RK438 step instead of 304 B for separate kernels).  Same arithmetic as
// k_v_rhs + k_rk_combine + k_v_deposit, so the fused and unfused drivers agree bitwise.
struct StageParams {
    const double *k1, *k2, *k3;
    double a1, a2, a3;
    double c1, c2, c3, cs;
    double* kout;      // k_s, or nullptr on the last stage
    double* qout;      // q_next for the moments pass (nullptr: not needed); on the last stage this is v itself
    double dt, nu;
    int uw;            // uniform particle weight w0: the weight array is not read
    double w0;
};

// MAXT < 1024 (the lane-private replica grids of the default 41-knot basis leave room for 20 warps): the
// register budget of the absent warps pays for a software-pipelined load of the next iteration's lines;
// MAXT == 1024 (64 registers) only prefetches them into L2.
template <int K, int VAR, int STAGE, int MAXT>
__global__ void __launch_bounds__(MAXT, 1)
k_lb_stage(const double* v, const double* __restrict__ w, long np, VCell m, const double* __restrict__ cellpoly,
           const double* __restrict__ poly, const double* __restrict__ mom, const StageParams S, int npar, int rep_log2,
           double* __restrict__ out, const FinishParams F)
{
    extern __shared__ double smem[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
    const int npoly = m.ncell * PolyRow<K>::stride;
    double* psh = smem;
    double* grid = smem + npoly;
    const int gsz = npar << rep_log2;
    const int gtotal = gsz * nwarps;
    double* scratch = grid + gtotal;
    for (int i = threadIdx.x; i < gtotal; i += blockDim.x) grid[i] = 0.0;
    load_poly<K>(psh, poly, m.ncell);
    __syncthreads();
    double* wg = grid + warp * gsz;
    const int rep = lane & ((1 << rep_log2) - 1);
    const double A1 = mom[5], A2 = mom[6];

    // STAGE s uses k_1 .. k_{s-1} (compile time): no pointer tests in the particle loop
    // phase A of a particle (reads the polynomial table only); the replica-grid read-modify-writes of the
    // particles of an iteration follow together, so that their dependency chains overlap (see k_vp_pass)
    auto one = [&](double vp, double wp, double k1, double k2, double k3, bool& active, double& ks, double& qn, int& cc,
                   double (&val)[K]) {
        double q = vp;
        if (STAGE >= 2) {
            double t = S.a1 * k1;
            if (STAGE >= 3) t = fma(S.a2, k2, t);
            if (STAGE >= 4) t = fma(S.a3, k3, t);
            q = fma(S.dt, t, vp);
        }
        double f, df;
        eval_f_df<K>(psh, m, q, f, df);
        ks = -S.nu * (df + fma(A2, q, A1) * f);
        double t;
        if (STAGE >= 2) {
            t = S.c1 * k1;
            if (STAGE >= 3) t = fma(S.c2, k2, t);
            if (STAGE >= 4) t = fma(S.c3, k3, t);
            t = fma(S.cs, ks, t);
        } else {
            t = S.cs * ks;
        }
        qn = fma(S.dt, t, vp);
        active = vdeposit_prepare<K, false>(qn, wp, active, m, cellpoly, cc, val);
    };

    const unsigned npairs = (unsigned)(np >> 1);
    const unsigned stride = gridDim.x * blockDim.x;
    const unsigned gtid = blockIdx.x * blockDim.x + threadIdx.x;
    const unsigned iters = (npairs + stride - 1) / stride;    // uniform trip count (warp-collective scatter variants)
    unsigned q = gtid;
    constexpr bool PIPE = (MAXT < 1024);
    struct Lines { double2 vv, ww, a, b, c; };
    auto load = [&](Lines& L, unsigned qq) {
        L.vv = L.ww = L.a = L.b = L.c = make_double2(0., 0.);
        if (qq < npairs) {
            L.vv = ld_stream2(v + 2 * (size_t)qq);
            L.ww = S.uw ? make_double2(S.w0, S.w0) : ld_stream2(w + 2 * (size_t)qq);
            if (STAGE >= 2) L.a = ld_stream2(S.k1 + 2 * (size_t)qq);
            if (STAGE >= 3) L.b = ld_stream2(S.k2 + 2 * (size_t)qq);
            if (STAGE >= 4) L.c = ld_stream2(S.k3 + 2 * (size_t)qq);
        }
    };
    Lines cur, nxt;
    if (PIPE) load(cur, q);
    for (unsigned it = 0; it < iters; ++it, q += stride) {
        const bool active = q < npairs;
        // pull the lines of a later iteration into L2 (one lane per 128-byte line = 8 pairs): the next one when the
        // pass has no registers for software-pipelined loads, the one after next otherwise
        const unsigned qp = q + (PIPE ? 2 : 1) * stride;
        if ((lane & 7) == 0 && qp < npairs) {
            const size_t e = 2 * (size_t)qp;
            prefetch_l2(v + e);
            if (!S.uw) prefetch_l2(w + e);
            if (STAGE >= 2) prefetch_l2(S.k1 + e);
            if (STAGE >= 3) prefetch_l2(S.k2 + e);
            if (STAGE >= 4) prefetch_l2(S.k3 + e);
        }
        if (PIPE) load(nxt, q + stride);
        else load(cur, q);
        double2 ks, qn;
        int c0, c1;
        double val0[K], val1[K];
        bool act0 = active, act1 = active;
        one(cur.vv.x, cur.ww.x, cur.a.x, cur.b.x, cur.c.x, act0, ks.x, qn.x, c0, val0);
        one(cur.vv.y, cur.ww.y, cur.a.y, cur.b.y, cur.c.y, act1, ks.y, qn.y, c1, val1);
        scatter<K, VAR>(wg, rep_log2, rep, lane, c0, val0, act0);
        scatter<K, VAR>(wg, rep_log2, rep, lane, c1, val1, act1);
        if (active) {
            if (STAGE < 4) st_stream2(S.kout + 2 * (size_t)q, ks);
            if (S.qout) st_stream2(S.qout + 2 * (size_t)q, qn);
        }
        if (PIPE) cur = nxt;
    }
    if ((np & 1) && blockIdx.x == 0 && warp == 0) {
        const bool active = (lane == 0);
        const long p = np - 1;
        double ks = 0., qn = 0.;
        int c0;
        double val0[K];
        bool act0 = active;
        one(active ? v[p] : 0.0, active ? (S.uw ? S.w0 : w[p]) : 0.0, (active && STAGE >= 2) ? S.k1[p] : 0.0,
            (active && STAGE >= 3) ? S.k2[p] : 0.0, (active && STAGE >= 4) ? S.k3[p] : 0.0, act0, ks, qn, c0, val0);
        scatter<K, VAR>(wg, rep_log2, rep, lane, c0, val0, act0);
        if (active) {
            if (STAGE < 4) S.kout[p] = ks;
            if (S.qout) S.qout[p] = qn;
        }
    }
    flush_grid<VAR>(grid, scratch, out, npar, 0, rep_log2, nwarps, npar);
    if (F.mode != FINISH_NONE) finish_grid(F, out, gridDim.x, npar, grid, scratch);
}

// Streaming helper: each thread handles U pairs (4 particles for U = 2) per iteration, all loads issued
// before any use, so that enough bytes are in flight for these low-byte passes.
#define VM_STREAM_PAIRS 2
#define VM_MOMENT_PAIRS 2      // (4 pairs in flight measured no faster: the pass is not load-latency bound)

// five unweighted particle sums: [sum f, sum v f, sum v^2 f, sum f', sum v f']
// mom = [5 sums, A1, A2]; with `ticket` != nullptr the last CTA to finish sums the per-CTA rows in the
// order of k_reduce_rows8 and evaluates compute_coefficients -- two launches fewer per right-hand side.
__device__ __forceinline__ void clb_coefficients(double* __restrict__ mom)
{
    const double n = mom[0], nu = mom[1], ne = mom[2];
    const double B1 = -mom[3], B2 = -mom[4];
    const double den = n * ne - nu * nu;
    mom[5] = (ne * B1 - nu * B2) / den;
    mom[6] = -(nu * B1 - n * B2) / den;
}

template <int K>
__global__ void __launch_bounds__(512, 2)
k_v_moments(const double* __restrict__ v, long np, VCell m, const double* __restrict__ poly, double* __restrict__ out,
            unsigned* ticket, double* __restrict__ mom, const FinishParams X /* xchg fields only */)
{
    extern __shared__ double psh[];
    load_poly<K>(psh, poly, m.ncell);
    __syncthreads();
    double s[5] = {0., 0., 0., 0., 0.};
    auto acc = [&](double vp) {
        double f, df;
        eval_f_df<K>(psh, m, vp, f, df);
        s[0] += f;
        s[1] = fma(vp, f, s[1]);
        s[2] = fma(vp * vp, f, s[2]);
        s[3] += df;
        s[4] = fma(vp, df, s[4]);
    };
    constexpr int U = VM_MOMENT_PAIRS;
    const long npairs = np >> 1;
    const long stride = (long)gridDim.x * blockDim.x;
    const long gtid = (long)blockIdx.x * blockDim.x + threadIdx.x;
    for (long base = gtid; base < npairs; base += U * stride) {
        double2 c[U];
#pragma unroll
        for (int u = 0; u < U; ++u)
            if (base + u * stride < npairs) c[u] = ld_stream2(v + 2 * (base + u * stride));
#pragma unroll
        for (int u = 0; u < U; ++u)
            if (base + u * stride < npairs) { acc(c[u].x); acc(c[u].y); }
    }
    if ((np & 1) && gtid == 0) acc(v[np - 1]);
    __shared__ double red[5][32];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
#pragma unroll
    for (int q = 0; q < 5; ++q) {
        const double t = warp_sum(s[q]);
        if (lane == 0) red[q][warp] = t;
    }
    __syncthreads();
    if (threadIdx.x < 8) {
        double t = 0.0;
        if (threadIdx.x < 5)
            for (int q = 0; q < nwarps; ++q) t += red[threadIdx.x][q];
        out[(size_t)blockIdx.x * 8 + threadIdx.x] = t;
    }
    if (ticket) {
        __shared__ int s_last;
        __shared__ double fin[32][8];
        __threadfence();
        __syncthreads();
        if (threadIdx.x == 0) s_last = (atomicAdd(ticket, 1u) == gridDim.x - 1u);
        __syncthreads();
        if (!s_last) return;
        __threadfence();
        const int nrows = gridDim.x;
        if (threadIdx.x < 256) {                     // same order as k_reduce_rows8
            const int c = threadIdx.x & 7, ch = threadIdx.x >> 3;
            const int len = (nrows + 31) / 32;
            double t = 0.0;
            for (int r = ch * len; r < min(nrows, (ch + 1) * len); ++r) t += __ldcg(out + (size_t)r * 8 + c);
            fin[ch][c] = t;
        }
        __syncthreads();
        __shared__ double xv[8];
        if (threadIdx.x < 8) {
            double t = 0.0;
            for (int q = 0; q < 32; ++q) t += fin[q][threadIdx.x];
            if (threadIdx.x < 5 && !X.xchg) mom[threadIdx.x] = t;
            xv[threadIdx.x] = threadIdx.x < 5 ? t : 0.0;
        }
        if (X.xchg) exchange_ll(X, xv, 8, mom);      // the five sums of all ranks, rank order (mom[5..7] are rewritten below)
        __syncthreads();
        if (threadIdx.x == 0) {
            __threadfence();
            clb_coefficients(mom);
            *ticket = 0u;
        }
    }
}

// A1, A2 of compute_coefficients (lenard_bernstein_conservative.jl:13-18) from the five sums
__global__ void k_clb_coeffs(double* __restrict__ mom, int conservative)
{
    if (threadIdx.x == 0) {
        if (conservative) {
            clb_coefficients(mom);
        } else {
            mom[5] = 0.0;   // LB: vdot = -nu (f' + v f)
            mom[6] = 1.0;
        }
    }
}

// vdot_p = -nu (f'(v_p) + (A1 + A2 v_p) f(v_p))
template <int K>
__global__ void __launch_bounds__(512, 2)
k_v_rhs(const double* __restrict__ v, long np, VCell m, const double* __restrict__ poly,
        const double* __restrict__ mom, double nu, double* __restrict__ vdot)
{
    extern __shared__ double psh[];
    load_poly<K>(psh, poly, m.ncell);
    __syncthreads();
    const double A1 = mom[5], A2 = mom[6];
    auto rhs = [&](double vp) {
        double f, df;
        eval_f_df<K>(psh, m, vp, f, df);
        return -nu * (df + fma(A2, vp, A1) * f);
    };
    constexpr int U = VM_STREAM_PAIRS;
    const long npairs = np >> 1;
    const long stride = (long)gridDim.x * blockDim.x;
    const long gtid = (long)blockIdx.x * blockDim.x + threadIdx.x;
    for (long base = gtid; base < npairs; base += U * stride) {
        double2 c[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const long q = base + u * stride;
            if (q < npairs) c[u] = ld_stream2(v + 2 * q);
            if ((threadIdx.x & 7) == 0 && q + U * stride < npairs) prefetch_l2(v + 2 * (q + U * stride));
        }
#pragma unroll
        for (int u = 0; u < U; ++u)
            if (base + u * stride < npairs)
                st_stream2(vdot + 2 * (base + u * stride), make_double2(rhs(c[u].x), rhs(c[u].y)));
    }
    if ((np & 1) && gtid == 0) vdot[np - 1] = rhs(v[np - 1]);
}

// f_s and f_s' at arbitrary points (plotting / tests)
template <int K>
__global__ void k_v_eval(const double* __restrict__ v, long np, VCell m, const double* __restrict__ poly,
                         double* __restrict__ f, double* __restrict__ df)
{
    extern __shared__ double psh[];
    load_poly<K>(psh, poly, m.ncell);
    __syncthreads();
    const long stride = (long)gridDim.x * blockDim.x;
    for (long p = (long)blockIdx.x * blockDim.x + threadIdx.x; p < np; p += stride) {
        double a, b;
        eval_f_df<K>(psh, m, v[p], a, b);
        f[p] = a;
        df[p] = b;
    }
}

// q = v + dt (a1 k1 + a2 k2 + a3 k3 + a4 k4)   (RK stage assembly / final update; null k's skipped)
__global__ void __launch_bounds__(512, 2)
k_rk_combine(const double* v, const double* __restrict__ k1, const double* __restrict__ k2,
             const double* __restrict__ k3, const double* __restrict__ k4, double a1, double a2, double a3, double a4,
             double dt, long np, double* q /* may alias v (final update) */)
{
    const long npairs = np >> 1;
    const long stride = (long)gridDim.x * blockDim.x;
    const long gtid = (long)blockIdx.x * blockDim.x + threadIdx.x;
    for (long p = gtid; p < npairs; p += stride) {
        const double2 vv = ld_stream2(v + 2 * p);
        double2 s = ld_stream2(k1 + 2 * p);
        s.x *= a1; s.y *= a1;
        if (k2) { const double2 t = ld_stream2(k2 + 2 * p); s.x = fma(a2, t.x, s.x); s.y = fma(a2, t.y, s.y); }
        if (k3) { const double2 t = ld_stream2(k3 + 2 * p); s.x = fma(a3, t.x, s.x); s.y = fma(a3, t.y, s.y); }
        if (k4) { const double2 t = ld_stream2(k4 + 2 * p); s.x = fma(a4, t.x, s.x); s.y = fma(a4, t.y, s.y); }
        st_stream2(q + 2 * p, make_double2(fma(dt, s.x, vv.x), fma(dt, s.y, vv.y)));
    }
    if ((np & 1) && gtid == 0) {
        const long p = np - 1;
        double s = a1 * k1[p];
        if (k2) s = fma(a2, k2[p], s);
        if (k3) s = fma(a3, k3[p], s);
        if (k4) s = fma(a4, k4[p], s);
        q[p] = fma(dt, s, v[p]);
    }
}

// [sum v, sum v^2] (script diagnostics, lenard_bernstein_conservative.jl:49-50)
__global__ void __launch_bounds__(512, 2) k_v_sums(const double* __restrict__ v, long np, double* __restrict__ out)
{
    double s1 = 0.0, s2 = 0.0;
    const long stride = (long)gridDim.x * blockDim.x;
    for (long p = (long)blockIdx.x * blockDim.x + threadIdx.x; p < np; p += stride) {
        const double vp = ld_stream(v + p);
        s1 += vp;
        s2 = fma(vp, vp, s2);
    }
    __shared__ double red[2][32];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
    s1 = warp_sum(s1); s2 = warp_sum(s2);
    if (lane == 0) { red[0][warp] = s1; red[1][warp] = s2; }
    __syncthreads();
    if (threadIdx.x < 8) {
        double t = 0.0;
        if (threadIdx.x < 2)
            for (int q = 0; q < nwarps; ++q) t += red[threadIdx.x][q];
        out[(size_t)blockIdx.x * 8 + threadIdx.x] = t;
    }
}

__global__ void k_v_store_diag(double* __restrict__ row, double t, const double* __restrict__ sums)
{
    if (threadIdx.x == 0) { row[0] = t; row[1] = sums[0]; row[2] = sums[1]; row[3] = 0.0; }
}

__global__ void __launch_bounds__(256) k_reduce_rows8(const double* __restrict__ rows, int nrows, double* __restrict__ out)
{
    // out[c] = sum_r rows[r][c], c < 8: 32 chunks of rows, then the chunks in order
    __shared__ double red[32][8];
    const int c = threadIdx.x & 7, ch = threadIdx.x >> 3;
    const int len = (nrows + 31) / 32;
    double s = 0.0;
    for (int r = ch * len; r < min(nrows, (ch + 1) * len); ++r) s += rows[(size_t)r * 8 + c];
    red[ch][c] = s;
    __syncthreads();
    if (threadIdx.x < 8) {
        double t = 0.0;
        for (int q = 0; q < 32; ++q) t += red[q][threadIdx.x];
        out[threadIdx.x] = t;
    }
}

// ================================================================ host ======
void vm_reduce_rows(vm_ctx* ctx, const double* rows, int nrows, int ncols, double* out);   // vm_field.cu

namespace {

VCell vcell(const vm_vspline* s)
{
    VCell m;
    m.inv_h = (double)((long double)s->ncell / ((long double)s->b - (long double)s->a));
    m.off = -s->a * m.inv_h;
    m.ncell = s->ncell;
    m.ncell_d = (double)s->ncell;
    m.slack = 2e-15 * (double)s->ncell;
    m.k = s->order;
    return m;
}

// the per-cell polynomial table of a large basis exceeds the 48 KB a kernel gets without opting in
template <typename Kern>
void vsmem_optin(Kern kern, vm_ctx* ctx, size_t smem)
{
    if (smem > 48 * 1024) VM_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    (void)ctx;
}

void geometry(vm_ctx* ctx, int* grid, int* threads)
{
    vm_launch_geometry(ctx, grid, threads);
    if (*threads > 512) *threads = 512;
}

template <int K, int VAR, int U, int MAXT>
void launch_vdep_tier(vm_vspline* s, const DepositPlan& pl, const double* v, const double* w, long np, double* out,
                      const FinishParams& F)
{
    vm_ctx* ctx = s->ctx;
    static size_t configured[64] = {};
    size_t& conf = configured[ctx->device & 63];
    if (pl.smem > conf) {
        VM_CUDA(cudaFuncSetAttribute(k_v_deposit<K, VAR, U, MAXT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)pl.smem));
        conf = pl.smem;
    }
    k_v_deposit<K, VAR, U, MAXT><<<pl.grid, pl.threads, pl.smem, ctx->stream>>>(v, w, np, vcell(s), s->cellpoly, s->npar,
                                                                               pl.rep_log2, out, F, s->uw, s->w0);
    VM_LAUNCHED(ctx);
}

template <int K, int VAR>
void launch_vdep_inst(vm_vspline* s, const DepositPlan& pl, const double* v, const double* w, long np, double* out,
                      const FinishParams& F)
{
    if constexpr (VAR == VAR_PRIV) {          // few lane-private warps: deeper software pipeline (pairs == 1: A/B switch back)
        const int per_sm = pl.threads * (pl.grid / s->ctx->sm_count);
        if (per_sm <= 320 && s->ctx->pairs != 1) return launch_vdep_tier<K, VAR, 4, 320>(s, pl, v, w, np, out, F);
        if (per_sm <= 640 && s->ctx->pairs != 1) return launch_vdep_tier<K, VAR, 2, 640>(s, pl, v, w, np, out, F);
    }
    launch_vdep_tier<K, VAR, 2, 1024>(s, pl, v, w, np, out, F);
}

template <int K>
void launch_vdep_var(vm_vspline* s, const DepositPlan& pl, const double* v, const double* w, long np, double* out,
                     const FinishParams& F)
{
    switch (pl.var) {
        case VAR_PRIV: launch_vdep_inst<K, VAR_PRIV>(s, pl, v, w, np, out, F); break;
        case VAR_MATCH: launch_vdep_inst<K, VAR_MATCH>(s, pl, v, w, np, out, F); break;
        case VAR_XOR: launch_vdep_inst<K, VAR_XOR>(s, pl, v, w, np, out, F); break;
        default: launch_vdep_inst<K, VAR_ATOMIC>(s, pl, v, w, np, out, F); break;
    }
}

#define VM_ORDER_SWITCH(order, ...)                                                         \
    switch (order) {                                                                        \
        case 2: { constexpr int K = 2; __VA_ARGS__; } break;                                \
        case 3: { constexpr int K = 3; __VA_ARGS__; } break;                                \
        case 4: { constexpr int K = 4; __VA_ARGS__; } break;                                \
        case 5: { constexpr int K = 5; __VA_ARGS__; } break;                                \
        case 6: { constexpr int K = 6; __VA_ARGS__; } break;                                \
        default: throw vm_error(VM_ERR_UNSUPPORTED, "spline order must be in 2..6");        \
    }

// plan + finish parameters shared by the projection deposit and the fused RK stage pass
struct VDepSetup {
    DepositPlan pl;
    FinishParams F;
    double* out;
};

VDepSetup vdep_setup(vm_vspline* s, int extra_doubles)
{
    vm_ctx* ctx = s->ctx;
    VDepSetup d{};
    d.pl = plan_deposit(ctx, s->npar, 0, extra_doubles, VM_DEPOSIT_DETERMINISTIC);
    d.out = vm_partials(ctx, (size_t)d.pl.grid * s->npar);
    const size_t gdoubles = ((size_t)s->npar << d.pl.rep_log2) * (size_t)(d.pl.threads / 32);
    if (s->npar <= VM_FUSE_MAX_N && !ctx->no_fuse && d.pl.var != VAR_ATOMIC && gdoubles >= (size_t)2 * s->npar + 1 &&
        d.pl.threads >= s->npar) {     // the finish works with one thread per basis function
        d.F.mode = FINISH_REDUCE;           // last CTA sums the per-CTA rows in a fixed order
        d.F.ticket = ctx->ticket;
        d.F.rhs = s->rhs;
        if (ctx->nranks == 1 || vm_xchg_setup(ctx, d.F)) {   // single GPU, or ranks connected through peer memory:
            d.F.mode = FINISH_REDUCE_VSOLVE;                 // (exchange,) mass solve and polynomial table too
            d.F.minv = s->minv; d.F.cellpoly = s->cellpoly; d.F.coef = s->coef; d.F.poly = s->poly;
            d.F.nv = s->nv; d.F.off = s->bc ? 1 : 0; d.F.ncell = s->ncell; d.F.k = s->order;
        }
    }
    return d;
}

// rows (or the fused reduction) -> all-reduce -> M^{-1} -> per-cell polynomials
void after_deposit(vm_vspline* s, const VDepSetup& d)
{
    vm_ctx* ctx = s->ctx;
    if (d.F.mode == FINISH_REDUCE_VSOLVE) return;      // already done by the pass's last CTA
    if (d.F.mode == FINISH_NONE) vm_reduce_rows(ctx, d.out, d.pl.grid, s->npar, s->rhs);
    vm_allreduce_sum(ctx, s->rhs, (size_t)s->npar);
    const int off = s->bc ? 1 : 0;
    k_v_solve<<<(s->nv + 7) / 8, 256, 0, ctx->stream>>>(s->minv, s->rhs, s->nv, off, s->npar, s->coef);
    VM_LAUNCHED(ctx);
    const int tot = s->ncell * s->order;
    k_v_poly<<<(tot + 255) / 256, 256, 0, ctx->stream>>>(s->coef, s->cellpoly, s->ncell, s->order, s->poly);
    VM_LAUNCHED(ctx);
}

// projection: deposit -> fixed-order reduce -> all-reduce -> M^{-1} -> per-cell polynomials
void project_dev(vm_vspline* s, const double* v, const double* w, long np)
{
    vm_ctx* ctx = s->ctx;
    VDepSetup d = vdep_setup(s, 0);
    vm_prof_mark(ctx);
    VM_ORDER_SWITCH(s->order, launch_vdep_var<K>(s, d.pl, v, w, np, d.out, d.F));
    vm_prof_mark(ctx);
    after_deposit(s, d);
}

template <int K, int VAR, int STAGE, int MAXT>
void launch_stage_inst(vm_vspline* s, const VDepSetup& d, const double* v, const double* w, long np, const StageParams& S)
{
    vm_ctx* ctx = s->ctx;
    static size_t configured[64] = {};
    size_t& conf = configured[ctx->device & 63];
    if (d.pl.smem > conf) {
        VM_CUDA(cudaFuncSetAttribute(k_lb_stage<K, VAR, STAGE, MAXT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)d.pl.smem));
        conf = d.pl.smem;
    }
    k_lb_stage<K, VAR, STAGE, MAXT><<<d.pl.grid, d.pl.threads, d.pl.smem, ctx->stream>>>(
        v, w, np, vcell(s), s->cellpoly, s->poly, s->moments, S, s->npar, d.pl.rep_log2, d.out, d.F);
    VM_LAUNCHED(ctx);
}

// the stage pass exists in the lane-private and the match-grouped flavour (the latter handles any replica count)
template <int K, int STAGE>
void launch_stage_var(vm_vspline* s, const VDepSetup& d, const double* v, const double* w, long np, const StageParams& S)
{
    const bool pipe = d.pl.threads <= 640 && s->ctx->pairs != 1;      // pairs == 1: A/B switch back to the L2-prefetch-only loop
    if (d.pl.var == VAR_PRIV) {
        if (pipe) launch_stage_inst<K, VAR_PRIV, STAGE, 640>(s, d, v, w, np, S);
        else launch_stage_inst<K, VAR_PRIV, STAGE, 1024>(s, d, v, w, np, S);
    } else launch_stage_inst<K, VAR_MATCH, STAGE, 1024>(s, d, v, w, np, S);
}

template <int K>
void launch_stage(vm_vspline* s, const VDepSetup& d, const double* v, const double* w, long np, const StageParams& S, int stage)
{
    switch (stage) {
        case 1: launch_stage_var<K, 1>(s, d, v, w, np, S); break;
        case 2: launch_stage_var<K, 2>(s, d, v, w, np, S); break;
        case 3: launch_stage_var<K, 3>(s, d, v, w, np, S); break;
        default: launch_stage_var<K, 4>(s, d, v, w, np, S); break;
    }
}

void moments_dev(vm_vspline* s, const double* v, long np, int conservative)
{
    vm_ctx* ctx = s->ctx;
    if (conservative) {
        s->lb_coeffs_set = false;                     // moments[5..6] are about to hold A1, A2 of the conservative model
        int grid, threads;
        geometry(ctx, &grid, &threads);
        if (threads < 256) threads = 256;             // the fused finish uses 256 threads
        double* out = vm_partials(ctx, (size_t)grid * 8);
        const size_t smem = (size_t)s->ncell * (s->order | 1) * sizeof(double);
        const bool fuse = (ctx->nranks == 1 || ctx->peers_connected) && !ctx->no_fuse;
        unsigned* ticket = fuse ? ctx->ticket : nullptr;
        FinishParams X{};
        if (fuse) vm_xchg_setup(ctx, X);
        VM_ORDER_SWITCH(s->order, vsmem_optin(k_v_moments<K>, ctx, smem);
                        k_v_moments<K><<<grid, threads, smem, ctx->stream>>>(v, np, vcell(s), s->poly, out, ticket, s->moments, X));
        VM_LAUNCHED(ctx);
        if (fuse) return;                             // sums + A1, A2 done by the last CTA
        k_reduce_rows8<<<1, 256, 0, ctx->stream>>>(out, grid, s->moments);
        VM_LAUNCHED(ctx);
        vm_allreduce_sum(ctx, s->moments, 5);
        k_clb_coeffs<<<1, 32, 0, ctx->stream>>>(s->moments, 1);
        VM_LAUNCHED(ctx);
    } else if (!s->lb_coeffs_set) {
        k_clb_coeffs<<<1, 32, 0, ctx->stream>>>(s->moments, 0);   // LB: A1 = 0, A2 = 1 (set once)
        VM_LAUNCHED(ctx);
        s->lb_coeffs_set = true;
    }
}

void rhs_dev(vm_vspline* s, const double* v, const double* w, long np, double nu, int conservative, double* vdot)
{
    vm_ctx* ctx = s->ctx;
    project_dev(s, v, w, np);
    moments_dev(s, v, np, conservative);
    int grid, threads;
    geometry(ctx, &grid, &threads);
    const size_t smem = (size_t)s->ncell * (s->order | 1) * sizeof(double);
    VM_ORDER_SWITCH(s->order, vsmem_optin(k_v_rhs<K>, ctx, smem);
                    k_v_rhs<K><<<grid, threads, smem, ctx->stream>>>(v, np, vcell(s), s->poly, s->moments, nu, vdot));
    VM_LAUNCHED(ctx);
}

double* work_array(vm_particles* p, int i)
{
    if (!p->work[i]) VM_CUDA(cudaMalloc(&p->work[i], (size_t)(p->n > 0 ? p->n : 1) * sizeof(double)));
    return p->work[i];
}

void check_pair(vm_vspline* s, vm_particles* p, const char* who)
{
    if (!s || !p) throw vm_error(VM_ERR_INVALID, std::string(who) + ": NULL handle");
    if (s->ctx != p->ctx) throw vm_error(VM_ERR_INVALID, std::string(who) + ": spline and particles belong to different contexts");
    s->uw = vm_particles_uniform_weight(p, &s->w0) ? 1 : 0;   // one weight for all particles: w is not streamed
}

}  // namespace

extern "C" {

int vm_vspline_create(vm_ctx* ctx, double vmin, double vmax, int nknots, int order, int bc, vm_vspline** out)
{
    VM_API_BEGIN(ctx)
    VM_REQUIRE(ctx != nullptr && out != nullptr, "vm_vspline_create: NULL argument");
    *out = nullptr;
    VM_REQUIRE(vmax > vmin, "vm_vspline_create: empty domain");
    if (order < VM_MIN_ORDER || order > VM_MAX_ORDER) throw vm_error(VM_ERR_UNSUPPORTED, "vm_vspline_create: order must be in 2..6");
    VM_REQUIRE(bc == 0 || bc == 1, "vm_vspline_create: bc must be 0 (none) or 1 (Dirichlet)");
    VM_REQUIRE(nknots >= 2 && nknots + order - 2 <= VM_MAX_NBASIS, "vm_vspline_create: nknots out of range");
    VM_REQUIRE(nknots + order - 2 - (bc ? 2 : 0) >= 1, "vm_vspline_create: empty basis");
    using vmhost::ld;
    vm_vspline* s = new vm_vspline();
    try {
        const int k = order;
        s->ctx = ctx;
        s->device = ctx->device;
        s->a = vmin; s->b = vmax; s->order = k; s->nknots = nknots; s->bc = bc;
        s->ncell = nknots - 1;
        s->npar = nknots + k - 2;
        s->nv = s->npar - (bc ? 2 : 0);
        s->h = (vmax - vmin) / s->ncell;
        const ld h = ((ld)vmax - (ld)vmin) / (ld)s->ncell;
        auto brk = [&](int i) -> ld {
            if (i <= 0) return (ld)vmin;
            if (i >= nknots - 1) return (ld)vmax;
            return (ld)vmin + (ld)i * h;
        };
        // exact polynomial pieces per cell and the parent mass matrix
        std::vector<double> cp((size_t)s->ncell * k * k);
        std::vector<ld> Mpar((size_t)s->npar * s->npar, 0);
        for (int c = 0; c < s->ncell; ++c) {
            ld tl[2 * vmhost::MAXK], A[vmhost::MAXK * vmhost::MAXK];
            for (int i = 0; i < 2 * k; ++i) tl[i] = brk(c + i - k + 1);
            const ld hc = brk(c + 1) - brk(c);
            vmhost::cell_polys(tl, k, brk(c), hc, A);
            for (int j = 0; j < k * k; ++j) cp[(size_t)c * k * k + j] = (double)A[j];
            for (int j1 = 0; j1 < k; ++j1)
                for (int j2 = 0; j2 < k; ++j2)
                    Mpar[(size_t)(c + j1) * s->npar + (c + j2)] += hc * vmhost::poly_dot(A + j1 * k, A + j2 * k, k);
        }
        const int off = bc ? 1 : 0, nv = s->nv;
        std::vector<ld> M((size_t)nv * nv), Minv;
        for (int i = 0; i < nv; ++i)
            for (int j = 0; j < nv; ++j) M[(size_t)i * nv + j] = Mpar[(size_t)(i + off) * s->npar + (j + off)];
        if (!vmhost::spd_banded_inverse(M, nv, k - 1, Minv)) throw vm_error(VM_ERR_INVALID, "vm_vspline_create: mass matrix is not positive definite");
        s->mass.resize((size_t)nv * nv);
        std::vector<double> minv((size_t)nv * nv);
        for (size_t i = 0; i < (size_t)nv * nv; ++i) { s->mass[i] = (double)M[i]; minv[i] = (double)Minv[i]; }

        VM_CUDA(cudaMalloc(&s->rhs, (size_t)s->npar * sizeof(double)));
        VM_CUDA(cudaMalloc(&s->coef, (size_t)s->npar * sizeof(double)));
        VM_CUDA(cudaMalloc(&s->minv, minv.size() * sizeof(double)));
        VM_CUDA(cudaMalloc(&s->cellpoly, cp.size() * sizeof(double)));
        VM_CUDA(cudaMalloc(&s->poly, (size_t)s->ncell * k * sizeof(double)));
        VM_CUDA(cudaMalloc(&s->moments, 16 * sizeof(double)));
        VM_CUDA(cudaMemsetAsync(s->rhs, 0, (size_t)s->npar * sizeof(double), ctx->stream));
        VM_CUDA(cudaMemsetAsync(s->coef, 0, (size_t)s->npar * sizeof(double), ctx->stream));
        VM_CUDA(cudaMemsetAsync(s->poly, 0, (size_t)s->ncell * k * sizeof(double), ctx->stream));
        VM_CUDA(cudaMemsetAsync(s->moments, 0, 16 * sizeof(double), ctx->stream));
        VM_CUDA(cudaMemcpyAsync(s->minv, minv.data(), minv.size() * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
        VM_CUDA(cudaMemcpyAsync(s->cellpoly, cp.data(), cp.size() * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
        VM_CUDA(cudaStreamSynchronize(ctx->stream));
        // the evaluation kernels keep the per-cell polynomial table (row stride order | 1) in shared memory, two CTAs per SM
        VM_REQUIRE((size_t)s->ncell * (k | 1) * sizeof(double) <= (ctx->smem_optin < 100 * 1024 ? ctx->smem_optin : 100 * 1024),
                   "vm_vspline_create: polynomial table exceeds shared memory (too many knots for this order)");
    } catch (...) {
        vm_vspline_destroy(s);
        throw;
    }
    *out = s;
    VM_API_END
}

int vm_vspline_destroy(vm_vspline* s)
{
    if (!s) return VM_OK;
    vm_child_quiesce(s->ctx, s->device);
    cudaFree(s->rhs); cudaFree(s->coef); cudaFree(s->minv); cudaFree(s->cellpoly); cudaFree(s->poly);
    cudaFree(s->moments); cudaFree(s->diag);
    delete s;
    return VM_OK;
}

int vm_vspline_size(vm_vspline* s) { return s ? s->nv : -1; }

int vm_vspline_get_coefficients(vm_vspline* s, double* host)
{
    VM_API_BEGIN(s ? s->ctx : nullptr)
    VM_REQUIRE(s != nullptr && host != nullptr, "vm_vspline_get_coefficients: NULL argument");
    VM_CUDA(cudaMemcpyAsync(host, s->coef + (s->bc ? 1 : 0), (size_t)s->nv * sizeof(double), cudaMemcpyDeviceToHost, s->ctx->stream));
    VM_CUDA(cudaStreamSynchronize(s->ctx->stream));
    VM_API_END
}

int vm_vspline_set_coefficients(vm_vspline* s, const double* host)
{
    VM_API_BEGIN(s ? s->ctx : nullptr)
    VM_REQUIRE(s != nullptr && host != nullptr, "vm_vspline_set_coefficients: NULL argument");
    vm_ctx* ctx = s->ctx;
    VM_CUDA(cudaMemsetAsync(s->coef, 0, (size_t)s->npar * sizeof(double), ctx->stream));
    VM_CUDA(cudaMemcpyAsync(s->coef + (s->bc ? 1 : 0), host, (size_t)s->nv * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
    const int tot = s->ncell * s->order;
    k_v_poly<<<(tot + 255) / 256, 256, 0, ctx->stream>>>(s->coef, s->cellpoly, s->ncell, s->order, s->poly);
    VM_LAUNCHED(ctx);
    VM_CUDA(cudaStreamSynchronize(ctx->stream));
    VM_API_END
}

int vm_vspline_get_rhs(vm_vspline* s, double* host)
{
    VM_API_BEGIN(s ? s->ctx : nullptr)
    VM_REQUIRE(s != nullptr && host != nullptr, "vm_vspline_get_rhs: NULL argument");
    VM_CUDA(cudaMemcpyAsync(host, s->rhs + (s->bc ? 1 : 0), (size_t)s->nv * sizeof(double), cudaMemcpyDeviceToHost, s->ctx->stream));
    VM_CUDA(cudaStreamSynchronize(s->ctx->stream));
    VM_API_END
}

int vm_vspline_get_mass_matrix(vm_vspline* s, double* host)
{
    VM_API_BEGIN(s ? s->ctx : nullptr)
    VM_REQUIRE(s != nullptr && host != nullptr, "vm_vspline_get_mass_matrix: NULL argument");
    std::memcpy(host, s->mass.data(), s->mass.size() * sizeof(double));
    VM_API_END
}

int vm_vproject(vm_vspline* s, vm_particles* p)
{
    VM_API_BEGIN(s ? s->ctx : nullptr)
    check_pair(s, p, "vm_vproject");
    project_dev(s, p->v, p->w, p->n);
    VM_API_END
}

// replacement velocities (stage values of a user-side integrator) go into a scratch array: the particle state
// itself is not touched, as in the reference, where `projection(v, dist, sdist)` only reads dist.particles.w
static double* upload_scratch_v(vm_particles* p, const double* v_host)
{
    double* q = work_array(p, 4);
    if (p->n > 0)
        VM_CUDA(cudaMemcpyAsync(q, v_host, (size_t)p->n * sizeof(double), cudaMemcpyHostToDevice, p->ctx->stream));
    return q;
}

int vm_vproject_at(vm_vspline* s, vm_particles* p, const double* v_host)
{
    VM_API_BEGIN(s ? s->ctx : nullptr)
    check_pair(s, p, "vm_vproject_at");
    VM_REQUIRE(v_host != nullptr || p->n == 0, "vm_vproject_at: v_host is NULL");
    project_dev(s, upload_scratch_v(p, v_host), p->w, p->n);
    VM_CUDA(cudaStreamSynchronize(s->ctx->stream));      // the host array is only borrowed for the call
    VM_API_END
}

int vm_lb_rhs_at(vm_vspline* s, vm_particles* p, const double* v_host, double nu, int conservative, double* vdot_host)
{
    VM_API_BEGIN(s ? s->ctx : nullptr)
    check_pair(s, p, "vm_lb_rhs_at");
    VM_REQUIRE(v_host != nullptr || p->n == 0, "vm_lb_rhs_at: v_host is NULL");
    vm_ctx* ctx = s->ctx;
    if (!p->a) VM_CUDA(cudaMalloc(&p->a, (size_t)(p->n > 0 ? p->n : 1) * sizeof(double)));
    rhs_dev(s, upload_scratch_v(p, v_host), p->w, p->n, nu, conservative, p->a);
    if (vdot_host && p->n > 0) VM_CUDA(cudaMemcpyAsync(vdot_host, p->a, (size_t)p->n * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    VM_CUDA(cudaStreamSynchronize(ctx->stream));
    VM_API_END
}

int vm_vmoments_at(vm_vspline* s, vm_particles* p, const double* v_host, double* out5, double* A2)
{
    VM_API_BEGIN(s ? s->ctx : nullptr)
    check_pair(s, p, "vm_vmoments_at");
    VM_REQUIRE(v_host != nullptr || p->n == 0, "vm_vmoments_at: v_host is NULL");
    vm_ctx* ctx = s->ctx;
    moments_dev(s, upload_scratch_v(p, v_host), p->n, 1);
    double* host = vm_pinned(ctx, 8);
    VM_CUDA(cudaMemcpyAsync(host, s->moments, 8 * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    VM_CUDA(cudaStreamSynchronize(ctx->stream));
    if (out5) for (int i = 0; i < 5; ++i) out5[i] = host[i];
    if (A2) { A2[0] = host[5]; A2[1] = host[6]; }
    VM_API_END
}

int vm_vspline_eval(vm_vspline* s, const double* v_host, long n, double* f_host, double* df_host)
{
    VM_API_BEGIN(s ? s->ctx : nullptr)
    VM_REQUIRE(s != nullptr && (n == 0 || v_host != nullptr), "vm_vspline_eval: NULL argument");
    if (n > 0) {
        vm_ctx* ctx = s->ctx;
        double* buf = nullptr;
        VM_CUDA(cudaMalloc(&buf, (size_t)3 * n * sizeof(double)));
        try {
            VM_CUDA(cudaMemcpyAsync(buf, v_host, (size_t)n * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
            int grid = (int)((n + 255) / 256);
            if (grid > ctx->sm_count * 4) grid = ctx->sm_count * 4;
            const size_t smem = (size_t)s->ncell * (s->order | 1) * sizeof(double);
            VM_ORDER_SWITCH(s->order, vsmem_optin(k_v_eval<K>, ctx, smem);
                            k_v_eval<K><<<grid, 256, smem, ctx->stream>>>(buf, n, vcell(s), s->poly, buf + n, buf + 2 * n));
            VM_LAUNCHED(ctx);
            if (f_host) VM_CUDA(cudaMemcpyAsync(f_host, buf + n, (size_t)n * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
            if (df_host) VM_CUDA(cudaMemcpyAsync(df_host, buf + 2 * n, (size_t)n * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
            VM_CUDA(cudaStreamSynchronize(ctx->stream));
        } catch (...) { cudaFree(buf); throw; }
        VM_CUDA(cudaFree(buf));
    }
    VM_API_END
}

int vm_vmoments(vm_vspline* s, vm_particles* p, double* out5, double* A2)
{
    VM_API_BEGIN(s ? s->ctx : nullptr)
    check_pair(s, p, "vm_vmoments");
    vm_ctx* ctx = s->ctx;
    moments_dev(s, p->v, p->n, 1);
    double* host = vm_pinned(ctx, 8);
    VM_CUDA(cudaMemcpyAsync(host, s->moments, 8 * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    VM_CUDA(cudaStreamSynchronize(ctx->stream));
    if (out5) for (int i = 0; i < 5; ++i) out5[i] = host[i];
    if (A2) { A2[0] = host[5]; A2[1] = host[6]; }
    VM_API_END
}

int vm_lb_rhs(vm_vspline* s, vm_particles* p, double nu, int conservative, double* vdot_host)
{
    VM_API_BEGIN(s ? s->ctx : nullptr)
    check_pair(s, p, "vm_lb_rhs");
    vm_ctx* ctx = s->ctx;
    if (!p->a) VM_CUDA(cudaMalloc(&p->a, (size_t)(p->n > 0 ? p->n : 1) * sizeof(double)));
    rhs_dev(s, p->v, p->w, p->n, nu, conservative, p->a);
    if (vdot_host) {
        if (p->n > 0) VM_CUDA(cudaMemcpyAsync(vdot_host, p->a, (size_t)p->n * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
        VM_CUDA(cudaStreamSynchronize(ctx->stream));
    }
    VM_API_END
}

int vm_lb_rk438_run(vm_vspline* s, vm_particles* p, double dt, int nsteps, double nu, int conservative,
                    int diag_every, double* diag_host)
{
    VM_API_BEGIN(s ? s->ctx : nullptr)
    check_pair(s, p, "vm_lb_rk438_run");
    VM_REQUIRE(nsteps >= 0 && diag_every >= 0, "vm_lb_rk438_run: bad argument");
    VM_REQUIRE(diag_every == 0 || diag_host != nullptr, "vm_lb_rk438_run: diag_host is NULL");
    vm_ctx* ctx = s->ctx;
    const long np = p->n;
    const int nrows = diag_every > 0 ? nsteps / diag_every + 1 : 0;
    if (nrows > s->diag_rows) {
        VM_CUDA(cudaStreamSynchronize(ctx->stream));
        if (s->diag) VM_CUDA(cudaFree(s->diag));
        s->diag = nullptr; s->diag_rows = 0;
        VM_CUDA(cudaMalloc(&s->diag, (size_t)nrows * 4 * sizeof(double)));
        s->diag_rows = nrows;
    }
    int grid, threads;
    geometry(ctx, &grid, &threads);
    int row = 0;
    auto record = [&](double t) {
        double* out = vm_partials(ctx, (size_t)grid * 8);
        k_v_sums<<<grid, threads, 0, ctx->stream>>>(p->v, np, out);
        VM_LAUNCHED(ctx);
        k_reduce_rows8<<<1, 256, 0, ctx->stream>>>(out, grid, s->moments + 8);
        VM_LAUNCHED(ctx);
        vm_allreduce_sum(ctx, s->moments + 8, 2);
        k_v_store_diag<<<1, 32, 0, ctx->stream>>>(s->diag + (size_t)row * 4, t, s->moments + 8);
        VM_LAUNCHED(ctx);
        ++row;
    };
    if (nrows > 0) record(0.0);
    if (nsteps > 0) {      // (an empty shard still runs every pass: it takes part in the other ranks' exchanges)
        double *k1 = work_array(p, 0), *k2 = work_array(p, 1), *k3 = work_array(p, 2), *q = work_array(p, 4);
        // classical 3/8 rule (GeometricIntegrators RK438): a21=1/3; a31=-1/3, a32=1; a41=1, a42=-1, a43=1
        const double a21 = 1.0 / 3.0, a31 = -1.0 / 3.0;
        if (ctx->no_fuse) {
            // reference-shaped driver: separate projection / moments / RHS / stage-assembly kernels
            double* k4 = work_array(p, 3);
            for (int st = 1; st <= nsteps; ++st) {
                rhs_dev(s, p->v, p->w, np, nu, conservative, k1);
                k_rk_combine<<<grid, threads, 0, ctx->stream>>>(p->v, k1, nullptr, nullptr, nullptr, a21, 0, 0, 0, dt, np, q);
                VM_LAUNCHED(ctx);
                rhs_dev(s, q, p->w, np, nu, conservative, k2);
                k_rk_combine<<<grid, threads, 0, ctx->stream>>>(p->v, k1, k2, nullptr, nullptr, a31, 1.0, 0, 0, dt, np, q);
                VM_LAUNCHED(ctx);
                rhs_dev(s, q, p->w, np, nu, conservative, k3);
                k_rk_combine<<<grid, threads, 0, ctx->stream>>>(p->v, k1, k2, k3, nullptr, 1.0, -1.0, 1.0, 0, dt, np, q);
                VM_LAUNCHED(ctx);
                rhs_dev(s, q, p->w, np, nu, conservative, k4);
                k_rk_combine<<<grid, threads, 0, ctx->stream>>>(p->v, k1, k2, k3, k4, 0.125, 0.375, 0.375, 0.125, dt, np, p->v);
                VM_LAUNCHED(ctx);
                if (diag_every > 0 && st % diag_every == 0) record(dt * st);
            }
        } else {
            // fused driver: one particle pass per stage (RHS + stage assembly + next projection's deposit),
            // plus the 8 B/particle moments pass for the conservative operator
            project_dev(s, p->v, p->w, np);
            moments_dev(s, p->v, np, conservative);
            const size_t extra = (size_t)s->ncell * (s->order | 1);    // polynomial table lives in shared memory too
            auto stage = [&](const double* pk1, const double* pk2, const double* pk3, double a1, double a2, double a3,
                             double c1, double c2, double c3, double cs, double* kout, double* qout) {
                StageParams S{};
                S.k1 = pk1; S.k2 = pk2; S.k3 = pk3;
                S.a1 = a1; S.a2 = a2; S.a3 = a3;
                S.c1 = c1; S.c2 = c2; S.c3 = c3; S.cs = cs;
                S.kout = kout; S.dt = dt; S.nu = nu;
                S.uw = s->uw; S.w0 = s->w0;
                S.qout = (conservative || qout == p->v) ? qout : nullptr;   // LB needs no stored stage state
                VDepSetup d = vdep_setup(s, (int)extra);
                if (d.pl.var == VAR_ATOMIC) throw vm_error(VM_ERR_UNSUPPORTED, "fused RK438 stage: no atomic deposit variant");
                const int stage_no = pk3 ? 4 : (pk2 ? 3 : (pk1 ? 2 : 1));
                vm_prof_mark(ctx);
                VM_ORDER_SWITCH(s->order, launch_stage<K>(s, d, p->v, p->w, np, S, stage_no));
                vm_prof_mark(ctx);
                after_deposit(s, d);
                moments_dev(s, qout, np, conservative);
            };
            for (int st = 1; st <= nsteps; ++st) {
                stage(nullptr, nullptr, nullptr, 0, 0, 0, 0, 0, 0, a21, k1, q);                    // k1, q2
                stage(k1, nullptr, nullptr, a21, 0, 0, a31, 0, 0, 1.0, k2, q);                      // k2, q3
                stage(k1, k2, nullptr, a31, 1.0, 0, 1.0, -1.0, 0, 1.0, k3, q);                      // k3, q4
                stage(k1, k2, k3, 1.0, -1.0, 1.0, 0.125, 0.375, 0.375, 0.125, nullptr, p->v);       // v_new (+ its projection)
                if (diag_every > 0 && st % diag_every == 0) record(dt * st);
            }
        }
    }
    if (nrows > 0) {
        double* host = vm_pinned(ctx, (size_t)nrows * 4);
        VM_CUDA(cudaMemcpyAsync(host, s->diag, (size_t)row * 4 * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
        VM_CUDA(cudaStreamSynchronize(ctx->stream));
        for (int i = 0; i < nrows * 4; ++i) diag_host[i] = i < row * 4 ? host[i] : 0.0;
    }
    VM_API_END
}

}  // extern "C"

// vm_internal.cuh -- shared host/device internals of libvlasov_b200.so (sm_100a, fp64).
#pragma once

#include <cuda_runtime.h>

#include <cstdint>
#include <cstdio>
#include <stdexcept>
#include <string>
#include <vector>

#include "vlasov_b200.h"

#define VM_MIN_ORDER 2
#define VM_MAX_ORDER 6
#define VM_MAX_NBASIS 4096
#define VM_FULL_MASK 0xffffffffu
#define VM_DIAG_COLS 4          // extra columns appended to a partial row: [sum w v^2, sum w v, sum w, spare]
#define VM_MAX_EVENTS 16
#define VM_MAX_PEERS 8          // ranks of the fused peer-memory exchange (one NVSwitch node)
#define VM_X_MAX_N 1024         // largest vector (basis functions) the fused peer exchange carries
// Inbox of the fused exchange: [2 sets][VM_MAX_PEERS ranks][VM_XSLOT_WORDS] 64-bit words.  Every double travels as
// two words {low 32 data bits | seq << 32}, {high 32 data bits | seq << 32} ("LL" protocol: data and flag arrive in
// the same 8-byte store, so the sender needs no fence and the receiver no separate flag).
#define VM_XSLOT_WORDS (2 * (VM_X_MAX_N + 8))
#define VM_XINBOX_WORDS (2 * VM_MAX_PEERS * VM_XSLOT_WORDS)
#define VM_MAX_GROUPS 63        // groups of CTAs in the two-level cross-CTA reduction (tickets 1 .. VM_MAX_GROUPS)
#define VM_GROUP_CTAS 16        // CTAs per group

struct vm_error : public std::runtime_error {
    int code;
    vm_error(int c, const std::string& m) : std::runtime_error(m), code(c) {}
};

#define VM_CUDA(expr)                                                                         \
    do {                                                                                      \
        cudaError_t e__ = (expr);                                                             \
        if (e__ != cudaSuccess)                                                               \
            throw vm_error(VM_ERR_CUDA, std::string(#expr) + ": " + cudaGetErrorString(e__)); \
    } while (0)

#define VM_REQUIRE(cond, msg)                                   \
    do {                                                        \
        if (!(cond)) throw vm_error(VM_ERR_INVALID, (msg));     \
    } while (0)

struct vm_ctx {
    int device = 0;
    cudaStream_t stream = nullptr;
    int sm_count = 0, cc_major = 0, cc_minor = 0;
    size_t smem_optin = 0;
    std::string last_error;
    unsigned long long launches = 0;
    cudaEvent_t events[VM_MAX_EVENTS] = {};
    cudaStream_t copy_stream = nullptr;      // device->host snapshot copies overlap with compute
    cudaEvent_t snap_ready = nullptr, snap_done = nullptr;
    // tuning (0 = auto)
    int ctas_per_sm = 0, threads_per_cta = 0, replicas = 0, profile = 0, no_fuse = 0, no_pdl = 0, force_match = 0, no_uniform_w = 0;
    int bankq = 0;      // bank-sorted large-mesh pass: 0 = auto (n >= 256), 1 = always, -1 = never
    int no_presolve = 0;       // 1: meshes above 128 cells keep the separate k_poisson_solve launch between fused passes (A/B)
    int af_replicas = 0;       // bank-steered replicas per CTA of the limb-atomic pass (0 = as many as fit, <= 32)
    int af = 0, af_ctas = 0;   // limb-atomic fixed-point pass: 0 = auto (from VM_AF_MIN_N cells), 1 = always, -1 = never; CTAs per SM (0 = auto)
    int pairs = 0, priv_min_warps = 0, no_repg = 0;   // pairs in flight per thread / fewest warps the lane-private deposit accepts
    // per-launch event brackets of the dominant kernel (profile == 1)
    std::vector<cudaEvent_t> prof_events;   // pairs: [2i] start, [2i+1] stop
    size_t prof_used = 0;                    // events in use since the last read
    // communicator (one process per GPU)
    void* nccl_comm = nullptr;
    int rank = 0, nranks = 1;
    unsigned* ticket = nullptr;          // device counters for the last-CTA finish: [0] grid, [1 + g] group g (zero between launches)
    // fused peer-memory exchange (vm_ctx_peer_connect): inbox = [2 sets][VM_MAX_PEERS][VM_XSLOT_WORDS] u64
    unsigned long long* inbox = nullptr;
    unsigned long long* peer_inbox[VM_MAX_PEERS] = {};
    bool peers_connected = false;
    unsigned long long xseq = 0;         // exchange sequence number (identical on all ranks)
    unsigned* xerr = nullptr;            // device: set when a peer wait timed out
    // scratch: per-CTA partial rows for the fixed-order reductions
    double* partials = nullptr;
    size_t partials_elems = 0;
    // pinned host staging for small read-backs
    double* pinned = nullptr;
    size_t pinned_elems = 0;
};

struct vm_particles {
    vm_ctx* ctx = nullptr;
    int device = 0;
    long n = 0;
    double *x = nullptr, *v = nullptr, *w = nullptr;
    double* a = nullptr;        // lazily allocated per-particle work array (E / vdot)
    double* work[5] = {};       // lazily allocated RK stage arrays
    double* snap = nullptr;     // lazily allocated staging copy of x and v for asynchronous snapshots
    bool snap_pending = false;
    // weight class: every sampler of the reference produces ONE weight for all particles (w = 1/N or L/N);
    // then the passes take w0 from the parameter block instead of streaming 8 B/particle from HBM
    bool w_dirty = true;        // w changed since the last classification
    bool uniform_w = false;
    double w0 = 0.0;
    // fixed-point scale of VM_DEPOSIT_FIXED (vm_particles_fixed_scale): valid until the weights or the communicator change
    bool fix_dirty = true;
    int fix_S = 0;
    bool fix_ok = false;        // the weights admit a fixed-point scale (finite, max |w| < 1e300)
    int fix_nranks = 0;
};

// Map a position to (first basis index, xi) on a uniform periodic grid.
struct CellMap {
    double inv_h;     // n / (b - a)
    double off;       // -a * inv_h
    int n;            // number of cells == number of periodic basis functions
    int bias;         // multiple of n plus index shift; makes the raw cell index positive
    unsigned inv_n;   // floor(2^32 / n), for the fast modulo
    int mask;         // n - 1 when n is a power of two (then (c + shift) & mask replaces the modulo), else -1
    int shift;        // index rotation reduced to [0, n)
};

struct vm_field {
    vm_ctx* ctx = nullptr;
    int device = 0;
    double a = 0, b = 0, h = 0;
    int order = 0, n = 0, shift = 0;
    CellMap map{};
    double *rhs = nullptr, *phi = nullptr, *dcoef = nullptr;  // device, n each
    bool rhs_global = true;                                   // rhs holds the sum over all ranks (false: this rank's deposit only)
    // meshes above 128 cells inside vm_vp_run: the solve of a deposit is done by the first CTAs of the NEXT fused pass
    // (vm_pass.cuh: pass_presolve) instead of a kernel of its own; true while rhs is newer than phi / dcoef
    bool solve_pending = false;
    unsigned* solve_count = nullptr;                          // device: tiles solved so far (monotonic)
    unsigned solve_target = 0;                                // host mirror: value of *solve_count after the last presolve
    double *ext_phi = nullptr, *ext_dcoef = nullptr;          // device: coefficient history of vm_vp_run_external (n x ext_cols)
    int ext_cols = 0;
    double *G = nullptr;                                      // device: circulant pseudo-inverse kernel (first column)
    double *stencil_s = nullptr;                              // device: stiffness stencil, 2k-1 entries
    double *diag = nullptr;                                   // device: [W, K, M, sum_w] rows
    int diag_rows = 0;
    double mass_st[VM_MAX_ORDER] = {}, stiff_st[VM_MAX_ORDER] = {};
};

struct vm_vspline {
    vm_ctx* ctx = nullptr;
    int device = 0;
    double a = 0, b = 0, h = 0;
    int order = 0, nknots = 0, ncell = 0, npar = 0, nv = 0, bc = 1;
    std::vector<double> mass;          // host nv x nv
    double *rhs = nullptr;             // device npar (parent indexing)
    double *coef = nullptr;            // device npar (parent indexing, zeros at dropped ends)
    double *minv = nullptr;            // device nv x nv dense inverse of the mass matrix
    double *cellpoly = nullptr;        // device ncell x k x k: A[c][j][m], B_{c+j} = sum_m A xi^m
    double *poly = nullptr;            // device ncell x k: f_s on cell c = sum_m poly[c][m] xi^m
    double *moments = nullptr;         // device 8: [5 sums, A1, A2, spare]
    double *diag = nullptr;            // device rows [t, sum v, sum v^2, spare]
    int diag_rows = 0;
    int uw = 0;                        // weight class of the particles of the current call (see vm_particles)
    double w0 = 0.0;
    bool lb_coeffs_set = false;        // moments[5..6] currently hold the LB constants (A1 = 0, A2 = 1)
};

void vm_set_error(vm_ctx* ctx, const std::string& msg);
void vm_use(vm_ctx* ctx);                         // cudaSetDevice
// Contexts are registered while they exist.  Host languages finalise handles in no particular order (Julia and
// Python at exit), so a child handle can outlive its context: destroying it then must not touch the context, and
// any other call on it reports VM_ERR_INVALID instead of dereferencing freed memory.
bool vm_ctx_alive(vm_ctx* ctx);
void vm_child_quiesce(vm_ctx* ctx, int device);   // device idle for this child: its context's streams, or the whole device
double* vm_partials(vm_ctx* ctx, size_t elems);   // grow-only device scratch
double* vm_pinned(vm_ctx* ctx, size_t elems);     // grow-only pinned host scratch
void vm_allreduce_sum(vm_ctx* ctx, double* dev, size_t count);  // no-op when nranks == 1
void vm_allreduce_max(vm_ctx* ctx, double* dev, size_t count);  // no-op when nranks == 1
int vm_particles_fixed_scale(vm_particles* p);                  // S of VM_DEPOSIT_FIXED: contributions are rounded to multiples of 2^-S
void vm_launch_geometry(vm_ctx* ctx, int* grid, int* threads);
void vm_prof_mark(vm_ctx* ctx);
void vm_check_peer_error(vm_ctx* ctx);
bool vm_particles_uniform_weight(vm_particles* p, double* w0);   // classifies w on first use after a change   // after a stream sync: throws if a peer exchange timed out   // records the next event of a start/stop pair when profiling is on

#define VM_API_BEGIN(ctxexpr)                                                                              \
    vm_ctx* ctx__ = (ctxexpr);                                                                             \
    try {                                                                                                  \
        if (ctx__ && !vm_ctx_alive(ctx__)) {                                                               \
            ctx__ = nullptr;                                                                               \
            throw vm_error(VM_ERR_INVALID, "the context of this handle has already been destroyed");       \
        }                                                                                                  \
        if (ctx__) vm_use(ctx__);

#define VM_API_END                                        \
    }                                                     \
    catch (const vm_error& e) {                           \
        vm_set_error(ctx__, e.what());                    \
        return e.code;                                    \
    }                                                     \
    catch (const std::bad_alloc&) {                       \
        vm_set_error(ctx__, "host allocation failed");    \
        return VM_ERR_NOMEM;                              \
    }                                                     \
    catch (const std::exception& e) {                     \
        vm_set_error(ctx__, e.what());                    \
        return VM_ERR_INVALID;                            \
    }                                                     \
    return VM_OK;

#define VM_LAUNCHED(ctx)                       \
    do {                                       \
        ++(ctx)->launches;                     \
        VM_CUDA(cudaGetLastError());           \
    } while (0)

// ------------------------------------------------------------------ device --
#ifdef __CUDACC__

// K nonzero uniform B-splines of order K at local coordinate xi in [0,1), increasing index, times a
// common factor w (the particle weight; pass 1.0 for plain values).  Orders 2-4 use closed forms that
// exploit the symmetry B_j(xi) = B_{K-1-j}(1 - xi) (14 fp64 instructions for the weighted cubic instead
// of 25 for the recursion); orders 5 and 6 use the de Boor recursion specialised to unit knot spacing.
template <int K>
__device__ __forceinline__ void bspline_uniform_w(double xi, double w, double (&N)[K])
{
    if (K == 1) {
        N[0] = w;
    } else if (K == 2) {
        N[1] = xi * w;
        N[0] = w - N[1];
    } else if (K == 3) {
        const double om = 1.0 - xi, wh = 0.5 * w;
        N[0] = (om * om) * wh;
        N[2] = (xi * xi) * wh;
        N[1] = fma(xi, om, 0.5) * w;
    } else if (K == 4) {
        const double om = 1.0 - xi, w6 = w * (1.0 / 6.0);
        const double x2 = xi * xi, o2 = om * om;
        N[0] = (o2 * om) * w6;
        N[3] = (x2 * xi) * w6;
        N[1] = fma(x2, fma(0.5, xi, -1.0), 2.0 / 3.0) * w;
        N[2] = fma(o2, fma(0.5, om, -1.0), 2.0 / 3.0) * w;
    } else {
        N[0] = w;
#pragma unroll
        for (int j = 1; j < K; ++j) {
            const double inv = 1.0 / (double)j;
            double saved = 0.0;
#pragma unroll
            for (int r = 0; r < j; ++r) {
                const double temp = N[r] * inv;
                const double right = (double)(r + 1) - xi;
                const double left = xi + (double)(j - r - 1);
                N[r] = fma(right, temp, saved);
                saved = left * temp;
            }
            N[j] = saved;
        }
    }
}

template <int K>
__device__ __forceinline__ void bspline_uniform(double xi, double (&N)[K])
{
    bspline_uniform_w<K>(xi, 1.0, N);
}

// (x - a)/h -> (cell index mod n incl. the index rotation, fractional coordinate in [0,1)).
// Two floor() flavours, chosen per kernel from A/B measurements on B200 (profiles/):
//  CONV = true : FRND.F64.FLOOR + F2I.F64 -- 2 instructions, but on the quarter-rate conversion pipe;
//                best where issue slots are scarce and there is one cell lookup per particle (deposit).
//  CONV = false: adding 1.5*2^52 leaves round-to-nearest(t) in the low mantissa word and one compare
//                turns the rounding into a floor -- 5 full-rate fp64 instructions; best for the fused
//                push+deposit pass (two lookups per particle saturate the conversion pipe).
// Valid for |t| < 2^30; the result is clamped in-bounds for garbage positions.
template <bool CONV, bool POW2 = false>
__device__ __forceinline__ void cell_of(const CellMap& m, double x, int& base, double& xi)
{
    const double t = fma(x, m.inv_h, m.off);
    int c;
    if (CONV) {
        const double r = floor(t);
        c = __double2int_rd(t);
        xi = t - r;
    } else {
        const double big = 6755399441055744.0;             // 1.5 * 2^52
        const double tm = t + big;
        double r = tm - big;                                // round-to-nearest-even(t), exact
        c = __double2loint(tm);
        if (r > t) { r -= 1.0; c -= 1; }
        xi = t - r;
    }
    if (POW2) {                                             // power-of-two grid: the modulo is a mask
        base = (c + m.shift) & m.mask;
    } else {
        const unsigned u = (unsigned)(c + m.bias);          // in [0, 2^31)
        const unsigned q = __umulhi(u, m.inv_n);
        unsigned rr = u - q * (unsigned)m.n;
        if (rr >= (unsigned)m.n) rr -= (unsigned)m.n;
        base = (int)min(rr, (unsigned)m.n - 1u);            // in-bounds even for garbage positions (|t| >= 2^30)
    }
}

__device__ __forceinline__ int wrap_add(int i, int j, int n)
{
    int r = i + j;
    return r >= n ? r - n : r;
}

// streaming 16-byte loads/stores (read once / written once per pass).  Loads do not allocate in L1
// (LDG.E.NA): the replica grids of mid-size meshes take up to 227 KB of the 256 KB L1/shared array, and what is
// left should not be churned by data that is never re-read.  Measured against the evict-first form
// (ld.global.cs, -DVM_LD_CS): -4 % at n_h = 32, -3 % at n_h = 64, equal at n_h = 16 / 128 and in the v-space passes.
#ifdef VM_LD_CS
#define VM_LD_STREAM "ld.global.cs"
#else
#define VM_LD_STREAM "ld.global.L1::no_allocate"
#endif
__device__ __forceinline__ double2 ld_stream2(const double* p)
{
    double2 r;
    asm volatile(VM_LD_STREAM ".v2.f64 {%0, %1}, [%2];" : "=d"(r.x), "=d"(r.y) : "l"(p));
    return r;
}
__device__ __forceinline__ void st_stream2(double* p, double2 v)
{
    asm volatile("st.global.cs.v2.f64 [%0], {%1, %2};" ::"l"(p), "d"(v.x), "d"(v.y) : "memory");
}
__device__ __forceinline__ double ld_stream(const double* p)
{
    double r;
    asm volatile(VM_LD_STREAM ".f64 %0, [%1];" : "=d"(r) : "l"(p));
    return r;
}
__device__ __forceinline__ void st_stream(double* p, double v)
{
    asm volatile("st.global.cs.f64 [%0], %1;" ::"l"(p), "d"(v) : "memory");
}

// L2 prefetch of the 128-byte line holding p (no register is tied up, unlike a software-pipelined load)
__device__ __forceinline__ void prefetch_l2(const void* p)
{
    asm volatile("prefetch.global.L2 [%0];" ::"l"(p));
}

// fixed-order warp tree sum (xor butterfly: every lane ends with the same bits)
__device__ __forceinline__ double warp_sum(double v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(VM_FULL_MASK, v, o);
    return v;
}

// phi = G (*) (rhs - mean(rhs))  (circular convolution with the pseudo-inverse kernel G), and
// dcoef_m = (phi_{m+1} - phi_m) / h, the coefficients of the derivative spline (E = -phi'): 32 outputs (one tile) by the
// first 256 threads of the calling CTA; ALL threads of the CTA must call (it synchronises the CTA).
// Each lane accumulates phi_i and phi_{i+1} with the same chunking over j, so the phi_{i+1} it uses
// is bit-identical to the phi_{i+1} its neighbour stores: dcoef is a pure function of the stored
// phi (same bits as k_dcoef_from_phi), as the reference's ExternalField/PoissonField test demands.
// Every tile recomputes the mean in the same fixed order, so all tiles (and all ranks) agree bitwise.
// r: n doubles, red: 2 * 8 * 33 doubles, wsum: 8 doubles of shared memory.  rhs is read with ld.cg (it may have been
// written by another SM's last CTA in the kernel before this one, or exchanged over NVLink).
__device__ __forceinline__ void poisson_solve_tile(int tile, const double* __restrict__ rhs, const double* __restrict__ G,
                                                   int n, double inv_h, double* __restrict__ phi, double* __restrict__ dcoef,
                                                   double* __restrict__ r, double* __restrict__ red, double* __restrict__ wsum)
{
    const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
    const bool work = t < 256;
    if (work) {
        double s = 0.0;
        for (int j = t; j < n; j += 256) s += __ldcg(rhs + j);
        s = warp_sum(s);
        if (lane == 0) wsum[warp] = s;
    }
    __syncthreads();
    double mean = 0.0;
#pragma unroll
    for (int q = 0; q < 8; ++q) mean += wsum[q];
    mean /= (double)n;
    if (work)
        for (int j = t; j < n; j += 256) r[j] = __ldcg(rhs + j) - mean;
    __syncthreads();
    const int i = tile * 32 + lane;
    if (work) {
        const int len = (n + 7) / 8;
        const int j0 = warp * len, j1 = min(n, j0 + len);
        double a0 = 0.0, a1 = 0.0;             // partial sums of phi_i and phi_{i+1}
        if (i < n) {
            int idx = i - j0;
            if (idx < 0) idx += n;
            double g1 = __ldg(G + (idx + 1 == n ? 0 : idx + 1));
            for (int j = j0; j < j1; ++j) {
                const double rj = r[j];
                const double g0 = __ldg(G + idx);
                a0 = fma(g0, rj, a0);
                a1 = fma(g1, rj, a1);
                g1 = g0;
                idx = (idx == 0) ? n - 1 : idx - 1;
            }
        }
        red[(0 * 8 + warp) * 33 + lane] = a0;
        red[(1 * 8 + warp) * 33 + lane] = a1;
    }
    __syncthreads();
    if (warp == 0 && i < n) {
        double p0 = 0.0, p1 = 0.0;
#pragma unroll
        for (int q = 0; q < 8; ++q) { p0 += red[(0 * 8 + q) * 33 + lane]; p1 += red[(1 * 8 + q) * 33 + lane]; }
        phi[i] = p0;
        dcoef[i] = (p1 - p0) * inv_h;
    }
}
#define VM_SOLVE_SCRATCH_DOUBLES(n) ((n) + 2 * 8 * 33 + 8)      // r + red + wsum

#endif  // __CUDACC__

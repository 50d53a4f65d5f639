// vm_pass_bq.cuh -- the fused x-space particle pass for LARGE meshes: bank-sorted deposition ("BANKQ").
//
// Why another layout.  A mesh of n_h >= 256 cells leaves no room for lane-private replica grids (32 copies per warp),
// and with fewer copies the K read-modify-writes of a warp of particles hit random rows: ~3-way bank conflicts on every
// LDS/STS (64-bit accesses are served per half-warp: 16 lanes into 16 bank pairs), plus collision handling
// (MATCH.ANY grouping + segmented reduce).  ncu (round 1): 72-75 shared-memory wavefronts per warp of particles,
// 1.1-1.2 ms per step at 1e8 particles = 41-44 % of the HBM roofline.
//
// The layout here: ONE replica grid per warp, and every lane only ever deposits particles whose first basis row b0
// satisfies b0 = lane (mod 32).  Then in tap j of a warp-wide read-modify-write lane l touches row b0_l + j, whose
// bank pair is (l + j) mod 16: the 16 lanes of a half-warp hit 16 distinct bank pairs -- conflict-free by construction,
// and two lanes can never touch the same row in the same tap (that would need b0_l = b0_l', i.e. l = l' mod 32).
// No collision detection, no shuffles, 2 wavefronts per LDS/STS: 16 per warp of particles, the lane-private cost.
// Particles reach "their" lane through 32 small FIFO queues per warp in shared memory (class c = b0 mod 32): the lane
// that computed a particle appends {xi, b0 / 32} packed into ONE 64-bit word to queue c (a conflicted STS, ~3-way), and
// lane c pops from queue c (entry-major layout: a warp-wide pop is conflict-free).  Arrivals per class per batch are
// Binomial(32, 1/32); a queue without room takes what fits and forces an extra pop round (lane efficiency 86 % at 16
// entries per class, tools/lsu_model.py --bankq).  The taps of one pop round are ordered by __syncwarp (tap j of one lane and tap
// j + 1 of its neighbour may be the same row); rounds are ordered by program order.
//
// Summation order: particle -> (CTA, warp, queue position) is static, so the result is bit-reproducible run to run
// like the other variants (different bits from them: a different order).
#pragma once
#include "vm_pass.cuh"

#ifndef VM_BQ_CAP_LOG2
#define VM_BQ_CAP_LOG2 4                       // entries per class queue (16)
#endif
// Most shared memory spent on copies of the gather table.  Measured (profiles/r02_bankq_ab.txt): up to 256 cells the
// 16 conflict-free copies pay (0.827 vs 0.848 ms per step); from 512 cells on the warps they displace are worth more
// than the bank conflicts they remove (512: 4 copies / 26 warps 0.893 ms vs 8 copies / 23 warps 0.956 ms; 1024: 2 copies /
// 17 warps 0.966 ms vs 4 copies / 15 warps 0.995 ms vs 1 copy / 18 warps 1.052 ms).
#ifndef VM_BQ_TABLE_BYTES
#define VM_BQ_TABLE_BYTES(n) ((n) <= 256 ? VM_GATHER_TABLE_MAX_BYTES : 17000)
#endif
#define VM_BQ_CAP (1 << VM_BQ_CAP_LOG2)
#define VM_BQ_QWORDS (32 * VM_BQ_CAP)          // 64-bit words per warp queue array
#define VM_BQ_TAILWORDS 16                     // 32 x u32 queue tails per warp (as 64-bit words)
#define VM_BQ_SINKWORDS 56                     // per-warp sink row (32 + K - 1 doubles + 15 of alignment slack): zeros of idle lanes land here
// The pass is instruction-issue bound (ncu: 76 % issue-active, 298 warp instructions per warp of particles in its first
// version), so the cell lookup uses the 2-instruction FRND/F2I form (quarter-rate conversion pipe, idle here) rather
// than the 5-instruction magic-number floor of the HBM-bound lane-private pass.
#ifndef VM_BQ_CONV
#define VM_BQ_CONV true
#endif
#ifndef VM_BQ_MAXW
#define VM_BQ_MAXW 32                          // most warps per CTA (register budget per thread = 65536 / (32 * warps))
#endif

struct BqParams {
    int gshift;        // log2(copies of the gather table): entry m of copy c at dsh[(m << gshift) + c], lane l reads copy l & (copies - 1)
    int per_warp;      // 64-bit words of shared memory per warp after the replica grids: tails + queue (+ queue of weights when they are streamed)
};

// 32-bit shared-memory addressing: the generic-pointer forms cost a CTA-window computation (S2UR CgaCtaId, ULEA) and
// 64-bit address arithmetic per access, and this pass is instruction-issue bound.
__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ double lds_f64(unsigned a)
{
    double v;
    asm volatile("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"(a) : "memory");
    return v;
}
__device__ __forceinline__ void sts_f64(unsigned a, double v) { asm volatile("st.shared.f64 [%0], %1;" ::"r"(a), "d"(v) : "memory"); }
__device__ __forceinline__ unsigned long long lds_u64(unsigned a)
{
    unsigned long long v;
    asm volatile("ld.shared.u64 %0, [%1];" : "=l"(v) : "r"(a) : "memory");
    return v;
}
__device__ __forceinline__ void sts_u64(unsigned a, unsigned long long v) { asm volatile("st.shared.u64 [%0], %1;" ::"r"(a), "l"(v) : "memory"); }
__device__ __forceinline__ unsigned lds_u32(unsigned a)
{
    unsigned v;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(a) : "memory");
    return v;
}
__device__ __forceinline__ void sts_u32(unsigned a, unsigned v) { asm volatile("st.shared.u32 [%0], %1;" ::"r"(a), "r"(v) : "memory"); }

// phi'(x) from the (replicated) table with a run-time copy count
// (s_tab = shared address of the table + (lane & (copies - 1)) * 8, stride_b = 8 << gshift bytes)
template <int K>
__device__ __forceinline__ double gather_dphi_rt(unsigned s_tab, int b0, double xi, int gshift, int stride_b)
{
    double Nd[K - 1 > 0 ? K - 1 : 1];
    bspline_uniform<(K - 1 > 0 ? K - 1 : 1)>(xi, Nd);
    unsigned d = s_tab + ((unsigned)b0 << (gshift + 3));
    double t[K - 1 > 0 ? K - 1 : 1];
#pragma unroll
    for (int j = 0; j < K - 1; ++j) {
        t[j] = lds_f64(d);
        d += stride_b;
    }
    double s = 0.0;
#pragma unroll
    for (int j = 0; j < K - 1; ++j) s = fma(Nd[j], t[j], s);
    return s;
}

// xi in [0, 1) is a multiple of 2^-52 (it is t - floor(t) with |t| >= 1; cell 0 loses sub-2^-52 bits, <= 1.1e-16
// absolute), so xi + 1 carries it exactly in its mantissa field: 52 bits of xi + 12 bits for b0 >> 5 in one word.
__device__ __forceinline__ unsigned long long bq_pack(double xi, int hi)
{
    xi = fmin(xi, 0x1.fffffffffffffp-1);                        // xi == 1.0 (t a hair below an integer) would alias to 0
    const unsigned long long bits = (unsigned long long)__double_as_longlong(xi + 1.0);
    return (bits << 12) | (unsigned long long)hi;
}
__device__ __forceinline__ void bq_unpack(unsigned long long word, double& xi, unsigned& hi)
{
    hi = (unsigned)word & 0xfffu;
    xi = __longlong_as_double((long long)((word >> 12) | 0x3ff0000000000000ull)) - 1.0;
}

// UW: all particles carry the weight P.w0 (no weight queue, no weight stream).
// FIXED: 64-bit fixed-point accumulation (VM_DEPOSIT_FIXED), see vm_deposit.cuh.
template <int K, int MODE, bool SPLIT, bool POW2, bool UW, bool FIXED, int MAXT>
__global__ void __launch_bounds__((MAXT == 1024 ? VM_BQ_MAXW * 32 : MAXT), 1)
k_vp_pass_bq(double* __restrict__ x, double* __restrict__ v, const double* __restrict__ w,
             const double* __restrict__ dcoef, double* __restrict__ out, const PassParams P, const FinishParams F,
             const BqParams Q)
{
    extern __shared__ double smem[];
    const int n = P.map.n;
    constexpr int GHOST = K - 1;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
    double* dsh = smem;
    double* grid = smem + (MODE == MODE_PUSH_DEPOSIT ? ((n + K) << Q.gshift) : 0);
    const int gsz = n + GHOST;                        // one replica per warp
    const int gtotal = gsz * nwarps;
    unsigned long long* wbase = (unsigned long long*)(grid + gtotal) + (size_t)warp * Q.per_warp;
    unsigned* tails = (unsigned*)wbase;               // tails[c]: entries ever queued into class c (written by one lane per batch)
    double* sink = (double*)(wbase + VM_BQ_TAILWORDS);   // 32 + K - 1 doubles: where lanes without an entry add their zeros
    unsigned long long* qbase = wbase + VM_BQ_TAILWORDS + VM_BQ_SINKWORDS;
    double* scratch = grid + gtotal + (size_t)nwarps * Q.per_warp;
    for (int i = threadIdx.x; i < gtotal; i += blockDim.x) grid[i] = 0.0;
    tails[lane] = 0u;
    sink[lane] = 0.0;
    if (lane < VM_BQ_SINKWORDS - 32) sink[32 + lane] = 0.0;
    asm volatile("griddepcontrol.wait;" ::: "memory");
    if (MODE == MODE_PUSH_DEPOSIT) {
        const int ext = K - 2 > 0 ? K - 2 : 0, copies = 1 << Q.gshift;
        for (int i = threadIdx.x; i < (n + ext) * copies; i += blockDim.x) {
            const int m = i >> Q.gshift;
            dsh[i] = dcoef[m < n ? m : m - n];
        }
    }
    __syncthreads();
    const double w0 = P.w0;
    const int gshift = Q.gshift, stride_b = 8 << gshift;
    const unsigned s_tab = smem_u32(dsh) + (lane & ((1 << gshift) - 1)) * 8;
    const unsigned s_rows = smem_u32(grid + warp * gsz) + lane * 8;      // row (hi << 5 | lane) of my warp's grid at s_rows + hi * 256
    // the sink slot of lane l must sit in the SAME bank pair as the rows lane l owns (rows_base + l + j): an idle lane's
    // access then never conflicts with an active lane's (ncu, first version: 37 % of all shared wavefronts of the pass were
    // such 2-way conflicts) -- shift the sink row by the bank-pair offset between the two regions
    const unsigned s_sink = smem_u32(sink) + lane * 8 + ((((s_rows - lane * 8) - smem_u32(sink)) >> 3) & 15u) * 8;
    const unsigned s_tails = smem_u32(tails);
    const unsigned s_q = smem_u32(qbase);                                // entry e of class c at s_q + (e & (CAP - 1)) * 256 + c * 8
    constexpr unsigned WOFF = VM_BQ_QWORDS * 8;                          // byte offset of the weight queue

    // lane c owns class c: it alone pops queue c (head = entries it has consumed); writers advance tails[c]
    unsigned head = 0u;
    const unsigned lt_mask = (1u << lane) - 1u;
    unsigned lane_inv[5];                              // all ones where bit b of the lane index is 0
#pragma unroll
    for (int bit = 0; bit < 5; ++bit) lane_inv[bit] = ((lane >> bit) & 1) ? 0u : VM_FULL_MASK;

    // Pop up to TWO entries per lane and apply their read-modify-writes tap by tap: the two chains LDS -> DADD -> STS of
    // a tap are independent (different rows: the two entries of one class differ in b0 >> 5, and equal ones are merged
    // first).  No predication: a lane without an entry adds zeros to its own slot of the sink row.
    auto pop_round = [&]() {
        const unsigned avail = lds_u32(s_tails + lane * 4) - head;       // conflict-free 32-bit read
        const bool h1 = avail >= 1u;
        bool h2 = avail >= 2u;
        const unsigned e1 = s_q + lane * 8 + ((head & (VM_BQ_CAP - 1)) << 8), e2 = s_q + lane * 8 + (((head + 1u) & (VM_BQ_CAP - 1)) << 8);
        const unsigned long long word1 = lds_u64(e1), word2 = lds_u64(e2);     // (stale entries where h1 / h2 are false)
        double wq1 = w0, wq2 = w0;
        if (!UW) { wq1 = lds_f64(e1 + WOFF); wq2 = lds_f64(e2 + WOFF); }
        head += min(avail, 2u);
        double xi1, xi2;
        unsigned hi1, hi2;
        bq_unpack(word1, xi1, hi1);
        bq_unpack(word2, xi2, hi2);
        double val1[K], val2[K];
        bspline_uniform_w<K>(xi1, h1 ? wq1 : 0.0, val1);
        bspline_uniform_w<K>(xi2, h2 ? wq2 : 0.0, val2);
        if (FIXED) {                       // round each contribution once, then integer arithmetic only
#pragma unroll
            for (int j = 0; j < K; ++j) {
                val1[j] = __longlong_as_double(fix_of(val1[j], P.fixscale));
                val2[j] = __longlong_as_double(fix_of(val2[j], P.fixscale));
            }
        }
        if (h2 && hi1 == hi2) {            // same rows: one read-modify-write for both (queue order: entry 1 first)
#pragma unroll
            for (int j = 0; j < K; ++j) val1[j] = acc_add<FIXED>(val1[j], val2[j]);
#pragma unroll
            for (int j = 0; j < K; ++j) val2[j] = 0.0;
            h2 = false;
        }
        const unsigned a1 = h1 ? s_rows + (hi1 << 8) : s_sink, a2 = h2 ? s_rows + (hi2 << 8) : s_sink;
        // (h1 false implies h2 false: both then add zeros to the sink, twice the same address, zeros either way)
#pragma unroll
        for (int j = 0; j < K; ++j) {
            const double r1 = lds_f64(a1 + 8 * j), r2 = lds_f64(a2 + 8 * j);
            sts_f64(a1 + 8 * j, acc_add<FIXED>(r1, val1[j]));
            sts_f64(a2 + 8 * j, acc_add<FIXED>(r2, val2[j]));
            __syncwarp();                 // tap j of this lane and tap j + 1 of a neighbour may be the same row
        }
    };

    // queue one particle per lane (cell base b0, local coordinate xi, weight wp).  Normally one trip; a class
    // without room takes what fits and the rest retries after a pop round (all particles in one cell: 32 trips).
    auto push_batch = [&](int b0, double xi, double wp, bool active) {
        const int cls = b0 & 31;
        const unsigned long long word = bq_pack(xi, b0 >> 5);
        // lanes of this batch that queue into the same class
#ifdef VM_BQ_MATCH
        const unsigned same = __match_any_sync(VM_FULL_MASK, cls);
#else
        // one ballot per class bit.  Lane c (the owner of class c) intersects them into the member mask of ITS class --
        // its own bits are loop-invariant (lane_inv) -- and every lane then fetches the mask of its particle's class
        // from the owner.
        unsigned mine = VM_FULL_MASK;
#pragma unroll
        for (int bit = 0; bit < 5; ++bit) {
            const unsigned bal = __ballot_sync(VM_FULL_MASK, (cls & (1 << bit)) != 0);
            mine &= bal ^ lane_inv[bit];
        }
        const unsigned same = __shfl_sync(VM_FULL_MASK, mine, cls);
#endif
        bool pending = active;
        for (;;) {
            const unsigned peers = same & __ballot_sync(VM_FULL_MASK, pending);
            const unsigned t = lds_u32(s_tails + cls * 4);              // same address within a class: broadcast
            const unsigned h = __shfl_sync(VM_FULL_MASK, head, cls);    // the owner's head
            const int room = VM_BQ_CAP - (int)(t - h);
            const int rank = __popc(peers & lt_mask);
            __syncwarp();                                               // all lanes have read the tails
            if (pending && rank < room) {
                const unsigned slot = s_q + (((t + rank) & (VM_BQ_CAP - 1)) << 8) + cls * 8;
                sts_u64(slot, word);
                if (!UW) sts_f64(slot + WOFF, wp);
                if (rank == 0) sts_u32(s_tails + cls * 4, t + min(__popc(peers), room));
                pending = false;
            }
            __syncwarp();                  // entries and tails visible to the owning lanes
            if (!__any_sync(VM_FULL_MASK, pending)) break;
            pop_round();
        }
    };

    // gather, kick, drifts, new cell of one particle (registers and the read-only field table only)
    auto advance = [&](double& xp, double& vp, int& b0, double& xi) {
        if (MODE == MODE_PUSH_DEPOSIT) {
            cell_of<VM_BQ_CONV, POW2>(P.map, xp, b0, xi);
            const double dphi = gather_dphi_rt<K>(s_tab, b0, xi, gshift, stride_b);
            vp = __dadd_rn(vp, __dmul_rn(P.kick, dphi));
            if (SPLIT) vp = __dadd_rn(vp, __dmul_rn(P.kick2, dphi));
        }
        if (MODE != MODE_DEPOSIT) xp = __dadd_rn(xp, __dmul_rn(P.drift1, vp));
        if (MODE == MODE_PUSH_DEPOSIT) xp = __dadd_rn(xp, __dmul_rn(P.drift2, vp));
        cell_of<VM_BQ_CONV, POW2>(P.map, xp, b0, xi);
    };

    const unsigned npairs = (unsigned)(P.n >> 1);
    const unsigned stride = gridDim.x * blockDim.x;
    const unsigned gtid = blockIdx.x * blockDim.x + threadIdx.x;
    const unsigned iters = (npairs + stride - 1) / stride;     // uniform trip count: the queues are warp-collective

    struct Lines { double2 x, v, w; };
    auto load = [&](Lines& L, unsigned q) {
        L.w = make_double2(w0, w0);
        if (q < npairs) {
            L.x = ld_stream2(x + 2 * (size_t)q);
            if (MODE != MODE_DEPOSIT) L.v = ld_stream2(v + 2 * (size_t)q);
            if (!UW) L.w = ld_stream2(w + 2 * (size_t)q);
        }
    };
    auto work = [&](Lines& L, unsigned q) {
        const bool active = q < npairs;
        int b0a, b0b;
        double xia, xib;
        advance(L.x.x, L.v.x, b0a, xia);
        advance(L.x.y, L.v.y, b0b, xib);
        if (active && MODE != MODE_DEPOSIT) {
            st_stream2(x + 2 * (size_t)q, L.x);
            if (MODE == MODE_PUSH_DEPOSIT) st_stream2(v + 2 * (size_t)q, L.v);
        }
        push_batch(b0a, xia, L.w.x, active);
        push_batch(b0b, xib, L.w.y, active);
        pop_round();                                           // two particles queued per lane, up to two popped
    };
    Lines A, B;
    A.x = A.v = B.x = B.v = make_double2(0., 0.);
    unsigned q = gtid;
    load(A, q);
    for (unsigned it = 0; it < iters; it += 2, q += 2 * stride) {    // unrolled over two buffer sets: no register moves
        load(B, q + stride);
        work(A, q);
        load(A, q + 2 * stride);
        work(B, q + stride);                                   // all-inactive when iters is odd (one idle half-iteration)
    }
    if (MODE != MODE_DEPOSIT && VM_PASS_EARLY_LOAD) __threadfence();     // same protocol as k_vp_pass: stores visible before the trigger
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    if ((P.n & 1) && blockIdx.x == 0 && warp == 0) {           // odd particle count: last particle, lane 0 of one warp
        const bool active = (lane == 0);
        double xp = 0., vp = 0., wp = 0., xi = 0.;
        int b0 = 0;
        if (active) {
            xp = x[P.n - 1];
            if (MODE != MODE_DEPOSIT) vp = v[P.n - 1];
            wp = UW ? w0 : w[P.n - 1];
            advance(xp, vp, b0, xi);
            if (MODE != MODE_DEPOSIT) {
                x[P.n - 1] = xp;
                if (MODE == MODE_PUSH_DEPOSIT) v[P.n - 1] = vp;
            }
        }
        push_batch(b0, xi, wp, active);
    }
    while (__any_sync(VM_FULL_MASK, head != lds_u32(s_tails + lane * 4))) pop_round();     // drain
    flush_grid<VAR_MATCH, FIXED>(grid, scratch, out, n, GHOST, 0, nwarps, P.ncols);
    if (F.mode != FINISH_NONE) finish_grid(F, out, gridDim.x, n, grid, scratch);
}

// ================================================================ host ======
// Shared memory of the bank-sorted pass: gather table (fused mode) + per warp one replica grid and the queues.
struct BqPlan {
    int warps, gshift;
    size_t smem;
    BqParams Q;
};

inline bool plan_bq(vm_ctx* ctx, int n, int order, int pass_mode, bool uniform_w, BqPlan* out)
{
    const size_t budget = (ctx->smem_optin < 227 * 1024 - 1024) ? ctx->smem_optin : (size_t)(227 * 1024 - 1024);
    const int per_warp_q = VM_BQ_TAILWORDS + VM_BQ_SINKWORDS + VM_BQ_QWORDS * (uniform_w ? 1 : 2);
    const size_t per_warp = ((size_t)(n + order - 1) + per_warp_q) * sizeof(double);
    int gshift = 0;
    if (pass_mode == MODE_PUSH_DEPOSIT)                       // as many table copies (<= 16) as fit in ~33 KB
        while (gshift < 4 && ((size_t)(n + order) << (gshift + 1)) * sizeof(double) <= (size_t)VM_BQ_TABLE_BYTES(n)) ++gshift;
    for (;;) {
        const size_t table = pass_mode == MODE_PUSH_DEPOSIT ? ((size_t)(n + order) << gshift) * sizeof(double) : 0;
        int warps = VM_BQ_MAXW;
        while (warps >= 4 && table + warps * (per_warp + 32 * sizeof(double)) > budget) --warps;
        if (warps >= 8 || (warps >= 4 && gshift == 0)) {
            out->warps = warps; out->gshift = gshift;
            out->smem = table + warps * (per_warp + 32 * sizeof(double));
            out->Q.gshift = gshift; out->Q.per_warp = per_warp_q;
            return true;
        }
        if (gshift == 0) return false;
        --gshift;                                             // trade table copies for warps
    }
}

template <int K, int MODE, bool SPLIT, bool POW2, bool UW, bool FIXED, int MAXT>
void launch_bq_inst(vm_ctx* ctx, const BqPlan& bp, double* x, double* v, const double* w, const double* dcoef,
                    double* out, const PassParams& P, const FinishParams& F)
{
    static size_t configured[64] = {};
    size_t& conf = configured[ctx->device & 63];
    if (bp.smem > conf) {
        VM_CUDA(cudaFuncSetAttribute(k_vp_pass_bq<K, MODE, SPLIT, POW2, UW, FIXED, MAXT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bp.smem));
        conf = bp.smem;
    }
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(ctx->sm_count);
    cfg.blockDim = dim3(bp.warps * 32);
    cfg.dynamicSmemBytes = bp.smem;
    cfg.stream = ctx->stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = ctx->no_pdl ? 0 : 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    VM_CUDA(cudaLaunchKernelEx(&cfg, k_vp_pass_bq<K, MODE, SPLIT, POW2, UW, FIXED, MAXT>, x, v, w, dcoef, out, P, F, bp.Q));
    ++ctx->launches;
}

template <int K, int MODE, bool POW2, bool UW, bool FIXED>
void launch_bq_fx(vm_ctx* ctx, const BqPlan& bp, double* x, double* v, const double* w, const double* dcoef,
                  double* out, const PassParams& P, const FinishParams& F)
{
    const bool split = (MODE == MODE_PUSH_DEPOSIT) && P.kick2 != 0.0;
    if (bp.warps <= 16) {                      // few warps: 128 registers per thread
        if (split) launch_bq_inst<K, MODE, (MODE == MODE_PUSH_DEPOSIT), POW2, UW, FIXED, 512>(ctx, bp, x, v, w, dcoef, out, P, F);
        else launch_bq_inst<K, MODE, false, POW2, UW, FIXED, 512>(ctx, bp, x, v, w, dcoef, out, P, F);
    } else {
        if (split) launch_bq_inst<K, MODE, (MODE == MODE_PUSH_DEPOSIT), POW2, UW, FIXED, 1024>(ctx, bp, x, v, w, dcoef, out, P, F);
        else launch_bq_inst<K, MODE, false, POW2, UW, FIXED, 1024>(ctx, bp, x, v, w, dcoef, out, P, F);
    }
}

template <int K, int MODE, bool POW2, bool UW>
void launch_bq_uw(vm_ctx* ctx, const BqPlan& bp, double* x, double* v, const double* w, const double* dcoef,
                  double* out, const PassParams& P, const FinishParams& F)
{
    if (P.fixscale != 0.0) launch_bq_fx<K, MODE, POW2, UW, true>(ctx, bp, x, v, w, dcoef, out, P, F);
    else launch_bq_fx<K, MODE, POW2, UW, false>(ctx, bp, x, v, w, dcoef, out, P, F);
}

template <int K, int MODE>
void launch_bq(vm_ctx* ctx, const BqPlan& bp, double* x, double* v, const double* w, const double* dcoef,
               double* out, const PassParams& P, const FinishParams& F)
{
    // power-of-two meshes: the periodic wrap is a mask; uniform weights: no weight stream, no weight queue.
    // (the generic combination POW2 = false serves every mesh with per-particle weights: fewer instantiations)
    if (P.uw) {
        if (P.map.mask >= 0) launch_bq_uw<K, MODE, true, true>(ctx, bp, x, v, w, dcoef, out, P, F);
        else launch_bq_uw<K, MODE, false, true>(ctx, bp, x, v, w, dcoef, out, P, F);
    } else {
        PassParams G = P;
        G.map.mask = -1;
        launch_bq_uw<K, MODE, false, false>(ctx, bp, x, v, w, dcoef, out, G, F);
    }
}

// vm_deposit.cuh -- shared-memory replica-grid deposition machinery used by the x-space
// (vm_push.cu) and v-space (vm_vspline.cu) particle passes.
//
// Deposition never uses shared-memory atomics on the hot variants (fp64 shared atomics are CAS
// loops): every warp owns private replica grids in shared memory and resolves intra-warp
// collisions by grouping lanes by cell (__match_any_sync) and reducing each group in lane order,
// so the result is bit-reproducible for a fixed launch geometry.
#pragma once
#include "vm_internal.cuh"

// Row pitch (32-bit words) of one replica of the limb-atomic layout (VAR_AF): the smallest pitch >= rows that is congruent
// to 32/R modulo 32, so that replica r is shifted by r * 32/R banks against replica 0.
__host__ __device__ inline int vm_af_pitch(int rows, int rep_log2)
{
    const int want = (32 >> rep_log2) & 31;
    return rows + ((want - rows) & 31);
}

// most shared memory a limb-atomic plan takes while a smaller one exists: the rest of the 256 KB array is the L1 the
// streaming loads are staged in (vm_pass.cuh: plan_af)
#ifndef VM_AF_SMEM_CAP
#define VM_AF_SMEM_CAP (156 * 1024)
#endif

enum { VAR_PRIV = 0, VAR_MATCH = 1, VAR_ATOMIC = 2, VAR_XOR = 3, VAR_AF = 5 };   // (4 is the bank-sorted pass in vm_pass_plan.variant)

// ------------------------------------------------ order-independent sums -----
// VM_DEPOSIT_FIXED: every contribution w * B_j(xi) is rounded ONCE to a 64-bit fixed-point integer (scale 2^S, S
// from an exact, sharding-independent bound on sum |w|) and all further additions -- replica grids, CTA rows, ranks --
// are integer additions, which commute: the deposited vector has the same bits for every launch geometry, deposit
// layout and number of GPUs.  The accumulators live in the same 8-byte words as the fp64 ones (bit patterns).
#ifndef VM_AF_SKIP0
#define VM_AF_SKIP0 0     // 1: limb-atomic layout without the high-limb atomic in lanes whose high limb + carry is zero -- fewer bank-conflict
                          // replays, but ptxas turns the predicated red into a branch per tap: measured 5 % SLOWER (profiles/r02b_af_ab.txt)
#endif
#define VM_FIX_MAGIC 6755399441055744.0      // 1.5 * 2^52: fma(v, scale, MAGIC) holds rint(v * scale) in its low mantissa bits
__device__ __forceinline__ long long fix_of(double v, double scale)          // |v * scale| < 2^51
{
    return __double_as_longlong(fma(v, scale, VM_FIX_MAGIC)) - __double_as_longlong(VM_FIX_MAGIC);
}
template <bool FIXED>
__device__ __forceinline__ double acc_add(double a, double b)
{
    return FIXED ? __longlong_as_double(__double_as_longlong(a) + __double_as_longlong(b)) : a + b;
}
__device__ __forceinline__ double acc_add_rt(double a, double b, int fixed)
{
    return fixed ? __longlong_as_double(__double_as_longlong(a) + __double_as_longlong(b)) : a + b;
}

// ---------------------------------------------------------------- scatter ---
// Replica grids carry `ghost` extra rows after the n real ones, so a particle's K consecutive
// basis indices b0 .. b0+K-1 never wrap inside the hot loop (periodic grids: ghost = K-1, folded back
// onto rows 0..K-2 when the grid is flushed; clamped v-space grids: ghost = 0).
//
// Add val[j] to row b0 + j, j < K, in this warp's replica grid.  Inactive lanes must carry val == 0.
template <int K, int VAR, bool FIXED = false>
__device__ __forceinline__ void scatter(double* __restrict__ wg, int rep_log2, int rep, int lane,
                                        int b0, const double (&val)[K], bool active, double fixscale = 0.0)
{
    static_assert(!FIXED || VAR == VAR_PRIV || VAR == VAR_AF, "fixed-point accumulation exists in the lane-private, bank-sorted and limb-atomic layouts");
    static_assert(FIXED || VAR != VAR_AF, "the limb-atomic layout accumulates fixed-point integers");
    if (VAR == VAR_AF) {
        // Grids shared by ALL warps of the CTA, each row a 64-bit fixed-point integer stored as two 32-bit limbs in two
        // arrays (lo[] then hi[]: consecutive rows in consecutive banks).  64-bit and fp64 shared-memory atomics are CAS
        // loops on sm_100a (ATOMS.CAST.SPIN), 32-bit integer adds are native (ATOMS.ADD): add the low limb with a
        // returning atomic, derive the carry from the returned value, add high limb + carry with a second one.  Every
        // limb update is atomic and every wrap of a low limb is carried exactly once by the thread that caused it, so
        // after all adds (hi:lo) is the exact sum mod 2^64 in ANY order.
        // Bank steering: R = 2^rep_log2 replicas of the grid per CTA, replica r shifted by r * 32/R banks (row pitch
        // RP = 32/R mod 32).  Lane l deposits a particle with first row b0 into replica ((l - b0) mod 32) / (32/R): its
        // word then sits in bank l - delta, 0 <= delta < 32/R, and tap j in bank l - delta + j.  With R = 32 the 32 lanes
        // of every atomic hit 32 distinct banks -- ONE wavefront per instruction, no collision inside a warp ever -- the
        // lane-private layout, but shared by all warps because the adds are atomic: 256 B per mesh cell per CTA instead
        // of per warp.  (Measured with a single replica: 3.8 wavefronts per atomic on random rows, the pass sat at 95 %
        // of the LSU wavefront peak; tools/microbench/atoms.cu.)  wg = lo words, rep = row pitch RP in words.
        if (active) {
            const unsigned r = ((unsigned)(lane - b0) & 31u) >> (5 - rep_log2);
            const unsigned s_lo = (unsigned)__cvta_generic_to_shared(wg) + ((r * (unsigned)rep + (unsigned)b0) << 2);
            const unsigned s_hi = s_lo + (((unsigned)rep << rep_log2) << 2);
            unsigned xl[K], xh[K], old[K];
#pragma unroll
            for (int j = 0; j < K; ++j) {
                // fix_of by limbs: the low word of the magic constant is zero, so the low limb is the low word of the
                // fma itself and only the high word needs the subtraction
                const double fx = fma(val[j], fixscale, VM_FIX_MAGIC);
                xl[j] = (unsigned)__double2loint(fx);
                xh[j] = (unsigned)__double2hiint(fx) - 0x43380000u;
            }
#pragma unroll
            for (int j = 0; j < K; ++j)          // the K returning atomics in flight together
                asm volatile("atom.shared.add.u32 %0, [%1], %2;" : "=r"(old[j]) : "r"(s_lo + 4u * j), "r"(xl[j]) : "memory");
#pragma unroll
            for (int j = 0; j < K; ++j) {
#if VM_AF_SKIP0
                asm volatile("{\n\t.reg .pred p;\n\t.reg .u32 t, h;\n\t"
                             "add.cc.u32 t, %1, %2;\n\taddc.u32 h, %3, 0;\n\t"
                             "setp.ne.u32 p, h, 0;\n\t@p red.shared.add.u32 [%0], h;\n\t}"
                             :: "r"(s_hi + 4u * j), "r"(old[j]), "r"(xl[j]), "r"(xh[j]) : "memory");
#else
                asm volatile("{\n\t.reg .u32 t, h;\n\t"
                             "add.cc.u32 t, %1, %2;\n\taddc.u32 h, %3, 0;\n\tred.shared.add.u32 [%0], h;\n\t}"
                             :: "r"(s_hi + 4u * j), "r"(old[j]), "r"(xl[j]), "r"(xh[j]) : "memory");
#endif
            }
        }
        return;
    }
    if (VAR == VAR_PRIV) {
        // one private column per lane: no collisions, no branches, immediate-offset LDS/DADD/STS
        if (FIXED) {
            long long* a = (long long*)wg + (b0 << 5) + lane;
#pragma unroll
            for (int j = 0; j < K; ++j) a[j * 32] += fix_of(val[j], fixscale);
        } else {
            double* a = wg + (b0 << 5) + lane;
#pragma unroll
            for (int j = 0; j < K; ++j) a[j * 32] += val[j];
        }
        return;
    }
    if (VAR == VAR_XOR) {
        // R = 2^rep_log2 >= 4 replicas per warp: the 32/R lanes {l, l^R, l^2R, ...} share a replica column.
        // Collisions inside such a group are found with 32/R - 1 xor-shuffle rounds of the cell index
        // (cheaper than MATCH.ANY, whose cost grows with the number of distinct keys in the warp); the
        // lowest colliding lane adds its peers' values in a fixed round order and does the plain RMW.
        const int key = active ? b0 : ~lane;             // inactive lanes never match anybody
        const int R = 1 << rep_log2;
        double acc[K];
#pragma unroll
        for (int j = 0; j < K; ++j) acc[j] = val[j];
        bool dead = false;
        for (int d = R; d < 32; d += R) {                // warp-uniform trip count
            const int pk = __shfl_xor_sync(VM_FULL_MASK, key, d);
            const bool same = (pk == key);
            if (same && ((lane ^ d) < lane)) dead = true;
            if (__any_sync(VM_FULL_MASK, same)) {
#pragma unroll
                for (int j = 0; j < K; ++j) {
                    const double t = __shfl_xor_sync(VM_FULL_MASK, val[j], d);
                    if (same) acc[j] += t;
                }
            }
        }
        double* a = wg + (b0 << rep_log2) + rep;
#pragma unroll
        for (int j = 0; j < K; ++j) {
            if (!dead && active) a[j << rep_log2] += acc[j];
            __syncwarp();
        }
        return;
    }
    // group the lanes that target the same (cell, replica): in-warp sort-by-cell
    const unsigned key = active ? (((unsigned)b0 << 5) | (unsigned)rep) : (0x80000000u | (unsigned)lane);
    const unsigned peers = __match_any_sync(VM_FULL_MASK, key);
    const int leader = __ffs(peers) - 1;
    const bool is_leader = (lane == leader);
    double acc[K];
#pragma unroll
    for (int j = 0; j < K; ++j) acc[j] = val[j];
    // segmented reduce in increasing lane order (fixed order => bit-reproducible)
    unsigned todo = is_leader ? (peers & ~(1u << lane)) : 0u;
    while (__any_sync(VM_FULL_MASK, todo != 0u)) {
        const int src = todo ? (__ffs(todo) - 1) : lane;
#pragma unroll
        for (int j = 0; j < K; ++j) {
            const double t = __shfl_sync(VM_FULL_MASK, val[j], src);
            if (todo) acc[j] += t;
        }
        todo &= todo - 1u;
    }
    double* a = wg + (b0 << rep_log2) + rep;
    if (VAR == VAR_MATCH) {
#pragma unroll
        for (int j = 0; j < K; ++j) {
            if (is_leader && active) a[j << rep_log2] += acc[j];
            __syncwarp();
        }
    } else {  // VAR_ATOMIC: one warp-aggregated shared atomic per distinct cell
        if (is_leader && active) {
#pragma unroll
            for (int j = 0; j < K; ++j) atomicAdd(a + (j << rep_log2), acc[j]);
        }
    }
}

// Sum all replica copies of every row in a fixed order (ghost rows folded onto rows 0..ghost-1) and
// emit the CTA's partial row.  Bank-conflict-free in both phases (the first version summed with a
// 256-byte stride across lanes: 16-way conflicts, 10 us per launch):
//   phase 1  element-wise sum of the per-warp grids into warp 0's grid (consecutive threads, consecutive words)
//   phase 2  one warp per group of 32/R rows: lanes read 32 consecutive words and the R replicas of each row
//            are combined with a segmented xor butterfly
template <int VAR, bool FIXED = false>
__device__ __forceinline__ void flush_grid(double* __restrict__ grid, double* __restrict__ scratch,
                                           double* __restrict__ out, int n, int ghost, int rep_log2, int nwarps,
                                           int ncols)
{
    const int T = blockDim.x, t = threadIdx.x, lane = t & 31, warp = t >> 5;
    const int rows = n + ghost;
    const int gsz = rows << rep_log2;
    const int R = 1 << rep_log2;
    __syncthreads();
    if (VAR != VAR_ATOMIC) {                       // phase 1 (the atomic variant has a single per-CTA grid)
        for (int e = t; e < gsz; e += T) {
            double s = grid[e];
            for (int wq = 1; wq < nwarps; ++wq) s = acc_add<FIXED>(s, grid[wq * gsz + e]);
            grid[e] = s;
        }
        __syncthreads();
    }
    // phase 2: row totals into scratch[0 .. rows)   (rows <= blockDim.x is not required: scratch is reused in chunks)
    const int rows_per_warp = 32 >> rep_log2;       // rows covered by one 32-word read
    for (int base = 0; base < rows; base += T) {    // chunk of up to T rows (one trip unless rows > blockDim.x)
        const int chunk_rows = min(rows - base, T);
        for (int g = warp * rows_per_warp; g < chunk_rows; g += (T >> 5) * rows_per_warp) {
            const int row = base + g + (lane >> rep_log2);
            double s = (g + (lane >> rep_log2) < chunk_rows) ? grid[((base + g) << rep_log2) + lane] : 0.0;
            for (int o = R >> 1; o > 0; o >>= 1) s = acc_add<FIXED>(s, __shfl_xor_sync(VM_FULL_MASK, s, o));
            if ((lane & (R - 1)) == 0 && g + (lane >> rep_log2) < chunk_rows) scratch[row - base] = s;
        }
        __syncthreads();
        // fold the ghost rows and emit: real row i gets ghost row n + i (i < ghost); needs both in this chunk or
        // a second look-up, so ghost rows are read straight from the replica sums when they fall outside
        for (int i = base + t; i < min(base + chunk_rows, n); i += T) {
            double s = scratch[i - base];
            if (i < ghost) {
                const int gr = n + i;               // ghost row index
                double gsum;
                if (gr >= base && gr < base + chunk_rows) gsum = scratch[gr - base];
                else {
                    gsum = 0.0;        // (all-zero bits: also the fixed-point zero)
                    for (int r = 0; r < R; ++r) gsum = acc_add<FIXED>(gsum, grid[(gr << rep_log2) + r]);
                }
                s = acc_add<FIXED>(s, gsum);
            }
            if (VAR == VAR_ATOMIC) atomicAdd(out + i, s);
            else out[(size_t)blockIdx.x * ncols + i] = s;
        }
        __syncthreads();
    }
}

// Limb-atomic layout (VAR_AF): join the limbs of every row, sum the R replicas, fold the ghost rows and emit the CTA's
// partial row -- as the 64-bit integer itself when the fused finish continues in integers (as_bits), converted to fp64
// otherwise.  pitch = row pitch of a replica in words; the hi limbs follow the R * pitch lo limbs.
__device__ __forceinline__ void flush_limbs(const unsigned* __restrict__ lo, double* __restrict__ out, int n, int ghost,
                                            int rep_log2, int pitch, int ncols, bool as_bits, double inv_scale)
{
    const int R = 1 << rep_log2;
    const unsigned* hi = lo + (pitch << rep_log2);
    __syncthreads();
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
        long long s = 0;
        for (int r = 0; r < R; ++r) {
            const int e = r * pitch + i;
            s += (long long)(((unsigned long long)hi[e] << 32) | lo[e]);
            if (i < ghost) s += (long long)(((unsigned long long)hi[e + n] << 32) | lo[e + n]);
        }
        out[(size_t)blockIdx.x * ncols + i] = as_bits ? __longlong_as_double(s) : (double)s * inv_scale;
    }
    __syncthreads();
}

// ------------------------------------------------ fused cross-CTA finish ----
// The last CTA to finish (atomic ticket) sums all per-CTA rows in a fixed order, exchanges the result with the
// other ranks over NVLink peer memory and (small grids) solves the periodic Poisson system / the v-space mass
// system -- the reduce, all-reduce and solve launches of a step disappear.
//   n <= VM_FUSE_MAX_N   one level: the last CTA of the grid reduces all rows, exchanges and solves
//   n <= VM_X_MAX_N      two levels: the last CTA of every group of VM_GROUP_CTAS CTAs reduces its group's rows
//                        (groups finish in parallel), the last group finisher reduces the group rows and
//                        exchanges; the O(n^2) solve then runs as its own multi-CTA kernel (one SM would need
//                        >= 8.5 us of fp64 issue for n = 1024)
#define VM_FUSE_MAX_N 128
enum { FINISH_NONE = 0, FINISH_REDUCE = 1, FINISH_REDUCE_SOLVE = 2, FINISH_REDUCE_VSOLVE = 4 };

__device__ __forceinline__ void st_volatile_v2_u64(unsigned long long* p, unsigned long long a, unsigned long long b)
{
    asm volatile("st.volatile.global.v2.u64 [%0], {%1, %2};" ::"l"(p), "l"(a), "l"(b) : "memory");
}
__device__ __forceinline__ void ld_volatile_v2_u64(const unsigned long long* p, unsigned long long& a, unsigned long long& b)
{
    asm volatile("ld.volatile.global.v2.u64 {%0, %1}, [%2];" : "=l"(a), "=l"(b) : "l"(p) : "memory");
}

struct FinishParams {
    int mode;
    int two_level;          // rows are reduced per group of VM_GROUP_CTAS CTAs first (n > VM_FUSE_MAX_N)
    unsigned* ticket;       // device counters, zero on entry, reset by the last CTA: [0] grid, [1 + g] group g
    double* rhs;            // n: reduced deposit (sum over all ranks when xchg is set)
    double* grows;          // two_level: ceil(grid / VM_GROUP_CTAS) x n group rows
    const double* G;        // n: circulant pseudo-inverse kernel      (FINISH_REDUCE_SOLVE)
    double* phi;            // n
    double* dcoef;          // n
    double inv_h;
    // xchg: all-gather of the partial grids through NVLink peer memory ("LL" words: data + sequence number)
    int fixed;              // the rows hold 64-bit fixed-point integers (VM_DEPOSIT_FIXED): integer sums, converted at the end
    double inv_scale;       // 2^-S
    int xchg, nranks, rank;
    unsigned long long seq;             // exchange number (identical on all ranks); slot set = seq & 1
    unsigned long long* peer[VM_MAX_PEERS];   // every rank's inbox as mapped on this device (peer[rank] is this rank's own)
    unsigned* err;
    // FINISH_REDUCE_VSOLVE (v-space projection): coef = Minv * rhs, then one polynomial per cell
    const double* minv;                 // nv x nv
    const double* cellpoly;             // ncell x k x k
    double* coef;                       // npar (parent indexing)
    double* poly;                       // ncell x k
    int nv, off, ncell, k;
};

// Fused collective: every rank writes its n partial sums v[0..n) into slot [set][rank] of EVERY rank's inbox with
// plain 16-byte P2P stores over NVLink, each 8-byte word carrying 32 data bits and the 32-bit sequence number of
// this exchange -- no fence, no separate flag, one one-way NVLink latency.  Every rank then polls its own inbox
// until the words of all ranks carry the current sequence number and sums them in rank order: the same bits on
// every rank.  Two slot sets: a peer can be at most one exchange ahead (it needs this rank's words for the next).
// All threads of the CTA must call; v (shared memory) holds the global sums on return (after a __syncthreads).
__device__ __forceinline__ void exchange_ll(const FinishParams& F, double* __restrict__ v, int n, double* __restrict__ gout)
{
    const int T = blockDim.x, t = threadIdx.x;
    const int set = (int)(F.seq & 1ull);
    const unsigned long long tag = ((F.seq % 4294967295ull) + 1ull) << 32;     // never 0: the inbox starts zeroed
    __syncthreads();
    for (int idx = t; idx < n * F.nranks; idx += T) {
        const int r = idx / n, i = idx - r * n;
        const unsigned long long bits = (unsigned long long)__double_as_longlong(v[i]);
        st_volatile_v2_u64(F.peer[r] + (size_t)(set * VM_MAX_PEERS + F.rank) * VM_XSLOT_WORDS + 2 * i,
                           (bits & 0xffffffffull) | tag, (bits >> 32) | tag);
    }
    __syncthreads();                                   // v is overwritten below
    const unsigned long long* inbox = F.peer[F.rank] + (size_t)set * VM_MAX_PEERS * VM_XSLOT_WORDS;
    for (int i = t; i < n; i += T) {
        unsigned long long lo[VM_MAX_PEERS], hi[VM_MAX_PEERS];
        const long long t0 = clock64();
        bool ok;
        do {                                           // all ranks' words of this element in flight together
            ok = true;
#pragma unroll
            for (int r = 0; r < VM_MAX_PEERS; ++r)
                if (r < F.nranks) ld_volatile_v2_u64(inbox + (size_t)r * VM_XSLOT_WORDS + 2 * i, lo[r], hi[r]);
#pragma unroll
            for (int r = 0; r < VM_MAX_PEERS; ++r)
                if (r < F.nranks) ok = ok && ((lo[r] & 0xffffffff00000000ull) == tag) && ((hi[r] & 0xffffffff00000000ull) == tag);
            if (!ok && clock64() - t0 > (1ll << 35)) { atomicExch(F.err, 1u); break; }   // ~17 s: give up, host reports
        } while (!ok);
        double s = 0.0;
#pragma unroll
        for (int r = 0; r < VM_MAX_PEERS; ++r)
            if (r < F.nranks) s = acc_add_rt(s, __longlong_as_double((long long)((lo[r] & 0xffffffffull) | (hi[r] << 32))), F.fixed);
        if (F.fixed) s = (double)__double_as_longlong(s) * F.inv_scale;     // the one conversion back to fp64
        v[i] = s;
        gout[i] = s;
    }
    __syncthreads();
}

// Small-grid solves on the reduced (and exchanged) vector r_sh[0..n) held in shared memory.
__device__ __forceinline__ void finish_solve(const FinishParams& F, int n, double* __restrict__ sm_a)
{
    const int T = blockDim.x, t = threadIdx.x;
    double* r_sh = sm_a;            // rhs (then rhs - mean)
    double* phi_sh = sm_a + n;
    double* g_sh = sm_a + 2 * n + 1;   // pseudo-inverse kernel G staged in shared memory by the caller
    if (F.mode == FINISH_REDUCE_SOLVE) {
        __syncthreads();
        if (t < 32) {                // mean in a fixed order: strided lane sums, then the xor tree
            double s = 0.0;
            for (int j = t; j < n; j += 32) s += r_sh[j];
            s = warp_sum(s);
            if (t == 0) sm_a[2 * n] = s / (double)n;
        }
        __syncthreads();
        const double mean = sm_a[2 * n];
        __syncthreads();
        if (t < n) r_sh[t] -= mean;
        __syncthreads();
        if (t < n) {
            double a = 0.0;
            int idx = t;             // G[(t - j) mod n]
            for (int j = 0; j < n; ++j) {
                a = fma(g_sh[idx], r_sh[j], a);
                idx = (idx == 0) ? n - 1 : idx - 1;
            }
            phi_sh[t] = a;
            F.phi[t] = a;
        }
        __syncthreads();
        if (t < n) F.dcoef[t] = (phi_sh[t + 1 == n ? 0 : t + 1] - phi_sh[t]) * F.inv_h;
    }
    if (F.mode == FINISH_REDUCE_VSOLVE) {
        // same arithmetic (and bits) as k_v_solve + k_v_poly: one warp per matrix row, fixed xor tree
        __syncthreads();
        double* c_sh = sm_a + n;                     // parent-indexed coefficients
        const int lane = t & 31, wid = t >> 5, nw = T >> 5;
        if (t < n) c_sh[t] = 0.0;
        __syncthreads();
        for (int i = wid; i < F.nv; i += nw) {
            double s = 0.0;
            for (int j = lane; j < F.nv; j += 32) s = fma(__ldg(F.minv + (size_t)i * F.nv + j), r_sh[F.off + j], s);
            s = warp_sum(s);
            if (lane == 0) c_sh[F.off + i] = s;
        }
        __syncthreads();
        if (t < n) F.coef[t] = c_sh[t];
        for (int idx = t; idx < F.ncell * F.k; idx += T) {
            const int c = idx / F.k, mm = idx - c * F.k;
            double s = 0.0;
            for (int j = 0; j < F.k; ++j) s = fma(c_sh[c + j], __ldg(F.cellpoly + ((size_t)c * F.k + j) * F.k + mm), s);
            F.poly[idx] = s;
        }
    }
}

// One level (n <= VM_FUSE_MAX_N, blockDim.x >= n).
__device__ __forceinline__ void finish_last_cta(const FinishParams& F, const double* rows, int nrows, int n,
                                                double* __restrict__ sm_a /* >= 3n + 1 doubles, free */,
                                                double* __restrict__ scratch /* blockDim.x doubles */)
{
    __shared__ int s_last;
    __threadfence();                                   // publish this CTA's row
    __syncthreads();
    if (threadIdx.x == 0) s_last = (atomicAdd(F.ticket, 1u) == gridDim.x - 1u);
    __syncthreads();
    if (!s_last) return;
    __threadfence();
    const int T = blockDim.x, t = threadIdx.x;
    double* r_sh = sm_a;
    double* g_sh = sm_a + 2 * n + 1;   // the convolution must not chase n dependent global loads (measured: 5 us of a 17 us step floor)
    if (F.mode == FINISH_REDUCE_SOLVE && t < n) g_sh[t] = __ldg(F.G + t);
    int P = 1;
    while (2 * P * n <= T && 2 * P <= nrows) P *= 2;
    if (t < n * P) {
        const int i = t % n, part = t / n;
        double s = 0.0;
        for (int r = part; r < nrows; r += 8 * P) {      // 8 independent L2 loads in flight, summed in row order
            double vals[8];
#pragma unroll
            for (int u = 0; u < 8; ++u) {
                const int rr = r + u * P;
                vals[u] = (rr < nrows) ? __ldcg(rows + (size_t)rr * n + i) : 0.0;
            }
#pragma unroll
            for (int u = 0; u < 8; ++u) s = acc_add_rt(s, vals[u], F.fixed);
        }
        scratch[part * n + i] = s;
    }
    __syncthreads();
    if (t < n) {
        double s = 0.0;
        for (int part = 0; part < P; ++part) s = acc_add_rt(s, scratch[part * n + t], F.fixed);
        if (!F.xchg) {
            if (F.fixed) s = (double)__double_as_longlong(s) * F.inv_scale;
            F.rhs[t] = s;
        }
        r_sh[t] = s;
    }
    if (F.xchg) exchange_ll(F, r_sh, n, F.rhs);
    finish_solve(F, n, sm_a);
    if (t == 0) *F.ticket = 0u;      // ready for the next launch on this stream
}

// Two levels (VM_FUSE_MAX_N < n <= VM_X_MAX_N): reduce (+ exchange) only; any CTA shape.
__device__ __forceinline__ void finish_two_level(const FinishParams& F, const double* rows, int nrows, int n,
                                                 double* __restrict__ sm_a /* >= n doubles, free */)
{
    __shared__ int s_last2;
    const int T = blockDim.x, t = threadIdx.x;
    const int g = blockIdx.x / VM_GROUP_CTAS, ngroups = (nrows + VM_GROUP_CTAS - 1) / VM_GROUP_CTAS;
    const int r0 = g * VM_GROUP_CTAS, cnt = min(nrows - r0, VM_GROUP_CTAS);
    __threadfence();                                   // publish this CTA's row
    __syncthreads();
    if (t == 0) s_last2 = (atomicAdd(F.ticket + 1 + g, 1u) == (unsigned)cnt - 1u);
    __syncthreads();
    if (!s_last2) return;
    __threadfence();
    for (int i = t; i < n; i += T) {                   // level 1: this group's rows, all loads in flight, row order
        double vals[VM_GROUP_CTAS];
#pragma unroll
        for (int u = 0; u < VM_GROUP_CTAS; ++u) vals[u] = (u < cnt) ? __ldcg(rows + (size_t)(r0 + u) * n + i) : 0.0;
        double s = 0.0;
#pragma unroll
        for (int u = 0; u < VM_GROUP_CTAS; ++u) s = acc_add_rt(s, vals[u], F.fixed);
        F.grows[(size_t)g * n + i] = s;
    }
    __threadfence();
    __syncthreads();
    if (t == 0) {
        F.ticket[1 + g] = 0u;
        s_last2 = (atomicAdd(F.ticket, 1u) == (unsigned)ngroups - 1u);
    }
    __syncthreads();
    if (!s_last2) return;
    __threadfence();
    for (int i = t; i < n; i += T) {                   // level 2: the group rows in group order
        double s = 0.0;
        for (int q0 = 0; q0 < ngroups; q0 += 8) {
            double vals[8];
#pragma unroll
            for (int u = 0; u < 8; ++u) vals[u] = (q0 + u < ngroups) ? __ldcg(F.grows + (size_t)(q0 + u) * n + i) : 0.0;
#pragma unroll
            for (int u = 0; u < 8; ++u) s = acc_add_rt(s, vals[u], F.fixed);
        }
        if (!F.xchg) {
            if (F.fixed) s = (double)__double_as_longlong(s) * F.inv_scale;
            F.rhs[i] = s;
        }
        sm_a[i] = s;
    }
    if (F.xchg) exchange_ll(F, sm_a, n, F.rhs);
    if (t == 0) *F.ticket = 0u;
}

__device__ __forceinline__ void finish_grid(const FinishParams& F, const double* rows, int nrows, int n,
                                            double* __restrict__ sm_a, double* __restrict__ scratch)
{
    if (F.two_level) finish_two_level(F, rows, nrows, n, sm_a);
    else finish_last_cta(F, rows, nrows, n, sm_a, scratch);
}

// ================================================================ host ======
// fused peer exchange available on this context: fill the exchange fields of F and take the next sequence number
inline bool vm_xchg_setup(vm_ctx* ctx, FinishParams& F)
{
    if (ctx->nranks <= 1 || !ctx->peers_connected) return false;
    F.xchg = 1;
    F.nranks = ctx->nranks; F.rank = ctx->rank; F.seq = ++ctx->xseq;
    F.err = ctx->xerr;
    for (int r = 0; r < ctx->nranks; ++r) F.peer[r] = ctx->peer_inbox[r];
    return true;
}

struct DepositPlan {
    int var, rep_log2, grid, threads;
    size_t smem;
};

// Choose CTA shape + replica count so that the replica grids fit in shared memory.
// Measured on B200 (profiles/): the lane-private variant (32 replicas per warp, no collision handling)
// beats the match-based variant by ~2.7x even when it only leaves room for 6-12 warps per SM, so it is
// preferred whenever at least VM_PRIV_MIN_WARPS warps fit; otherwise the match variant runs with as many
// replicas per warp as fit next to 32 resident warps.
#define VM_PRIV_MIN_WARPS 6
#define VM_PRIV_MIN_WARPS_PUSH 6      // (12 before the deep-pipeline tiers: n_h = 128 now runs lane-private with 6 warps, 8 pairs each)

// Pairs of particles each thread keeps in flight in the lane-private passes, given the resident threads per
// SM the replica grids leave room for (measured, profiles/README.md): fewer warps -> deeper software pipeline.
inline int vm_auto_pairs(bool deposit_only, int threads_per_sm)
{
    if (deposit_only) return threads_per_sm <= 256 ? 8 : (threads_per_sm <= 512 ? 4 : 2);
    return threads_per_sm <= 192 ? 8 : (threads_per_sm <= 448 ? 4 : 1);
}
inline DepositPlan plan_deposit(vm_ctx* ctx, int n_real, int ghost, int extra_doubles, int mode,
                                int priv_min_warps = VM_PRIV_MIN_WARPS)
{
    const int sm = ctx->sm_count;
    const int n = n_real + ghost;   // rows per replica grid
    const size_t sm_total = 227 * 1024;   // usable shared memory per SM on sm_100
    auto budget_of = [&](int ctas) {
        size_t b = sm_total / ctas - 1024;                   // 1 KB per-CTA system reservation
        return b > ctx->smem_optin ? ctx->smem_optin : b;
    };
    auto fixed_of = [&](int threads) { return ((size_t)extra_doubles + (size_t)threads) * sizeof(double); };
    const bool tuned = ctx->ctas_per_sm > 0 || ctx->threads_per_cta > 0 || ctx->replicas > 0;

    if (mode != VM_DEPOSIT_ATOMIC && !tuned) {
        // lane-private first: 2 CTAs x 16 warps if that fits, else one CTA with as many warps as fit
        const size_t per_warp = (size_t)n * 32 * sizeof(double);
        if (2 * (fixed_of(512) + 16 * per_warp) <= 2 * budget_of(2)) {
            return DepositPlan{VAR_PRIV, 5, sm * 2, 512, fixed_of(512) + 16 * per_warp};
        }
        const size_t b1 = budget_of(1);
        int warps = 32;
        while (warps >= priv_min_warps && fixed_of(warps * 32) + warps * per_warp > b1) --warps;
        if (warps >= priv_min_warps)
            return DepositPlan{VAR_PRIV, 5, sm, warps * 32, fixed_of(warps * 32) + warps * per_warp};
    }

    if (mode != VM_DEPOSIT_ATOMIC && !tuned) {
        // next best (measured): 16 or 8 replicas per warp with xor-shuffle collision handling, provided
        // at least 24 warps stay resident; below that MATCH.ANY grouping with 32 warps wins again.  (Running the
        // xor variant with 6 warps and the deep software pipeline of the lane-private variant, or with 4 replicas,
        // measured 2-4x slower at n_h = 200 .. 1024: its shuffles and warp syncs serialise within a warp.)
        for (int rl = 4; rl >= 3; --rl) {
            const size_t per_warp = ((size_t)n << rl) * sizeof(double);
            const size_t b1 = budget_of(1);
            int warps = 32;
            while (warps >= 24 && fixed_of(warps * 32) + warps * per_warp > b1) --warps;
            if (warps >= 24) return DepositPlan{VAR_XOR, rl, sm, warps * 32, fixed_of(warps * 32) + warps * per_warp};
        }
    }

    struct Try { int ctas, threads; };
    std::vector<Try> tries;
    if (ctx->ctas_per_sm > 0 || ctx->threads_per_cta > 0) {
        tries.push_back({ctx->ctas_per_sm > 0 ? ctx->ctas_per_sm : 2, ctx->threads_per_cta > 0 ? ctx->threads_per_cta : 512});
    } else {
        tries = {{2, 512}, {1, 512}, {1, 256}, {1, 128}, {1, 64}};
    }
    for (const Try& t : tries) {
        const int nwarps = t.threads / 32;
        const size_t budget = budget_of(t.ctas);
        const size_t fixed = fixed_of(t.threads);
        if (budget <= fixed) continue;
        const size_t avail = (budget - fixed) / sizeof(double);
        DepositPlan pl{};
        pl.grid = sm * t.ctas;
        pl.threads = t.threads;
        if (mode == VM_DEPOSIT_ATOMIC) {
            int rl = 5;
            while (rl > 0 && ((size_t)n << rl) > avail) --rl;
            if (((size_t)n << rl) > avail) continue;
            pl.var = VAR_ATOMIC;
            pl.rep_log2 = rl;
            pl.smem = fixed + ((size_t)n << rl) * sizeof(double);
            return pl;
        }
        const size_t per_warp = avail / nwarps;
        int rl = 5;
        if (ctx->replicas > 0) { rl = 0; while ((1 << rl) < ctx->replicas) ++rl; }
        while (rl > 0 && ((size_t)n << rl) > per_warp && ctx->replicas == 0) --rl;
        if (((size_t)n << rl) > per_warp) continue;
        pl.var = (rl == 5) ? VAR_PRIV : ((rl >= 3 && !ctx->force_match) ? VAR_XOR : VAR_MATCH);
        pl.rep_log2 = rl;
        pl.smem = fixed + ((size_t)n << rl) * nwarps * sizeof(double);
        return pl;
    }
    throw vm_error(VM_ERR_UNSUPPORTED, "deposit: replica grids do not fit in shared memory for this n_basis/tuning");
}


// vm_fill.cu -- device-side synthetic particle loads reproducing the *distributions* of the
// reference's src/examples/*.jl samplers (not their RNG streams: the reference draws from Julia's
// unseeded global RNG).  Counter-based Philox4x32-10 keyed by (seed, global particle
// index), so a load does not depend on how particles are sharded over GPUs.
#include "vm_internal.cuh"

namespace {

__device__ __forceinline__ uint4 philox4x32_10(uint4 c, uint2 k)
{
    const unsigned M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        const unsigned hi0 = __umulhi(M0, c.x), lo0 = M0 * c.x;
        const unsigned hi1 = __umulhi(M1, c.z), lo1 = M1 * c.z;
        c = make_uint4(hi1 ^ c.y ^ k.x, lo1, hi0 ^ c.w ^ k.y, lo0);
        k.x += W0;
        k.y += W1;
    }
    return c;
}

// uniform double in (0,1) from 53 random bits
__device__ __forceinline__ double u01(unsigned hi, unsigned lo)
{
    const unsigned long long bits = ((unsigned long long)(hi >> 5) << 26) | (unsigned long long)(lo >> 6);
    return ((double)bits + 0.5) * (1.0 / 9007199254740992.0);
}

struct Draw { double u[4]; };

__device__ __forceinline__ Draw draw4(unsigned long long seed, unsigned long long g)
{
    const uint2 key = make_uint2((unsigned)seed, (unsigned)(seed >> 32));
    const uint4 a = philox4x32_10(make_uint4((unsigned)g, (unsigned)(g >> 32), 0u, 0u), key);
    const uint4 b = philox4x32_10(make_uint4((unsigned)g, (unsigned)(g >> 32), 1u, 0u), key);
    Draw d;
    d.u[0] = u01(a.x, a.y); d.u[1] = u01(a.z, a.w); d.u[2] = u01(b.x, b.y); d.u[3] = u01(b.z, b.w);
    return d;
}

struct FillParams {
    int kind;
    double p[8];
    unsigned long long seed;
    long first, total, n;
    double xmax;   // VM_FILL_NORMAL: ceil(max |z|) over the global population
};

// max |z_x| over ALL global particles (every rank computes the same number without communication)
__global__ void k_fill_absmax(FillParams P, unsigned long long* out)
{
    double m = 0.0;
    const long stride = (long)gridDim.x * blockDim.x;
    for (long g = (long)blockIdx.x * blockDim.x + threadIdx.x; g < P.total; g += stride) {
        const Draw d = draw4(P.seed, (unsigned long long)g);
        m = fmax(m, fabs(normcdfinv(d.u[0])));
    }
    for (int o = 16; o > 0; o >>= 1) m = fmax(m, __shfl_xor_sync(VM_FULL_MASK, m, o));
    if ((threadIdx.x & 31) == 0) atomicMax(out, (unsigned long long)__double_as_longlong(m));
}

// Solve x -/+ (eps/kappa) sin(kappa x) = u L for x in [0, L): inverse CDF of (1 -/+ eps cos(kappa x)) / L
__device__ __forceinline__ double inv_cdf_cos(double u, double L, double eps, double kappa, double sign)
{
    const double target = u * L;
    double x = target;
    for (int it = 0; it < 40; ++it) {
        const double F = x + sign * (eps / kappa) * sin(kappa * x) - target;
        const double dF = 1.0 + sign * eps * cos(kappa * x);
        const double dx = F / dF;
        x -= dx;
        if (fabs(dx) <= 1e-15 * L) break;
    }
    return fmin(fmax(x, 0.0), nextafter(L, 0.0));
}

__global__ void __launch_bounds__(256) k_fill(FillParams P, double* __restrict__ x, double* __restrict__ v,
                                              double* __restrict__ w)
{
    const long stride = (long)gridDim.x * blockDim.x;
    const double invN = 1.0 / (double)P.total;
    for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < P.n; i += stride) {
        const long g = P.first + i;
        const Draw d = draw4(P.seed, (unsigned long long)g);
        double xp, vp, wp = invN;
        switch (P.kind) {
            case VM_FILL_NORMAL: {              // normal.jl:15-33
                const double z = normcdfinv(d.u[0]);
                xp = (z + P.xmax) / (2.0 * P.xmax) * (P.p[1] - P.p[0]) + P.p[0];
                vp = normcdfinv(d.u[1]);
            } break;
            case VM_FILL_BUMP_ON_TAIL: {        // bumpontail.jl:43-75; params eps, kappa, alpha, sigma, v0
                const double L = 6.283185307179586476925286766559 / P.p[1];
                xp = inv_cdf_cos(d.u[0], L, P.p[0], P.p[1], -1.0);
                vp = normcdfinv(d.u[1]);
                if (d.u[2] > 1.0 - P.p[2]) vp = vp * P.p[3] + P.p[4];
                wp = L * invN;
            } break;
            case VM_FILL_LANDAU: {              // (1 + eps cos kx) Maxwellian; params eps, kappa
                const double L = 6.283185307179586476925286766559 / P.p[1];
                xp = inv_cdf_cos(d.u[0], L, P.p[0], P.p[1], +1.0);
                vp = normcdfinv(d.u[1]);
                wp = L * invN;
            } break;
            case VM_FILL_DOUBLE_MAXWELLIAN: {   // doublemaxwellian.jl:14-35; params xlo, xhi, shift
                xp = d.u[0] * (P.p[1] - P.p[0]) + P.p[0];
                vp = normcdfinv(d.u[1]) + ((g < P.total / 2) ? P.p[2] : -P.p[2]);
            } break;
            case VM_FILL_UNIFORM: {             // uniform.jl; params xlo, xhi, vlo, vhi
                xp = d.u[0] * (P.p[1] - P.p[0]) + P.p[0];
                vp = d.u[1] * (P.p[3] - P.p[2]) + P.p[2];
            } break;
            case VM_FILL_SHIFTED_NORMAL_V: {    // shiftednormalv.jl; params xlo, xhi, shift
                xp = d.u[0] * (P.p[1] - P.p[0]) + P.p[0];
                vp = normcdfinv(d.u[1]) + P.p[2];
            } break;
            default: {                          // VM_FILL_SHIFTED_UNIFORM; params xlo, xhi, vlo, vhi, shift
                xp = d.u[0] * (P.p[1] - P.p[0]) + P.p[0];
                vp = d.u[1] * (P.p[3] - P.p[2]) + P.p[2] + P.p[4];
            } break;
        }
        x[i] = xp;
        v[i] = vp;
        w[i] = wp;
    }
}

// ---------------------------------------------------------------- Sobol loads ----
// draw!(dist, f_x, ::BumpOnTail, ::AcceptRejectSampling / ::ImportanceSampling) (bumpontail.jl:43-75 / 90-121): proposals
// are the points of a 2-D Sobol sequence (Sobol.jl: Joe-Kuo direction numbers, Antonov-Saleev Gray-code order, the
// all-zero point left out), y1 -> x0 = y1 L, y2 -> v = sqrt(2) erfinv(2 y2 - 1); accept-reject keeps a proposal when
// rand <= f_x(x0) / (1 + eps), importance sampling keeps every proposal with the weight f_x(x0) L / N; a second rand
// moves the velocity to the tail population.  The reference's `rand` is Julia's unseeded global RNG: Philox keyed by
// (seed, proposal index) here.  Point k (1-based) = XOR over the set bits b of gray(k) = k ^ (k >> 1) of the direction
// numbers: dimension 1 V_b = 2^(31-b) (the bit reversal of gray(k)), dimension 2 m_0 = 1, m_b = 2 m_{b-1} ^ m_{b-1}
// (primitive polynomial x + 1), V_b = m_b 2^(31-b).  Sequential in the reference (particle n = n-th accepted proposal);
// here: count the accepted proposals per block of VM_SOBOL_BLOCK, scan, and compact in order -- every rank scans ALL
// proposals (no storage, ~1 ms per 1e8) and writes its own slice, so the load does not depend on the sharding.
#define VM_SOBOL_BLOCK 2048            // proposals per CTA trip: 256 threads x 8 consecutive proposals

__device__ __forceinline__ void sobol2(unsigned long long k, double& y1, double& y2)
{
    const unsigned g = (unsigned)(k ^ (k >> 1));
    unsigned x2 = 0u, m = 1u, bits = g;
    for (int b = 0; bits != 0u; ++b, bits >>= 1) {
        if (bits & 1u) x2 ^= m << (31 - b);
        m ^= m << 1;                                   // m_{b+1} = 2 m_b ^ m_b
    }
    y1 = (double)__brev(g) * (1.0 / 4294967296.0);
    y2 = (double)x2 * (1.0 / 4294967296.0);
}

struct SobolParams {
    double eps, kappa, alpha, sigma, v0, L;
    unsigned long long seed, skip;     // proposal j (0-based) is Sobol point skip + 1 + j
    long nprop;                        // proposals examined
    long first, n, total;              // this rank writes accepted ordinals [first, first + n) of `total`
    int importance;                    // 1: keep every proposal, weight f_x L / N
};

__device__ __forceinline__ bool sobol_proposal(const SobolParams& S, long j, double& xp, double& vp, double& wp)
{
    double y1, y2;
    sobol2(S.skip + 1ull + (unsigned long long)j, y1, y2);
    const Draw d = draw4(S.seed, (unsigned long long)j);
    xp = y1 * S.L;
    const double fx = 1.0 - S.eps * cos(S.kappa * xp);
    // sqrt(2) erfinv(2 y - 1) == normcdfinv(y); y2 = 0 cannot occur (it needs gray(k) = 0)
    vp = normcdfinv(y2);
    if (d.u[1] > 1.0 - S.alpha) vp = vp * S.sigma + S.v0;
    wp = (S.importance ? fx : 1.0) * S.L / (double)S.total;
    return S.importance || !(d.u[0] > fx / (1.0 + S.eps));
}

__global__ void __launch_bounds__(256) k_sobol_count(SobolParams S, unsigned* __restrict__ cnt, long nblocks)
{
    __shared__ unsigned wsum[8];
    for (long blk = blockIdx.x; blk < nblocks; blk += gridDim.x) {
        unsigned c = 0u;
        const long j0 = blk * VM_SOBOL_BLOCK + (long)threadIdx.x * 8;
        for (int u = 0; u < 8; ++u) {
            double xp, vp, wp;
            if (j0 + u < S.nprop && sobol_proposal(S, j0 + u, xp, vp, wp)) ++c;
        }
        for (int o = 16; o > 0; o >>= 1) c += __shfl_xor_sync(VM_FULL_MASK, c, o);
        if ((threadIdx.x & 31) == 0) wsum[threadIdx.x >> 5] = c;
        __syncthreads();
        if (threadIdx.x == 0) {
            unsigned t = 0u;
            for (int q = 0; q < 8; ++q) t += wsum[q];
            cnt[blk] = t;
        }
        __syncthreads();
    }
}

// exclusive scan of cnt[0..nblocks) into base[0..nblocks] (base[nblocks] = total), one CTA
__global__ void __launch_bounds__(1024) k_sobol_scan(const unsigned* __restrict__ cnt, unsigned long long* __restrict__ base, long nblocks)
{
    __shared__ unsigned long long part[1024];
    const long per = (nblocks + 1023) / 1024;
    const long b0 = (long)threadIdx.x * per, b1 = min(nblocks, b0 + per);
    unsigned long long s = 0ull;
    for (long b = b0; b < b1; ++b) s += cnt[b];
    part[threadIdx.x] = s;
    __syncthreads();
    if (threadIdx.x == 0) {
        unsigned long long run = 0ull;
        for (int t = 0; t < 1024; ++t) { const unsigned long long v = part[t]; part[t] = run; run += v; }
        base[nblocks] = run;
    }
    __syncthreads();
    unsigned long long run = part[threadIdx.x];
    for (long b = b0; b < b1; ++b) { base[b] = run; run += cnt[b]; }
}

__global__ void __launch_bounds__(256) k_sobol_emit(SobolParams S, const unsigned long long* __restrict__ base, long nblocks,
                                                    double* __restrict__ x, double* __restrict__ v, double* __restrict__ w)
{
    __shared__ unsigned wsum[8];
    const unsigned long long lo = (unsigned long long)S.first, hi = lo + (unsigned long long)S.n;
    for (long blk = blockIdx.x; blk < nblocks; blk += gridDim.x) {
        const unsigned long long bb = base[blk];
        if (bb >= hi || base[blk + 1] <= lo) continue;          // (uniform per CTA: no barrier is skipped by part of it)
        double xs[8], vs[8], ws[8];
        unsigned mask = 0u, c = 0u;
        const long j0 = blk * VM_SOBOL_BLOCK + (long)threadIdx.x * 8;
        for (int u = 0; u < 8; ++u)
            if (j0 + u < S.nprop && sobol_proposal(S, j0 + u, xs[u], vs[u], ws[u])) { mask |= 1u << u; ++c; }
        // ordinal of this thread's first accepted proposal inside the block: scan over the threads in index order
        unsigned incl = c;
        const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
        for (int o = 1; o < 32; o <<= 1) {
            const unsigned t = __shfl_up_sync(VM_FULL_MASK, incl, o);
            if (lane >= o) incl += t;
        }
        if (lane == 31) wsum[warp] = incl;
        __syncthreads();
        unsigned before = incl - c;
        for (int q = 0; q < warp; ++q) before += wsum[q];
        unsigned long long ord = bb + before;
        for (int u = 0; u < 8; ++u) {
            if (mask & (1u << u)) {
                if (ord >= lo && ord < hi) { x[ord - lo] = xs[u]; v[ord - lo] = vs[u]; w[ord - lo] = ws[u]; }
                ++ord;
            }
        }
        __syncthreads();
    }
}

// fills p with accepted ordinals [first, first + p->n) of total_n; throws if the proposals run out
void fill_sobol(vm_particles* p, const double* prm, unsigned long long seed, long first, long total_n, bool importance)
{
    vm_ctx* ctx = p->ctx;
    SobolParams S{};
    S.eps = prm[0]; S.kappa = prm[1]; S.alpha = prm[2]; S.sigma = prm[3]; S.v0 = prm[4];
    VM_REQUIRE(S.kappa > 0.0 && S.eps >= 0.0 && S.eps < 1.0, "vm_particles_fill: Sobol bump-on-tail load needs kappa > 0 and 0 <= eps < 1");
    S.L = 6.283185307179586476925286766559 / S.kappa;
    S.seed = seed; S.first = first; S.n = p->n; S.total = total_n; S.importance = importance ? 1 : 0;
    // skip(s, 2N) of Sobol.jl skips the largest power of two <= 2N + 1 points unless exact = true (its documented default);
    // prm[5] >= 0 overrides the count (a fixture of the reference can pin it)
    if (prm[5] >= 0.0) S.skip = (unsigned long long)prm[5];
    else { S.skip = 1ull; while (2ull * S.skip <= 2ull * (unsigned long long)total_n + 1ull) S.skip *= 2ull; }
    if (p->n == 0) return;
    for (double margin = 1.02;; margin *= 2.0) {
        S.nprop = importance ? total_n : (long)((double)total_n * (1.0 + S.eps) * margin) + 65536;
        VM_REQUIRE(S.skip + (unsigned long long)S.nprop < (1ull << 32), "vm_particles_fill: Sobol load beyond 2^32 sequence points");
        const long nblocks = (S.nprop + VM_SOBOL_BLOCK - 1) / VM_SOBOL_BLOCK;
        unsigned* cnt = nullptr;
        unsigned long long* base = nullptr;
        VM_CUDA(cudaMalloc(&cnt, (size_t)nblocks * sizeof(unsigned)));
        VM_CUDA(cudaMalloc(&base, (size_t)(nblocks + 1) * sizeof(unsigned long long)));
        const int grid = ctx->sm_count * 8;
        k_sobol_count<<<grid, 256, 0, ctx->stream>>>(S, cnt, nblocks);
        VM_LAUNCHED(ctx);
        k_sobol_scan<<<1, 1024, 0, ctx->stream>>>(cnt, base, nblocks);
        VM_LAUNCHED(ctx);
        unsigned long long accepted = 0ull;
        VM_CUDA(cudaMemcpyAsync(&accepted, base + nblocks, sizeof(accepted), cudaMemcpyDeviceToHost, ctx->stream));
        VM_CUDA(cudaStreamSynchronize(ctx->stream));
        const bool enough = accepted >= (unsigned long long)total_n;
        if (enough) {
            k_sobol_emit<<<grid, 256, 0, ctx->stream>>>(S, base, nblocks, p->x, p->v, p->w);
            VM_LAUNCHED(ctx);
            VM_CUDA(cudaStreamSynchronize(ctx->stream));
        }
        VM_CUDA(cudaFree(cnt));
        VM_CUDA(cudaFree(base));
        if (enough) break;
        VM_REQUIRE(margin < 16.0, "vm_particles_fill: Sobol accept-reject load did not accept enough proposals");
    }
    p->w_dirty = true;
    p->fix_dirty = true;
}

}  // namespace

extern "C" int vm_particles_fill(vm_particles* p, int kind, const double* params, int nparams,
                                 unsigned long long seed, long first_index, long total_n)
{
    VM_API_BEGIN(p ? p->ctx : nullptr)
    VM_REQUIRE(p != nullptr, "vm_particles_fill: NULL handle");
    static const int need[9] = {2, 5, 3, 4, 3, 5, 2, 6, 6};
    VM_REQUIRE(kind >= 0 && kind <= 8, "vm_particles_fill: unknown kind");
    VM_REQUIRE(nparams == need[kind] && params != nullptr, "vm_particles_fill: wrong number of parameters for this kind");
    VM_REQUIRE(first_index >= 0 && total_n >= first_index + p->n, "vm_particles_fill: shard exceeds the global population");
    vm_ctx* ctx = p->ctx;
    if (kind == VM_FILL_BUMP_ON_TAIL_SOBOL || kind == VM_FILL_BUMP_ON_TAIL_SOBOL_IS) {
        fill_sobol(p, params, seed, first_index, total_n, kind == VM_FILL_BUMP_ON_TAIL_SOBOL_IS);
        return VM_OK;
    }
    FillParams P{};
    P.kind = kind;
    for (int i = 0; i < nparams; ++i) P.p[i] = params[i];
    P.seed = seed; P.first = first_index; P.total = total_n; P.n = p->n;
    P.xmax = 1.0;
    const int grid = ctx->sm_count * 8;
    if (kind == VM_FILL_NORMAL && total_n > 0) {
        unsigned long long* dmax = nullptr;
        VM_CUDA(cudaMalloc(&dmax, sizeof(unsigned long long)));
        VM_CUDA(cudaMemsetAsync(dmax, 0, sizeof(unsigned long long), ctx->stream));
        k_fill_absmax<<<grid, 256, 0, ctx->stream>>>(P, dmax);
        VM_LAUNCHED(ctx);
        unsigned long long bits = 0;
        VM_CUDA(cudaMemcpyAsync(&bits, dmax, sizeof(bits), cudaMemcpyDeviceToHost, ctx->stream));
        VM_CUDA(cudaStreamSynchronize(ctx->stream));
        VM_CUDA(cudaFree(dmax));
        double m;
        memcpy(&m, &bits, sizeof(m));
        P.xmax = ceil(m);                      // normal.jl:19 xmax = ceil(maximum(abs.(x0)))
        if (!(P.xmax > 0.0)) P.xmax = 1.0;
    }
    if (p->n > 0) {
        k_fill<<<grid, 256, 0, ctx->stream>>>(P, p->x, p->v, p->w);
        VM_LAUNCHED(ctx);
        p->w_dirty = true;
        p->fix_dirty = true;
    }
    VM_API_END
}

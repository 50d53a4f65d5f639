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

}  // namespace

extern "C" int vm_particles_fill(vm_particles* p, int kind, const double* params, int nparams,
                                 unsigned long long seed, long first_index, long total_n)
{
    VM_API_BEGIN(p ? p->ctx : nullptr)
    VM_REQUIRE(p != nullptr, "vm_particles_fill: NULL handle");
    static const int need[7] = {2, 5, 3, 4, 3, 5, 2};
    VM_REQUIRE(kind >= 0 && kind <= 6, "vm_particles_fill: unknown kind");
    VM_REQUIRE(nparams == need[kind] && params != nullptr, "vm_particles_fill: wrong number of parameters for this kind");
    VM_REQUIRE(first_index >= 0 && total_n >= first_index + p->n, "vm_particles_fill: shard exceeds the global population");
    vm_ctx* ctx = p->ctx;
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

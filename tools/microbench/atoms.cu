// atoms.cu -- how fast is a 64-bit fixed-point deposit built from native 32-bit shared-memory atomics on sm_100a?
// (64-bit and fp64 shared atomicAdd compile to a CAS loop -- ATOMS.CAST.SPIN.64 -- and were 8x too slow in round 1.)
// One CTA-wide grid of n rows, a lo[] and a hi[] array of 32-bit words; a particle adds K = 4 contributions:
//   old = atomicAdd(&lo[r], xlo); carry = (old + xlo) < old; atomicAdd(&hi[r], xhi + carry)      (two native ATOMS)
// Rows are pseudo-random per lane (the large-mesh case: no locality), two particles per thread and trip.
// Variants: 0 = full (returning lo + carry + hi), 1 = 8 non-returning adds, 2 = lo only (4 adds),
//           3 = racy LDS/DADD/STS on one fp64 grid (timing floor of the un-handled single replica),
//           4 = full with lo/hi interleaved in one 64-bit word (16 banks per limb instead of 32)
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o atoms atoms.cu && ./atoms
#include <cuda_runtime.h>

#include <cstdio>
#include <cstdlib>

#define CK(x)                                                                                  \
    do {                                                                                       \
        cudaError_t e = (x);                                                                   \
        if (e != cudaSuccess) { fprintf(stderr, "%s: %s\n", #x, cudaGetErrorString(e)); exit(1); } \
    } while (0)

template <int VARIANT>
__global__ void __launch_bounds__(1024, 1) k_atoms(unsigned long long* out, int iters, int nmask)
{
    extern __shared__ unsigned smem[];
    const int n = nmask + 1, rows = n + 4;
    unsigned* lo = smem;
    unsigned* hi = smem + rows;
    double* g = (double*)smem;
    for (int i = threadIdx.x; i < 2 * rows; i += blockDim.x) smem[i] = 0u;
    __syncthreads();
    unsigned s = (threadIdx.x + blockIdx.x * blockDim.x) * 2654435761u + 12345u;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int p = 0; p < 2; ++p) {
            s = s * 1664525u + 1013904223u;
            const int b0 = (int)((s >> 9) & (unsigned)nmask);
            const unsigned xlo = s | 0x80000000u, xhi = (s >> 28);
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const int r = b0 + j;
                if (VARIANT == 0) {
                    const unsigned old = atomicAdd(lo + r, xlo + j);
                    const unsigned c = (old + (xlo + j)) < old ? 1u : 0u;
                    atomicAdd(hi + r, xhi + c);
                } else if (VARIANT == 1) {
                    atomicAdd(lo + r, xlo + j);
                    atomicAdd(hi + r, xhi);
                } else if (VARIANT == 2) {
                    atomicAdd(lo + r, xlo + j);
                } else if (VARIANT == 3) {
                    g[r] += (double)(int)xhi;
                } else {
                    const unsigned old = atomicAdd(smem + 2 * r, xlo + j);
                    const unsigned c = (old + (xlo + j)) < old ? 1u : 0u;
                    atomicAdd(smem + 2 * r + 1, xhi + c);
                }
            }
        }
    }
    __syncthreads();
    unsigned long long t = 0;
    for (int i = threadIdx.x; i < rows; i += blockDim.x) t += ((unsigned long long)hi[i] << 32) | lo[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = t;
}

template <int VARIANT>
static void run(const char* name, int sms, int clk_khz, int threads, int ctas_per_sm, int n, unsigned long long* out, bool last)
{
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    const int iters = 1 << 13;
    const size_t smem = (size_t)(n + 4) * 8 + 64;
    k_atoms<VARIANT><<<sms * ctas_per_sm, threads, smem>>>(out, 64, n - 1);
    CK(cudaEventRecord(e0));
    k_atoms<VARIANT><<<sms * ctas_per_sm, threads, smem>>>(out, iters, n - 1);
    CK(cudaEventRecord(e1));
    CK(cudaEventSynchronize(e1));
    float ms;
    CK(cudaEventElapsedTime(&ms, e0, e1));
    // warps of particles per SM = ctas_per_sm * (threads / 32) * iters * 2
    const double wp = (double)ctas_per_sm * (threads / 32) * iters * 2.0;
    const double clk = ms * 1e-3 * clk_khz * 1e3;
    printf("  {\"variant\": \"%s\", \"n\": %d, \"threads\": %d, \"ctas_per_sm\": %d, \"clk_per_warp_of_particles_per_sm\": %.2f}%s\n",
           name, n, threads, ctas_per_sm, clk / wp, last ? "" : ",");
}

int main()
{
    cudaDeviceProp prop;
    CK(cudaGetDeviceProperties(&prop, 0));
    const int sms = prop.multiProcessorCount;
    int clk_khz = 0;
    CK(cudaDeviceGetAttribute(&clk_khz, cudaDevAttrClockRate, 0));
    unsigned long long* out;
    CK(cudaMalloc(&out, (size_t)sms * 2 * 1024 * sizeof(unsigned long long)));
    printf("{\"device\": \"%s\", \"budget_clk_at_hbm_roofline_32B\": %.1f, \"rows\": [\n", prop.name,
           32.0 * 32.0 / (6547.2e9 / sms / (clk_khz * 1e3)));
    const int ns[4] = {128, 256, 1024, 4096};
    for (int i = 0; i < 4; ++i) {
        const int n = ns[i];
        run<0>("full_lo_hi_carry", sms, clk_khz, 1024, 1, n, out, false);
        run<0>("full_lo_hi_carry", sms, clk_khz, 1024, 2, n, out, false);
        run<0>("full_lo_hi_carry", sms, clk_khz, 512, 1, n, out, false);
        run<1>("eight_adds_no_return", sms, clk_khz, 1024, 1, n, out, false);
        run<1>("eight_adds_no_return", sms, clk_khz, 1024, 2, n, out, false);
        run<2>("four_adds_lo_only", sms, clk_khz, 1024, 2, n, out, false);
        run<3>("racy_lds_dadd_sts", sms, clk_khz, 1024, 2, n, out, false);
        run<4>("full_interleaved_words", sms, clk_khz, 1024, 2, n, out, i == 3);
    }
    printf("]}\n");
    return 0;
}

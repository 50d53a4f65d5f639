// peaks.cu -- device ceilings the particle passes are measured against besides HBM (SURVEY 8d asks for the fp64
// number; profiles/README.md section 7 explains why the shared-memory data pipe matters).  Stand-alone program:
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o peaks peaks.cu && ./peaks
// Prints one JSON object: fp64 FMA rate, shared-memory wavefront rate (64-bit conflict-free read-modify-write, the
// access pattern of the lane-private deposit), HBM copy rate.  NOT RUN in round 1 (written after the GPU budget
// of the round was spent); run it first thing in round 2 and put the numbers into profiles/.
#include <cuda_runtime.h>

#include <cstdio>
#include <cstdlib>

#define CK(x)                                                                                  \
    do {                                                                                       \
        cudaError_t e = (x);                                                                   \
        if (e != cudaSuccess) { fprintf(stderr, "%s: %s\n", #x, cudaGetErrorString(e)); exit(1); } \
    } while (0)

// 8 independent DFMA chains per thread
__global__ void __launch_bounds__(1024, 1) k_dfma(double* out, int iters, double a, double b)
{
    double c0 = threadIdx.x, c1 = c0 + 1, c2 = c0 + 2, c3 = c0 + 3, c4 = c0 + 4, c5 = c0 + 5, c6 = c0 + 6, c7 = c0 + 7;
    for (int i = 0; i < iters; ++i) {
        c0 = fma(c0, a, b); c1 = fma(c1, a, b); c2 = fma(c2, a, b); c3 = fma(c3, a, b);
        c4 = fma(c4, a, b); c5 = fma(c5, a, b); c6 = fma(c6, a, b); c7 = fma(c7, a, b);
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = ((c0 + c1) + (c2 + c3)) + ((c4 + c5) + (c6 + c7));
}

// lane-private 64-bit read-modify-write on K = 4 rows of a replica grid: per warp and trip 4 LDS.64 + 4 STS.64
// = 16 wavefronts (2 per instruction), bank-conflict-free
__global__ void __launch_bounds__(1024, 1) k_lsu(double* out, int iters, int rows)
{
    extern __shared__ double grid[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    double* wg = grid + (size_t)warp * rows * 32;
    for (int i = lane; i < rows * 32; i += 32) wg[i] = 0.0;
    __syncwarp();
    unsigned s = threadIdx.x * 2654435761u + blockIdx.x;
    for (int i = 0; i < iters; ++i) {
        s = s * 1664525u + 1013904223u;
        const int b0 = (int)((s >> 8) % (unsigned)(rows - 3));
        double* a = wg + b0 * 32 + lane;
        a[0] += 1.0; a[32] += 2.0; a[64] += 3.0; a[96] += 4.0;
    }
    __syncwarp();
    double t = 0.0;
    for (int i = lane; i < rows * 32; i += 32) t += wg[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = t;
}

__global__ void __launch_bounds__(1024, 1) k_copy(const double2* __restrict__ in, double2* __restrict__ out, size_t n)
{
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) out[i] = in[i];
}

static float timed(cudaEvent_t e0, cudaEvent_t e1)
{
    float ms;
    CK(cudaEventSynchronize(e1));
    CK(cudaEventElapsedTime(&ms, e0, e1));
    return ms;
}

int main()
{
    cudaDeviceProp prop;
    CK(cudaGetDeviceProperties(&prop, 0));
    const int sms = prop.multiProcessorCount;
    int clk_khz = 0;
    CK(cudaDeviceGetAttribute(&clk_khz, cudaDevAttrClockRate, 0));
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    double* out;
    CK(cudaMalloc(&out, (size_t)sms * 1024 * sizeof(double)));

    // fp64
    const int it_f = 1 << 16;
    k_dfma<<<sms, 1024>>>(out, 1024, 1.0000001, 1e-9);
    CK(cudaEventRecord(e0));
    k_dfma<<<sms, 1024>>>(out, it_f, 1.0000001, 1e-9);
    CK(cudaEventRecord(e1));
    const float ms_f = timed(e0, e1);
    const double dfma = (double)sms * 1024 * 8.0 * it_f;
    const double tflops = 2.0 * dfma / (ms_f * 1e-3) / 1e12;

    // shared-memory wavefronts: 16 warps x 19 rows (the n_h = 16 cubic replica grid) per CTA, 2 CTAs per SM
    const int rows = 19, warps = 16, it_l = 1 << 16;
    const size_t smem = (size_t)warps * rows * 32 * sizeof(double);
    CK(cudaFuncSetAttribute(k_lsu, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    k_lsu<<<2 * sms, warps * 32, smem>>>(out, 256, rows);
    CK(cudaEventRecord(e0));
    k_lsu<<<2 * sms, warps * 32, smem>>>(out, it_l, rows);
    CK(cudaEventRecord(e1));
    const float ms_l = timed(e0, e1);
    const double wavefronts = 2.0 * sms * warps * (double)it_l * 16.0;
    const double wf_per_clk_sm = wavefronts / sms / (ms_l * 1e-3 * clk_khz * 1e3);

    // HBM copy, 2 GiB each way
    const size_t n = (size_t)1 << 27;   // double2 elements = 2 GiB
    double2 *a, *b;
    CK(cudaMalloc(&a, n * sizeof(double2))); CK(cudaMalloc(&b, n * sizeof(double2)));
    CK(cudaMemset(a, 0, n * sizeof(double2)));
    k_copy<<<sms * 2, 1024>>>(a, b, n);
    CK(cudaEventRecord(e0));
    for (int r = 0; r < 5; ++r) k_copy<<<sms * 2, 1024>>>(a, b, n);
    CK(cudaEventRecord(e1));
    const float ms_c = timed(e0, e1) / 5;
    const double gbs = 2.0 * n * sizeof(double2) / (ms_c * 1e-3) / 1e9;

    printf("{\"device\": \"%s\", \"sms\": %d, \"clock_khz_nominal\": %d, \"fp64_fma_tflops\": %.2f, "
           "\"dfma_per_clk_per_sm\": %.1f, \"smem_rmw_wavefronts_per_clk_per_sm\": %.3f, \"hbm_copy_gbs\": %.1f}\n",
           prop.name, sms, clk_khz, tflops, dfma / sms / (ms_f * 1e-3 * clk_khz * 1e3), wf_per_clk_sm, gbs);
    return 0;
}

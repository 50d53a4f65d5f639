// vm_field.cu -- periodic spline field: setup (stencils, circulant pseudo-inverse), fixed-order
// reduction of per-CTA partial rows, replicated Poisson solve, field energy, accessors.
//
// Replaces PoissonSolvers.jl's Potential / PoissonSolverPBSplines at the reference's call sites
// (src/models/vlasov_poisson.jl:12-15, src/electric_field.jl:39-51).
#include <cstring>

#include "vm_internal.cuh"
#include "vm_spline_host.hpp"

// out[c] = sum_r rows[r][c] in a fixed order (chunks of rows, then chunks in order):
// identical bits for identical inputs, independent of scheduling.
__global__ void __launch_bounds__(256) k_reduce_rows(const double* __restrict__ rows, int nrows, int ncols,
                                                     double* __restrict__ out)
{
    __shared__ double red[8][33];
    const int o = threadIdx.x & 31, ch = threadIdx.x >> 5;
    const int col = blockIdx.x * 32 + o;
    const int len = (nrows + 7) / 8;
    double s = 0.0;
    if (col < ncols) {
        const int r0 = ch * len, r1 = min(nrows, r0 + len);
        for (int r = r0; r < r1; ++r) s += rows[(size_t)r * ncols + col];
    }
    red[ch][o] = s;
    __syncthreads();
    if (ch == 0 && col < ncols) {
        double t = 0.0;
#pragma unroll
        for (int q = 0; q < 8; ++q) t += red[q][o];
        out[col] = t;
    }
}

// The Poisson solve as a kernel of its own: one tile of 32 outputs per CTA (poisson_solve_tile, vm_internal.cuh).
__global__ void __launch_bounds__(256) k_poisson_solve(const double* __restrict__ rhs, const double* __restrict__ G,
                                                       int n, double inv_h,
                                                       double* __restrict__ phi, double* __restrict__ dcoef)
{
    extern __shared__ double r[];          // n
    __shared__ double red[2 * 8 * 33];
    __shared__ double wsum[8];
    poisson_solve_tile(blockIdx.x, rhs, G, n, inv_h, phi, dcoef, r, red, wsum);
}

// dcoef_m = (phi_{m+1} - phi_m) / h for externally prescribed coefficients (ExternalField)
__global__ void k_dcoef_from_phi(const double* __restrict__ phi, int n, double inv_h, double* __restrict__ dcoef)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) dcoef[i] = (phi[i + 1 == n ? 0 : i + 1] - phi[i]) * inv_h;
}

// the same for every column of an n x ncols coefficient history (blockIdx.y = column)
__global__ void k_dcoef_from_phi_cols(const double* __restrict__ phi, int n, double inv_h, double* __restrict__ dcoef)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const size_t o = (size_t)blockIdx.y * n;
    if (i < n) dcoef[o + i] = (phi[o + (i + 1 == n ? 0 : i + 1)] - phi[o + i]) * inv_h;
}

// ---- double-double helpers for the energy quadratic form --------------------------------------
__device__ __forceinline__ void two_sum(double a, double b, double& s, double& e)
{
    s = a + b;
    const double bb = s - a;
    e = (a - (s - bb)) + (b - bb);
}
__device__ __forceinline__ void dd_add(double& hi, double& lo, double xh, double xl)
{
    double s, e;
    two_sum(hi, xh, s, e);
    e += lo + xl;
    hi = s + e;
    lo = e - (hi - s);
}
__device__ __forceinline__ void dd_add_prod(double& hi, double& lo, double a, double b)
{
    const double p = a * b;
    const double pe = fma(a, b, -p);
    dd_add(hi, lo, p, pe);
}

// W = 1/2 phi^T S phi with S the circulant stiffness matrix (stencil st[0..k-1]); accumulated in
// double-double so the quadratic form does not lose digits to cancellation.
__global__ void __launch_bounds__(256) k_field_energy(const double* __restrict__ phi, const double* __restrict__ st,
                                                      int k, int n, double* __restrict__ out)
{
    __shared__ double rh[256], rl[256];
    double hi = 0.0, lo = 0.0;
    for (int i = threadIdx.x; i < n; i += 256) {
        double th = 0.0, tl = 0.0;
        dd_add_prod(th, tl, st[0], phi[i]);
        for (int d = 1; d < k; ++d) {
            int ip = (i + d) % n, im = ((i - d) % n + n) % n;
            dd_add_prod(th, tl, st[d], phi[ip]);
            dd_add_prod(th, tl, st[d], phi[im]);
        }
        // (th + tl) * phi_i
        const double p = th * phi[i];
        const double pe = fma(th, phi[i], -p) + tl * phi[i];
        dd_add(hi, lo, p, pe);
    }
    rh[threadIdx.x] = hi;
    rl[threadIdx.x] = lo;
    __syncthreads();
    for (int sft = 128; sft > 0; sft >>= 1) {
        if (threadIdx.x < sft) {
            double h = rh[threadIdx.x], l = rl[threadIdx.x];
            dd_add(h, l, rh[threadIdx.x + sft], rl[threadIdx.x + sft]);
            rh[threadIdx.x] = h;
            rl[threadIdx.x] = l;
        }
        __syncthreads();
    }
    if (threadIdx.x == 0) out[0] = 0.5 * (rh[0] + rl[0]);
}

// row = [W / chi^2, 1/2 sum w v^2, sum w v, sum w]
__global__ void k_store_diag(double* __restrict__ row, const double* __restrict__ energy, const double* __restrict__ wv,
                             double inv_chi2)
{
    if (threadIdx.x == 0) {
        row[0] = energy[0] * inv_chi2;
        row[1] = 0.5 * wv[0];
        row[2] = wv[1];
        row[3] = wv[2];
    }
}

// --------------------------------------------------------------- internals --
void vm_reduce_rows(vm_ctx* ctx, const double* rows, int nrows, int ncols, double* out)
{
    k_reduce_rows<<<(ncols + 31) / 32, 256, 0, ctx->stream>>>(rows, nrows, ncols, out);
    VM_LAUNCHED(ctx);
}

void vm_field_reduce_rows(vm_field* f, const double* rows, int nrows, int ncols, double* out)
{
    vm_reduce_rows(f->ctx, rows, nrows, ncols, out);
}

void vm_field_solve_local(vm_field* f, bool allreduce)
{
    vm_ctx* ctx = f->ctx;
    f->solve_pending = false;
    // rhs may already hold the sum over the ranks (fused peer exchange, an earlier solve, vm_vp_run): reducing it
    // again would scale phi by nranks -- update!(potential) of the reference can be repeated safely, so can this
    if (allreduce && ctx->nranks > 1 && !f->rhs_global) {
        vm_allreduce_sum(ctx, f->rhs, (size_t)f->n);
        f->rhs_global = true;
    }
    k_poisson_solve<<<(f->n + 31) / 32, 256, (size_t)f->n * sizeof(double), ctx->stream>>>(f->rhs, f->G, f->n, f->map.inv_h,
                                                                                          f->phi, f->dcoef);
    VM_LAUNCHED(ctx);
}

// device scalars appended to the rhs allocation: [n .. n+3] = wv sums, [n+4] = energy
double* vm_field_wv(vm_field* f) { return f->rhs + f->n; }
static double* field_energy_ptr(vm_field* f) { return f->rhs + f->n + VM_DIAG_COLS; }

void vm_field_energy_dev(vm_field* f, const double* phi)
{
    vm_ctx* ctx = f->ctx;
    k_field_energy<<<1, 256, 0, ctx->stream>>>(phi ? phi : f->phi, f->stencil_s, f->order, f->n, field_energy_ptr(f));
    VM_LAUNCHED(ctx);
}

// ExternalField (src/electric_field.jl:55-77): coefficient history coeffs[:, ts], column-major n x ncols, kept on
// the device together with the derivative-spline coefficients of every column.
void vm_field_ext_upload(vm_field* f, const double* coeffs_host, int ncols)
{
    vm_ctx* ctx = f->ctx;
    const size_t elems = (size_t)f->n * ncols;
    if (ncols > f->ext_cols) {
        VM_CUDA(cudaStreamSynchronize(ctx->stream));
        cudaFree(f->ext_phi); cudaFree(f->ext_dcoef);
        f->ext_phi = f->ext_dcoef = nullptr;
        f->ext_cols = 0;
        VM_CUDA(cudaMalloc(&f->ext_phi, elems * sizeof(double)));
        VM_CUDA(cudaMalloc(&f->ext_dcoef, elems * sizeof(double)));
        f->ext_cols = ncols;
    }
    VM_CUDA(cudaMemcpyAsync(f->ext_phi, coeffs_host, elems * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
    k_dcoef_from_phi_cols<<<dim3((f->n + 255) / 256, ncols), 256, 0, ctx->stream>>>(f->ext_phi, f->n, f->map.inv_h, f->ext_dcoef);
    VM_LAUNCHED(ctx);
    VM_CUDA(cudaStreamSynchronize(ctx->stream));           // the host matrix is only borrowed for the call
}

void vm_field_ext_select(vm_field* f, int col)
{
    vm_ctx* ctx = f->ctx;
    const size_t nb = (size_t)f->n * sizeof(double);
    VM_CUDA(cudaMemcpyAsync(f->phi, f->ext_phi + (size_t)col * f->n, nb, cudaMemcpyDeviceToDevice, ctx->stream));
    VM_CUDA(cudaMemcpyAsync(f->dcoef, f->ext_dcoef + (size_t)col * f->n, nb, cudaMemcpyDeviceToDevice, ctx->stream));
}

double* vm_field_diag_rows(vm_field* f, int rows)
{
    if (rows > f->diag_rows) {
        VM_CUDA(cudaStreamSynchronize(f->ctx->stream));
        if (f->diag) VM_CUDA(cudaFree(f->diag));
        f->diag = nullptr;
        f->diag_rows = 0;
        VM_CUDA(cudaMalloc(&f->diag, (size_t)rows * 4 * sizeof(double)));
        f->diag_rows = rows;
    }
    return f->diag;
}

void vm_field_store_diag(vm_field* f, int row, double chi)
{
    vm_ctx* ctx = f->ctx;
    k_store_diag<<<1, 32, 0, ctx->stream>>>(f->diag + (size_t)row * 4, field_energy_ptr(f), vm_field_wv(f),
                                            1.0 / (chi * chi));
    VM_LAUNCHED(ctx);
}

static void upload(vm_ctx* ctx, double* dst, const std::vector<double>& src)
{
    VM_CUDA(cudaMemcpyAsync(dst, src.data(), src.size() * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
    VM_CUDA(cudaStreamSynchronize(ctx->stream));
}

extern "C" {

int vm_field_create(vm_ctx* ctx, double a, double b, int order, int n_basis, int index_shift, vm_field** out)
{
    VM_API_BEGIN(ctx)
    VM_REQUIRE(ctx != nullptr && out != nullptr, "vm_field_create: NULL argument");
    *out = nullptr;
    VM_REQUIRE(b > a, "vm_field_create: empty domain");
    if (order < VM_MIN_ORDER || order > VM_MAX_ORDER) throw vm_error(VM_ERR_UNSUPPORTED, "vm_field_create: order must be in 2..6");
    VM_REQUIRE(n_basis >= order && n_basis <= VM_MAX_NBASIS, "vm_field_create: n_basis must be in order..4096");
    using vmhost::ld;
    const int n = n_basis, k = order;
    vm_field* f = new vm_field();
    try {
        f->ctx = ctx;
        f->device = ctx->device;
        f->a = a; f->b = b; f->order = k; f->n = n; f->shift = index_shift;
        f->h = (b - a) / n;
        f->map.inv_h = (double)((ld)n / ((ld)b - (ld)a));
        f->map.off = -a * f->map.inv_h;
        f->map.n = n;
        int sh = ((index_shift % n) + n) % n;
        f->map.bias = n * ((1 << 30) / n) + sh;
        f->map.inv_n = (unsigned)(0x100000000ull / (unsigned long long)n);
        f->map.mask = ((n & (n - 1)) == 0) ? n - 1 : -1;
        f->map.shift = sh;

        ld mass[vmhost::MAXK], stiff[vmhost::MAXK];
        vmhost::uniform_stencils(k, mass, stiff);
        const ld h = ((ld)b - (ld)a) / (ld)n;
        std::vector<double> st(k);
        ld sts[vmhost::MAXK];
        for (int d = 0; d < k; ++d) {
            f->mass_st[d] = (double)(mass[d] * h);
            sts[d] = stiff[d] / h;
            f->stiff_st[d] = (double)sts[d];
            st[d] = f->stiff_st[d];
        }
        std::vector<ld> G;
        if (!vmhost::circulant_pinv(sts, k, n, G)) throw vm_error(VM_ERR_INVALID, "vm_field_create: stiffness matrix is not positive semi-definite");
        std::vector<double> Gd(n);
        for (int m = 0; m < n; ++m) Gd[m] = (double)G[m];
        const size_t nb = (size_t)n * sizeof(double);
        VM_CUDA(cudaMalloc(&f->rhs, nb + (VM_DIAG_COLS + 4) * sizeof(double)));
        VM_CUDA(cudaMalloc(&f->phi, nb));
        VM_CUDA(cudaMalloc(&f->dcoef, nb));
        VM_CUDA(cudaMalloc(&f->G, nb));
        VM_CUDA(cudaMalloc(&f->stencil_s, k * sizeof(double)));
        VM_CUDA(cudaMalloc(&f->solve_count, 2 * sizeof(unsigned)));          // [0] tiles solved, [1] wait timed out
        VM_CUDA(cudaMemsetAsync(f->solve_count, 0, 2 * sizeof(unsigned), ctx->stream));
        VM_CUDA(cudaMemsetAsync(f->rhs, 0, nb + (VM_DIAG_COLS + 4) * sizeof(double), ctx->stream));
        VM_CUDA(cudaMemsetAsync(f->phi, 0, nb, ctx->stream));
        VM_CUDA(cudaMemsetAsync(f->dcoef, 0, nb, ctx->stream));
        upload(ctx, f->G, Gd);
        upload(ctx, f->stencil_s, st);
    } catch (...) {
        vm_field_destroy(f);
        throw;
    }
    *out = f;
    VM_API_END
}

int vm_field_destroy(vm_field* f)
{
    if (!f) return VM_OK;
    vm_child_quiesce(f->ctx, f->device);
    cudaFree(f->rhs); cudaFree(f->phi); cudaFree(f->dcoef); cudaFree(f->G);
    cudaFree(f->stencil_s); cudaFree(f->diag); cudaFree(f->ext_phi); cudaFree(f->ext_dcoef); cudaFree(f->solve_count);
    delete f;
    return VM_OK;
}

// after a stream sync: a fused pass that gave up waiting for the in-pass solve (pass_presolve) left a flag
void vm_field_check_solve_error(vm_field* f)
{
    unsigned e = 0;
    VM_CUDA(cudaMemcpy(&e, f->solve_count + 1, sizeof(e), cudaMemcpyDeviceToHost));
    if (e) throw vm_error(VM_ERR_CUDA, "in-pass Poisson solve timed out (not all CTAs of the fused pass resident?); set tuning no_presolve = 1");
}

static void get_vec(vm_field* f, const double* dev, double* host)
{
    VM_CUDA(cudaMemcpyAsync(host, dev, (size_t)f->n * sizeof(double), cudaMemcpyDeviceToHost, f->ctx->stream));
    VM_CUDA(cudaStreamSynchronize(f->ctx->stream));
    vm_field_check_solve_error(f);
}

int vm_field_get_rhs(vm_field* f, double* host)
{
    VM_API_BEGIN(f ? f->ctx : nullptr)
    VM_REQUIRE(f != nullptr && host != nullptr, "vm_field_get_rhs: NULL argument");
    get_vec(f, f->rhs, host);
    VM_API_END
}

int vm_field_get_coefficients(vm_field* f, double* host)
{
    VM_API_BEGIN(f ? f->ctx : nullptr)
    VM_REQUIRE(f != nullptr && host != nullptr, "vm_field_get_coefficients: NULL argument");
    get_vec(f, f->phi, host);
    VM_API_END
}

int vm_field_set_coefficients(vm_field* f, const double* host)
{
    VM_API_BEGIN(f ? f->ctx : nullptr)
    VM_REQUIRE(f != nullptr && host != nullptr, "vm_field_set_coefficients: NULL argument");
    vm_ctx* ctx = f->ctx;
    VM_CUDA(cudaMemcpyAsync(f->phi, host, (size_t)f->n * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
    k_dcoef_from_phi<<<(f->n + 255) / 256, 256, 0, ctx->stream>>>(f->phi, f->n, f->map.inv_h, f->dcoef);
    VM_LAUNCHED(ctx);
    VM_CUDA(cudaStreamSynchronize(ctx->stream));
    VM_API_END
}

int vm_field_get_stencils(vm_field* f, double* mass_k, double* stiff_k)
{
    VM_API_BEGIN(f ? f->ctx : nullptr)
    VM_REQUIRE(f != nullptr, "vm_field_get_stencils: NULL handle");
    for (int d = 0; d < f->order; ++d) {
        if (mass_k) mass_k[d] = f->mass_st[d];
        if (stiff_k) stiff_k[d] = f->stiff_st[d];
    }
    VM_API_END
}

int vm_field_solve(vm_field* f)
{
    VM_API_BEGIN(f ? f->ctx : nullptr)
    VM_REQUIRE(f != nullptr, "vm_field_solve: NULL handle");
    vm_field_solve_local(f, true);
    VM_API_END
}

int vm_field_energy(vm_field* f, double* W)
{
    VM_API_BEGIN(f ? f->ctx : nullptr)
    VM_REQUIRE(f != nullptr && W != nullptr, "vm_field_energy: NULL argument");
    vm_ctx* ctx = f->ctx;
    vm_field_energy_dev(f, nullptr);
    double* host = vm_pinned(ctx, 1);
    VM_CUDA(cudaMemcpyAsync(host, field_energy_ptr(f), sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    VM_CUDA(cudaStreamSynchronize(ctx->stream));
    *W = host[0];
    VM_API_END
}

}  // extern "C"

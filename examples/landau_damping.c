/* landau_damping.c -- the C ABI used from plain C (no Python, no torch): linear Landau damping of a
 * (1 + eps cos(kappa x)) Maxwellian with cubic splines, the device-side load and the fused time loop.
 *
 *   gcc -std=c11 -Iinclude examples/landau_damping.c -o landau -Lvlasovmethods.jl_b200 -lvlasov_b200 -lm \
 *       -Wl,-rpath,$PWD/vlasovmethods.jl_b200
 *   ./landau [particles] [n_basis] [steps]
 *
 * Prints t, field energy W, kinetic energy K, momentum M every 10 steps (the quantities of save_timestep!,
 * src/vlasov_poisson.jl:58-67) and the fitted damping rate (theory for kappa = 0.5: gamma = -0.1533).
 * Exit status 3 when no CUDA device is present (the library has no CPU path). */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>

#include "vlasov_b200.h"

#define CHECK(call, ctx)                                                                   \
    do {                                                                                   \
        int rc_ = (call);                                                                  \
        if (rc_ != VM_OK) {                                                                \
            fprintf(stderr, "%s failed (%d): %s\n", #call, rc_, vm_last_error(ctx));        \
            return rc_ == VM_ERR_NO_DEVICE ? 3 : 1;                                        \
        }                                                                                  \
    } while (0)

int main(int argc, char** argv)
{
    const long n_part = argc > 1 ? atol(argv[1]) : 4000000L;
    const int n_basis = argc > 2 ? atoi(argv[2]) : 32;
    const int n_steps = argc > 3 ? atoi(argv[3]) : 150;
    const double kappa = 0.5, eps = 0.05, dt = 0.1;
    const double length = 2.0 * acos(-1.0) / kappa;
    const int diag_every = 1;      /* W oscillates with period pi/omega = 2.2: the maxima need every step */

    if (vm_abi_version() != VM_ABI_VERSION) {
        fprintf(stderr, "header/library ABI mismatch\n");
        return 1;
    }
    vm_ctx* ctx = NULL;
    CHECK(vm_ctx_create(0, &ctx), NULL);

    vm_particles* p = NULL;
    vm_field* f = NULL;
    CHECK(vm_particles_create(ctx, n_part, &p), ctx);
    CHECK(vm_field_create(ctx, 0.0, length, 4, n_basis, 0, &f), ctx);
    const double params[2] = {eps, kappa};
    CHECK(vm_particles_fill(p, VM_FILL_LANDAU, params, 2, 20240601ULL, 0, n_part), ctx);

    const int rows = n_steps / diag_every + 1;
    double* diag = (double*)calloc((size_t)rows * 4, sizeof(double));
    if (!diag) return 1;
    CHECK(vm_vp_run(f, p, dt, n_steps, diag_every, 0, 1.0, diag), ctx);

    /* W(t) ~ exp(2 gamma t) between its local maxima: least-squares slope of log W over the maxima */
    double st = 0, sl = 0, stt = 0, stl = 0;
    int m = 0;
    for (int i = 0; i < rows; ++i) {
        const double t = i * diag_every * dt, W = diag[4 * i];
        if (i % 10 == 0) printf("t = %6.2f   W = %.6e   K = %.9f   M = % .3e\n", t, W, diag[4 * i + 1], diag[4 * i + 2]);
        const int is_max = i > 0 && i + 1 < rows && W > diag[4 * (i - 1)] && W > diag[4 * (i + 1)];
        if (is_max && t < 14.0) {
            const double l = log(W);
            st += t; sl += l; stt += t * t; stl += t * l; ++m;
        }
    }
    if (m >= 2) printf("damping rate from %d maxima of W: gamma = %.4f (linear theory -0.1533)\n", m,
                       0.5 * (m * stl - st * sl) / (m * stt - st * st));
    free(diag);
    CHECK(vm_field_destroy(f), ctx);
    CHECK(vm_particles_destroy(p), ctx);
    CHECK(vm_ctx_destroy(ctx), NULL);
    return 0;
}

/*
 * vm_oracle.c -- CPU ORACLE (test infrastructure, NOT product code).
 * See vm_oracle.h for scope, parity status ("parity unpinned" for the
 * third-party arithmetic) and usage rules.
 *
 * Plain C restatement of the VlasovMethods.jl particle hot path; every public
 * function cites the reference file:line it follows (paths relative to
 * /root/reference).  Sums that the reference accumulates sequentially in
 * Float64 are accumulated here in long double (x87 80-bit) so the oracle is
 * at least as accurate as the Julia CPU path.  The vmo_baseline_* functions
 * at the bottom are the *timed* CPU baseline: same loop structure as the
 * reference, plain double, optional OpenMP.
 */
#include "vm_oracle.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

typedef long double ld;
#define VMO_MAXK 8

/* ====================================================================== */
/* B-spline basis (BSplineKit.jl `evaluate_all`, de Boor recursion)       */
/* ====================================================================== */

void vmo_eval_all(const double* t, int s, int k, double x, double* N)
{
    double left[VMO_MAXK + 1], right[VMO_MAXK + 1];
    N[0] = 1.0;
    for (int j = 1; j < k; ++j) {
        left[j] = x - t[s + 1 - j];
        right[j] = t[s + j] - x;
        double saved = 0.0;
        for (int r = 0; r < j; ++r) {
            double den = right[r + 1] + left[j - r];
            double temp = (den != 0.0) ? N[r] / den : 0.0;
            N[r] = saved + right[r + 1] * temp;
            saved = left[j - r] * temp;
        }
        N[j] = saved;
    }
}

void vmo_eval_all_deriv(const double* t, int s, int k, double x, double* dN)
{
    /* B'_{m,k} = (k-1) [ B_{m,k-1}/(t_{m+k-1}-t_m) - B_{m+1,k-1}/(t_{m+k}-t_{m+1}) ] */
    double Nl[VMO_MAXK];
    if (k < 2) { dN[0] = 0.0; return; }
    vmo_eval_all(t, s, k - 1, x, Nl); /* Nl[j'] = B_{s-k+2+j', k-1} */
    for (int j = 0; j < k; ++j) {
        int m = s - k + 1 + j;
        double acc = 0.0;
        if (j >= 1) {
            double den = t[m + k - 1] - t[m];
            if (den != 0.0) acc += Nl[j - 1] / den;
        }
        if (j <= k - 2) {
            double den = t[m + k] - t[m + 1];
            if (den != 0.0) acc -= Nl[j] / den;
        }
        dN[j] = (double)(k - 1) * acc;
    }
}

/* searchsortedlast on the uniform breakpoints a + c*h, c = 0..ncell */
static int find_cell(double a, double h, int ncell, double x)
{
    int lo = 0, hi = ncell;
    while (hi - lo > 1) {
        int mid = (lo + hi) >> 1;
        if (a + mid * h <= x) lo = mid; else hi = mid;
    }
    return lo;
}

static double reduce_periodic(double a, double b, double x)
{
    double L = b - a;
    double xr = x - L * floor((x - a) / L);
    if (xr < a) xr += L;
    if (xr >= b) xr -= L;
    if (xr < a) xr = a;
    return xr;
}

int vmo_periodic_eval(double a, double b, int n, int k, double x, double* N, double* dN)
{
    double h = (b - a) / n;
    double xr = reduce_periodic(a, b, x);
    int c = find_cell(a, h, n, xr);
    double tl[2 * VMO_MAXK];
    for (int i = 0; i < 2 * k; ++i) tl[i] = a + (double)(c - k + 1 + i) * h;
    if (N) vmo_eval_all(tl, k - 1, k, xr, N);
    if (dN) vmo_eval_all_deriv(tl, k - 1, k, xr, dN);
    return c;
}

static double breakpoint(double a, double b, int nknots, int i)
{
    if (i <= 0) return a;
    if (i >= nknots - 1) return b;
    return a + (double)i * ((b - a) / (double)(nknots - 1));
}

int vmo_clamped_eval(double a, double b, int nknots, int k, double x, double* N, double* dN)
{
    if (!(x >= a && x <= b)) return -1;
    int ncell = nknots - 1;
    double h = (b - a) / ncell;
    int c = (x >= b) ? ncell - 1 : find_cell(a, h, ncell, x);
    /* make sure breakpoint(c) <= x < breakpoint(c+1) under the exact knots */
    while (c > 0 && x < breakpoint(a, b, nknots, c)) --c;
    while (c < ncell - 1 && x >= breakpoint(a, b, nknots, c + 1)) ++c;
    double tl[2 * VMO_MAXK];
    for (int i = 0; i < 2 * k; ++i) tl[i] = breakpoint(a, b, nknots, c + i - k + 1);
    if (N) vmo_eval_all(tl, k - 1, k, x, N);
    if (dN) vmo_eval_all_deriv(tl, k - 1, k, x, dN);
    return c;
}

static inline int pmod(int i, int n) { int r = i % n; return r < 0 ? r + n : r; }

/* ====================================================================== */
/* Gauss-Legendre + Galerkin matrices (BSplineKit.galerkin_matrix)        */
/* ====================================================================== */

static void gauss_legendre(int m, ld* xs, ld* ws)
{
    const ld PI = 3.141592653589793238462643383279502884L;
    for (int i = 0; i < m; ++i) {
        ld z = cosl(PI * (i + 0.75L) / (m + 0.5L));
        ld pp = 1;
        for (int it = 0; it < 100; ++it) {
            ld p1 = 1, p2 = 0;
            for (int j = 0; j < m; ++j) {
                ld p3 = p2; p2 = p1;
                p1 = ((2 * j + 1) * z * p2 - j * p3) / (j + 1);
            }
            pp = m * (z * p1 - p2) / (z * z - 1);
            ld dz = p1 / pp;
            z -= dz;
            if (fabsl(dz) < 1e-19L) break;
        }
        xs[i] = z;
        ws[i] = 2 / ((1 - z * z) * pp * pp);
    }
}

static void periodic_galerkin(double a, double b, int n, int k, int shift, int deriv, double* out)
{
    ld* acc = (ld*)calloc((size_t)n * n, sizeof(ld));
    ld xs[VMO_MAXK], ws[VMO_MAXK];
    gauss_legendre(k, xs, ws);
    double h = (b - a) / n;
    for (int c = 0; c < n; ++c) {
        double tl[2 * VMO_MAXK];
        for (int i = 0; i < 2 * k; ++i) tl[i] = a + (double)(c - k + 1 + i) * h;
        for (int q = 0; q < k; ++q) {
            double xq = (double)(a + (c + 0.5L * (1 + xs[q])) * (ld)h);
            double N[VMO_MAXK];
            if (deriv) vmo_eval_all_deriv(tl, k - 1, k, xq, N); else vmo_eval_all(tl, k - 1, k, xq, N);
            ld wq = 0.5L * h * ws[q];
            for (int j1 = 0; j1 < k; ++j1)
                for (int j2 = 0; j2 < k; ++j2)
                    acc[(size_t)pmod(c + j1 + shift, n) * n + pmod(c + j2 + shift, n)] += wq * N[j1] * N[j2];
        }
    }
    for (size_t i = 0; i < (size_t)n * n; ++i) out[i] = (double)acc[i];
    free(acc);
}

void vmo_periodic_mass(double a, double b, int n, int k, int shift, double* M)
{ periodic_galerkin(a, b, n, k, shift, 0, M); }

void vmo_periodic_stiffness(double a, double b, int n, int k, int shift, double* S)
{ periodic_galerkin(a, b, n, k, shift, 1, S); }

void vmo_dirichlet_mass(double a, double b, int nknots, int k, double* M)
{
    int ncell = nknots - 1, npar = nknots + k - 2, nv = npar - 2;
    ld* acc = (ld*)calloc((size_t)nv * nv, sizeof(ld));
    ld xs[VMO_MAXK], ws[VMO_MAXK];
    gauss_legendre(k, xs, ws);
    for (int c = 0; c < ncell; ++c) {
        double x0 = breakpoint(a, b, nknots, c), x1 = breakpoint(a, b, nknots, c + 1);
        for (int q = 0; q < k; ++q) {
            double xq = (double)(x0 + 0.5L * (1 + xs[q]) * ((ld)x1 - x0));
            double N[VMO_MAXK];
            int cc = vmo_clamped_eval(a, b, nknots, k, xq, N, NULL);
            (void)cc;
            ld wq = 0.5L * ((ld)x1 - x0) * ws[q];
            for (int j1 = 0; j1 < k; ++j1) {
                int p1 = c + j1; if (p1 == 0 || p1 == npar - 1) continue;
                for (int j2 = 0; j2 < k; ++j2) {
                    int p2 = c + j2; if (p2 == 0 || p2 == npar - 1) continue;
                    acc[(size_t)(p1 - 1) * nv + (p2 - 1)] += wq * N[j1] * N[j2];
                }
            }
        }
    }
    for (size_t i = 0; i < (size_t)nv * nv; ++i) M[i] = (double)acc[i];
    free(acc);
}

/* ====================================================================== */
/* x-space: deposit, Poisson solve, gather, splitting flows               */
/* ====================================================================== */

void vmo_deposit_periodic(const double* x, const double* w, long np,
                          double a, double b, int n, int k, int shift, double* rhs)
{
    /* src/projections/potential.jl:4  b .= 0 */
    ld* acc = (ld*)calloc((size_t)n, sizeof(ld));
    for (long p = 0; p < np; ++p) {                 /* :10 for (x, w) in zip(points, weights) */
        double N[VMO_MAXK], bs[VMO_MAXK];
        int c = vmo_periodic_eval(a, b, n, k, x[p], N, NULL);   /* :11 ilast, bs = basis(x) */
        int ilast = c + k - 1 + shift;
        for (int d = 0; d < k; ++d) bs[d] = N[k - 1 - d];       /* bs[d] = b_{ilast-d} */
        for (int d = 0; d < k; ++d) {                           /* :15 for (di, bi) in pairs(bs) */
            int i = ilast - d;                                  /* :16 i = ilast + 1 - di */
            acc[pmod(i, n)] += (ld)w[p] * (ld)bs[d];            /* :17 b[i] += w * bi (PeriodicVector wrap) */
        }
    }
    for (int i = 0; i < n; ++i) rhs[i] = (double)acc[i];
    free(acc);
}

/* dense LU with partial pivoting, long double; solves A x = b in place (b -> x) */
static int lu_solve_ld(ld* A, int n, ld* bvec)
{
    int* piv = (int*)malloc(sizeof(int) * n);
    for (int c = 0; c < n; ++c) {
        int pr = c; ld best = fabsl(A[(size_t)c * n + c]);
        for (int r = c + 1; r < n; ++r) { ld v = fabsl(A[(size_t)r * n + c]); if (v > best) { best = v; pr = r; } }
        piv[c] = pr;
        if (best == 0) { free(piv); return -1; }
        if (pr != c) {
            for (int j = 0; j < n; ++j) { ld t = A[(size_t)c * n + j]; A[(size_t)c * n + j] = A[(size_t)pr * n + j]; A[(size_t)pr * n + j] = t; }
            ld t = bvec[c]; bvec[c] = bvec[pr]; bvec[pr] = t;
        }
        ld inv = 1 / A[(size_t)c * n + c];
        for (int r = c + 1; r < n; ++r) {
            ld f = A[(size_t)r * n + c] * inv;
            if (f == 0) continue;
            A[(size_t)r * n + c] = f;
            for (int j = c + 1; j < n; ++j) A[(size_t)r * n + j] -= f * A[(size_t)c * n + j];
            bvec[r] -= f * bvec[c];
        }
    }
    for (int r = n - 1; r >= 0; --r) {
        ld s = bvec[r];
        for (int j = r + 1; j < n; ++j) s -= A[(size_t)r * n + j] * bvec[j];
        bvec[r] = s / A[(size_t)r * n + r];
    }
    free(piv);
    return 0;
}

void vmo_poisson_solve(const double* S, int n, const double* rhs, double* phi)
{
    ld* A = (ld*)malloc(sizeof(ld) * (size_t)n * n);
    ld* bvec = (ld*)malloc(sizeof(ld) * n);
    ld mean = 0;
    for (int i = 0; i < n; ++i) mean += rhs[i];
    mean /= n;
    for (int i = 0; i < n; ++i) bvec[i] = (ld)rhs[i] - mean;          /* P*rhs, P = I - 11^T/n */
    for (size_t i = 0; i < (size_t)n * n; ++i) A[i] = (ld)S[i] + 1.0L / n; /* S + R regularisation */
    lu_solve_ld(A, n, bvec);
    /* enforce the gauge sum(phi)=0 exactly up to rounding */
    ld m2 = 0; for (int i = 0; i < n; ++i) m2 += bvec[i]; m2 /= n;
    for (int i = 0; i < n; ++i) phi[i] = (double)(bvec[i] - m2);
    free(A); free(bvec);
}

void vmo_eval_dphi(const double* x, long np, double a, double b, int n, int k, int shift,
                   const double* phi, double* dphi)
{
    for (long p = 0; p < np; ++p) {
        double dN[VMO_MAXK];
        int c = vmo_periodic_eval(a, b, n, k, x[p], NULL, dN);
        ld s = 0;
        for (int j = 0; j < k; ++j) s += (ld)phi[pmod(c + j + shift, n)] * (ld)dN[j];
        dphi[p] = (double)s;
    }
}

double vmo_field_energy(const double* S, int n, const double* phi)
{
    ld e = 0;
    for (int i = 0; i < n; ++i) {
        ld r = 0;
        for (int j = 0; j < n; ++j) r += (ld)S[(size_t)i * n + j] * (ld)phi[j];
        e += (ld)phi[i] * r;
    }
    return (double)(0.5L * e);
}

void vmo_s_advection(double* x, const double* v, long np, double dt)
{
    /* src/models/vlasov_poisson.jl:53-58  z[1,i] = zbar[1,i] + (t-tbar)*zbar[2,i] */
    for (long p = 0; p < np; ++p) x[p] = x[p] + dt * v[p];
}

void vmo_s_acceleration(const double* x, double* v, const double* w, long np, double dt,
                        double a, double b, int n, int k, int shift, const double* S,
                        const double* x_src)
{
    /* src/models/vlasov_poisson.jl:61-67: update_potential! (:12-15) then
     * z[2,i] = zbar[2,i] - (t-tbar) * phi(zbar[1,i], Derivative(1)) */
    double* rhs = (double*)malloc(sizeof(double) * n);
    double* phi = (double*)malloc(sizeof(double) * n);
    double* dphi = (double*)malloc(sizeof(double) * (size_t)np);
    vmo_deposit_periodic(x_src ? x_src : x, w, np, a, b, n, k, shift, rhs);
    vmo_poisson_solve(S, n, rhs, phi);
    vmo_eval_dphi(x, np, a, b, n, k, shift, phi, dphi);
    for (long p = 0; p < np; ++p) v[p] = v[p] - dt * dphi[p];
    free(rhs); free(phi); free(dphi);
}

void vmo_vp_strang_step(double* x, double* v, const double* w, long np, double dt,
                        double a, double b, int n, int k, int shift, const double* S,
                        const double* x_src)
{
    double hdt = 0.5 * dt;
    vmo_s_advection(x, v, np, hdt);
    vmo_s_acceleration(x, v, w, np, hdt, a, b, n, k, shift, S, x_src);
    vmo_s_acceleration(x, v, w, np, hdt, a, b, n, k, shift, S, x_src);
    vmo_s_advection(x, v, np, hdt);
}

static void save_timestep(const double* x, const double* v, const double* w, long np, double chi,
                          double a, double b, int n, int k, int shift, const double* S,
                          double* diag_row, double* phi_row)
{
    /* update!(efield, x, w, t) then save_timestep! (src/vlasov_poisson.jl:58-67) */
    double* rhs = (double*)malloc(sizeof(double) * n);
    double* phi = (double*)malloc(sizeof(double) * n);
    vmo_deposit_periodic(x, w, np, a, b, n, k, shift, rhs);
    vmo_poisson_solve(S, n, rhs, phi);
    ld K = 0, M = 0;
    for (long p = 0; p < np; ++p) { K += (ld)w[p] * v[p] * v[p]; M += (ld)w[p] * v[p]; }
    diag_row[0] = vmo_field_energy(S, n, phi) / (chi * chi);   /* energy(ScaledField) electric_field.jl:33 */
    diag_row[1] = (double)(0.5L * K);
    diag_row[2] = (double)M;
    if (phi_row) memcpy(phi_row, phi, sizeof(double) * n);
    free(rhs); free(phi);
}

void vmo_integrate_vp(double* x, double* v, const double* w, long np, double dt, double chi,
                      int nt, int nsave, double a, double b, int n, int k, int shift,
                      const double* S, double* diag, double* phi_hist)
{
    double Dt = dt * chi;                                      /* src/vlasov_poisson.jl:80 */
    double* rhs = (double*)malloc(sizeof(double) * n);
    double* phi = (double*)malloc(sizeof(double) * n);
    double* acc = (double*)malloc(sizeof(double) * (size_t)np);
    int ts = 0;
    save_timestep(x, v, w, np, chi, a, b, n, k, shift, S, diag, phi_hist);   /* :88-91 */
    for (int it = 1; it <= nt; ++it) {                         /* :94 */
        for (long p = 0; p < np; ++p) x[p] += 0.5 * Dt * v[p]; /* :99 */
        /* :102 efield(a, x, w, t) = update! (deposit + solve) ; efield! ; a ./= chi^2 */
        vmo_deposit_periodic(x, w, np, a, b, n, k, shift, rhs);
        vmo_poisson_solve(S, n, rhs, phi);
        vmo_eval_dphi(x, np, a, b, n, k, shift, phi, acc);
        for (long p = 0; p < np; ++p) acc[p] = -acc[p] / (chi * chi);
        for (long p = 0; p < np; ++p) v[p] += Dt * acc[p];     /* :105 */
        for (long p = 0; p < np; ++p) x[p] += 0.5 * Dt * v[p]; /* :108 */
        if (nsave > 0 && it % nsave == 0) {                    /* :110-114 */
            ++ts;
            save_timestep(x, v, w, np, chi, a, b, n, k, shift, S, diag + 3 * (size_t)ts,
                          phi_hist ? phi_hist + (size_t)n * ts : NULL);
        }
    }
    free(rhs); free(phi); free(acc);
}

/* integrate_vp! driven by an ExternalField: update!(f, x, w, t) sets ts = round(t / dt_c) and
 * phi = coeffs[:, ts] (src/electric_field.jl:66-69; no deposit, no solve), efield! gathers, ScaledField divides
 * by chi^2 (:26-29); loop body src/vlasov_poisson.jl:94-115.  coeffs is column-major n x ncols (time index 0
 * first, as the OffsetMatrix of :62).  diag rows [W, K, M] with W = energy(::ExternalField) (:75) / chi^2. */
void vmo_integrate_vp_external(double* x, double* v, const double* w, long np, double dt, double chi,
                               int nt, int nsave, double a, double b, int n, int k, int shift,
                               const double* S, const double* coeffs, int ncols, double dt_c, double* diag)
{
    double Dt = dt * chi;
    double* acc = (double*)malloc(sizeof(double) * (size_t)np);
    int ts = 0, row = 0;
    (void)ncols;
    for (int it = 0; it <= nt; ++it) {
        const double t = it * dt;
        if (it > 0) {
            for (long p = 0; p < np; ++p) x[p] += 0.5 * Dt * v[p];
            ts = (int)nearbyint(t / dt_c);                      /* Julia round(): ties to even */
            vmo_eval_dphi(x, np, a, b, n, k, shift, coeffs + (size_t)n * ts, acc);
            for (long p = 0; p < np; ++p) acc[p] = -acc[p] / (chi * chi);
            for (long p = 0; p < np; ++p) v[p] += Dt * acc[p];
            for (long p = 0; p < np; ++p) x[p] += 0.5 * Dt * v[p];
        }
        if (nsave > 0 && it % nsave == 0) {
            ld K = 0, M = 0;
            ts = (int)nearbyint(t / dt_c);
            for (long p = 0; p < np; ++p) { K += (ld)w[p] * v[p] * v[p]; M += (ld)w[p] * v[p]; }
            diag[3 * row + 0] = vmo_field_energy(S, n, coeffs + (size_t)n * ts) / (chi * chi);
            diag[3 * row + 1] = (double)(0.5L * K);
            diag[3 * row + 2] = (double)M;
            ++row;
        }
    }
    free(acc);
}

/* lorentz_force!(zdot, t, z, params) (src/models/vlasov_poisson.jl:23-29): update_potential!(model) -- deposit
 * from x_src (NULL: the state itself, i.e. the self-consistent reading) and solve -- then
 * zdot[1,i] = z[2,i], zdot[2,i] = -phi(z[1,i], Derivative(1)). */
void vmo_lorentz_force(const double* x, const double* v, const double* w, long np, double a, double b,
                       int n, int k, int shift, const double* S, const double* x_src,
                       double* xdot, double* vdot)
{
    double* rhs = (double*)malloc(sizeof(double) * n);
    double* phi = (double*)malloc(sizeof(double) * n);
    vmo_deposit_periodic(x_src ? x_src : x, w, np, a, b, n, k, shift, rhs);
    vmo_poisson_solve(S, n, rhs, phi);
    vmo_eval_dphi(x, np, a, b, n, k, shift, phi, vdot);
    for (long p = 0; p < np; ++p) { xdot[p] = v[p]; vdot[p] = -vdot[p]; }
    free(rhs); free(phi);
}

/* One classical RK4 step of zdot = lorentz_force(z): what a generic (unsplit) Runge-Kutta driver of
 * GeometricIntegrators does with the vector field above.  Stage sums in the textbook order. */
void vmo_vp_rk4_step(double* x, double* v, const double* w, long np, double dt, double a, double b,
                     int n, int k, int shift, const double* S)
{
    size_t N = (size_t)np;
    double* buf = (double*)malloc(sizeof(double) * N * 10);
    double *xs = buf, *vs = buf + N, *kx[4], *kv[4];
    const double c[4] = {0.0, 0.5, 0.5, 1.0};
    for (int s = 0; s < 4; ++s) { kx[s] = buf + (2 + 2 * s) * N; kv[s] = buf + (3 + 2 * s) * N; }
    for (int s = 0; s < 4; ++s) {
        for (size_t p = 0; p < N; ++p) {
            xs[p] = s ? x[p] + (c[s] * dt) * kx[s - 1][p] : x[p];
            vs[p] = s ? v[p] + (c[s] * dt) * kv[s - 1][p] : v[p];
        }
        vmo_lorentz_force(xs, vs, w, np, a, b, n, k, shift, S, NULL, kx[s], kv[s]);
    }
    for (size_t p = 0; p < N; ++p) {
        x[p] += (dt / 6.0) * (kx[0][p] + 2.0 * kx[1][p] + 2.0 * kx[2][p] + kx[3][p]);
        v[p] += (dt / 6.0) * (kv[0][p] + 2.0 * kv[1][p] + 2.0 * kv[2][p] + kv[3][p]);
    }
    free(buf);
}

/* ====================================================================== */
/* v-space: projection, moments, LB / CLB right-hand sides, RK438         */
/* ====================================================================== */

/* banded Cholesky (half bandwidth bw) on dense row-major storage, long double */
static void banded_cholesky_solve(const double* M, int n, int bw, ld* rhs)
{
    ld* Lm = (ld*)calloc((size_t)n * n, sizeof(ld));
    for (int i = 0; i < n; ++i) {
        int j0 = i - bw < 0 ? 0 : i - bw;
        for (int j = j0; j <= i; ++j) {
            ld s = M[(size_t)i * n + j];
            int k0 = (i - bw > j - bw ? i - bw : j - bw); if (k0 < 0) k0 = 0;
            for (int q = k0; q < j; ++q) s -= Lm[(size_t)i * n + q] * Lm[(size_t)j * n + q];
            Lm[(size_t)i * n + j] = (i == j) ? sqrtl(s) : s / Lm[(size_t)j * n + j];
        }
    }
    for (int i = 0; i < n; ++i) {           /* L y = b */
        ld s = rhs[i];
        int j0 = i - bw < 0 ? 0 : i - bw;
        for (int j = j0; j < i; ++j) s -= Lm[(size_t)i * n + j] * rhs[j];
        rhs[i] = s / Lm[(size_t)i * n + i];
    }
    for (int i = n - 1; i >= 0; --i) {      /* L^T x = y */
        ld s = rhs[i];
        int j1 = i + bw >= n ? n - 1 : i + bw;
        for (int j = i + 1; j <= j1; ++j) s -= Lm[(size_t)j * n + i] * rhs[j];
        rhs[i] = s / Lm[(size_t)i * n + i];
    }
    free(Lm);
}

void vmo_vproject(const double* v, const double* w, long np, double a, double b,
                  int nknots, int k, const double* M, double* coef, double* rhs_out)
{
    int npar = nknots + k - 2, nv = npar - 2;
    ld* rhs = (ld*)calloc((size_t)nv, sizeof(ld));               /* distribution.jl:36 zeros */
    for (long p = 0; p < np; ++p) {                              /* :40 */
        double N[VMO_MAXK];
        int c = vmo_clamped_eval(a, b, nknots, k, v[p], N, NULL);/* :41 basis(velocities[p]) */
        if (c < 0) continue;                                     /* outside domain: all basis values 0 */
        for (int d = 0; d < k; ++d) {                            /* :45-48, decreasing index like bs */
            int par = c + k - 1 - d;
            if (par == 0 || par == npar - 1) continue;           /* dropped by the Dirichlet recombination */
            rhs[par - 1] += (ld)N[k - 1 - d] * (ld)w[p];         /* :47 rhs[i] += bi * w[1,p] */
        }
    }
    if (rhs_out) for (int i = 0; i < nv; ++i) rhs_out[i] = (double)rhs[i];
    banded_cholesky_solve(M, nv, k - 1, rhs);                    /* :52 ldiv!(coefficients, mass_fact, rhs) */
    for (int i = 0; i < nv; ++i) coef[i] = (double)rhs[i];
    free(rhs);
}

static void spline_point(double a, double b, int nknots, int k, const double* coef, double v,
                         ld* f, ld* df)
{
    int npar = nknots + k - 2;
    double N[VMO_MAXK], dN[VMO_MAXK];
    int c = vmo_clamped_eval(a, b, nknots, k, v, N, dN);
    ld sf = 0, sd = 0;
    if (c >= 0) {
        for (int j = 0; j < k; ++j) {
            int par = c + j;
            if (par == 0 || par == npar - 1) continue;
            sf += (ld)coef[par - 1] * (ld)N[j];
            sd += (ld)coef[par - 1] * (ld)dN[j];
        }
    }
    *f = sf; *df = sd;
}

void vmo_vspline_eval(const double* v, long np, double a, double b, int nknots, int k,
                      const double* coef, double* f, double* df)
{
    for (long p = 0; p < np; ++p) {
        ld sf, sd;
        spline_point(a, b, nknots, k, coef, v[p], &sf, &sd);
        if (f) f[p] = (double)sf;
        if (df) df[p] = (double)sd;
    }
}

void vmo_vmoments(const double* v, long np, double a, double b, int nknots, int k,
                  const double* coef, double* out5)
{
    /* density.jl:43-52: sum(moment.(vp) .* spline.(vp)) -- UNWEIGHTED particle sums */
    ld m[5] = {0, 0, 0, 0, 0};
    for (long p = 0; p < np; ++p) {
        ld f, df;
        spline_point(a, b, nknots, k, coef, v[p], &f, &df);
        m[0] += f; m[1] += (ld)v[p] * f; m[2] += (ld)v[p] * v[p] * f;
        m[3] += df; m[4] += (ld)v[p] * df;
    }
    for (int i = 0; i < 5; ++i) out5[i] = (double)m[i];
}

void vmo_clb_coefficients(const double* m5, double* A1, double* A2)
{
    /* lenard_bernstein_conservative.jl:13-18 */
    ld n = m5[0], nu = m5[1], neps = m5[2];
    ld B1 = -(ld)m5[3], B2 = -(ld)m5[4];
    *A1 = (double)((neps * B1 - nu * B2) / (n * neps - nu * nu));
    *A2 = (double)(-(nu * B1 - n * B2) / (n * neps - nu * nu));
}

void vmo_lb_rhs(const double* v, const double* w, long np, double a, double b, int nknots,
                int k, const double* M, double nu, int conservative, double* vdot,
                double* coef_out, double* A_out)
{
    int nv = nknots + k - 4;
    double* coef = (double*)malloc(sizeof(double) * nv);
    vmo_vproject(v, w, np, a, b, nknots, k, M, coef, NULL);      /* fs = projection(v, idist, dist) */
    double A1 = 0.0, A2 = 1.0;                                   /* LB: vdot = -nu (f' + v f) */
    if (conservative) {
        double m5[5];
        vmo_vmoments(v, np, a, b, nknots, k, coef, m5);          /* compute_coefficients */
        vmo_clb_coefficients(m5, &A1, &A2);
    }
    for (long p = 0; p < np; ++p) {
        ld f, df;
        spline_point(a, b, nknots, k, coef, v[p], &f, &df);
        vdot[p] = (double)(-(ld)nu * (df + ((ld)A1 + (ld)A2 * v[p]) * f));
    }
    if (coef_out) memcpy(coef_out, coef, sizeof(double) * nv);
    if (A_out) { A_out[0] = A1; A_out[1] = A2; }
    free(coef);
}

void vmo_lb_rk438_step(double* v, const double* w, long np, double dt, double a, double b,
                       int nknots, int k, const double* M, double nu, int conservative)
{
    /* classical 3/8 rule: c=(0,1/3,2/3,1); a21=1/3; a31=-1/3,a32=1; a41=1,a42=-1,a43=1;
     * b=(1/8,3/8,3/8,1/8)  (GeometricIntegrators RK438) */
    size_t nb = sizeof(double) * (size_t)np;
    double *k1 = malloc(nb), *k2 = malloc(nb), *k3 = malloc(nb), *k4 = malloc(nb), *q = malloc(nb);
    vmo_lb_rhs(v, w, np, a, b, nknots, k, M, nu, conservative, k1, NULL, NULL);
    for (long p = 0; p < np; ++p) q[p] = v[p] + dt * (k1[p] / 3.0);
    vmo_lb_rhs(q, w, np, a, b, nknots, k, M, nu, conservative, k2, NULL, NULL);
    for (long p = 0; p < np; ++p) q[p] = v[p] + dt * (-k1[p] / 3.0 + k2[p]);
    vmo_lb_rhs(q, w, np, a, b, nknots, k, M, nu, conservative, k3, NULL, NULL);
    for (long p = 0; p < np; ++p) q[p] = v[p] + dt * (k1[p] - k2[p] + k3[p]);
    vmo_lb_rhs(q, w, np, a, b, nknots, k, M, nu, conservative, k4, NULL, NULL);
    for (long p = 0; p < np; ++p)
        v[p] = (double)((ld)v[p] + (ld)dt * (((ld)k1[p] + 3.0L * k2[p] + 3.0L * k3[p] + (ld)k4[p]) / 8.0L));
    free(k1); free(k2); free(k3); free(k4); free(q);
}

/* ====================================================================== */
/* Timed CPU baseline: "C restatement of the reference algorithm"         */
/* plain double, reference loop structure, optional OpenMP                */
/* ====================================================================== */

int vmo_max_threads(void)
{
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

static void bl_deposit(const double* x, const double* w, long np, double a, double b, int n, int k,
                       int shift, double* rhs, int nthreads)
{
    memset(rhs, 0, sizeof(double) * n);
#ifdef _OPENMP
    #pragma omp parallel num_threads(nthreads)
#endif
    {
        double* loc = (double*)calloc((size_t)n, sizeof(double));
#ifdef _OPENMP
        #pragma omp for schedule(static)
#endif
        for (long p = 0; p < np; ++p) {
            double N[VMO_MAXK];
            int c = vmo_periodic_eval(a, b, n, k, x[p], N, NULL);
            for (int d = 0; d < k; ++d) loc[pmod(c + k - 1 + shift - d, n)] += w[p] * N[k - 1 - d];
        }
#ifdef _OPENMP
        #pragma omp critical
#endif
        for (int i = 0; i < n; ++i) rhs[i] += loc[i];
        free(loc);
    }
    (void)nthreads;
}

static void bl_dense_solve(const double* Ainv, int n, const double* bvec, double* out)
{
    for (int i = 0; i < n; ++i) {
        double s = 0;
        for (int j = 0; j < n; ++j) s += Ainv[(size_t)i * n + j] * bvec[j];
        out[i] = s;
    }
}

static void bl_kick(const double* x, double* v, const double* w, long np, double dt, double a, double b,
                    int n, int k, int shift, const double* Ainv, double* rhs, double* phi, int nthreads)
{
    /* s_acceleration!: deposit + solve + per-particle phi'(x) evaluation */
    bl_deposit(x, w, np, a, b, n, k, shift, rhs, nthreads);
    double mean = 0; for (int i = 0; i < n; ++i) mean += rhs[i]; mean /= n;
    for (int i = 0; i < n; ++i) rhs[i] -= mean;
    bl_dense_solve(Ainv, n, rhs, phi);
#ifdef _OPENMP
    #pragma omp parallel for schedule(static) num_threads(nthreads)
#endif
    for (long p = 0; p < np; ++p) {
        double dN[VMO_MAXK];
        int c = vmo_periodic_eval(a, b, n, k, x[p], NULL, dN);
        double s = 0;
        for (int j = 0; j < k; ++j) s += phi[pmod(c + j + shift, n)] * dN[j];
        v[p] = v[p] - dt * s;
    }
}

void vmo_baseline_vp_steps(double* x, double* v, const double* w, long np, double dt, int nsteps,
                           double a, double b, int n, int k, int shift, int nthreads)
{
    /* factor once (the reference factors at construction), solve per kick */
    double* S = (double*)malloc(sizeof(double) * (size_t)n * n);
    double* Ainv = (double*)malloc(sizeof(double) * (size_t)n * n);
    vmo_periodic_stiffness(a, b, n, k, shift, S);
    {
        ld* col = (ld*)malloc(sizeof(ld) * n);
        ld* A = (ld*)malloc(sizeof(ld) * (size_t)n * n);
        for (int j = 0; j < n; ++j) {
            for (size_t i = 0; i < (size_t)n * n; ++i) A[i] = (ld)S[i] + 1.0L / n;
            for (int i = 0; i < n; ++i) col[i] = (i == j);
            lu_solve_ld(A, n, col);
            for (int i = 0; i < n; ++i) Ainv[(size_t)i * n + j] = (double)col[i];
        }
        free(col); free(A);
    }
    double* rhs = (double*)malloc(sizeof(double) * n);
    double* phi = (double*)malloc(sizeof(double) * n);
    double hdt = 0.5 * dt;
    for (int s = 0; s < nsteps; ++s) {
        /* Strang = A(dt/2) B(dt/2) B(dt/2) A(dt/2), each B re-deposits and re-solves */
#ifdef _OPENMP
        #pragma omp parallel for schedule(static) num_threads(nthreads)
#endif
        for (long p = 0; p < np; ++p) x[p] = x[p] + hdt * v[p];
        bl_kick(x, v, w, np, hdt, a, b, n, k, shift, Ainv, rhs, phi, nthreads);
        bl_kick(x, v, w, np, hdt, a, b, n, k, shift, Ainv, rhs, phi, nthreads);
#ifdef _OPENMP
        #pragma omp parallel for schedule(static) num_threads(nthreads)
#endif
        for (long p = 0; p < np; ++p) x[p] = x[p] + hdt * v[p];
    }
    free(S); free(Ainv); free(rhs); free(phi);
}

void vmo_baseline_lb_rhs(const double* v, const double* w, long np, double a, double b, int nknots,
                         int k, double nu, int conservative, double* vdot, int nthreads)
{
    int npar = nknots + k - 2, nv = npar - 2;
    double* M = (double*)malloc(sizeof(double) * (size_t)nv * nv);
    vmo_dirichlet_mass(a, b, nknots, k, M);
    double* rhs = (double*)calloc((size_t)nv, sizeof(double));
#ifdef _OPENMP
    #pragma omp parallel num_threads(nthreads)
#endif
    {
        double* loc = (double*)calloc((size_t)nv, sizeof(double));
#ifdef _OPENMP
        #pragma omp for schedule(static)
#endif
        for (long p = 0; p < np; ++p) {
            double N[VMO_MAXK];
            int c = vmo_clamped_eval(a, b, nknots, k, v[p], N, NULL);
            if (c < 0) continue;
            for (int j = 0; j < k; ++j) {
                int par = c + j;
                if (par == 0 || par == npar - 1) continue;
                loc[par - 1] += N[j] * w[p];
            }
        }
#ifdef _OPENMP
        #pragma omp critical
#endif
        for (int i = 0; i < nv; ++i) rhs[i] += loc[i];
        free(loc);
    }
    ld* r = (ld*)malloc(sizeof(ld) * nv);
    for (int i = 0; i < nv; ++i) r[i] = rhs[i];
    banded_cholesky_solve(M, nv, k - 1, r);
    double* coef = (double*)malloc(sizeof(double) * nv);
    for (int i = 0; i < nv; ++i) coef[i] = (double)r[i];
    double A1 = 0.0, A2 = 1.0;
    if (conservative) {
        /* five separate particle passes, as density.jl does */
        double m5[5];
        for (int which = 0; which < 5; ++which) {
            double s = 0;
#ifdef _OPENMP
            #pragma omp parallel for schedule(static) reduction(+:s) num_threads(nthreads)
#endif
            for (long p = 0; p < np; ++p) {
                ld f, df;
                spline_point(a, b, nknots, k, coef, v[p], &f, &df);
                double val = (which < 3) ? (double)f : (double)df;
                double mom = (which == 0 || which == 3) ? 1.0 : ((which == 1 || which == 4) ? v[p] : v[p] * v[p]);
                s += mom * val;
            }
            m5[which] = s;
        }
        vmo_clb_coefficients(m5, &A1, &A2);
    }
#ifdef _OPENMP
    #pragma omp parallel for schedule(static) num_threads(nthreads)
#endif
    for (long p = 0; p < np; ++p) {
        ld f, df, f2, df2;
        spline_point(a, b, nknots, k, coef, v[p], &f, &df2);   /* fs.(v)   */
        spline_point(a, b, nknots, k, coef, v[p], &f2, &df);   /* dfdv.(v) */
        vdot[p] = -nu * ((double)df + (A1 + A2 * v[p]) * (double)f);
    }
    free(M); free(rhs); free(r); free(coef);
    (void)nthreads;
}

/*
 * vm_oracle.h -- CPU ORACLE for the VlasovMethods.jl particle hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing in the product path (the package
 * vlasovmethods.jl_b200/ and its CUDA library) may include, link, import or
 * execute anything under oracle/.  Only tests/, __graft_entry__.smoke() and
 * bench.py's cpu_baseline / --impl reference legs use it, and only as the
 * checker or as the timed CPU baseline.
 *
 * PARITY STATUS: "parity unpinned" for the third-party arithmetic.
 * The reference (JuliaPlasma/VlasovMethods.jl v0.2.1, 100 % Julia) cannot be
 * run here (no Julia toolchain), and the B-spline / Poisson / integrator
 * arithmetic lives in un-vendored packages with no Manifest pin:
 *   BSplineKit.jl 0.14-0.17, PoissonSolvers.jl 0.3.5, ParticleMethods.jl 0.1,
 *   GeometricIntegrators.jl 0.13 (reference Project.toml:30-50).
 * The reference's own tests hold no golden vectors for this path
 * (test/projections_tests.jl:6-34 is a 5e-2 statistical check, restated in
 * tests/).  This oracle therefore restates the *published algorithms* of those
 * packages at the reference's call sites and is pinned by known-answer tests
 * (tests/test_oracle_kat.py): scipy.interpolate.BSpline cross-checks, closed
 * form circulant mass/stiffness stencils, a manufactured Poisson solution,
 * partition of unity, Dirichlet-basis facts, Maxwellian fixed point of LB.
 *
 * Each function cites the reference file:line it follows
 * (paths relative to /root/reference).
 */
#ifndef VM_ORACLE_H
#define VM_ORACLE_H

#ifdef __cplusplus
extern "C" {
#endif

/* ---------------------------------------------------------------- basis -- */

/* All nonzero order-k B-splines at x for the knot span s (t[s] <= x < t[s+1]).
 * N[j] = B_{s-k+1+j,k}(x), j = 0..k-1 (increasing basis index).
 * BSplineKit `evaluate_all` (call site src/projections/potential.jl:11) returns
 * the same numbers in DEcreasing index order (ilast, ilast-1, ...). */
void vmo_eval_all(const double* t, int s, int k, double x, double* N);

/* d/dx of the same k functions. */
void vmo_eval_all_deriv(const double* t, int s, int k, double x, double* dN);

/* Periodic uniform basis: n functions on [a,b), order k, index rotation
 * `shift`: a particle in cell c touches basis indices (c+j+shift) mod n.
 * Returns the cell c and writes N[0..k-1].  x is reduced into [a,b) first. */
int vmo_periodic_eval(double a, double b, int n, int k, double x, double* N, double* dN);

/* Clamped basis on breakpoints LinRange(a,b,nknots), order k
 * (src/distributions/spline_distribution.jl:24-25).  Returns the cell
 * (0..nknots-2) or -1 when x is outside [a,b]; N[j] = parent B_{c+j}(x). */
int vmo_clamped_eval(double a, double b, int nknots, int k, double x, double* N, double* dN);

/* ----------------------------------------------------------- matrices ---- */
/* Dense n x n (row-major) Galerkin matrices by Gauss-Legendre quadrature
 * (k nodes per cell), as BSplineKit.galerkin_matrix does
 * (call site src/distributions/spline_distribution.jl:10). */
void vmo_periodic_mass(double a, double b, int n, int k, int shift, double* M);
void vmo_periodic_stiffness(double a, double b, int n, int k, int shift, double* S);
/* Dirichlet-recombined clamped basis: nv = nknots + k - 4 functions. */
void vmo_dirichlet_mass(double a, double b, int nknots, int k, double* M);

/* ------------------------------------------------- x-space (Vlasov-Poisson) */
/* rhs_i = sum_p w_p B_i(x_p), periodic index wrap, rhs zeroed first.
 * Follows src/projections/potential.jl:2-22 (loop :10-19).
 * Accumulates in long double (extended precision) -- a better reference
 * than the sequential Float64 sum of the Julia code. */
void vmo_deposit_periodic(const double* x, const double* w, long np,
                          double a, double b, int n, int k, int shift, double* rhs);

/* Periodic Poisson solve, call site src/models/vlasov_poisson.jl:14 /
 * src/electric_field.jl:45 (PoissonSolvers.jl, absent): find phi with
 *   S phi = rhs - mean(rhs),  sum(phi) = 0
 * via dense LU of (S + 11^T/n) in long double. */
void vmo_poisson_solve(const double* S, int n, const double* rhs, double* phi);

/* phi'(x_p) = sum_j phi_j B_j'(x_p): src/models/vlasov_poisson.jl:27,48,65. */
void vmo_eval_dphi(const double* x, long np, double a, double b, int n, int k, int shift,
                   const double* phi, double* dphi);

/* W = 1/2 phi^T S phi  (src/electric_field.jl:47). */
double vmo_field_energy(const double* S, int n, const double* phi);

/* Exact sub-flows of the splitting: src/models/vlasov_poisson.jl:53-67. */
void vmo_s_advection(double* x, const double* v, long np, double dt);
void vmo_s_acceleration(const double* x, double* v, const double* w, long np, double dt,
                        double a, double b, int n, int k, int shift, const double* S,
                        const double* x_src /* deposit source; NULL -> x (self-consistent),
                                               else frozen model particles */);

/* One Strang step as GeometricIntegrators composes the two exact flows
 * (src/models/vlasov_poisson.jl:73-89, src/methods/splitting.jl:41):
 * A(dt/2) B(dt/2) B(dt/2) A(dt/2). */
void vmo_vp_strang_step(double* x, double* v, const double* w, long np, double dt,
                        double a, double b, int n, int k, int shift, const double* S,
                        const double* x_src);

/* Legacy leapfrog loop src/vlasov_poisson.jl:70-119 with ScaledField chi
 * (src/electric_field.jl:21-35).  diag has (nt/nsave + 1) rows of [W,K,M]
 * (save_timestep! :58-67).  phi_hist (optional) gets n doubles per row. */
void vmo_integrate_vp(double* x, double* v, const double* w, long np, double dt, double chi,
                      int nt, int nsave, double a, double b, int n, int k, int shift,
                      const double* S, double* diag, double* phi_hist);

/* integrate_vp! with an ExternalField (src/electric_field.jl:55-77, src/vlasov_poisson.jl:94-115) */
void vmo_integrate_vp_external(double* x, double* v, const double* w, long np, double dt, double chi,
                               int nt, int nsave, double a, double b, int n, int k, int shift,
                               const double* S, const double* coeffs, int ncols, double dt_c, double* diag);
/* lorentz_force! (src/models/vlasov_poisson.jl:23-29) and one RK4 step of it */
void vmo_lorentz_force(const double* x, const double* v, const double* w, long np, double a, double b,
                       int n, int k, int shift, const double* S, const double* x_src,
                       double* xdot, double* vdot);
void vmo_vp_rk4_step(double* x, double* v, const double* w, long np, double dt, double a, double b,
                     int n, int k, int shift, const double* S);

/* ------------------------------------------- v-space (Lenard-Bernstein) --- */
/* projection(v, dist, sdist): src/projections/distribution.jl:35-55.
 * coef[nv] = M^{-1} (sum_p w_p phi_i(v_p)); also returns the raw rhs if
 * rhs_out != NULL. Banded Cholesky as spline_distribution.jl:11 + ldiv! :52. */
void vmo_vproject(const double* v, const double* w, long np, double a, double b,
                  int nknots, int k, const double* M, double* coef, double* rhs_out);

/* f_s(v_p) and f_s'(v_p) for the Dirichlet spline with coefficients coef. */
void vmo_vspline_eval(const double* v, long np, double a, double b, int nknots, int k,
                      const double* coef, double* f, double* df);

/* Five unweighted particle sums, src/projections/density.jl:6-52:
 * out = [sum f, sum v f, sum v^2 f, sum f', sum v f']. */
void vmo_vmoments(const double* v, long np, double a, double b, int nknots, int k,
                  const double* coef, double* out5);

/* A1, A2 of src/models/lenard_bernstein_conservative.jl:11-21. */
void vmo_clb_coefficients(const double* m5, double* A1, double* A2);

/* LB_rhs! (src/models/lenard_bernstein.jl:20-30) when conservative == 0,
 * CLB_rhs! (src/models/lenard_bernstein_conservative.jl:24-36) otherwise. */
void vmo_lb_rhs(const double* v, const double* w, long np, double a, double b, int nknots,
                int k, const double* M, double nu, int conservative, double* vdot,
                double* coef_out /* nv, optional */, double* A_out /* 2, optional */);

/* One RK438 (3/8 rule) step of the LB/CLB velocity ODE as driven by
 * src/methods/geometric_integrator.jl:31-32. */
void vmo_lb_rk438_step(double* v, const double* w, long np, double dt, double a, double b,
                       int nknots, int k, const double* M, double nu, int conservative);

/* ------------------------------------------------ timed CPU baseline ------ */
/* "C restatement of the reference algorithm" (BASELINE.md section 2): plain
 * double accumulators, per-particle Cox-de Boor with knot search, 2 deposits
 * + 2 solves per Strang step, separate passes.  nthreads = 1 reproduces the
 * single-threaded reference; > 1 uses OpenMP with thread-private grids. */
void vmo_baseline_vp_steps(double* x, double* v, const double* w, long np, double dt, int nsteps,
                           double a, double b, int n, int k, int shift, int nthreads);
void vmo_baseline_lb_rhs(const double* v, const double* w, long np, double a, double b, int nknots,
                         int k, double nu, int conservative, double* vdot, int nthreads);
int  vmo_max_threads(void);

#ifdef __cplusplus
}
#endif
#endif

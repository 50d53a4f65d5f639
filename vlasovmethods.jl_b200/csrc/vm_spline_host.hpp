// vm_spline_host.hpp -- host-side (setup-time) B-spline algebra in long double.
// Piecewise-polynomial form of B-splines on arbitrary knots, exact Galerkin integrals,
// circulant pseudo-inverse of the periodic stiffness matrix, banded SPD inverse.
#pragma once

#include <cmath>
#include <vector>

namespace vmhost {

typedef long double ld;
constexpr int MAXK = 8;

// Polynomial pieces of the k B-splines that are nonzero on the knot span [t[k-1], t[k]) of the
// 2k-knot window t: B_{j}(x0 + hc*xi) = sum_m A[j*k + m] xi^m, j = 0..k-1 (increasing index).
// Cox-de Boor recursion carried out on polynomial coefficients.
inline void cell_polys(const ld* t, int k, ld x0, ld hc, ld* A)
{
    const int s = k - 1;
    ld N[MAXK][MAXK] = {};
    N[0][0] = 1;
    for (int j = 1; j < k; ++j) {
        ld saved[MAXK] = {};
        for (int r = 0; r < j; ++r) {
            // right = (t[s+r+1] - x0) - hc*xi ; left = (x0 - t[s+1-(j-r)]) + hc*xi
            const ld r0 = t[s + r + 1] - x0, l0 = x0 - t[s + 1 - (j - r)];
            const ld den = t[s + r + 1] - t[s + 1 - (j - r)];
            ld temp[MAXK] = {};
            if (den != 0) for (int m = 0; m < j; ++m) temp[m] = N[r][m] / den;
            ld nr[MAXK] = {}, ns[MAXK] = {};
            for (int m = 0; m < j; ++m) {
                nr[m] += r0 * temp[m];
                nr[m + 1] -= hc * temp[m];
                ns[m] += l0 * temp[m];
                ns[m + 1] += hc * temp[m];
            }
            for (int m = 0; m <= j; ++m) { N[r][m] = saved[m] + nr[m]; saved[m] = ns[m]; }
        }
        for (int m = 0; m <= j; ++m) N[j][m] = saved[m];
    }
    for (int j = 0; j < k; ++j)
        for (int m = 0; m < k; ++m) A[j * k + m] = N[j][m];
}

// int_0^1 P(xi) Q(xi) dxi for coefficient arrays of length k
inline ld poly_dot(const ld* p, const ld* q, int k)
{
    ld s = 0;
    for (int a = 0; a < k; ++a)
        for (int b = 0; b < k; ++b) s += p[a] * q[b] / (ld)(a + b + 1);
    return s;
}

// int_0^1 P'(xi) Q'(xi) dxi
inline ld poly_ddot(const ld* p, const ld* q, int k)
{
    ld s = 0;
    for (int a = 1; a < k; ++a)
        for (int b = 1; b < k; ++b) s += (ld)a * (ld)b * p[a] * q[b] / (ld)(a + b - 1);
    return s;
}

// Uniform cardinal B-splines of order k, unit spacing: mass[d] = int B_0 B_d, stiff[d] = int B_0' B_d'
// (multiply mass by h and divide stiff by h for spacing h).
inline void uniform_stencils(int k, ld* mass, ld* stiff)
{
    ld t[2 * MAXK], A[MAXK * MAXK];
    for (int i = 0; i < 2 * k; ++i) t[i] = (ld)(i - (k - 1));
    cell_polys(t, k, 0, 1, A);
    for (int d = 0; d < k; ++d) {
        ld m = 0, s = 0;
        for (int j = 0; j + d < k; ++j) {
            m += poly_dot(A + j * k, A + (j + d) * k, k);
            s += poly_ddot(A + j * k, A + (j + d) * k, k);
        }
        mass[d] = m;
        stiff[d] = s;
    }
}

// Pseudo-inverse (gauge sum = 0) of the symmetric circulant matrix with stencil st[0..k-1]
// (centre, +-1, ...): returns its first column G (length n).
inline bool circulant_pinv(const ld* st, int k, int n, std::vector<ld>& G)
{
    const ld PI = 3.141592653589793238462643383279502884L;
    std::vector<ld> row(n, 0), cs(n);
    row[0] += st[0];
    for (int d = 1; d < k; ++d) { row[d % n] += st[d]; row[(n - (d % n)) % n] += st[d]; }
    for (int q = 0; q < n; ++q) cs[q] = cosl(2 * PI * (ld)q / (ld)n);
    std::vector<ld> lam(n);
    for (int m = 0; m < n; ++m) {
        ld s = 0;
        for (int j = 0; j < n; ++j) s += row[j] * cs[(size_t)((long long)m * j % n)];
        lam[m] = s;
    }
    G.assign(n, 0);
    for (int m = 1; m < n; ++m)
        if (!(lam[m] > 0)) return false;
    for (int j = 0; j < n; ++j) {
        ld s = 0;
        for (int m = 1; m < n; ++m) s += cs[(size_t)((long long)m * j % n)] / lam[m];
        G[j] = s / (ld)n;
    }
    return true;
}

// Dense inverse of a symmetric positive definite banded matrix (half bandwidth bw), row-major n x n.
inline bool spd_banded_inverse(const std::vector<ld>& M, int n, int bw, std::vector<ld>& inv)
{
    std::vector<ld> L((size_t)n * n, 0);
    for (int i = 0; i < n; ++i) {
        for (int j = (i - bw < 0 ? 0 : i - bw); j <= i; ++j) {
            ld s = M[(size_t)i * n + j];
            int q0 = i - bw; if (q0 < 0) q0 = 0;
            for (int q = q0; q < j; ++q) s -= L[(size_t)i * n + q] * L[(size_t)j * n + q];
            if (i == j) { if (!(s > 0)) return false; L[(size_t)i * n + j] = sqrtl(s); }
            else L[(size_t)i * n + j] = s / L[(size_t)j * n + j];
        }
    }
    inv.assign((size_t)n * n, 0);
    std::vector<ld> y(n);
    for (int c = 0; c < n; ++c) {
        for (int i = 0; i < n; ++i) {
            ld s = (i == c) ? 1 : 0;
            for (int j = (i - bw < 0 ? 0 : i - bw); j < i; ++j) s -= L[(size_t)i * n + j] * y[j];
            y[i] = s / L[(size_t)i * n + i];
        }
        for (int i = n - 1; i >= 0; --i) {
            ld s = y[i];
            int j1 = i + bw >= n ? n - 1 : i + bw;
            for (int j = i + 1; j <= j1; ++j) s -= L[(size_t)j * n + i] * y[j];
            y[i] = s / L[(size_t)i * n + i];
        }
        for (int i = 0; i < n; ++i) inv[(size_t)i * n + c] = y[i];
    }
    return true;
}

}  // namespace vmhost

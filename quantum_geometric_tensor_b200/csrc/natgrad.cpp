// natgrad.cpp — regularised natural-gradient step  x = (G + lambda I)^-1 g  on the real P x P metric.
//
// Semantics of compute_regularized_natural_gradient and its helpers
// (reference src/quantum_geometric/core/quantum_geometric_gradient.c:2721-2964):
//   * kappa = sigma_max / sigma_min over the positive singular values (:2763-2774), infinity when
//     sigma_min <= 1e-15;
//   * adaptive regularisation lambda = max(lambda, 1e-6 * sqrt(kappa)) when kappa > threshold (:2898-2912);
//   * Tikhonov G + lambda I, solve (:2920-2959); if the regularised matrix is singular and the fallback is
//     enabled, SVD pseudo-inverse of G with the singular-value cutoff (:2800-2885).
// P is a few hundred at most, so this stays on the host; one symmetric
// eigen-decomposition G = V diag(w) V^T (Householder + QL) serves the condition number, the solve and the pseudo-inverse
// (for a symmetric matrix the singular values are |w|).
#include <cmath>
#include <cstring>
#include <limits>
#include <vector>

#include "ctx.hpp"

namespace {

// Symmetric eigen-decomposition A = V diag(w) V^T: Householder reduction to tridiagonal form followed by
// implicit-shift QL iterations (the classic tred2 / tql2 pair), O(n^3) with a small constant — a 256 x 256
// metric takes a few milliseconds.  A (row-major) is destroyed; V receives the eigenvectors as columns.
// sqrt(a^2 + b^2) without std::hypot's slow path; scaled only when squaring could leave the normal range
inline double pythag(double a, double b) {
    const double x = std::fabs(a), y = std::fabs(b);
    const double m = x > y ? x : y;
    if (m > 1e-140 && m < 1e140) return std::sqrt(a * a + b * b);
    return std::hypot(a, b);
}

// dot product with four independent accumulators (fixed summation order; lets the compiler keep four
// multiply-add chains in flight without value-changing flags)
inline double dot4(const double* a, const double* b, int n) {
    double s0 = 0.0, s1 = 0.0, s2 = 0.0, s3 = 0.0;
    int k = 0;
    for (; k + 4 <= n; k += 4) { s0 += a[k] * b[k]; s1 += a[k + 1] * b[k + 1]; s2 += a[k + 2] * b[k + 2]; s3 += a[k + 3] * b[k + 3]; }
    for (; k < n; k++) s0 += a[k] * b[k];
    return (s0 + s1) + (s2 + s3);
}

// Eigenvalues only: the same Householder reduction with the matrix-vector product and the rank-2 update of every
// step running over contiguous row prefixes of the lower triangle (the classic form below walks columns);
// d / e receive the tridiagonal matrix in the layout the QL loop expects.
void tridiagonalize_values(std::vector<double>& A, int n, std::vector<double>& d, std::vector<double>& e) {
    std::vector<double> u(n), pv(n);
    for (int i = n - 1; i >= 1; i--) {
        const int l = i - 1;
        double* ai = &A[(size_t)i * n];
        if (l == 0) { e[i] = ai[0]; continue; }
        double scale = 0.0;
        for (int k = 0; k <= l; k++) scale += std::fabs(ai[k]);
        if (scale == 0.0) { e[i] = ai[l]; continue; }
        double h = 0.0;
        for (int k = 0; k <= l; k++) { u[k] = ai[k] / scale; h += u[k] * u[k]; }
        const double f = u[l];
        const double g = f >= 0.0 ? -std::sqrt(h) : std::sqrt(h);
        e[i] = scale * g;
        h -= f * g;
        u[l] = f - g;
        // p = A u / h from the lower triangle only: row j contributes a dot product to p[j] and an axpy to p[0..j-1],
        // both over the contiguous row prefix; the rank-2 update then touches the lower triangle only (half the traffic)
        for (int j = 0; j <= l; j++) pv[j] = 0.0;
        for (int j = 0; j <= l; j++) {
            const double* aj = &A[(size_t)j * n];
            const double uj = u[j];
            pv[j] += dot4(aj, u.data(), j) + aj[j] * uj;
            for (int k = 0; k < j; k++) pv[k] += aj[k] * uj;
        }
        double K = 0.0;
        for (int j = 0; j <= l; j++) { pv[j] /= h; K += pv[j] * u[j]; }
        K /= (h + h);
        for (int j = 0; j <= l; j++) pv[j] -= K * u[j];
        for (int j = 0; j <= l; j++) {
            double* aj = &A[(size_t)j * n];
            const double uj = u[j], qj = pv[j];
            for (int k = 0; k <= j; k++) aj[k] -= uj * pv[k] + qj * u[k];
        }
    }
    e[0] = 0.0;
    for (int i = 0; i < n; i++) d[i] = A[(size_t)i * n + i];
}

void sym_eigh(std::vector<double>& A, int n, std::vector<double>& w, std::vector<double>& V, bool want_vectors) {
    std::vector<double> d(n), e(n);
    auto a = [&](int i, int j) -> double& { return A[(size_t)i * n + j]; };
    if (!want_vectors) tridiagonalize_values(A, n, d, e);
    else
    // --- Householder tridiagonalisation, accumulating the transformation in A
    { for (int i = n - 1; i >= 1; i--) {
        const int l = i - 1;
        double h = 0.0, scale = 0.0;
        if (l > 0) {
            for (int k = 0; k <= l; k++) scale += std::fabs(a(i, k));
            if (scale == 0.0) e[i] = a(i, l);
            else {
                for (int k = 0; k <= l; k++) { a(i, k) /= scale; h += a(i, k) * a(i, k); }
                double f = a(i, l);
                double g = f >= 0.0 ? -std::sqrt(h) : std::sqrt(h);
                e[i] = scale * g;
                h -= f * g;
                a(i, l) = f - g;
                f = 0.0;
                for (int j = 0; j <= l; j++) {
                    a(j, i) = a(i, j) / h;
                    g = 0.0;
                    for (int k = 0; k <= j; k++) g += a(j, k) * a(i, k);
                    for (int k = j + 1; k <= l; k++) g += a(k, j) * a(i, k);
                    e[j] = g / h;
                    f += e[j] * a(i, j);
                }
                const double hh = f / (h + h);
                for (int j = 0; j <= l; j++) {
                    f = a(i, j);
                    e[j] = g = e[j] - hh * f;
                    for (int k = 0; k <= j; k++) a(j, k) -= f * e[k] + g * a(i, k);
                }
            }
        } else {
            e[i] = a(i, l);
        }
        d[i] = h;
    }
    d[0] = 0.0; e[0] = 0.0; }
    if (want_vectors) for (int i = 0; i < n; i++) {
        const int l = i - 1;
        if (d[i] != 0.0) {
            for (int j = 0; j <= l; j++) {
                double g = 0.0;
                for (int k = 0; k <= l; k++) g += a(i, k) * a(k, j);
                for (int k = 0; k <= l; k++) a(k, j) -= g * a(k, i);
            }
        }
        d[i] = a(i, i);
        a(i, i) = 1.0;
        for (int j = 0; j <= l; j++) a(j, i) = a(i, j) = 0.0;
    }
    // --- implicit QL on the tridiagonal matrix (d diagonal, e sub-diagonal); the accumulated transformation is
    //     kept transposed (rows = vectors) so that every rotation touches two contiguous rows
    if (want_vectors)
        for (int i = 0; i < n; i++) for (int j = i + 1; j < n; j++) std::swap(a(i, j), a(j, i));
    for (int i = 1; i < n; i++) e[i - 1] = e[i];
    e[n - 1] = 0.0;
    for (int l = 0; l < n; l++) {
        int iter = 0, m;
        do {
            for (m = l; m < n - 1; m++) {
                const double dd = std::fabs(d[m]) + std::fabs(d[m + 1]);
                if (std::fabs(e[m]) <= 2.3e-16 * dd) break;
            }
            if (m != l) {
                if (iter++ == 200) break;
                double g = (d[l + 1] - d[l]) / (2.0 * e[l]);
                double r = std::sqrt(g * g + 1.0);
                g = d[m] - d[l] + e[l] / (g + (g >= 0.0 ? std::fabs(r) : -std::fabs(r)));
                double sn = 1.0, cs = 1.0, p = 0.0;
                int i;
                for (i = m - 1; i >= l; i--) {
                    double f = sn * e[i];
                    const double bb = cs * e[i];
                    e[i + 1] = (r = pythag(f, g));
                    if (r == 0.0) { d[i + 1] -= p; e[m] = 0.0; break; }
                    sn = f / r; cs = g / r;
                    g = d[i + 1] - p;
                    r = (d[i] - g) * sn + 2.0 * cs * bb;
                    d[i + 1] = g + (p = sn * r);
                    g = cs * r - bb;
                    if (want_vectors) {
                        double* r1 = &A[(size_t)(i + 1) * n];
                        double* r0 = &A[(size_t)i * n];
                        for (int k = 0; k < n; k++) {
                            f = r1[k];
                            r1[k] = sn * r0[k] + cs * f;
                            r0[k] = cs * r0[k] - sn * f;
                        }
                    }
                }
                if (r == 0.0 && i >= l) continue;
                d[l] -= p; e[l] = g; e[m] = 0.0;
            }
        } while (m != l);
    }
    w = d;
    if (want_vectors) {
        V.resize((size_t)n * n);
        for (int i = 0; i < n; i++) for (int j = 0; j < n; j++) V[(size_t)j * n + i] = A[(size_t)i * n + j];
    }
}

// (G + lambda I) x = g by Cholesky; false when the matrix is not positive definite
bool cholesky_solve(const double* G, double lambda, const double* g, int n, double* x) {
    std::vector<double> L((size_t)n * n);
    for (int i = 0; i < n; i++) {
        for (int j = 0; j <= i; j++) {
            double s = 0.5 * (G[(size_t)i * n + j] + G[(size_t)j * n + i]) + (i == j ? lambda : 0.0);
            s -= dot4(&L[(size_t)i * n], &L[(size_t)j * n], j);
            if (i == j) {
                if (!(s > 0.0)) return false;
                L[(size_t)i * n + i] = std::sqrt(s);
            } else {
                L[(size_t)i * n + j] = s / L[(size_t)j * n + j];
            }
        }
    }
    for (int i = 0; i < n; i++) {
        double s = g[i];
        for (int k = 0; k < i; k++) s -= L[(size_t)i * n + k] * x[k];
        x[i] = s / L[(size_t)i * n + i];
    }
    for (int i = n - 1; i >= 0; i--) {
        double s = x[i];
        for (int k = i + 1; k < n; k++) s -= L[(size_t)k * n + i] * x[k];
        x[i] = s / L[(size_t)i * n + i];
    }
    return true;
}

}  // namespace

extern "C" int qgt_b200_natural_gradient(qgt_b200_ctx* ctx, const double* metric, const double* grad, size_t num_params,
                                         const qgt_b200_natgrad_config* cfg_in, double* out, double* lambda_used) {
    (void)ctx;
    if (!metric || !grad || !out || num_params == 0) return qgt::fail(QGT_B200_ERR_INVALID_ARG, "metric/grad/out is NULL or P == 0");
    qgt_b200_natgrad_config cfg = {1e-4, 1e8, 1, 1, 1e-10};   // get_default_natural_gradient_config, gradient.c:2721
    if (cfg_in) cfg = *cfg_in;
    const int n = (int)num_params;
    std::vector<double> A((size_t)n * n), w, V;
    auto load = [&]() {
        for (int i = 0; i < n; i++)
            for (int j = 0; j < n; j++) A[(size_t)i * n + j] = 0.5 * (metric[(size_t)i * n + j] + metric[(size_t)j * n + i]);
    };
    double lambda = cfg.regularization;
    if (cfg.adaptive) {
        load();
        sym_eigh(A, n, w, V, false);                    // eigenvalues only: singular values of a symmetric matrix = |w|
        double smax = 0.0;
        for (int i = 0; i < n; i++) smax = std::max(smax, std::fabs(w[i]));
        // Condition number over the numerically non-zero spectrum: singular values at or below
        // max(singular_cutoff, n * eps * sigma_max) - the ones the pseudo-inverse branch below (and the reference's,
        // gradient.c:2800-2885) treats as zero - do not count.  A rank-deficient metric (redundant ansatz parameters,
        // QAOA) then keeps the condition number of its range instead of kappa = 1e16 / lambda = 100, which would turn
        // the natural gradient into 0.01 * gradient.  The reference itself sets kappa = INFINITY below
        // sigma_min = 1e-15 (gradient.c:2770-2774), i.e. lambda = infinity: BASELINE.md section 4 #17; its
        // single-precision SVD never resolves sigma below ~1e-7 * sigma_max anyway.
        const double cut = std::max(cfg.singular_cutoff, (double)n * 2.220446049250313e-16 * smax);
        double smin = smax;
        for (int i = 0; i < n; i++) { const double sv = std::fabs(w[i]); if (sv > cut && sv < smin) smin = sv; }
        const double kappa = smin > 0.0 ? smax / smin : 1.0;
        if (kappa > cfg.condition_threshold) {
            const double al = 1e-6 * std::sqrt(kappa);
            if (al > lambda) lambda = al;
        }
    }
    if (lambda_used) *lambda_used = lambda;
    if (cholesky_solve(metric, lambda, grad, n, out)) return QGT_B200_OK;
    if (!cfg.pseudoinverse_fallback) return qgt::fail(-55 /* QGT_ERROR_MATRIX_SINGULAR */, "regularised metric is not positive definite");
    // SVD pseudo-inverse of G with the singular-value cutoff (gradient.c:2800-2885)
    load();
    sym_eigh(A, n, w, V, true);
    std::vector<double> coef(n, 0.0);
    for (int k = 0; k < n; k++) {
        double sdot = 0.0;
        for (int i = 0; i < n; i++) sdot += V[(size_t)i * n + k] * grad[i];
        coef[k] = std::fabs(w[k]) > cfg.singular_cutoff ? sdot / w[k] : 0.0;
    }
    for (int i = 0; i < n; i++) {
        double sdot = 0.0;
        for (int k = 0; k < n; k++) sdot += V[(size_t)i * n + k] * coef[k];
        out[i] = sdot;
    }
    return QGT_B200_OK;
}

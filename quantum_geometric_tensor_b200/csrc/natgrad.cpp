// natgrad.cpp — regularised natural-gradient step  x = (G + lambda I)^-1 g  on the real P x P metric.
//
// Semantics of compute_regularized_natural_gradient and its helpers
// (reference src/quantum_geometric/core/quantum_geometric_gradient.c:2721-2964):
//   * kappa = sigma_max / sigma_min over the positive singular values (:2763-2774), infinity when
//     sigma_min <= 1e-15;
//   * adaptive regularisation lambda = max(lambda, 1e-6 * sqrt(kappa)) when kappa > threshold (:2898-2912);
//   * Tikhonov G + lambda I, solve (:2920-2959); if the regularised matrix is singular and the fallback is
//     enabled, SVD pseudo-inverse of G with the singular-value cutoff (:2800-2885).
// P is a few hundred at most, so this stays on the host; one symmetric Jacobi
// eigen-decomposition G = V diag(w) V^T serves the condition number, the solve and the pseudo-inverse
// (for a symmetric matrix the singular values are |w|).
#include <cmath>
#include <cstring>
#include <limits>
#include <vector>

#include "ctx.hpp"

namespace {

// cyclic Jacobi; A (n x n, symmetric, row-major) is destroyed, V receives the eigenvectors as columns
void jacobi_eigh(std::vector<double>& A, int n, std::vector<double>& w, std::vector<double>& V) {
    V.assign((size_t)n * n, 0.0);
    for (int i = 0; i < n; i++) V[(size_t)i * n + i] = 1.0;
    for (int sweep = 0; sweep < 64; sweep++) {
        double off = 0.0, diag = 0.0;
        for (int p = 0; p < n; p++) {
            diag += A[(size_t)p * n + p] * A[(size_t)p * n + p];
            for (int q = p + 1; q < n; q++) off += A[(size_t)p * n + q] * A[(size_t)p * n + q];
        }
        if (off <= 1e-32 * (diag + off) || off == 0.0) break;
        for (int p = 0; p < n - 1; p++) {
            for (int q = p + 1; q < n; q++) {
                const double apq = A[(size_t)p * n + q];
                if (apq == 0.0) continue;
                const double theta = (A[(size_t)q * n + q] - A[(size_t)p * n + p]) / (2.0 * apq);
                const double t = (theta >= 0 ? 1.0 : -1.0) / (std::fabs(theta) + std::sqrt(theta * theta + 1.0));
                const double c = 1.0 / std::sqrt(t * t + 1.0), s = t * c;
                for (int k = 0; k < n; k++) {
                    const double akp = A[(size_t)k * n + p], akq = A[(size_t)k * n + q];
                    A[(size_t)k * n + p] = c * akp - s * akq;
                    A[(size_t)k * n + q] = s * akp + c * akq;
                }
                for (int k = 0; k < n; k++) {
                    const double apk = A[(size_t)p * n + k], aqk = A[(size_t)q * n + k];
                    A[(size_t)p * n + k] = c * apk - s * aqk;
                    A[(size_t)q * n + k] = s * apk + c * aqk;
                }
                for (int k = 0; k < n; k++) {
                    const double vkp = V[(size_t)k * n + p], vkq = V[(size_t)k * n + q];
                    V[(size_t)k * n + p] = c * vkp - s * vkq;
                    V[(size_t)k * n + q] = s * vkp + c * vkq;
                }
            }
        }
    }
    w.resize(n);
    for (int i = 0; i < n; i++) w[i] = A[(size_t)i * n + i];
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
    for (int i = 0; i < n; i++)
        for (int j = 0; j < n; j++) A[(size_t)i * n + j] = 0.5 * (metric[(size_t)i * n + j] + metric[(size_t)j * n + i]);
    jacobi_eigh(A, n, w, V);

    double lambda = cfg.regularization;
    if (cfg.adaptive) {
        double smax = std::fabs(w[0]), smin = std::fabs(w[0]);
        for (int i = 1; i < n; i++) {
            const double s = std::fabs(w[i]);
            if (s > smax) smax = s;
            if (s < smin && s > 0) smin = s;
        }
        const double kappa = smin > 1e-15 ? smax / smin : std::numeric_limits<double>::infinity();
        if (kappa > cfg.condition_threshold) {
            const double al = 1e-6 * std::sqrt(kappa);
            if (al > lambda) lambda = al;
        }
    }
    if (lambda_used) *lambda_used = lambda;

    // coefficients of grad in the eigenbasis
    std::vector<double> coef(n, 0.0);
    for (int k = 0; k < n; k++) {
        double s = 0.0;
        for (int i = 0; i < n; i++) s += V[(size_t)i * n + k] * grad[i];
        coef[k] = s;
    }
    bool singular = false;
    for (int k = 0; k < n; k++) if (std::fabs(w[k] + lambda) < 1e-300) singular = true;
    if (singular) {
        if (!cfg.pseudoinverse_fallback) return qgt::fail(-55 /* QGT_ERROR_MATRIX_SINGULAR */, "regularised metric is singular");
        for (int k = 0; k < n; k++) coef[k] = std::fabs(w[k]) > cfg.singular_cutoff ? coef[k] / w[k] : 0.0;
    } else {
        for (int k = 0; k < n; k++) coef[k] /= (w[k] + lambda);
    }
    for (int i = 0; i < n; i++) {
        double s = 0.0;
        for (int k = 0; k < n; k++) s += V[(size_t)i * n + k] * coef[k];
        out[i] = s;
    }
    return QGT_B200_OK;
}

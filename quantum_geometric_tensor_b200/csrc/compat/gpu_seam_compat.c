/*
 * gpu_seam_compat.c — the QGT-on-GPU seam of the reference:
 *   include/quantum_geometric/hardware/quantum_geometric_tensor_gpu.h:38-81  (GPUContext with function-pointer hooks)
 *   src/quantum_geometric/hardware/quantum_geometric_tensor_gpu.c:101-330    (validate -> alloc -> H2D -> execute hook -> D2H)
 *
 * The reference never assigns the hooks (its own test passes a zeroed GPUContext, tests/test_quantum_geometric_tensor_gpu.c:66-75)
 * and defines no arithmetic behind them.  Here qgt_b200_context_init() fills them in, and the three entry points keep the
 * reference's host-side sequence and error codes.  Data convention of this build (the reference has none):
 *
 *   state   rows x cols ComplexFloat, row-major: row 0 = psi (normalised, checked against cfg->precision), rows 1..rows-1 =
 *           the derivative columns d_mu psi, so P = rows - 1 parameters on a 2^n = cols dimensional state;
 *   output  a buffer of the same rows x cols size (as the reference allocates it), zero-filled, carrying at its start
 *             metric      P x P row-major, .real = Re Q_ab           (Fubini-Study metric)
 *             curvature   P x P row-major, .real = Im Q_ab           (Berry curvature, core convention)
 *             connection  P entries        A_a = i <psi|d_a psi>     (Berry connection)
 *           with Q_ab = <d_a psi|d_b psi> - <d_a psi|psi><psi|d_b psi>, from the same device Gram kernel as qgt_b200_gram.
 *
 * Deviations from the reference, deliberate: the norm check covers psi (row 0), not all rows x cols entries, and uses
 * max(cfg->precision, 1e-5) because float data cannot meet the default 1e-10 (the reference's check would reject every
 * float state); hooks left NULL by the caller are populated instead of being called through a NULL pointer.
 */
#include <math.h>
#include <stdlib.h>
#include <string.h>

#include "compat_common.h"

QGTConfig qgt_default_config(void) {
    QGTConfig c = {1e-10, true, true, 2, 3};          /* quantum_geometric_tensor_gpu.c:8-14 */
    return c;
}

const char* qgt_error_string(qgt_error_t e) {
    switch (e) {
    case 0: return "Success";
    case -1: return "Invalid parameter";
    case -2: return "Memory allocation failed";
    case -3: return "Dimension mismatch";
    case -4: return "Invalid state";
    case -5: return "Validation failed";
    case -6: return "Hardware failure";
    case -7: return "Not implemented";
    case -15: return "Internal error";
    default: return "Unknown error";
    }
}

static size_t seam_block_size(void) { return 256; }

enum { SEAM_METRIC, SEAM_CONNECTION, SEAM_CURVATURE };

/* device pointers in, device pointers out */
static qgt_error_t seam_execute(int what, void* d_state, void* d_out, size_t rows, size_t cols) {
    if (!d_state || !d_out || rows < 2 || cols == 0) return QGT_ERROR_INVALID_PARAMETER;
    qgt_b200_ctx* ctx = qgt_compat_ctx();
    if (!ctx) return -6;
    const size_t P = rows - 1;
    const size_t nout = what == SEAM_CONNECTION ? P : P * P;
    if (nout > rows * cols) return QGT_ERROR_DIMENSION_MISMATCH;
    const float* psi = (const float*)d_state;
    const float* cols_ptr = psi + 2 * cols;
    ComplexFloat* host = (ComplexFloat*)calloc(nout, sizeof *host);
    if (!host) return QGT_ERROR_MEMORY_ALLOCATION;
    int rc = 0;
    if (what == SEAM_CONNECTION) {
        for (size_t a = 0; a < P && !rc; a++) {
            float ip[2];
            rc = qgt_b200_c64_inner_product(ctx, psi, cols_ptr + 2 * a * cols, cols, ip, NULL);
            host[a].real = -ip[1]; host[a].imag = ip[0];          /* i * <psi|d_a psi> */
        }
    } else {
        double* m = (double*)malloc(P * P * sizeof(double));
        if (!m) { free(host); return QGT_ERROR_MEMORY_ALLOCATION; }
        rc = what == SEAM_METRIC ? qgt_b200_gram_c64(ctx, psi, cols_ptr, cols, P, m, NULL, NULL)
                                 : qgt_b200_gram_c64(ctx, psi, cols_ptr, cols, P, NULL, m, NULL);
        if (!rc) for (size_t i = 0; i < P * P; i++) host[i].real = (float)m[i];
        free(m);
    }
    if (!rc) rc = qgt_b200_memset_async(d_out, 0, rows * cols * sizeof(ComplexFloat), qgt_b200_ctx_stream(ctx));
    if (!rc) rc = qgt_b200_memcpy_async(d_out, host, nout * sizeof(ComplexFloat), qgt_b200_ctx_stream(ctx));
    if (!rc) rc = qgt_b200_stream_synchronize(qgt_b200_ctx_stream(ctx));
    free(host);
    if (rc) { qgt_compat_set_error("GPUContext execute hook", rc); return -6; }
    return QGT_SUCCESS;
}

static qgt_error_t seam_execute_metric(void* s, void* o, size_t r, size_t c) { return seam_execute(SEAM_METRIC, s, o, r, c); }
static qgt_error_t seam_execute_connection(void* s, void* o, size_t r, size_t c) { return seam_execute(SEAM_CONNECTION, s, o, r, c); }
static qgt_error_t seam_execute_curvature(void* s, void* o, size_t r, size_t c) { return seam_execute(SEAM_CURVATURE, s, o, r, c); }

qgt_error_t qgt_b200_context_init(GPUContext* ctx) {
    if (!ctx) return QGT_ERROR_INVALID_PARAMETER;
    qgt_b200_ctx* c = qgt_compat_ctx();
    ctx->is_available = c != NULL;
    ctx->malloc = gpu_malloc;
    ctx->free = qgt_gpu_free_buffer;
    ctx->memcpy_to_device = gpu_memcpy_host_to_device;
    ctx->memcpy_from_device = gpu_memcpy_device_to_host;
    ctx->get_optimal_block_size = seam_block_size;
    ctx->cuda.stream = c ? qgt_b200_ctx_stream(c) : NULL;
    ctx->cuda.module = NULL;
    ctx->cuda.execute_metric = seam_execute_metric;
    ctx->cuda.execute_connection = seam_execute_connection;
    ctx->cuda.execute_curvature = seam_execute_curvature;
    return c ? QGT_SUCCESS : -6;
}

/* the reference's host sequence (quantum_geometric_tensor_gpu.c:101-179), hooks taken from the context */
static qgt_error_t seam_run(int what, GPUContext* ctx, const ComplexFloat* state, ComplexFloat* out, size_t rows, size_t cols, const QGTConfig* cfg) {
    if (!ctx || !state || !out || !cfg) return QGT_ERROR_INVALID_PARAMETER;
    if (rows < 2 || cols == 0) return QGT_ERROR_INVALID_PARAMETER;
    double norm = 0.0;
    for (size_t i = 0; i < cols; i++) norm += (double)state[i].real * state[i].real + (double)state[i].imag * state[i].imag;
    const double tol = cfg->precision > 1e-5 ? cfg->precision : 1e-5;
    if (fabs(norm - 1.0) > tol) return QGT_ERROR_INVALID_STATE;
    if (!ctx->is_available) return -6;                               /* QGT_ERROR_HARDWARE_FAILURE */
    if (!ctx->malloc || !ctx->free || !ctx->memcpy_to_device || !ctx->memcpy_from_device || !ctx->cuda.execute_metric ||
        !ctx->cuda.execute_connection || !ctx->cuda.execute_curvature) {
        const bool avail = ctx->is_available;
        if (qgt_b200_context_init(ctx) != QGT_SUCCESS || !avail) return -6;
    }
    const size_t bytes = rows * cols * sizeof(ComplexFloat);
    void *d_state = NULL, *d_out = NULL;
    if (ctx->malloc(&d_state, bytes) != QGT_SUCCESS) return QGT_ERROR_MEMORY_ALLOCATION;
    if (ctx->malloc(&d_out, bytes) != QGT_SUCCESS) { ctx->free(d_state); return QGT_ERROR_MEMORY_ALLOCATION; }
    qgt_error_t err = ctx->memcpy_to_device(d_state, state, bytes);
    if (err == QGT_SUCCESS) {
        err = what == SEAM_METRIC ? ctx->cuda.execute_metric(d_state, d_out, rows, cols)
            : what == SEAM_CONNECTION ? ctx->cuda.execute_connection(d_state, d_out, rows, cols)
                                      : ctx->cuda.execute_curvature(d_state, d_out, rows, cols);
    }
    if (err == QGT_SUCCESS) err = ctx->memcpy_from_device(out, d_out, bytes);
    ctx->free(d_state);
    ctx->free(d_out);
    if (err == QGT_ERROR_DIMENSION_MISMATCH || err == QGT_ERROR_INVALID_PARAMETER) return err;
    return err == QGT_SUCCESS ? QGT_SUCCESS : -6;
}

qgt_error_t compute_quantum_metric_gpu(GPUContext* ctx, const ComplexFloat* state, ComplexFloat* metric, size_t rows, size_t cols, const QGTConfig* cfg) {
    return seam_run(SEAM_METRIC, ctx, state, metric, rows, cols, cfg);
}
qgt_error_t compute_quantum_connection_gpu(GPUContext* ctx, const ComplexFloat* state, ComplexFloat* connection, size_t rows, size_t cols, const QGTConfig* cfg) {
    return seam_run(SEAM_CONNECTION, ctx, state, connection, rows, cols, cfg);
}
qgt_error_t compute_quantum_curvature_gpu(GPUContext* ctx, const ComplexFloat* state, ComplexFloat* curvature, size_t rows, size_t cols, const QGTConfig* cfg) {
    return seam_run(SEAM_CURVATURE, ctx, state, curvature, rows, cols, cfg);
}

/* test_seams.c — the reference's back-end seams through the C ABI:
 *   1. ComputeBackendOps: our table registers with the REFERENCE's registry (oracle/_ref/libcompute_ref.so = the
 *      unmodified supercomputer/compute_backend.c + backends/compute_cpu.c + compute_simd.c), the engine selects it, and
 *      every quantum_* slot is compared with the reference's CPU backend on the same inputs (host and device buffers);
 *   2. the QGT-on-GPU seam (GPUContext hooks + compute_quantum_{metric,connection,curvature}_gpu) against a plain
 *      triple loop;
 *   3. gate objects + the network API in the order the reference's tests/test_quantum_geometric_minimal.c uses them;
 *   4. parameter shifts / derivative columns (core/quantum_parameter_shift.h).
 * Usage: test_seams [--host-only] */
#include "quantum_geometric/supercomputer/compute_backend.h"
#include "quantum_geometric/hardware/quantum_geometric_tensor_gpu.h"
#include "quantum_geometric/core/quantum_geometric_gpu.h"
#include "quantum_geometric/core/quantum_geometric_tensor_network.h"
#include "quantum_geometric/core/quantum_gate_operations.h"
#include "quantum_geometric/core/quantum_parameter_shift.h"
#include "quantum_geometric/core/numerical_backend.h"
#include "quantum_geometric/hardware/quantum_simulator.h"
#include <complex.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#define CHECK(c) do { if (!(c)) { fprintf(stderr, "CHECK failed %s:%d: %s\n", __FILE__, __LINE__, #c); exit(1); } } while (0)

static unsigned long long rng_state = 0x9E3779B97F4A7C15ull;
static float frand(void) {
    rng_state = rng_state * 6364136223846793005ull + 1442695040888963407ull;
    return (float)((rng_state >> 40) / (double)(1ull << 24)) - 0.5f;
}
static void fill(float* v, size_t n) { for (size_t i = 0; i < n; i++) v[i] = frand(); }
static double max_diff(const float* a, const float* b, size_t n) {
    double m = 0;
    for (size_t i = 0; i < n; i++) { double d = fabs((double)a[i] - b[i]); if (d > m) m = d; }
    return m;
}
static double max_abs(const float* a, size_t n) {
    double m = 0;
    for (size_t i = 0; i < n; i++) if (fabs(a[i]) > m) m = fabs(a[i]);
    return m;
}

/* ---- 1. ComputeBackendOps ------------------------------------------------------------------------------------------- */
static void test_registration(int have_gpu) {
    const ComputeBackendInfo* mine = compute_get_backend_by_type(COMPUTE_BACKEND_CUDA);
    CHECK(mine && strcmp(mine->name, "b200") == 0 && mine->priority == 100);       /* the constructor registered the table */
    CHECK(mine->ops == qgt_b200_compute_backend_ops());
    CHECK(qgt_b200_register_compute_backend() == COMPUTE_SUCCESS);                  /* idempotent: the registry updates in place */
    int cuda_entries = 0;
    for (int i = 0; i < compute_get_backend_count(); i++) if (compute_get_backend_info(i)->type == COMPUTE_BACKEND_CUDA) cuda_entries++;
    CHECK(cuda_entries == 1);
    CHECK(compute_get_backend_by_type(COMPUTE_BACKEND_CPU) != NULL);                 /* the reference's own CPU backend is there too */
    CHECK(compute_backend_available(COMPUTE_BACKEND_CUDA) == (have_gpu != 0));
    ComputeDistributedConfig cfg;
    memset(&cfg, 0, sizeof cfg);
    cfg.num_nodes = 1; cfg.devices_per_node = 1; cfg.preferred_backend = COMPUTE_BACKEND_CUDA; cfg.allow_fallback = false;
    ComputeEngine* e = compute_engine_init(&cfg);
    if (!have_gpu) {
        CHECK(e == NULL);                              /* no device: no engine, never a host fallback of ours */
        cfg.allow_fallback = true;
        e = compute_engine_init(&cfg);
        CHECK(e && compute_engine_get_backend_type(e) == COMPUTE_BACKEND_CPU);
        compute_engine_cleanup(e);
        return;
    }
    CHECK(e && compute_engine_get_backend_type(e) == COMPUTE_BACKEND_CUDA);
    CHECK(compute_engine_get_ops(e) == mine->ops);
    cfg.preferred_backend = COMPUTE_BACKEND_AUTO; cfg.allow_fallback = true;         /* priority 100 beats the CPU backend */
    ComputeEngine* e2 = compute_engine_init(&cfg);
    CHECK(e2 && compute_engine_get_backend_type(e2) == COMPUTE_BACKEND_CUDA);
    compute_engine_cleanup(e2);
    compute_engine_cleanup(e);
}

static void test_backend_ops(void) {
    ComputeDistributedConfig cfg;
    memset(&cfg, 0, sizeof cfg);
    cfg.num_nodes = 1; cfg.devices_per_node = 1; cfg.preferred_backend = COMPUTE_BACKEND_CUDA;
    ComputeEngine* eng = compute_engine_init(&cfg);
    CHECK(eng);
    const ComputeBackendOps* g = compute_engine_get_ops(eng);
    ComputeBackend* gb = compute_engine_get_backend(eng);
    cfg.preferred_backend = COMPUTE_BACKEND_CPU;
    ComputeEngine* ceng = compute_engine_init(&cfg);
    CHECK(ceng && compute_engine_get_backend_type(ceng) == COMPUTE_BACKEND_CPU);
    const ComputeBackendOps* r = compute_engine_get_ops(ceng);
    ComputeBackend* rb = compute_engine_get_backend(ceng);

    int ndev = 0; size_t mem = 0;
    CHECK(g->get_capabilities(gb, &ndev, &mem) == COMPUTE_SUCCESS && ndev >= 1 && mem > (size_t)1 << 30);

    /* argument checks as the reference's */
    float dummy[8] = {0};
    CHECK(g->quantum_normalize(gb, NULL, 4, NULL) == COMPUTE_ERROR_INVALID_ARGUMENT);
    CHECK(g->quantum_normalize(gb, dummy, 0, NULL) == COMPUTE_ERROR_INVALID_ARGUMENT);
    CHECK(g->quantum_unitary(gb, dummy, 0, dummy, 2, NULL) == COMPUTE_ERROR_INVALID_ARGUMENT);
    CHECK(g->quantum_inner_product(gb, dummy, dummy, NULL, 4, NULL) == COMPUTE_ERROR_INVALID_ARGUMENT);
    CHECK(g->execute(gb, NULL, NULL, NULL) == COMPUTE_ERROR_INVALID_ARGUMENT);

    for (int pass = 0; pass < 3; pass++) {
        const size_t n = pass == 0 ? 64 : pass == 1 ? 1000 : ((size_t)1 << 16);      /* 1000: not a power of two, ragged tail */
        float *a = malloc(2 * n * sizeof(float)), *b = malloc(2 * n * sizeof(float)), *obs = malloc(n * sizeof(float));
        float *x = malloc(2 * n * sizeof(float)), *y = malloc(2 * n * sizeof(float));
        CHECK(a && b && obs && x && y);
        fill(a, 2 * n); fill(b, 2 * n); fill(obs, n);
        const double scale = sqrt((double)n);

        /* normalize (host buffers) */
        memcpy(x, a, 2 * n * sizeof(float)); memcpy(y, a, 2 * n * sizeof(float));
        CHECK(g->quantum_normalize(gb, x, n, NULL) == COMPUTE_SUCCESS);
        CHECK(r->quantum_normalize(rb, y, n, NULL) == COMPUTE_SUCCESS);
        CHECK(max_diff(x, y, 2 * n) < 1e-5 * max_abs(y, 2 * n) + 1e-7);

        /* inner product, gradient (= <backward|forward>), diagonal expectation */
        float gi[2], ri[2];
        CHECK(g->quantum_inner_product(gb, gi, a, b, n, NULL) == COMPUTE_SUCCESS);
        CHECK(r->quantum_inner_product(rb, ri, a, b, n, NULL) == COMPUTE_SUCCESS);
        CHECK(fabs(gi[0] - ri[0]) < 2e-5 * scale && fabs(gi[1] - ri[1]) < 2e-5 * scale);
        CHECK(g->quantum_gradient(gb, gi, a, b, n, NULL) == COMPUTE_SUCCESS);
        CHECK(r->quantum_gradient(rb, ri, a, b, n, NULL) == COMPUTE_SUCCESS);
        CHECK(fabs(gi[0] - ri[0]) < 2e-5 * scale && fabs(gi[1] - ri[1]) < 2e-5 * scale);
        float ge = 0, re = 0;
        CHECK(g->quantum_expectation(gb, &ge, a, obs, n, NULL) == COMPUTE_SUCCESS);
        CHECK(r->quantum_expectation(rb, &re, a, obs, n, NULL) == COMPUTE_SUCCESS);
        CHECK(fabs(ge - re) < 2e-5 * scale);

        /* the same through device buffers from alloc(): in place, stream ordered, read back with memcpy */
        float* da = g->alloc(gb, 2 * n * sizeof(float), COMPUTE_MEM_DEVICE);
        float* db = g->alloc(gb, 2 * n * sizeof(float), COMPUTE_MEM_DEVICE);
        float* dres = g->alloc(gb, 2 * sizeof(float), COMPUTE_MEM_DEVICE);
        CHECK(da && db && dres);
        ComputeStream* st = g->create_stream(gb);
        CHECK(st);
        CHECK(g->memcpy(gb, da, COMPUTE_MEM_DEVICE, a, COMPUTE_MEM_HOST, 2 * n * sizeof(float), st) == COMPUTE_SUCCESS);
        CHECK(g->memcpy(gb, db, COMPUTE_MEM_DEVICE, b, COMPUTE_MEM_HOST, 2 * n * sizeof(float), st) == COMPUTE_SUCCESS);
        CHECK(g->quantum_inner_product(gb, dres, da, db, n, st) == COMPUTE_SUCCESS);
        CHECK(g->quantum_normalize(gb, da, n, st) == COMPUTE_SUCCESS);
        float hres[2];
        CHECK(g->memcpy(gb, hres, COMPUTE_MEM_HOST, dres, COMPUTE_MEM_DEVICE, sizeof hres, st) == COMPUTE_SUCCESS);
        CHECK(g->memcpy(gb, x, COMPUTE_MEM_HOST, da, COMPUTE_MEM_DEVICE, 2 * n * sizeof(float), st) == COMPUTE_SUCCESS);
        ComputeEvent* ev = g->create_event(gb);
        CHECK(ev && g->record_event(gb, ev, st) == COMPUTE_SUCCESS && g->wait_event(gb, NULL, ev) == COMPUTE_SUCCESS);
        CHECK(g->synchronize_stream(gb, st) == COMPUTE_SUCCESS);
        CHECK(r->quantum_inner_product(rb, ri, a, b, n, NULL) == COMPUTE_SUCCESS);
        CHECK(fabs(hres[0] - ri[0]) < 2e-5 * scale && fabs(hres[1] - ri[1]) < 2e-5 * scale);
        CHECK(max_diff(x, y, 2 * n) < 1e-5 * max_abs(y, 2 * n) + 1e-7);
        CHECK(g->memset(gb, da, 0, 2 * n * sizeof(float), st) == COMPUTE_SUCCESS);
        CHECK(g->memcpy(gb, x, COMPUTE_MEM_HOST, da, COMPUTE_MEM_DEVICE, 2 * n * sizeof(float), NULL) == COMPUTE_SUCCESS);
        CHECK(max_abs(x, 2 * n) == 0.0);
        g->destroy_event(gb, ev);
        g->destroy_stream(gb, st);
        g->free(gb, da, COMPUTE_MEM_DEVICE); g->free(gb, db, COMPUTE_MEM_DEVICE); g->free(gb, dres, COMPUTE_MEM_DEVICE);
        free(a); free(b); free(obs); free(x); free(y);
    }

    /* dense unitary (the reference's meaning of quantum_unitary): 128 x 128 */
    {
        const size_t n = 128;
        float *u = malloc(2 * n * n * sizeof(float)), *x = malloc(2 * n * sizeof(float)), *y = malloc(2 * n * sizeof(float));
        fill(u, 2 * n * n); fill(x, 2 * n); memcpy(y, x, 2 * n * sizeof(float));
        CHECK(g->quantum_unitary(gb, x, n, u, n, NULL) == COMPUTE_SUCCESS);
        CHECK(r->quantum_unitary(rb, y, n, u, n, NULL) == COMPUTE_SUCCESS);
        CHECK(max_diff(x, y, 2 * n) < 2e-5 * max_abs(y, 2 * n));
        /* through execute() */
        ComputeQuantumOp op; memset(&op, 0, sizeof op);
        op.type = QUANTUM_OP_UNITARY; op.output_data = x; op.output_size = n; op.parameters = u; op.param_size = n;
        memcpy(x, y, 2 * n * sizeof(float));
        ComputeExecutionPlan* plan = g->create_plan(gb, &op);
        CHECK(plan && plan->num_partitions == 1);
        CHECK(g->execute(gb, &op, plan, NULL) == COMPUTE_SUCCESS);
        g->destroy_plan(gb, plan);
        op.output_data = y;
        CHECK(r->execute(rb, &op, NULL, NULL) == COMPUTE_SUCCESS);
        CHECK(max_diff(x, y, 2 * n) < 2e-5 * max_abs(y, 2 * n));
        free(u); free(x); free(y);
    }

    /* gate-level execute: a 2-qubit matrix on qubits (5, 2) of a 10-qubit register vs the dense expansion on the reference backend */
    {
        const size_t nq = 10, n = (size_t)1 << nq;
        const size_t tq[2] = {5, 2};
        float gate[32], *x = malloc(2 * n * sizeof(float)), *y = malloc(2 * n * sizeof(float)), *u = calloc(2 * n * n, sizeof(float));
        fill(gate, 32); fill(x, 2 * n); memcpy(y, x, 2 * n * sizeof(float));
        for (size_t row = 0; row < n; row++) {
            const size_t ra = ((row >> tq[0]) & 1) | (((row >> tq[1]) & 1) << 1);
            const size_t base = row & ~(((size_t)1 << tq[0]) | ((size_t)1 << tq[1]));
            for (size_t ca = 0; ca < 4; ca++) {
                const size_t col = base | ((ca & 1) << tq[0]) | (((ca >> 1) & 1) << tq[1]);
                u[2 * (row * n + col)] = gate[2 * (ra * 4 + ca)];
                u[2 * (row * n + col) + 1] = gate[2 * (ra * 4 + ca) + 1];
            }
        }
        ComputeQuantumOp op; memset(&op, 0, sizeof op);
        op.type = QUANTUM_OP_UNITARY; op.output_data = x; op.output_size = n; op.parameters = gate; op.param_size = 4;
        op.num_qubits = nq; op.target_qubits = (size_t*)tq; op.num_targets = 2;
        CHECK(g->execute(gb, &op, NULL, NULL) == COMPUTE_SUCCESS);
        CHECK(r->quantum_unitary(rb, y, n, u, n, NULL) == COMPUTE_SUCCESS);
        CHECK(max_diff(x, y, 2 * n) < 1e-5 * max_abs(y, 2 * n));
        /* 1-, 3- and 4-qubit gates against the same kind of expansion, built with our dense product (checked above) */
        for (int k = 1; k <= 4; k++) {
            if (k == 2) continue;
            const size_t tk[4] = {7, 0, 9, 3};
            const size_t gd = (size_t)1 << k;
            float* gm = malloc(2 * gd * gd * sizeof(float));
            fill(gm, 2 * gd * gd); fill(x, 2 * n); memcpy(y, x, 2 * n * sizeof(float));
            memset(u, 0, 2 * n * n * sizeof(float));
            size_t mask = 0;
            for (int j = 0; j < k; j++) mask |= (size_t)1 << tk[j];
            for (size_t row = 0; row < n; row++) {
                size_t ra = 0;
                for (int j = 0; j < k; j++) ra |= ((row >> tk[j]) & 1) << j;
                for (size_t ca = 0; ca < gd; ca++) {
                    size_t col = row & ~mask;
                    for (int j = 0; j < k; j++) col |= ((ca >> j) & 1) << tk[j];
                    u[2 * (row * n + col)] = gm[2 * (ra * gd + ca)];
                    u[2 * (row * n + col) + 1] = gm[2 * (ra * gd + ca) + 1];
                }
            }
            op.parameters = gm; op.param_size = gd; op.target_qubits = (size_t*)tk; op.num_targets = (size_t)k; op.output_data = x;
            CHECK(g->execute(gb, &op, NULL, NULL) == COMPUTE_SUCCESS);
            CHECK(r->quantum_unitary(rb, y, n, u, n, NULL) == COMPUTE_SUCCESS);
            CHECK(max_diff(x, y, 2 * n) < 2e-5 * max_abs(y, 2 * n));
            free(gm);
        }
        op.num_targets = 5;
        CHECK(g->execute(gb, &op, NULL, NULL) == COMPUTE_ERROR_NOT_IMPLEMENTED);
        free(x); free(y); free(u);
    }

    /* tensor contraction C[m x k] = A[m x n] B[n x k], ragged sizes */
    {
        const size_t m = 37, n = 70, k = 45;
        float *A = malloc(2 * m * n * sizeof(float)), *B = malloc(2 * n * k * sizeof(float));
        float *C = malloc(2 * m * k * sizeof(float)), *Cr = malloc(2 * m * k * sizeof(float));
        fill(A, 2 * m * n); fill(B, 2 * n * k);
        CHECK(g->quantum_tensor_contract(gb, C, A, B, m, n, k, NULL) == COMPUTE_SUCCESS);
        CHECK(r->quantum_tensor_contract(rb, Cr, A, B, m, n, k, NULL) == COMPUTE_SUCCESS);
        CHECK(max_diff(C, Cr, 2 * m * k) < 1e-5 * max_abs(Cr, 2 * m * k));
        size_t dims[3] = {m, n, k};
        ComputeQuantumOp op; memset(&op, 0, sizeof op);
        op.type = QUANTUM_OP_TENSOR_CONTRACT; op.output_data = C; op.input_data = A; op.parameters = B; op.dims = dims; op.num_dims = 3;
        memset(C, 0, 2 * m * k * sizeof(float));
        CHECK(g->execute(gb, &op, NULL, NULL) == COMPUTE_SUCCESS);
        CHECK(max_diff(C, Cr, 2 * m * k) < 1e-5 * max_abs(Cr, 2 * m * k));
        free(A); free(B); free(C); free(Cr);
    }

    /* collectives with one rank are the reference's single-node copies */
    {
        double s[4] = {1, 2, 3, 4}, d[4] = {0};
        CHECK(g->allreduce(gb, s, d, 4, COMPUTE_DTYPE_FLOAT64, COMPUTE_REDUCE_SUM) == COMPUTE_SUCCESS && d[3] == 4.0);
        memset(d, 0, sizeof d);
        CHECK(g->allgather(gb, s, d, 4, COMPUTE_DTYPE_FLOAT64) == COMPUTE_SUCCESS && d[2] == 3.0);
        CHECK(g->broadcast(gb, s, 4, COMPUTE_DTYPE_FLOAT64, 0) == COMPUTE_SUCCESS && s[1] == 2.0);
        CHECK(g->barrier(gb) == COMPUTE_SUCCESS);
    }
    ComputeMetrics mt;
    CHECK(g->get_metrics(gb, &mt) == COMPUTE_SUCCESS && mt.operations_per_second >= 3.0 && mt.peak_memory_bytes > 0);
    CHECK(g->reset_metrics(gb) == COMPUTE_SUCCESS);
    compute_engine_cleanup(ceng);
    compute_engine_cleanup(eng);
    printf("ComputeBackendOps: every slot agrees with the reference's CPU backend\n");
}

/* ---- 2. QGT-on-GPU seam ---------------------------------------------------------------------------------------------- */
static void test_gpu_seam(void) {
    const size_t P = 5, D = 512, rows = P + 1, cols = D;
    ComplexFloat* buf = malloc(rows * cols * sizeof *buf);
    ComplexFloat* out = malloc(rows * cols * sizeof *out);
    fill((float*)buf, 2 * rows * cols);
    double nrm = 0;
    for (size_t i = 0; i < D; i++) nrm += (double)buf[i].real * buf[i].real + (double)buf[i].imag * buf[i].imag;
    for (size_t i = 0; i < D; i++) { buf[i].real = (float)(buf[i].real / sqrt(nrm)); buf[i].imag = (float)(buf[i].imag / sqrt(nrm)); }
    /* reference triple loop (differential_geometry.c:2832-2856 in double) */
    double qre[25], qim[25], are[5], aim[5];
    for (size_t a = 0; a < P; a++) {
        double vr = 0, vi = 0;          /* <psi|d_a psi> */
        for (size_t i = 0; i < D; i++) {
            const ComplexFloat p = buf[i], d = buf[(a + 1) * D + i];
            vr += (double)p.real * d.real + (double)p.imag * d.imag;
            vi += (double)p.real * d.imag - (double)p.imag * d.real;
        }
        are[a] = -vi; aim[a] = vr;      /* i * v */
    }
    for (size_t a = 0; a < P; a++) for (size_t b = 0; b < P; b++) {
        double sr = 0, si = 0;
        for (size_t i = 0; i < D; i++) {
            const ComplexFloat x = buf[(a + 1) * D + i], y = buf[(b + 1) * D + i];
            sr += (double)x.real * y.real + (double)x.imag * y.imag;
            si += (double)x.real * y.imag - (double)x.imag * y.real;
        }
        /* <d_a|psi> = conj(v_a), <psi|d_b> = v_b, with v = -i A */
        const double var = aim[a], vai = -are[a], vbr = aim[b], vbi = -are[b];
        qre[a * P + b] = sr - (var * vbr + vai * vbi);
        qim[a * P + b] = si - (var * vbi - vai * vbr);
    }
    CHECK(qg_gpu_init() == QG_GPU_SUCCESS);
    GPUContext ctx;
    memset(&ctx, 0, sizeof ctx);
    ctx.is_available = true;                                        /* a zeroed context, as the reference's test builds it */
    QGTConfig cfg = qgt_default_config();
    CHECK(cfg.precision == 1e-10 && cfg.error_correction == 2 && cfg.optimization_level == 3);
    CHECK(compute_quantum_metric_gpu(NULL, buf, out, rows, cols, &cfg) == QGT_ERROR_INVALID_PARAMETER);
    CHECK(compute_quantum_metric_gpu(&ctx, buf, out, rows, cols, NULL) == QGT_ERROR_INVALID_PARAMETER);
    CHECK(compute_quantum_metric_gpu(&ctx, buf, out, rows, cols, &cfg) == QGT_SUCCESS);
    CHECK(ctx.malloc && ctx.cuda.execute_metric);                   /* the hooks were filled in */
    for (size_t i = 0; i < P * P; i++) CHECK(fabs(out[i].real - qre[i]) < 1e-5 && out[i].imag == 0.0f);
    for (size_t i = P * P; i < rows * cols; i++) CHECK(out[i].real == 0.0f && out[i].imag == 0.0f);
    CHECK(compute_quantum_curvature_gpu(&ctx, buf, out, rows, cols, &cfg) == QGT_SUCCESS);
    for (size_t i = 0; i < P * P; i++) CHECK(fabs(out[i].real - qim[i]) < 1e-5);
    CHECK(compute_quantum_connection_gpu(&ctx, buf, out, rows, cols, &cfg) == QGT_SUCCESS);
    for (size_t a = 0; a < P; a++) CHECK(fabs(out[a].real - are[a]) < 1e-5 && fabs(out[a].imag - aim[a]) < 1e-5);
    /* an explicit context, hooks driven directly with device buffers */
    GPUContext c2;
    CHECK(qgt_b200_context_init(&c2) == QGT_SUCCESS && c2.is_available && c2.get_optimal_block_size() == 256);
    void *ds = NULL, *dout = NULL;
    CHECK(c2.malloc(&ds, rows * cols * sizeof *buf) == QGT_SUCCESS && c2.malloc(&dout, rows * cols * sizeof *buf) == QGT_SUCCESS);
    CHECK(c2.memcpy_to_device(ds, buf, rows * cols * sizeof *buf) == QGT_SUCCESS);
    CHECK(c2.cuda.execute_metric(ds, dout, rows, cols) == QGT_SUCCESS);
    CHECK(c2.memcpy_from_device(out, dout, rows * cols * sizeof *buf) == QGT_SUCCESS);
    for (size_t i = 0; i < P * P; i++) CHECK(fabs(out[i].real - qre[i]) < 1e-5);
    c2.free(ds); c2.free(dout);
    /* an un-normalised psi is rejected like the reference does */
    buf[0].real += 0.5f;
    CHECK(compute_quantum_metric_gpu(&ctx, buf, out, rows, cols, &cfg) == QGT_ERROR_INVALID_STATE);
    ctx.is_available = false; buf[0].real -= 0.5f;
    CHECK(compute_quantum_metric_gpu(&ctx, buf, out, rows, cols, &cfg) == QGT_ERROR_HARDWARE_FAILURE);
    qg_gpu_cleanup();
    free(buf); free(out);
    printf("QGT-on-GPU seam: metric / curvature / connection agree with the triple loop\n");
}

/* ---- 3. gate objects + network API (host parts run without a GPU) ---------------------------------------------------------- */
static void test_gate_objects(void) {
    size_t q0 = 0;
    double half = 0.5;
    quantum_gate_t* rx = create_quantum_gate(GATE_TYPE_RX, &q0, 1, &half, 1);
    CHECK(rx && rx->is_parameterized && rx->num_parameters == 1 && rx->parameters[0] == 0.5 && rx->target_qubits[0] == 0);
    CHECK(fabs(rx->matrix[0].real - cos(0.25)) < 1e-7 && fabs(rx->matrix[1].imag + sin(0.25)) < 1e-7);   /* [c, -is; -is, c] */
    CHECK(create_quantum_gate(GATE_TYPE_RX, &q0, 1, NULL, 0) == NULL);           /* a rotation needs its angle */
    CHECK(create_quantum_gate(GATE_TYPE_CNOT, &q0, 1, NULL, 0) == NULL);         /* two-qubit kinds need two qubits */
    CHECK(create_quantum_gate(GATE_TYPE_H, NULL, 1, NULL, 0) == NULL);
    double one = 1.0;
    CHECK(update_gate_parameters(rx, &one, 1) && fabs(rx->matrix[0].real - cos(0.5)) < 1e-7);
    CHECK(shift_gate_parameters(rx, 0, -0.5) && rx->parameters[0] == 0.5);
    quantum_gate_t* cp = copy_quantum_gate(rx);
    CHECK(cp && cp->parameters != rx->parameters && cp->parameters[0] == 0.5 && cp->matrix[0].real == rx->matrix[0].real);
    quantum_gate_t* cn = create_cnot_gate(1, 0);
    CHECK(cn && cn->is_controlled && cn->num_qubits == 2 && cn->target_qubits[0] == 1 && cn->target_qubits[1] == 0);
    CHECK(cn->matrix[0].real == 1.0f && cn->matrix[2 * 4 + 3].real == 1.0f && cn->matrix[3 * 4 + 2].real == 1.0f && cn->matrix[2 * 4 + 2].real == 0.0f);
    quantum_gate_t* h = create_h_gate(0);
    CHECK(h && !h->is_parameterized && fabs(h->matrix[3].real + 0.70710678f) < 1e-6);
    destroy_quantum_gate(h); destroy_quantum_gate(cn); destroy_quantum_gate(cp); destroy_quantum_gate(rx);
    destroy_quantum_gate(NULL);
    numerical_config_t ncfg = {.type = NUMERICAL_BACKEND_CPU, .max_threads = 1};
    CHECK(initialize_numerical_backend(&ncfg) == NUMERICAL_SUCCESS && initialize_numerical_backend(NULL) == NUMERICAL_ERROR_INVALID_ARGUMENT);
    CHECK(strcmp(get_numerical_error_string(NUMERICAL_SUCCESS), "Success") == 0);
    shutdown_numerical_backend();
}

static void test_minimal_network(void) {
    /* the sequence of the reference's tests/test_quantum_geometric_minimal.c, with the return values read as the bools they are */
    quantum_geometric_tensor_network_t* n = create_quantum_geometric_tensor_network(2, 1, false, false);
    CHECK(n);
    size_t qubits[] = {0};
    double params[] = {0.5};
    quantum_gate_t* gate = create_quantum_gate(GATE_TYPE_RX, qubits, 1, params, 1);
    CHECK(gate && apply_quantum_gate(n, gate, qubits, 1));
    ComplexFloat q;
    CHECK(compute_quantum_geometric_tensor(n, 0, 0, &q));
    CHECK(fabs(q.real - 0.25) < 1e-6 && fabs(q.imag) < 1e-6);                    /* one rotation: Q = Var(X/2) = 1/4 */
    /* a gate built with two qubits and applied without an explicit qubit list keeps control and target apart */
    quantum_gate_t* ry = create_ry_gate(1, 0.9);
    quantum_gate_t* cn = create_cnot_gate(0, 1);
    CHECK(apply_quantum_gate(n, cn, NULL, 0) && apply_quantum_gate(n, ry, NULL, 0));
    ComplexFloat* psi = NULL; size_t dim = 0;
    CHECK(get_quantum_state(n, &psi, &dim) && dim == 4);
    /* RX(0.5) on q0, CNOT 0->1, RY(0.9) on q1, by hand */
    const double c = cos(0.25), s = sin(0.25), cy = cos(0.45), sy = sin(0.45);
    /* after RX: c|00> - i s|01> (bit 0 = qubit 0); after CNOT(control 0, target 1): c|00> - i s|11>; RY on qubit 1 */
    const double ere[4] = {c * cy, 0, c * sy, 0}, eim[4] = {0, s * sy, 0, -s * cy};
    for (int i = 0; i < 4; i++) CHECK(fabs(psi[i].real - ere[i]) < 1e-6 && fabs(psi[i].imag - eim[i]) < 1e-6);
    free(psi);

    /* 4. parameter shifts: exact column vs centred difference, shifted states, shift_parameter */
    ComplexFloat *col = NULL, *fd = NULL, *fw = NULL, *bw = NULL;
    size_t d1 = 0, d2 = 0, d3 = 0;
    for (size_t mu = 0; mu < 2; mu++) {
        CHECK(compute_higher_order_gradient(n, mu, NULL, 0, &col, &d1) && d1 == 4);
        CHECK(compute_centered_difference_gradient(n, mu, 1e-3, &fd, &d2) && d2 == 4);
        for (int i = 0; i < 4; i++) CHECK(fabs(col[i].real - fd[i].real) < 2e-4 && fabs(col[i].imag - fd[i].imag) < 2e-4);
        free(col); free(fd);
    }
    CHECK(!compute_higher_order_gradient(n, 2, NULL, 0, &col, &d1));             /* only two parameters */
    CHECK(compute_shifted_states(n, 0, M_PI / 2, &fw, &bw, &d3) && d3 == 4);
    /* parameter-shift rule for the state: psi(t+s) - psi(t-s) = 4 sin(s/2) d psi, s = pi/2 */
    CHECK(compute_higher_order_gradient(n, 0, NULL, 0, &col, &d1));
    for (int i = 0; i < 4; i++) {
        CHECK(fabs((fw[i].real - bw[i].real) - 4 * sin(M_PI / 4) * col[i].real) < 1e-5);
        CHECK(fabs((fw[i].imag - bw[i].imag) - 4 * sin(M_PI / 4) * col[i].imag) < 1e-5);
    }
    free(fw); free(bw); free(col);
    CHECK(compute_parameter_shift_gradient(n, 1, 0.01, &fd, &d2)); free(fd);
    double err = -1;
    CHECK(compute_gradient_with_error(n, 1, &fd, &err, &d2) && err >= 0 && err < 1e-3); free(fd);
    CHECK(shift_parameter(n, 0, 0.25) && !shift_parameter(n, 7, 0.1));
    CHECK(compute_quantum_geometric_tensor(n, 0, 0, &q) && fabs(q.real - 0.25) < 1e-6);      /* cache was invalidated and rebuilt */
    destroy_quantum_gate(gate); destroy_quantum_gate(ry); destroy_quantum_gate(cn);
    destroy_quantum_geometric_tensor_network(n);
    printf("network API + parameter shifts: ok\n");
}

/* ---- 5. depolarizing-noise trajectories of sim_execute_circuit (quantum_simulator.c:290-311, 525-529) ------------------ */
static unsigned long long lcg_state;
static double lcg(void) { lcg_state = lcg_state * 1103515245ull + 12345ull; return (double)(lcg_state & 0x7fffffff) / (double)0x7fffffff; }

static void test_noise_trajectory(void) {
    const uint32_t n = 5;
    const int ng = 40;
    gate_type_t kinds[40]; uint32_t tg[40], ct[40]; double ang[40];
    unsigned long long r = 777;
    for (int i = 0; i < ng; i++) {
        r = r * 6364136223846793005ull + 1442695040888963407ull;
        const int k = (int)((r >> 33) % 5);
        kinds[i] = k == 0 ? GATE_TYPE_H : k == 1 ? GATE_TYPE_RX : k == 2 ? GATE_TYPE_RZ : k == 3 ? GATE_TYPE_CNOT : GATE_TYPE_RY;
        tg[i] = (uint32_t)((r >> 40) % n); ct[i] = (tg[i] + 1 + (uint32_t)((r >> 50) % (n - 1))) % n;
        ang[i] = (double)((r >> 20) % 1000) / 200.0 - 2.5;
    }
    double model[3] = {0.6, 0.0, 0.0};
    struct SimulatorConfig cfg;
    memset(&cfg, 0, sizeof cfg);
    cfg.noise_model = model;
    SimulatorState* noisy = sim_init(n, 0, &cfg);
    SimulatorState* clean = sim_init(n, 0, NULL);
    CHECK(noisy && clean && noisy->active_noise.type == NOISE_DEPOLARIZING && noisy->active_noise.gate_error_rate == 0.6);
    CHECK(clean->active_noise.type == NOISE_NONE);
    SimulatorCircuit* c = sim_create_circuit(n, 0);
    SimulatorCircuit* replay = sim_create_circuit(n, 0);
    lcg_state = 4242;
    int injected = 0;
    for (int i = 0; i < ng; i++) {
        CHECK(sim_add_gate(c, kinds[i], tg[i], ct[i], &ang[i]));
        CHECK(sim_add_gate(replay, kinds[i], tg[i], ct[i], &ang[i]));
        if (lcg() < 0.6) {                       /* the same two draws per event, in the same order */
            const double pick = lcg();
            if (pick < 0.75) { CHECK(sim_add_gate(replay, pick < 0.25 ? GATE_TYPE_X : pick < 0.5 ? GATE_TYPE_Y : GATE_TYPE_Z, tg[i], 0, NULL)); injected++; }
        }
    }
    CHECK(injected > 5);
    qgt_compat_seed(4242);
    CHECK(sim_execute_circuit(noisy, c));
    CHECK(sim_execute_circuit(clean, replay));
    double nrm = 0, diff = 0;
    for (size_t i = 0; i < ((size_t)1 << n); i++) {
        nrm += creal(noisy->amplitudes[i] * conj(noisy->amplitudes[i]));
        const double d = cabs(noisy->amplitudes[i] - clean->amplitudes[i]);
        if (d > diff) diff = d;
    }
    CHECK(fabs(nrm - 1.0) < 1e-12 && diff < 1e-14);            /* Pauli errors keep the state normalised; trajectory = replay */
    /* and the noise did something: the noiseless circuit ends elsewhere */
    sim_reset_state(clean);
    CHECK(sim_execute_circuit(clean, c));
    diff = 0;
    for (size_t i = 0; i < ((size_t)1 << n); i++) { const double d = cabs(noisy->amplitudes[i] - clean->amplitudes[i]); if (d > diff) diff = d; }
    CHECK(diff > 1e-3);
    sim_cleanup_circuit(c); sim_cleanup_circuit(replay); sim_cleanup(noisy); sim_cleanup(clean);
    printf("depolarizing-noise trajectory: equals its replay\n");
}

int main(int argc, char** argv) {
    const int host_only = argc > 1 && strcmp(argv[1], "--host-only") == 0;
    test_gate_objects();
    test_registration(!host_only);
    if (host_only) { printf("all host-only seam checks passed\n"); return 0; }
    test_backend_ops();
    test_gpu_seam();
    test_minimal_network();
    test_noise_trajectory();
    printf("all seam checks passed\n");
    return 0;
}

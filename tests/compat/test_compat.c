/* test_compat.c — exercises the reference's C entry points (include/quantum_geometric/...) against the CUDA
 * library, in the style of the reference's own tests (standalone main, assert): the known answers of
 * tests/test_quantum_simulator_cpu.c (magnitudes to 1e-10), plus checks the reference tests never make
 * (1-qubit gates on a superposed register with target >= 1, the QGT of a circuit).
 * Usage: test_compat [--host-only]   (--host-only skips everything that needs a GPU) */
#include "quantum_geometric/hardware/quantum_simulator_cpu.h"
#include "quantum_geometric/hardware/quantum_simulator.h"
#include "quantum_geometric/core/quantum_geometric_tensor_network.h"
#include "quantum_geometric/core/quantum_geometric_metric.h"
#include "quantum_geometric/core/quantum_geometric_curvature.h"
#include "quantum_geometric/core/quantum_geometric_gradient.h"
#include "quantum_geometric/distributed/differential_geometry.h"
#include <assert.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#define S2 0.70710678118654752440
#define CHECK(c) do { if (!(c)) { fprintf(stderr, "CHECK failed %s:%d: %s\n", __FILE__, __LINE__, #c); exit(1); } } while (0)

static QuantumGate G(gate_type_t t, uint32_t target, uint32_t control, double p) {
    QuantumGate g = { .type = t, .target_qubit = target, .control_qubit = control, .parameter = p, .parameters = NULL, .num_parameters = 0 };
    return g;
}

static void run(double complex* st, size_t nq, const QuantumGate* gs, size_t ng) {
    CPUSimCircuit* c = cpu_sim_create_circuit(ng + 2);
    CHECK(c);
    for (size_t i = 0; i < ng; i++) cpu_sim_add_gate(c, &gs[i]);
    simulate_circuit_cpu(st, c, nq);
    cpu_sim_cleanup_circuit(c);
}

static void test_simulator_known_answers(void) {
    double complex s1[2], s2[4], s3[8];
    init_simulator_state(s1, 2);
    CHECK(fabs(cabs(s1[0]) - 1.0) < 1e-10 && cabs(s1[1]) < 1e-10);
    QuantumGate h = G(GATE_H, 0, 0, 0);
    run(s1, 1, &h, 1);
    CHECK(fabs(cabs(s1[0]) - S2) < 1e-10 && fabs(cabs(s1[1]) - S2) < 1e-10);              /* :64-67 */
    init_simulator_state(s2, 4);
    QuantumGate bell[2] = { G(GATE_H, 0, 0, 0), G(GATE_CNOT, 1, 0, 0) };
    run(s2, 2, bell, 2);
    CHECK(fabs(cabs(s2[0]) - S2) < 1e-10 && cabs(s2[1]) < 1e-10 && cabs(s2[2]) < 1e-10 && fabs(cabs(s2[3]) - S2) < 1e-10);  /* :131-134 */
    init_simulator_state(s1, 2);
    QuantumGate x = G(GATE_X, 0, 0, 0);
    run(s1, 1, &x, 1);
    CHECK(cabs(s1[0]) < 1e-10 && fabs(cabs(s1[1]) - 1.0) < 1e-10);                          /* :177-178 */
    init_simulator_state(s1, 2);
    QuantumGate hz[2] = { G(GATE_H, 0, 0, 0), G(GATE_Z, 0, 0, 0) };
    run(s1, 1, hz, 2);
    CHECK(fabs(cabs(s1[0]) - S2) < 1e-10 && fabs(cabs(s1[1]) - S2) < 1e-10);              /* :234-235 */
    CHECK(fabs(creal(s1[1]) + S2) < 1e-12);                                               /* and the phase the reference never checks */
    init_simulator_state(s1, 2);
    QuantumGate rx = G(GATE_RX, 0, 0, M_PI);
    run(s1, 1, &rx, 1);
    CHECK(cabs(s1[0]) < 1e-10 && fabs(cabs(s1[1]) - 1.0) < 1e-10);                          /* :281-282 */
    init_simulator_state(s3, 8);
    QuantumGate ghz[3] = { G(GATE_H, 0, 0, 0), G(GATE_CNOT, 1, 0, 0), G(GATE_CNOT, 2, 1, 0) };
    run(s3, 3, ghz, 3);
    CHECK(fabs(cabs(s3[0]) - S2) < 1e-10 && fabs(cabs(s3[7]) - S2) < 1e-10);              /* :404-408 */
    for (int i = 1; i < 7; i++) CHECK(cabs(s3[i]) < 1e-10);
    /* where the reference's 1-qubit kernel is wrong: target >= 1 on a superposed register */
    double complex s4[16];
    init_simulator_state(s4, 16);
    QuantumGate sup[8] = { G(GATE_RY, 0, 0, 0.3), G(GATE_RY, 1, 0, 0.5), G(GATE_RY, 2, 0, 0.7), G(GATE_RY, 3, 0, 0.9),
                           G(GATE_RX, 1, 0, 1.1), G(GATE_RX, 2, 0, 1.3), G(GATE_RX, 3, 0, 1.5), G(GATE_CRY, 3, 1, 0.4) };
    run(s4, 4, sup, 8);
    double nrm = 0;
    for (int i = 0; i < 16; i++) nrm += creal(s4[i] * conj(s4[i]));
    CHECK(fabs(nrm - 1.0) < 1e-12);
    double avg, mx;
    CPUSimCircuit* c = cpu_sim_create_circuit(4);
    cpu_sim_add_gate(c, &bell[0]); cpu_sim_add_gate(c, &bell[1]);
    cpu_sim_get_error_statistics(c, &avg, &mx);
    CHECK(fabs(mx - 0.005) < 1e-15 && fabs(avg - 0.003) < 1e-15 && c->num_qubits == 2);
    cpu_sim_cleanup_circuit(c);
    printf("  simulator_cpu entry points: ok\n");
}

static void test_sim_api(void) {
    SimulatorState* st = sim_init(3, 0, NULL);
    SimulatorCircuit* c = sim_create_circuit(3, 0);
    CHECK(st && c);
    double p[4] = { 0.7, 0, 0, 0 };
    CHECK(sim_add_gate(c, GATE_H, 0, 0, NULL) && sim_add_gate(c, GATE_CNOT, 1, 0, NULL) && sim_add_gate(c, GATE_RZ, 2, 0, p));
    CHECK(!sim_add_gate(c, GATE_RX, 0, 0, NULL));           /* a rotation without parameters is rejected */
    CHECK(sim_execute_circuit(st, c));
    double complex* sv = sim_get_statevector(st);
    CHECK(sv && fabs(cabs(sv[0]) - S2) < 1e-12 && fabs(cabs(sv[3]) - S2) < 1e-12);
    CHECK(fabs(carg(sv[0]) + 0.35) < 1e-12);                /* RZ(0.7) on |0>: phase -0.35 */
    free(sv);
    sim_reset_state(st);
    CHECK(creal(st->amplitudes[0]) == 1.0);
    sim_cleanup_circuit(c);
    sim_cleanup(st);
    printf("  sim_* entry points: ok\n");
}

static void add_rot(quantum_geometric_tensor_network_t* n, gate_type_t t, size_t q, double angle) {
    double p[1] = { angle };
    size_t tq[1] = { q };
    quantum_gate_t g;
    memset(&g, 0, sizeof g);
    g.type = t; g.num_qubits = 1; g.is_parameterized = true; g.target_qubits = tq; g.parameters = p; g.num_parameters = 1;
    CHECK(apply_quantum_gate(n, &g, NULL, 0));
}

static void add_cnot(quantum_geometric_tensor_network_t* n, size_t control, size_t target) {
    size_t qs[2] = { control, target };
    quantum_gate_t g;
    memset(&g, 0, sizeof g);
    g.type = GATE_CNOT; g.num_qubits = 2; g.is_controlled = true;
    CHECK(apply_quantum_gate(n, &g, qs, 2));
}

static void test_qgt_api(void) {
    /* single qubit RY(t1) RZ(t2): Q = [[1/4, i sin(t1)/4], [-i sin(t1)/4, sin^2(t1)/4]] */
    quantum_geometric_tensor_network_t* n = create_quantum_geometric_tensor_network(1, 1, false, true);
    CHECK(n);
    add_rot(n, GATE_RY, 0, 0.7);
    add_rot(n, GATE_RZ, 0, 1.3);
    ComplexFloat q;
    CHECK(compute_quantum_geometric_tensor(n, 0, 1, &q));
    CHECK(fabsf(q.real) < 1e-6f && fabsf(q.imag - 0.16105442f) < 1e-6f);
    double g11, b01;
    CHECK(compute_quantum_metric(n, 1, 1, &g11) && fabs(g11 - 0.10375411) < 1e-8);
    CHECK(compute_berry_curvature(n, 0, 1, &b01) && fabs(b01 - 0.25 * sin(0.7)) < 1e-12);
    CHECK(!compute_quantum_metric(n, 2, 0, &g11));
    destroy_quantum_geometric_tensor_network(n);

    /* 4-qubit two-layer ansatz: full matrix calls agree with the element calls; legacy ctor keeps its 16 cap */
    const size_t nq = 4, P = 16;
    n = create_quantum_geometric_tensor_network(nq, 2, false, true);
    double ang = 0.1;
    for (int layer = 0; layer < 2; layer++) {
        for (size_t qb = 0; qb < nq; qb++) add_rot(n, GATE_RY, qb, ang += 0.37);
        for (size_t qb = 0; qb < nq; qb++) add_rot(n, GATE_RZ, qb, ang += 0.21);
        for (size_t qb = 0; qb + 1 < nq; qb++) add_cnot(n, qb, qb + 1);
    }
    ComplexFloat* full = malloc(P * P * sizeof *full);
    CHECK(geometric_compute_full_qgt(full, n, P) == QGT_SUCCESS);
    CHECK(geometric_compute_full_qgt(full, n, P + 1) == QGT_ERROR_DIMENSION_MISMATCH);
    quantum_geometric_metric_t* m = NULL;
    quantum_geometric_curvature_t* cv = NULL;
    CHECK(geometric_create_metric(&m, GEOMETRIC_METRIC_FUBINI_STUDY, 17, HARDWARE_TYPE_CPU) == QGT_ERROR_INVALID_PARAMETER);
    CHECK(geometric_create_metric(&m, GEOMETRIC_METRIC_FUBINI_STUDY, P, HARDWARE_TYPE_CPU) == QGT_SUCCESS);
    CHECK(geometric_create_curvature(&cv, GEOMETRIC_CURVATURE_BERRY, P, HARDWARE_TYPE_CPU) == QGT_SUCCESS);
    CHECK(geometric_compute_fubini_study_metric(m, n, P) == QGT_SUCCESS);
    CHECK(geometric_compute_berry_curvature(cv, n, P) == QGT_SUCCESS);
    ComplexFloat* comp = malloc(P * P * sizeof *comp);
    CHECK(geometric_compose_qgt(comp, m, cv, P) == QGT_SUCCESS);
    for (size_t i = 0; i < P; i++)
        for (size_t j = 0; j < P; j++) {
            double gij, bij;
            CHECK(compute_quantum_metric(n, i, j, &gij) && compute_berry_curvature(n, i, j, &bij));
            CHECK(fabs(full[i * P + j].real - gij) < 1e-6 && fabs(full[i * P + j].imag - bij) < 1e-6);
            CHECK(comp[i * P + j].real == full[i * P + j].real && comp[i * P + j].imag == full[i * P + j].imag);
            CHECK(fabsf(m->components[i * P + j].real - m->components[j * P + i].real) < 1e-6f);      /* g symmetric */
            CHECK(fabsf(cv->components[i * P + j].real + cv->components[j * P + i].real) < 1e-6f);    /* Omega antisymmetric */
        }
    for (size_t i = 0; i < nq; i++) CHECK(fabsf(m->components[i * P + i].real - 0.25f) < 1e-6f);       /* first RY layer */
    ComplexFloat* sv = NULL; size_t dim = 0;
    CHECK(get_quantum_state(n, &sv, &dim) && dim == 16);
    double nrm = 0;
    for (size_t i = 0; i < dim; i++) nrm += (double)sv[i].real * sv[i].real + (double)sv[i].imag * sv[i].imag;
    CHECK(fabs(nrm - 1.0) < 1e-5);
    free(sv); free(full); free(comp);
    geometric_destroy_metric(m); geometric_destroy_curvature(cv);
    quantum_geometric_metric_t* big = NULL;
    CHECK(qgt_b200_alloc_metric(&big, 256) == QGT_SUCCESS && big->dimension == 256);
    geometric_destroy_metric(big);
    destroy_quantum_geometric_tensor_network(n);
    printf("  tensor-network / metric / curvature entry points: ok\n");
}

static void test_diffgeo(void) {
    const size_t dim = 512, P = 5;
    ComplexDouble* psi = malloc(dim * sizeof *psi);
    ComplexDouble* d = malloc(P * dim * sizeof *d);
    unsigned s = 12345;
    double nrm = 0;
    for (size_t i = 0; i < dim; i++) {
        s = s * 1103515245u + 12345u; psi[i].real = (double)(s >> 8) / (1 << 24) - 0.5;
        s = s * 1103515245u + 12345u; psi[i].imag = (double)(s >> 8) / (1 << 24) - 0.5;
        nrm += psi[i].real * psi[i].real + psi[i].imag * psi[i].imag;
    }
    for (size_t i = 0; i < dim; i++) { psi[i].real /= sqrt(nrm); psi[i].imag /= sqrt(nrm); }
    for (size_t i = 0; i < P * dim; i++) {
        s = s * 1103515245u + 12345u; d[i].real = ((double)(s >> 8) / (1 << 24) - 0.5) / 16;
        s = s * 1103515245u + 12345u; d[i].imag = ((double)(s >> 8) / (1 << 24) - 0.5) / 16;
    }
    double g[25], f[25];
    diffgeo_engine_t* e = diffgeo_engine_create();
    CHECK(e && diffgeo_compute_fubini_study(e, psi, dim, d, P, g) && diffgeo_compute_berry_curvature(e, psi, dim, d, P, f));
    for (size_t i = 0; i < P; i++)
        for (size_t j = 0; j < P; j++) {          /* the reference's triple loop, differential_geometry.c:2832-2856 */
            double complex in = 0, vi = 0, vj = 0;
            for (size_t k = 0; k < dim; k++) {
                double complex di = d[i * dim + k].real - I * d[i * dim + k].imag, dj = d[j * dim + k].real + I * d[j * dim + k].imag;
                double complex p = psi[k].real + I * psi[k].imag;
                in += di * dj; vi += di * p; vj += conj(p) * dj;
            }
            double complex q = in - vi * vj;
            CHECK(fabs(g[i * P + j] - creal(q)) < 1e-13 && fabs(f[i * P + j] + 2 * cimag(q)) < 1e-13);
        }
    diffgeo_engine_destroy(e);
    free(psi); free(d);
    printf("  diffgeo_compute_fubini_study / _berry_curvature: ok\n");
}

static void test_natural_gradient(void) {
    const size_t P = 6;
    ComplexFloat G[36], g[6], x[6];
    for (size_t i = 0; i < P; i++) {
        for (size_t j = 0; j < P; j++) { G[i * P + j].real = (i == j ? 0.25f : 0.02f / (1 + (float)(i > j ? i - j : j - i))); G[i * P + j].imag = 0; }
        g[i].real = 0.1f * (float)(i + 1); g[i].imag = -0.05f * (float)i;
    }
    natural_gradient_config_t cfg = get_default_natural_gradient_config();
    CHECK(fabsf(cfg.regularization_param - 1e-4f) < 1e-10f && cfg.use_adaptive_regularization);
    CHECK(compute_regularized_natural_gradient(g, G, x, P, &cfg));
    for (size_t i = 0; i < P; i++) {
        double re = 0, im = 0;
        for (size_t j = 0; j < P; j++) { double a = G[i * P + j].real + (i == j ? 1e-4 : 0.0); re += a * x[j].real; im += a * x[j].imag; }
        CHECK(fabs(re - g[i].real) < 1e-5 && fabs(im - g[i].imag) < 1e-5);
    }
    /* a complex Hermitian "metric" (the full Q = g + i Omega): the complex system is solved, as the reference's general
       inverse would; a non-Hermitian matrix is refused */
    ComplexFloat H[36], y[6];
    for (size_t i = 0; i < P; i++)
        for (size_t j = 0; j < P; j++) { H[i * P + j] = G[i * P + j]; H[i * P + j].imag = i == j ? 0.0f : (i < j ? 0.01f : -0.01f) * (float)(i + j); }
    CHECK(compute_regularized_natural_gradient(g, H, y, P, &cfg));
    for (size_t i = 0; i < P; i++) {
        double re = 0, im = 0;
        for (size_t j = 0; j < P; j++) {
            const double ar = H[i * P + j].real + (i == j ? 1e-4 : 0.0), ai = H[i * P + j].imag;
            re += ar * y[j].real - ai * y[j].imag; im += ar * y[j].imag + ai * y[j].real;
        }
        CHECK(fabs(re - g[i].real) < 1e-5 && fabs(im - g[i].imag) < 1e-5);
    }
    H[1].real += 0.1f;                                   /* H[0][1] != conj(H[1][0]) */
    CHECK(!compute_regularized_natural_gradient(g, H, y, P, &cfg));
    printf("  compute_regularized_natural_gradient: ok\n");
}

/* two threads drive independent circuits at the same time: each thread owns its own device context */
#include <pthread.h>
static void* thread_body(void* arg) {
    const int id = (int)(size_t)arg;
    for (int rep = 0; rep < 20; rep++) {
        const size_t nq = 10 + (size_t)id;
        double complex* st = malloc(((size_t)1 << nq) * sizeof *st);
        init_simulator_state(st, (size_t)1 << nq);
        QuantumGate gs[3] = { G(GATE_H, 0, 0, 0), G(GATE_CNOT, (uint32_t)nq - 1, 0, 0), G(GATE_RZ, (uint32_t)id, 0, 0.3 * (id + 1)) };
        run(st, nq, gs, 3);
        const size_t hi = ((size_t)1 << (nq - 1)) | 1;
        if (fabs(cabs(st[0]) - S2) > 1e-12 || fabs(cabs(st[hi]) - S2) > 1e-12) { free(st); return (void*)1; }
        free(st);
    }
    return NULL;
}
static void test_two_threads(void) {
    pthread_t t[2];
    void* r[2] = {NULL, NULL};
    for (size_t i = 0; i < 2; i++) CHECK(pthread_create(&t[i], NULL, thread_body, (void*)i) == 0);
    for (size_t i = 0; i < 2; i++) pthread_join(t[i], &r[i]);
    CHECK(r[0] == NULL && r[1] == NULL);
    printf("  two threads, two contexts: ok\n");
}

/* the device-memory seam, in the style of the reference's tests/test_quantum_geometric_gpu.c */
static void test_gpu_memory_seam(void) {
    CHECK(qg_gpu_init() == QG_GPU_SUCCESS);
    CHECK(qg_gpu_init() == QG_GPU_SUCCESS);                       /* idempotent */
    int count = 0;
    CHECK(qg_gpu_get_device_count(&count) == QG_GPU_SUCCESS && count >= 1);
    gpu_device_info_t info;
    CHECK(qg_gpu_get_device_info(0, &info) == QG_GPU_SUCCESS);
    CHECK(info.compute_capability_major == 10 && info.total_memory > ((size_t)1 << 30) && info.compute_units > 0 && strlen(info.name) > 0);
    CHECK(qg_gpu_get_device_info(count, &info) == QG_GPU_ERROR_INVALID_DEVICE);
    CHECK(qg_gpu_get_last_error() == QG_GPU_ERROR_INVALID_DEVICE);
    CHECK(qg_gpu_set_device(0) == QG_GPU_SUCCESS && qg_gpu_set_device(-1) == QG_GPU_ERROR_INVALID_DEVICE);

    enum { N = 4096 };
    double src[N], dst[N];
    for (int i = 0; i < N; i++) { src[i] = 0.5 * i - 7.0; dst[i] = 0.0; }
    gpu_buffer_t buf = {0};
    CHECK(qg_gpu_allocate(&buf, 0) == QG_GPU_ERROR_INVALID_VALUE);
    CHECK(qg_gpu_allocate(&buf, sizeof src) == QG_GPU_SUCCESS && buf.device_ptr && buf.size == sizeof src && !buf.is_pinned);
    CHECK(qg_gpu_memcpy_to_device(&buf, src, sizeof src + 8) == QG_GPU_ERROR_INVALID_VALUE);   /* larger than the buffer */
    CHECK(qg_gpu_memcpy_to_device(&buf, src, sizeof src) == QG_GPU_SUCCESS);
    CHECK(qg_gpu_memcpy_to_host(dst, &buf, sizeof dst) == QG_GPU_SUCCESS);
    CHECK(memcmp(src, dst, sizeof src) == 0);
    CHECK(qg_gpu_free(&buf) == QG_GPU_SUCCESS && buf.device_ptr == NULL && buf.size == 0);
    CHECK(qg_gpu_free(&buf) == QG_GPU_ERROR_INVALID_VALUE);

    gpu_buffer_t pin = {0};
    CHECK(qg_gpu_allocate_pinned(&pin, sizeof src) == QG_GPU_SUCCESS && pin.is_pinned);
    CHECK(qg_gpu_memcpy_to_device(&pin, src, sizeof src) == QG_GPU_SUCCESS);
    memset(dst, 0, sizeof dst);
    CHECK(qg_gpu_memcpy_to_host(dst, &pin, sizeof dst) == QG_GPU_SUCCESS && memcmp(src, dst, sizeof src) == 0);
    CHECK(qg_gpu_free(&pin) == QG_GPU_SUCCESS);

    void* raw = NULL;
    CHECK(gpu_malloc(&raw, sizeof src) == QGT_SUCCESS && raw);
    CHECK(gpu_malloc(NULL, 16) == QGT_ERROR_INVALID_PARAMETER);
    CHECK(gpu_memcpy_host_to_device(raw, src, sizeof src) == QGT_SUCCESS);
    memset(dst, 0, sizeof dst);
    CHECK(gpu_memcpy_device_to_host(dst, raw, sizeof dst) == QGT_SUCCESS && memcmp(src, dst, sizeof src) == 0);
    CHECK(qgt_gpu_free_buffer(raw) == QGT_SUCCESS);

    int s0 = -1, s1 = -1;
    CHECK(qg_gpu_create_stream(&s0) == QG_GPU_SUCCESS && s0 >= 0);
    CHECK(qg_gpu_create_stream(&s1) == QG_GPU_SUCCESS && s1 >= 0 && s1 != s0);
    CHECK(qg_gpu_synchronize_stream(s0) == QG_GPU_SUCCESS && qg_gpu_synchronize() == QG_GPU_SUCCESS);
    CHECK(qg_gpu_destroy_stream(s0) == QG_GPU_SUCCESS && qg_gpu_destroy_stream(s0) == QG_GPU_ERROR_INVALID_VALUE);
    CHECK(strcmp(qg_gpu_get_error_string(QG_GPU_ERROR_NO_DEVICE), "No GPU device available") == 0);
    qg_gpu_cleanup();                                              /* destroys the stream still registered */
    CHECK(qg_gpu_get_device_count(&count) == QG_GPU_ERROR_NOT_INITIALIZED);
    printf("  qg_gpu_* / gpu_malloc device-memory seam: ok\n");
}

/* measurement / expectation / sampling / circuit text format through the reference's sim_* names */
static void test_sim_measurement_api(void) {
    SimulatorState* st = sim_init(3, 3, NULL);
    SimulatorCircuit* c = sim_create_circuit(3, 3);
    CHECK(st && c);
    double t = 1.1;
    CHECK(sim_add_gate(c, GATE_H, 0, 0, NULL) && sim_add_gate(c, GATE_CNOT, 1, 0, NULL) && sim_add_gate(c, GATE_RY, 2, 0, &t));
    CHECK(sim_execute_circuit(st, c));
    /* <ZZZ> of (|00>+|11>)/sqrt2 x RY(t)|0>: parity of the Bell pair is even, so <ZZZ> = cos t */
    CHECK(fabs(sim_get_expectation_value(st, "Z") - cos(t)) < 1e-12);
    CHECK(sim_get_expectation_value(st, "X") == 0.0);                       /* unknown observables give 0 like the reference */
    qgt_compat_seed(12345);
    uint64_t* counts = sim_get_measurement_counts(st, 4000);
    CHECK(counts);
    uint64_t tot = 0;
    for (int i = 0; i < 8; i++) tot += counts[i];
    CHECK(tot == 4000);
    CHECK(counts[1] == 0 && counts[2] == 0 && counts[5] == 0 && counts[6] == 0);      /* qubits 0 and 1 are perfectly correlated */
    const double p000 = 0.5 * cos(t / 2) * cos(t / 2);
    CHECK(fabs((double)counts[0] / 4000.0 - p000) < 0.04);
    free(counts);
    /* measuring qubit 0 collapses qubit 1 with it */
    CHECK(sim_measure_qubit(st, 0, 0));
    bool* bits = sim_get_measurement_results(st);
    CHECK(bits);
    const int b0 = bits[0];
    free(bits);
    double nrm = 0;
    for (int i = 0; i < 8; i++) {
        nrm += creal(st->amplitudes[i] * conj(st->amplitudes[i]));
        if (((i & 1) != b0) || (((i >> 1) & 1) != b0)) CHECK(cabs(st->amplitudes[i]) < 1e-14);
    }
    CHECK(fabs(nrm - 1.0) < 1e-12);
    CHECK(sim_measure_all(st));
    int nonzero = 0;
    for (int i = 0; i < 8; i++) if (cabs(st->amplitudes[i]) > 1e-12) nonzero++;
    CHECK(nonzero == 1);
    /* text format round trip */
    const char* path = "/tmp/qgt_b200_compat_circuit.txt";
    CHECK(sim_save_circuit(c, path));
    SimulatorCircuit* c2 = sim_load_circuit(path);
    CHECK(c2);
    SimulatorState* s1 = sim_init(3, 0, NULL); SimulatorState* s2 = sim_init(3, 0, NULL);
    CHECK(sim_execute_circuit(s1, c) && sim_execute_circuit(s2, c2));
    for (int i = 0; i < 8; i++) CHECK(cabs(s1->amplitudes[i] - s2->amplitudes[i]) < 1e-15);
    sim_cleanup(s1); sim_cleanup(s2); sim_cleanup_circuit(c2);
    sim_cleanup_circuit(c); sim_cleanup(st);
    /* a circuit with a gate outside the supported set is refused as a whole (state untouched) */
    double complex s4[4];
    init_simulator_state(s4, 4);
    QuantumGate bad[2] = { G(GATE_H, 0, 0, 0), G(GATE_TYPE_ISWAP, 1, 0, 0) };
    run(s4, 2, bad, 2);
    CHECK(creal(s4[0]) == 1.0 && cabs(s4[1]) == 0.0);
    CHECK(strstr(qgt_compat_last_error(), "unsupported") != NULL);
    printf("sim measurement / text format ok\n");
}

int main(int argc, char** argv) {
    const int host_only = argc > 1 && !strcmp(argv[1], "--host-only");
    printf("compat entry points (%s)\n", host_only ? "host-only parts" : "with GPU");
    test_natural_gradient();
    if (host_only) {
        double complex st[4];
        init_simulator_state(st, 4);
        CHECK(creal(st[0]) == 1.0 && cabs(st[3]) == 0.0);
        QuantumGate h = G(GATE_H, 0, 0, 0);
        run(st, 2, &h, 1);                       /* no device: reports the error, leaves the state untouched */
        CHECK(creal(st[0]) == 1.0 && cabs(st[1]) == 0.0);
        CHECK(strstr(qgt_compat_last_error(), "no CPU fallback") != NULL);
        /* the device-memory seam never hands out host memory as "device" memory */
        CHECK(qg_gpu_init() == QG_GPU_ERROR_NO_DEVICE && qg_gpu_get_last_error() == QG_GPU_ERROR_NO_DEVICE);
        gpu_buffer_t b = {0};
        CHECK(qg_gpu_allocate(&b, 64) == QG_GPU_ERROR_INVALID_VALUE && b.device_ptr == NULL);   /* not initialised */
        void* raw = NULL;
        CHECK(gpu_malloc(&raw, 64) == QGT_ERROR_GPU_NOT_AVAILABLE && raw == NULL);
        printf("all host-only checks passed\n");
        return 0;
    }
    test_simulator_known_answers();
    test_sim_api();
    test_qgt_api();
    test_diffgeo();
    test_gpu_memory_seam();
    test_sim_measurement_api();
    test_two_threads();
    printf("all compat checks passed\n");
    return 0;
}

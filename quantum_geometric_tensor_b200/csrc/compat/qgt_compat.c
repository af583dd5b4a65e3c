/*
 * qgt_compat.c — the reference's QGT entry points served by the CUDA library.
 *
 *   diffgeo_compute_fubini_study / _berry_curvature        distributed/differential_geometry.c:2819-2906
 *   create_quantum_geometric_tensor_network, apply_quantum_gate, get_quantum_state,
 *   compute_quantum_geometric_tensor / _metric / _berry_curvature
 *                                                          core/quantum_geometric_tensor_network.c:27-209,585-880,1034-1227
 *   geometric_compute_fubini_study_metric                  core/quantum_geometric_metric.c:353-396
 *   geometric_compute_berry_curvature / compose / full_qgt core/quantum_geometric_curvature.c:201-320
 *   compute_regularized_natural_gradient                   core/quantum_geometric_gradient.c:2721-2964
 *
 * The reference's per-element routine re-simulates the circuit for every (mu, nu) (6 P^2 circuit runs for a
 * full tensor) on an engine that does not apply gates.  Here the
 * network object just records the gate list; the first query evaluates the whole Q on the GPU in one
 * qgt_b200_qgt call (complex double), caches it, and every element / matrix query reads the cache, narrowed
 * to the ComplexFloat of the legacy API.
 */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "compat_common.h"

/* ---- diffgeo ------------------------------------------------------------------------------------------------ */
struct diffgeo_engine { unsigned long metric_evaluations, curvature_evaluations; };

diffgeo_engine_t* diffgeo_engine_create(void) { return (diffgeo_engine_t*)calloc(1, sizeof(diffgeo_engine_t)); }
void diffgeo_engine_destroy(diffgeo_engine_t* e) { free(e); }

static bool diffgeo_run(const ComplexDouble* state, size_t dim, const ComplexDouble* d, size_t P, double* metric, double* berry) {
    qgt_b200_ctx* ctx = qgt_compat_ctx();
    if (!ctx) return false;
    int rc = qgt_b200_gram(ctx, (const double*)state, (const double*)d, dim, P, metric, berry, NULL);
    if (rc) { qgt_compat_set_error("qgt_b200_gram", rc); return false; }
    return true;
}

bool diffgeo_compute_fubini_study(diffgeo_engine_t* e, const ComplexDouble* state, size_t dim,
                                  const ComplexDouble* d, size_t P, double* metric_out) {
    if (!e || !state || !d || !metric_out || dim == 0 || P == 0) return false;
    if (!diffgeo_run(state, dim, d, P, metric_out, NULL)) return false;
    e->metric_evaluations++;
    return true;
}

bool diffgeo_compute_berry_curvature(diffgeo_engine_t* e, const ComplexDouble* state, size_t dim,
                                     const ComplexDouble* d, size_t P, double* curvature_out) {
    if (!e || !state || !d || !curvature_out || dim == 0 || P == 0) return false;
    if (!diffgeo_run(state, dim, d, P, NULL, curvature_out)) return false;
    for (size_t i = 0; i < P * P; i++) curvature_out[i] *= -2.0;       /* F = -2 Im Q (:2899-2900) */
    e->curvature_evaluations++;
    return true;
}

/* ---- the "tensor network" object: a recorded gate list + cached Q --------------------------------------------- */
typedef struct {
    qgt_b200_gate* gates;
    size_t num_gates, capacity;
    double* theta;
    size_t num_params, theta_cap;
    double* q;              /* cached P x P complex (re, im), NULL when stale */
    size_t q_params;
} qgtn_state;

static __thread char g_qgtn_err[256];
const char* get_quantum_geometric_tensor_network_error(void) { return g_qgtn_err; }
static void qgtn_err(const char* m) { snprintf(g_qgtn_err, sizeof g_qgtn_err, "%s", m); }

quantum_geometric_tensor_network_t* create_quantum_geometric_tensor_network(size_t num_qubits, size_t num_layers,
                                                                             bool is_distributed, bool use_hardware_acceleration) {
    if (num_qubits == 0 || num_qubits > 40) { qgtn_err("invalid number of qubits"); return NULL; }
    quantum_geometric_tensor_network_t* n = (quantum_geometric_tensor_network_t*)calloc(1, sizeof *n);
    qgtn_state* st = (qgtn_state*)calloc(1, sizeof *st);
    if (!n || !st) { free(n); free(st); qgtn_err("out of memory"); return NULL; }
    n->num_qubits = num_qubits; n->num_layers = num_layers; n->is_distributed = is_distributed;
    n->use_hardware_acceleration = use_hardware_acceleration;
    n->hardware_config.type = QGTN_BACKEND_SIMULATOR; n->hardware_config.supports_gradients = true;
    n->backend_state = st;
    return n;
}

void destroy_quantum_geometric_tensor_network(quantum_geometric_tensor_network_t* n) {
    if (!n) return;
    qgtn_state* st = (qgtn_state*)n->backend_state;
    if (st) { free(st->gates); free(st->theta); free(st->q); free(st); }
    free(n);
}

bool apply_quantum_gate(quantum_geometric_tensor_network_t* n, const quantum_gate_t* gate, const size_t* qubits, size_t num_qubits) {
    if (!n || !gate || !n->backend_state) { qgtn_err("invalid arguments"); return false; }
    qgtn_state* st = (qgtn_state*)n->backend_state;
    /* qubits[] (when given) overrides the gate's own targets: last entry = target, first = control for 2-qubit gates */
    uint32_t target = 0, control = 0;
    if (qubits && num_qubits >= 1) { target = (uint32_t)qubits[num_qubits - 1]; if (num_qubits >= 2) control = (uint32_t)qubits[0]; }
    else {
        /* create_quantum_gate (core/quantum_gate_operations.c:225-319) stores {control, target} of a two-qubit gate in
           target_qubits and leaves control_qubits NULL */
        if (gate->target_qubits) target = (uint32_t)gate->target_qubits[gate->control_qubits || gate->num_qubits < 2 ? 0 : gate->num_qubits - 1];
        else if (gate->qubits && gate->num_qubits) target = (uint32_t)gate->qubits[gate->num_qubits - 1];
        if (gate->control_qubits && gate->num_controls) control = (uint32_t)gate->control_qubits[0];
        else if (gate->target_qubits && gate->num_qubits >= 2) control = (uint32_t)gate->target_qubits[0];
        else if (gate->qubits && gate->num_qubits >= 2) control = (uint32_t)gate->qubits[0];
    }
    if (target >= n->num_qubits || control >= n->num_qubits) { qgtn_err("qubit index out of range"); return false; }
    const bool param = gate->is_parameterized && gate->parameters && gate->num_parameters > 0;
    qgt_b200_gate g;
    if (!qgt_compat_convert_gate(gate->type, target, control, gate->parameters ? gate->parameters[0] : 0.0,
                                 param ? (int)st->num_params : -1, &g)) { qgtn_err("unsupported gate type"); return false; }
    if (st->num_gates == st->capacity) {
        size_t nc = st->capacity ? 2 * st->capacity : 64;
        qgt_b200_gate* ng = (qgt_b200_gate*)realloc(st->gates, nc * sizeof *ng);
        if (!ng) { qgtn_err("out of memory"); return false; }
        st->gates = ng; st->capacity = nc;
    }
    if (g.param >= 0) {
        if (st->num_params == st->theta_cap) {
            size_t nc = st->theta_cap ? 2 * st->theta_cap : 64;
            double* nt = (double*)realloc(st->theta, nc * sizeof *nt);
            if (!nt) { qgtn_err("out of memory"); return false; }
            st->theta = nt; st->theta_cap = nc;
        }
        st->theta[st->num_params++] = gate->parameters[0];
    }
    st->gates[st->num_gates++] = g;
    free(st->q); st->q = NULL;                       /* cached tensor is stale */
    return true;
}

static void qgtn_circuit(const quantum_geometric_tensor_network_t* n, qgt_b200_circuit* c) {
    const qgtn_state* st = (const qgtn_state*)n->backend_state;
    memset(c, 0, sizeof *c);
    c->num_qubits = (int32_t)n->num_qubits; c->num_params = (int32_t)st->num_params;
    c->gates = st->gates; c->num_gates = st->num_gates;
}

bool get_quantum_state(const quantum_geometric_tensor_network_t* n, ComplexFloat** state_vector, size_t* dimension) {
    if (!n || !state_vector || !dimension || !n->backend_state) { qgtn_err("invalid arguments"); return false; }
    qgt_b200_ctx* ctx = qgt_compat_ctx();
    if (!ctx) { qgtn_err(qgt_compat_last_error()); return false; }
    const qgtn_state* st = (const qgtn_state*)n->backend_state;
    const size_t dim = (size_t)1 << n->num_qubits;
    double* amps = (double*)calloc(dim, 2 * sizeof(double));
    ComplexFloat* out = (ComplexFloat*)malloc(dim * sizeof *out);
    if (!amps || !out) { free(amps); free(out); qgtn_err("out of memory"); return false; }
    amps[0] = 1.0;
    qgt_b200_circuit c;
    qgtn_circuit(n, &c);
    int rc = qgt_b200_simulate_host(ctx, amps, (int)n->num_qubits, &c, st->theta);
    if (rc) { qgt_compat_set_error("get_quantum_state", rc); qgtn_err(qgt_compat_last_error()); free(amps); free(out); return false; }
    for (size_t i = 0; i < dim; i++) { out[i].real = (float)amps[2 * i]; out[i].imag = (float)amps[2 * i + 1]; }
    free(amps);
    *state_vector = out;                             /* caller frees, as in the reference */
    *dimension = dim;
    return true;
}

/* ---- parameter shifts and derivative columns (core/quantum_parameter_shift.c:151-662, core/quantum_geometric_gradient.c:1849-2406)
 * Parameter index = order of the parameterised gates (find_parameterized_gate, quantum_parameter_shift.c:278-337).  The
 * reference's versions re-run a circuit engine that does not apply gates; here the circuit runs
 * on the device and the states come back as ComplexFloat arrays the caller frees, as in the reference. */
bool shift_parameter(quantum_geometric_tensor_network_t* n, size_t param_idx, double shift_amount) {
    if (!n || !n->backend_state) { qgtn_err("invalid arguments"); return false; }
    qgtn_state* st = (qgtn_state*)n->backend_state;
    if (param_idx >= st->num_params) { qgtn_err("parameter index out of range"); return false; }
    st->theta[param_idx] += shift_amount;
    free(st->q); st->q = NULL;
    return true;
}

/* |psi(theta)> of the recorded circuit on the device */
static qgt_b200_state* qgtn_run(const quantum_geometric_tensor_network_t* n, const double* theta) {
    qgt_b200_ctx* ctx = qgt_compat_ctx();
    if (!ctx) { qgtn_err(qgt_compat_last_error()); return NULL; }
    qgt_b200_state* s = NULL;
    int rc = qgt_b200_state_create(ctx, (int)n->num_qubits, &s);
    if (!rc) rc = qgt_b200_state_init(s, QGT_B200_INIT_ZERO);
    if (!rc) {
        qgt_b200_circuit c;
        qgtn_circuit(n, &c);
        if (c.num_gates) rc = qgt_b200_apply_circuit(s, &c, theta);
    }
    if (rc) { qgt_compat_set_error("circuit run", rc); qgtn_err(qgt_compat_last_error()); qgt_b200_state_destroy(s); return NULL; }
    return s;
}

static ComplexFloat* qgtn_download(const qgt_b200_state* s, size_t dim) {
    ComplexFloat* out = (ComplexFloat*)malloc(dim * sizeof *out);
    if (!out) { qgtn_err("out of memory"); return NULL; }
    int rc = qgt_b200_state_download_c64(s, (float*)out);
    if (rc) { qgt_compat_set_error("state download", rc); qgtn_err(qgt_compat_last_error()); free(out); return NULL; }
    return out;
}

/* psi(theta + s e_mu) and psi(theta - s e_mu) on the device; the recorded parameters are left unchanged */
static bool qgtn_shifted_pair(const quantum_geometric_tensor_network_t* n, size_t param_idx, double shift,
                              qgt_b200_state** fwd, qgt_b200_state** bwd) {
    const qgtn_state* st = (const qgtn_state*)n->backend_state;
    if (param_idx >= st->num_params) { qgtn_err("parameter index out of range"); return false; }
    double* th = (double*)malloc((st->num_params ? st->num_params : 1) * sizeof(double));
    if (!th) { qgtn_err("out of memory"); return false; }
    memcpy(th, st->theta, st->num_params * sizeof(double));
    th[param_idx] = st->theta[param_idx] + shift;
    *fwd = qgtn_run(n, th);
    th[param_idx] = st->theta[param_idx] - shift;
    *bwd = *fwd ? qgtn_run(n, th) : NULL;
    free(th);
    if (!*fwd || !*bwd) { qgt_b200_state_destroy(*fwd); qgt_b200_state_destroy(*bwd); *fwd = *bwd = NULL; return false; }
    return true;
}

bool compute_shifted_states(quantum_geometric_tensor_network_t* n, size_t param_idx, double shift_amount,
                            ComplexFloat** forward_state, ComplexFloat** backward_state, size_t* dimension) {
    if (!n || !n->backend_state || !forward_state || !backward_state || !dimension) { qgtn_err("invalid arguments"); return false; }
    qgt_b200_state *f = NULL, *b = NULL;
    if (!qgtn_shifted_pair(n, param_idx, shift_amount, &f, &b)) return false;
    const size_t dim = (size_t)1 << n->num_qubits;
    ComplexFloat* fo = qgtn_download(f, dim);
    ComplexFloat* bo = fo ? qgtn_download(b, dim) : NULL;
    qgt_b200_state_destroy(f); qgt_b200_state_destroy(b);
    if (!fo || !bo) { free(fo); free(bo); return false; }
    *forward_state = fo; *backward_state = bo; *dimension = dim;
    return true;
}

/* (psi(theta + s) - psi(theta - s)) / (2 s): the formula of both reference routines (quantum_parameter_shift.c:195-196, 258-259) */
static bool qgtn_difference(const quantum_geometric_tensor_network_t* n, size_t param_idx, double s, ComplexFloat** gradient, size_t* dimension) {
    if (!n || !n->backend_state || !gradient || !dimension) { qgtn_err("invalid arguments"); return false; }
    if (s == 0.0) { qgtn_err("zero step"); return false; }
    qgt_b200_state *f = NULL, *b = NULL;
    if (!qgtn_shifted_pair(n, param_idx, s, &f, &b)) return false;
    int rc = qgt_b200_state_axpy(f, -1.0, 0.0, b);
    if (!rc) rc = qgt_b200_state_scale(f, 1.0 / (2.0 * s), 0.0);
    const size_t dim = (size_t)1 << n->num_qubits;
    ComplexFloat* out = rc ? NULL : qgtn_download(f, dim);
    if (rc) { qgt_compat_set_error("finite difference", rc); qgtn_err(qgt_compat_last_error()); }
    qgt_b200_state_destroy(f); qgt_b200_state_destroy(b);
    if (!out) return false;
    *gradient = out; *dimension = dim;
    return true;
}

bool compute_parameter_shift_gradient(const quantum_geometric_tensor_network_t* n, size_t param_idx, double shift_amount,
                                      ComplexFloat** gradient, size_t* dimension) {
    return qgtn_difference(n, param_idx, shift_amount, gradient, dimension);
}

bool compute_centered_difference_gradient(const quantum_geometric_tensor_network_t* n, size_t param_idx, double step_size,
                                          ComplexFloat** gradient, size_t* dimension) {
    return qgtn_difference(n, param_idx, step_size, gradient, dimension);
}

/* d_mu psi.  The reference combines shifted states with a formula that is not a derivative (quantum_geometric_gradient.c:2356-2360,
 * BASELINE.md §4 #6); this returns the exact column U_{>k} (-i/2 P) U_{<=k} |0> from the device, whatever the shift list says. */
bool compute_higher_order_gradient(const quantum_geometric_tensor_network_t* n, size_t param_idx, const double* shift_amounts,
                                   size_t num_shifts, ComplexFloat** gradient, size_t* dimension) {
    (void)shift_amounts; (void)num_shifts;
    if (!n || !n->backend_state || !gradient || !dimension) { qgtn_err("invalid arguments"); return false; }
    const qgtn_state* st = (const qgtn_state*)n->backend_state;
    if (param_idx >= st->num_params) { qgtn_err("parameter index out of range"); return false; }
    qgt_b200_ctx* ctx = qgt_compat_ctx();
    if (!ctx) { qgtn_err(qgt_compat_last_error()); return false; }
    qgt_b200_state* col = NULL;
    int rc = qgt_b200_state_create(ctx, (int)n->num_qubits, &col);
    if (!rc) {
        qgt_b200_circuit c;
        qgtn_circuit(n, &c);
        rc = qgt_b200_derivative(ctx, &c, st->theta, (int)param_idx, col);
    }
    const size_t dim = (size_t)1 << n->num_qubits;
    ComplexFloat* out = rc ? NULL : qgtn_download(col, dim);
    if (rc) { qgt_compat_set_error("qgt_b200_derivative", rc); qgtn_err(qgt_compat_last_error()); }
    qgt_b200_state_destroy(col);
    if (!out) return false;
    *gradient = out; *dimension = dim;
    return true;
}

/* error estimate = || D(h) - D(h/2) || of two centred differences (quantum_parameter_shift.c:17-149 compares two step sizes) */
bool compute_gradient_with_error(const quantum_geometric_tensor_network_t* n, size_t param_idx, ComplexFloat** gradient,
                                 double* error_estimate, size_t* dimension) {
    if (!gradient || !error_estimate || !dimension) { qgtn_err("invalid arguments"); return false; }
    ComplexFloat *g1 = NULL, *g2 = NULL;
    size_t d1 = 0, d2 = 0;
    if (!qgtn_difference(n, param_idx, 1e-2, &g1, &d1)) return false;
    if (!qgtn_difference(n, param_idx, 5e-3, &g2, &d2)) { free(g1); return false; }
    double e = 0.0;
    for (size_t i = 0; i < d1; i++) {
        const double dr = (double)g1[i].real - g2[i].real, di = (double)g1[i].imag - g2[i].imag;
        e += dr * dr + di * di;
    }
    free(g1);
    *gradient = g2; *error_estimate = sqrt(e); *dimension = d2;
    return true;
}

/* full Q of the recorded circuit, evaluated once and cached */
static const double* qgtn_full(const quantum_geometric_tensor_network_t* n) {
    qgtn_state* st = (qgtn_state*)n->backend_state;
    if (st->q && st->q_params == st->num_params) return st->q;
    qgt_b200_ctx* ctx = qgt_compat_ctx();
    if (!ctx) { qgtn_err(qgt_compat_last_error()); return NULL; }
    const size_t P = st->num_params;
    double* q = (double*)calloc(P * P ? P * P : 1, 2 * sizeof(double));
    if (!q) { qgtn_err("out of memory"); return NULL; }
    qgt_b200_circuit c;
    qgtn_circuit(n, &c);
    int rc = qgt_b200_qgt(ctx, &c, st->theta, NULL, NULL, q, NULL);
    if (rc) { qgt_compat_set_error("qgt_b200_qgt", rc); qgtn_err(qgt_compat_last_error()); free(q); return NULL; }
    free(st->q);
    st->q = q; st->q_params = P;
    return q;
}

bool compute_quantum_geometric_tensor(const quantum_geometric_tensor_network_t* n, size_t i, size_t j, ComplexFloat* result) {
    if (!n || !result || !n->backend_state) { qgtn_err("invalid arguments"); return false; }
    const size_t P = ((const qgtn_state*)n->backend_state)->num_params;
    if (i >= P || j >= P) { qgtn_err("parameter index out of range"); return false; }
    const double* q = qgtn_full(n);
    if (!q) return false;
    result->real = (float)q[2 * (i * P + j)]; result->imag = (float)q[2 * (i * P + j) + 1];
    return true;
}

bool compute_quantum_metric(const quantum_geometric_tensor_network_t* n, size_t i, size_t j, double* result) {
    if (!n || !result || !n->backend_state) { qgtn_err("invalid arguments"); return false; }
    const size_t P = ((const qgtn_state*)n->backend_state)->num_params;
    if (i >= P || j >= P) { qgtn_err("parameter index out of range"); return false; }
    const double* q = qgtn_full(n);
    if (!q) return false;
    *result = q[2 * (i * P + j)];                    /* Re Q (:1183-1203) */
    return true;
}

bool compute_berry_curvature(const quantum_geometric_tensor_network_t* n, size_t i, size_t j, double* result) {
    if (!n || !result || !n->backend_state) { qgtn_err("invalid arguments"); return false; }
    const size_t P = ((const qgtn_state*)n->backend_state)->num_params;
    if (i >= P || j >= P) { qgtn_err("parameter index out of range"); return false; }
    const double* q = qgtn_full(n);
    if (!q) return false;
    *result = q[2 * (i * P + j) + 1];                /* Im Q (:1205-1226) */
    return true;
}

/* ---- metric / curvature objects --------------------------------------------------------------------------------- */
static qgt_error_t alloc_metric(quantum_geometric_metric_t** m, geometric_metric_type_t type, size_t dim, HardwareType hw) {
    quantum_geometric_metric_t* x = (quantum_geometric_metric_t*)calloc(1, sizeof *x);
    if (!x) return QGT_ERROR_MEMORY_ALLOCATION;
    x->components = (ComplexFloat*)calloc(dim * dim, sizeof(ComplexFloat));
    if (!x->components) { free(x); return QGT_ERROR_MEMORY_ALLOCATION; }
    x->type = type; x->dimension = dim; x->is_symmetric = true; x->hardware = hw;
    *m = x;
    return QGT_SUCCESS;
}

qgt_error_t geometric_create_metric(quantum_geometric_metric_t** m, geometric_metric_type_t type, size_t dim, HardwareType hw) {
    if (!m || dim == 0 || dim > QGT_MAX_DIMENSIONS) return QGT_ERROR_INVALID_PARAMETER;     /* metric.c:14 */
    return alloc_metric(m, type, dim, hw);
}

qgt_error_t qgt_b200_alloc_metric(quantum_geometric_metric_t** m, size_t dim) {
    if (!m || dim == 0) return QGT_ERROR_INVALID_PARAMETER;
    return alloc_metric(m, GEOMETRIC_METRIC_FUBINI_STUDY, dim, HARDWARE_TYPE_CUDA);
}

void geometric_destroy_metric(quantum_geometric_metric_t* m) { if (m) { free(m->components); free(m); } }

static qgt_error_t alloc_curvature(quantum_geometric_curvature_t** c, geometric_curvature_type_t type, size_t dim, HardwareType hw) {
    quantum_geometric_curvature_t* x = (quantum_geometric_curvature_t*)calloc(1, sizeof *x);
    if (!x) return QGT_ERROR_MEMORY_ALLOCATION;
    x->components = (ComplexFloat*)calloc(dim * dim, sizeof(ComplexFloat));
    if (!x->components) { free(x); return QGT_ERROR_MEMORY_ALLOCATION; }
    x->type = type; x->dimension = dim; x->hardware = hw;
    *c = x;
    return QGT_SUCCESS;
}

qgt_error_t geometric_create_curvature(quantum_geometric_curvature_t** c, geometric_curvature_type_t type, size_t dim, HardwareType hw) {
    if (!c || dim == 0 || dim > QGT_MAX_DIMENSIONS) return QGT_ERROR_INVALID_PARAMETER;     /* curvature.c:14 */
    return alloc_curvature(c, type, dim, hw);
}

qgt_error_t qgt_b200_alloc_curvature(quantum_geometric_curvature_t** c, size_t dim) {
    if (!c || dim == 0) return QGT_ERROR_INVALID_PARAMETER;
    return alloc_curvature(c, GEOMETRIC_CURVATURE_BERRY, dim, HARDWARE_TYPE_CUDA);
}

void geometric_destroy_curvature(quantum_geometric_curvature_t* c) { if (c) { free(c->components); free(c); } }

static qgt_error_t check_params(const quantum_geometric_tensor_network_t* n, size_t num_params) {
    if (!n || !n->backend_state || num_params == 0) return QGT_ERROR_INVALID_PARAMETER;
    if (num_params != ((const qgtn_state*)n->backend_state)->num_params) return QGT_ERROR_DIMENSION_MISMATCH;
    return QGT_SUCCESS;
}

qgt_error_t geometric_compute_fubini_study_metric(quantum_geometric_metric_t* m, const quantum_geometric_tensor_network_t* n, size_t P) {
    if (!m || !m->components) return QGT_ERROR_INVALID_PARAMETER;
    qgt_error_t rc = check_params(n, P);
    if (rc) return rc;
    if (m->dimension != P) return QGT_ERROR_DIMENSION_MISMATCH;
    const double* q = qgtn_full(n);
    if (!q) return QGT_ERROR_INVALID_STATE;
    for (size_t i = 0; i < P * P; i++) { m->components[i].real = (float)q[2 * i]; m->components[i].imag = 0.0f; }
    m->type = GEOMETRIC_METRIC_FUBINI_STUDY; m->is_symmetric = true;
    return QGT_SUCCESS;
}

qgt_error_t geometric_compute_berry_curvature(quantum_geometric_curvature_t* c, const quantum_geometric_tensor_network_t* n, size_t P) {
    if (!c || !c->components) return QGT_ERROR_INVALID_PARAMETER;
    qgt_error_t rc = check_params(n, P);
    if (rc) return rc;
    if (c->dimension != P) return QGT_ERROR_DIMENSION_MISMATCH;
    const double* q = qgtn_full(n);
    if (!q) return QGT_ERROR_INVALID_STATE;
    bool flat = true;
    for (size_t i = 0; i < P * P; i++) {               /* Omega = Im Q, antisymmetric, stored in .real (curvature.c:201-247) */
        c->components[i].real = (float)q[2 * i + 1]; c->components[i].imag = 0.0f;
        if (fabs(q[2 * i + 1]) > 1e-12) flat = false;
    }
    c->type = GEOMETRIC_CURVATURE_BERRY; c->is_flat = flat;
    return QGT_SUCCESS;
}

qgt_error_t geometric_compute_berry_curvature_element(const quantum_geometric_tensor_network_t* n, size_t mu, size_t nu, float* result) {
    double v;
    if (!result) return QGT_ERROR_INVALID_PARAMETER;
    if (!compute_berry_curvature(n, mu, nu, &v)) return QGT_ERROR_INVALID_PARAMETER;
    *result = (float)v;
    return QGT_SUCCESS;
}

qgt_error_t geometric_compose_qgt(ComplexFloat* qgt, const quantum_geometric_metric_t* m, const quantum_geometric_curvature_t* c, size_t dim) {
    if (!qgt || !m || !c || !m->components || !c->components) return QGT_ERROR_INVALID_PARAMETER;
    if (m->dimension != dim || c->dimension != dim) return QGT_ERROR_DIMENSION_MISMATCH;
    for (size_t i = 0; i < dim * dim; i++) { qgt[i].real = m->components[i].real; qgt[i].imag = c->components[i].real; }   /* Q = g + i Omega */
    return QGT_SUCCESS;
}

qgt_error_t geometric_compute_full_qgt(ComplexFloat* qgt, const quantum_geometric_tensor_network_t* n, size_t P) {
    if (!qgt) return QGT_ERROR_INVALID_PARAMETER;
    qgt_error_t rc = check_params(n, P);
    if (rc) return rc;
    const double* q = qgtn_full(n);
    if (!q) return QGT_ERROR_INVALID_STATE;
    for (size_t i = 0; i < P * P; i++) { qgt[i].real = (float)q[2 * i]; qgt[i].imag = (float)q[2 * i + 1]; }
    return QGT_SUCCESS;
}

/* ---- natural gradient ----------------------------------------------------------------------------------------------- */
natural_gradient_config_t get_default_natural_gradient_config(void) {
    natural_gradient_config_t c = {1e-4f, 1e8f, true, true, 1e-10f};
    return c;
}

/* The metric may be the real symmetric Fubini-Study metric (the intended input) or a complex Hermitian matrix such as the
 * full Q = g + i Omega that geometric_compute_full_qgt produces (the reference inverts whatever complex matrix it is given,
 * quantum_geometric_gradient.c:2902-2950).  Hermitian input is solved exactly through its real symmetric embedding
 * [[Re, -Im], [Im, Re]]; a matrix that is not Hermitian (to 1e-5 of its largest entry, float data) is refused. */
bool compute_regularized_natural_gradient(const ComplexFloat* gradient, const ComplexFloat* metric, ComplexFloat* natural_gradient,
                                          size_t dimension, const natural_gradient_config_t* config) {
    if (!gradient || !metric || !natural_gradient || !config || dimension == 0) return false;
    const size_t P = dimension;
    double scale = 0.0, max_imag = 0.0, defect = 0.0;
    for (size_t i = 0; i < P; i++)
        for (size_t j = 0; j < P; j++) {
            const ComplexFloat a = metric[i * P + j], b = metric[j * P + i];
            scale = fmax(scale, fmax(fabs(a.real), fabs(a.imag)));
            max_imag = fmax(max_imag, fabs(a.imag));
            defect = fmax(defect, fmax(fabs((double)a.real - b.real), fabs((double)a.imag + b.imag)));
        }
    if (defect > 1e-5 * fmax(scale, 1e-30)) { qgt_compat_set_error("compute_regularized_natural_gradient: metric is not Hermitian", QGT_ERROR_INVALID_PARAMETER); return false; }
    const bool complex_metric = max_imag > 1e-7 * fmax(scale, 1e-30);
    const size_t N = complex_metric ? 2 * P : P;
    double* G = (double*)malloc(N * N * sizeof(double));
    double* g = (double*)malloc(2 * P * sizeof(double));
    double* x = (double*)malloc(2 * P * sizeof(double));
    if (!G || !g || !x) { free(G); free(g); free(x); return false; }
    for (size_t i = 0; i < P; i++) { g[i] = gradient[i].real; g[P + i] = gradient[i].imag; }
    qgt_b200_natgrad_config cfg = {config->regularization_param, config->condition_threshold, config->use_adaptive_regularization,
                                   config->use_pseudoinverse_fallback, config->singular_value_cutoff};
    int rc;
    if (!complex_metric) {
        for (size_t i = 0; i < P * P; i++) G[i] = metric[i].real;
        rc = qgt_b200_natural_gradient(NULL, G, g, P, &cfg, x, NULL);                /* real and imaginary parts: the solve is linear */
        if (!rc) rc = qgt_b200_natural_gradient(NULL, G, g + P, P, &cfg, x + P, NULL);
    } else {
        for (size_t i = 0; i < P; i++)
            for (size_t j = 0; j < P; j++) {
                const double re = 0.5 * ((double)metric[i * P + j].real + metric[j * P + i].real);
                const double im = 0.5 * ((double)metric[i * P + j].imag - metric[j * P + i].imag);
                G[i * N + j] = re;          G[i * N + P + j] = -im;
                G[(P + i) * N + j] = im;    G[(P + i) * N + P + j] = re;
            }
        rc = qgt_b200_natural_gradient(NULL, G, g, N, &cfg, x, NULL);               /* (Re x, Im x) of the complex system */
    }
    if (!rc) for (size_t i = 0; i < P; i++) { natural_gradient[i].real = (float)x[i]; natural_gradient[i].imag = (float)x[P + i]; }
    free(G); free(g); free(x);
    return rc == 0;
}

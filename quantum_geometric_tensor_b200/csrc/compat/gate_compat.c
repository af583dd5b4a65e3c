/*
 * gate_compat.c — gate objects of the reference's network API and the numerical-backend calls its programs make first.
 *
 *   create_quantum_gate / copy / update / shift / create_*_gate / destroy_quantum_gate
 *                                                   core/quantum_gate_operations.c:11-440 (tables :11-150)
 *   initialize_numerical_backend / shutdown / get_numerical_error_string
 *                                                   core/numerical_backend.h:9-41,146-147
 *
 * Host-only bookkeeping (a gate is a few numbers); the gates are consumed by apply_quantum_gate (qgt_compat.c), which
 * records them for the device.  tests/test_quantum_geometric_minimal.c of the reference builds against this.
 */
#include <math.h>
#include <stdlib.h>
#include <string.h>

#include "compat_common.h"

static ComplexFloat cf(double re, double im) { ComplexFloat z = {(float)re, (float)im}; return z; }

/* 2x2 tables: RX=[c,-is;-is,c], RY=[c,-s;s,c], RZ=diag(e^-i t/2, e^+i t/2) (quantum_gate_operations.c:11-47), fixed gates :50-130 */
static void one_qubit_matrix(gate_type_t type, double angle, ComplexFloat m[4]) {
    const double c = cos(angle / 2.0), s = sin(angle / 2.0), r = 1.0 / sqrt(2.0);
    m[0] = cf(1, 0); m[1] = cf(0, 0); m[2] = cf(0, 0); m[3] = cf(1, 0);
    switch (type) {
    case GATE_TYPE_RX: m[0] = cf(c, 0); m[1] = cf(0, -s); m[2] = cf(0, -s); m[3] = cf(c, 0); break;
    case GATE_TYPE_RY: m[0] = cf(c, 0); m[1] = cf(-s, 0); m[2] = cf(s, 0); m[3] = cf(c, 0); break;
    case GATE_TYPE_RZ: m[0] = cf(c, -s); m[3] = cf(c, s); break;
    case GATE_TYPE_X: m[0] = cf(0, 0); m[1] = cf(1, 0); m[2] = cf(1, 0); m[3] = cf(0, 0); break;
    case GATE_TYPE_Y: m[0] = cf(0, 0); m[1] = cf(0, -1); m[2] = cf(0, 1); m[3] = cf(0, 0); break;
    case GATE_TYPE_Z: m[3] = cf(-1, 0); break;
    case GATE_TYPE_H: m[0] = cf(r, 0); m[1] = cf(r, 0); m[2] = cf(r, 0); m[3] = cf(-r, 0); break;
    case GATE_TYPE_S: m[3] = cf(0, 1); break;
    case GATE_TYPE_T: m[3] = cf(cos(M_PI / 4), sin(M_PI / 4)); break;
    default: break;       /* identity, as the reference's fallback */
    }
}

static void fill_matrix(quantum_gate_t* g) {
    const size_t dim = (size_t)1 << g->num_qubits;
    memset(g->matrix, 0, dim * dim * sizeof(ComplexFloat));
    const double angle = g->parameters && g->num_parameters ? g->parameters[0] : 0.0;
    if (g->type == GATE_TYPE_CNOT || g->type == GATE_TYPE_CZ) {
        /* |0><0| (x) I + |1><1| (x) U with the control as the high index bit (:132-172) */
        ComplexFloat u[4];
        one_qubit_matrix(g->type == GATE_TYPE_CNOT ? GATE_TYPE_X : GATE_TYPE_Z, 0.0, u);
        for (size_t i = 0; i < dim / 2; i++) g->matrix[i * dim + i] = cf(1, 0);
        for (size_t i = 0; i < 2; i++)
            for (size_t j = 0; j < 2; j++) g->matrix[(dim / 2 + i) * dim + dim / 2 + j] = u[i * 2 + j];
    } else if (g->type == GATE_TYPE_SWAP && g->num_qubits == 2) {
        g->matrix[0] = cf(1, 0); g->matrix[1 * 4 + 2] = cf(1, 0); g->matrix[2 * 4 + 1] = cf(1, 0); g->matrix[15] = cf(1, 0);
    } else {
        ComplexFloat u[4];
        one_qubit_matrix(g->type, angle, u);
        if (dim == 2) memcpy(g->matrix, u, sizeof u);
        else { for (size_t i = 0; i < dim; i++) g->matrix[i * dim + i] = cf(1, 0); }
    }
}

quantum_gate_t* create_quantum_gate(gate_type_t type, const size_t* qubits, size_t num_qubits, const double* parameters, size_t num_parameters) {
    if (!qubits || num_qubits == 0 || num_qubits > 8 || (!parameters && num_parameters > 0)) return NULL;
    const bool rot = type == GATE_TYPE_RX || type == GATE_TYPE_RY || type == GATE_TYPE_RZ;
    const bool two = type == GATE_TYPE_CNOT || type == GATE_TYPE_CZ || type == GATE_TYPE_SWAP;
    if (rot && num_parameters != 1) return NULL;
    if (two && num_qubits != 2) return NULL;
    quantum_gate_t* g = (quantum_gate_t*)calloc(1, sizeof *g);
    if (!g) return NULL;
    const size_t dim = (size_t)1 << num_qubits;
    g->type = type;
    g->num_qubits = num_qubits;
    g->target_qubits = (size_t*)malloc(num_qubits * sizeof(size_t));
    g->matrix = (ComplexFloat*)malloc(dim * dim * sizeof(ComplexFloat));
    if (rot) g->parameters = (double*)malloc(sizeof(double));
    if (!g->target_qubits || !g->matrix || (rot && !g->parameters)) { destroy_quantum_gate(g); return NULL; }
    memcpy(g->target_qubits, qubits, num_qubits * sizeof(size_t));
    g->is_controlled = type == GATE_TYPE_CNOT || type == GATE_TYPE_CZ;
    if (rot) { g->is_parameterized = true; g->parameters[0] = parameters[0]; g->num_parameters = 1; }
    fill_matrix(g);
    return g;
}

quantum_gate_t* copy_quantum_gate(const quantum_gate_t* src) {
    if (!src) return NULL;
    quantum_gate_t* g = (quantum_gate_t*)calloc(1, sizeof *g);
    if (!g) return NULL;
    *g = *src;
    g->target_qubits = g->control_qubits = g->qubits = NULL; g->parameters = NULL; g->matrix = NULL; g->custom_data = NULL;
    const size_t dim = (size_t)1 << src->num_qubits;
    bool ok = true;
    if (src->target_qubits) { g->target_qubits = (size_t*)malloc(src->num_qubits * sizeof(size_t)); ok = ok && g->target_qubits; }
    if (src->control_qubits && src->num_controls) { g->control_qubits = (size_t*)malloc(src->num_controls * sizeof(size_t)); ok = ok && g->control_qubits; }
    if (src->parameters && src->num_parameters) { g->parameters = (double*)malloc(src->num_parameters * sizeof(double)); ok = ok && g->parameters; }
    if (src->matrix) { g->matrix = (ComplexFloat*)malloc(dim * dim * sizeof(ComplexFloat)); ok = ok && g->matrix; }
    if (!ok) { destroy_quantum_gate(g); return NULL; }
    if (g->target_qubits) memcpy(g->target_qubits, src->target_qubits, src->num_qubits * sizeof(size_t));
    if (g->control_qubits) memcpy(g->control_qubits, src->control_qubits, src->num_controls * sizeof(size_t));
    if (g->parameters) memcpy(g->parameters, src->parameters, src->num_parameters * sizeof(double));
    if (g->matrix) memcpy(g->matrix, src->matrix, dim * dim * sizeof(ComplexFloat));
    return g;
}

bool update_gate_parameters(quantum_gate_t* g, const double* parameters, size_t num_parameters) {
    if (!g || !g->is_parameterized || !parameters || num_parameters != g->num_parameters ||
        (g->type != GATE_TYPE_RX && g->type != GATE_TYPE_RY && g->type != GATE_TYPE_RZ)) return false;
    memcpy(g->parameters, parameters, num_parameters * sizeof(double));
    fill_matrix(g);
    return true;
}

bool shift_gate_parameters(quantum_gate_t* g, size_t param_idx, double shift_amount) {
    if (!g || !g->is_parameterized || !g->parameters || param_idx >= g->num_parameters) return false;
    g->parameters[param_idx] += shift_amount;
    fill_matrix(g);
    return true;
}

static quantum_gate_t* one(gate_type_t t, size_t q, const double* angle) { return create_quantum_gate(t, &q, 1, angle, angle ? 1 : 0); }
quantum_gate_t* create_rx_gate(size_t q, double a) { return one(GATE_TYPE_RX, q, &a); }
quantum_gate_t* create_ry_gate(size_t q, double a) { return one(GATE_TYPE_RY, q, &a); }
quantum_gate_t* create_rz_gate(size_t q, double a) { return one(GATE_TYPE_RZ, q, &a); }
quantum_gate_t* create_h_gate(size_t q) { return one(GATE_TYPE_H, q, NULL); }
quantum_gate_t* create_x_gate(size_t q) { return one(GATE_TYPE_X, q, NULL); }
quantum_gate_t* create_y_gate(size_t q) { return one(GATE_TYPE_Y, q, NULL); }
quantum_gate_t* create_z_gate(size_t q) { return one(GATE_TYPE_Z, q, NULL); }
quantum_gate_t* create_cnot_gate(size_t control, size_t target) { size_t q[2] = {control, target}; return create_quantum_gate(GATE_TYPE_CNOT, q, 2, NULL, 0); }
quantum_gate_t* create_cz_gate(size_t control, size_t target) { size_t q[2] = {control, target}; return create_quantum_gate(GATE_TYPE_CZ, q, 2, NULL, 0); }

void destroy_quantum_gate(quantum_gate_t* g) {
    if (!g) return;
    free(g->target_qubits); free(g->control_qubits); free(g->qubits); free(g->parameters); free(g->matrix);
    free(g);
}

/* ---- numerical backend: nothing to set up, the arithmetic of this layer runs in libqgt_b200 ------------------------ */
static bool g_numerical_up = false;

numerical_error_t initialize_numerical_backend(const numerical_config_t* config) {
    if (!config) return NUMERICAL_ERROR_INVALID_ARGUMENT;
    g_numerical_up = true;
    return NUMERICAL_SUCCESS;
}

void shutdown_numerical_backend(void) { g_numerical_up = false; }

const char* get_numerical_error_string(numerical_error_t e) {
    switch (e) {
    case NUMERICAL_SUCCESS: return "Success";
    case NUMERICAL_ERROR_INVALID_ARGUMENT: return "Invalid argument";
    case NUMERICAL_ERROR_MEMORY: return "Memory allocation failed";
    case NUMERICAL_ERROR_BACKEND: return "Backend error";
    case NUMERICAL_ERROR_COMPUTATION: return "Computation error";
    case NUMERICAL_ERROR_NOT_IMPLEMENTED: return "Not implemented";
    case NUMERICAL_ERROR_INVALID_STATE: return "Invalid state";
    default: return "Unknown error";
    }
}

/*
 * circuit_compat.c — the ComplexFloat circuit path of the reference ("path B"):
 *   quantum_circuit_create / destroy / reset, the gate builders, quantum_circuit_execute / measure / measure_all,
 *   validate / depth / gate_count              core/quantum_circuit_operations.c:692-1136, 1139-1223, 1245-1330
 *   init_quantum_state / quantum_state_reset / quantum_state_cleanup               :2570-2630
 *   per-gate loops                                                                   :113-288
 *
 * The flat gate list is lowered to one qgt_b200_circuit and run as fused sweeps on the device: the ComplexFloat
 * amplitudes are widened on the device, swept in complex double and narrowed back (16 * 2^n bytes cross PCIe each way
 * instead of one host pass per gate).  Gate conventions are the reference's for this path (:1147-1190), which differ from
 * its simulator by global phases: X = RX(pi) = -iX, Y = RY(pi) = -iY, Z = RZ(pi) = diag(-i, i), S = RZ(pi/2).
 * A phase gate recorded with an angle (quantum_circuit_phase) runs as RZ(angle); the reference records the angle and then
 * ignores it — with angle = pi/2 both agree.
 */
#include <math.h>
#include <stdlib.h>
#include <string.h>

#include "compat_common.h"

quantum_state* init_quantum_state(size_t num_qubits) {
    if (num_qubits == 0 || num_qubits > 40) return NULL;
    QuantumState* s = (QuantumState*)calloc(1, sizeof *s);
    if (!s) return NULL;
    s->num_qubits = num_qubits;
    s->dimension = (size_t)1 << num_qubits;
    s->amplitudes = (ComplexFloat*)calloc(s->dimension, sizeof(ComplexFloat));
    if (!s->amplitudes) { free(s); return NULL; }
    s->amplitudes[0].real = 1.0f;
    s->is_normalized = true;
    return s;
}

void quantum_state_reset(quantum_state* s) {
    if (!s || !s->amplitudes) return;
    memset(s->amplitudes, 0, s->dimension * sizeof(ComplexFloat));
    s->amplitudes[0].real = 1.0f;
    s->is_normalized = true;
}

void quantum_state_cleanup(quantum_state* s) {
    if (!s) return;
    free(s->amplitudes);
    free(s->workspace);
    free(s);
}

quantum_circuit_t* quantum_circuit_create(size_t num_qubits) {
    if (num_qubits == 0) return NULL;
    quantum_circuit_t* c = (quantum_circuit_t*)calloc(1, sizeof *c);
    if (!c) return NULL;
    c->num_qubits = num_qubits;
    c->max_gates = 1024;
    c->gates = (quantum_gate_t**)malloc(c->max_gates * sizeof(quantum_gate_t*));
    if (!c->gates) { free(c); return NULL; }
    return c;
}

static void free_gates(quantum_circuit_t* c) {
    for (size_t i = 0; i < c->num_gates; i++) {
        quantum_gate_t* g = c->gates[i];
        if (!g) continue;
        free(g->qubits); free(g->parameters); free(g->custom_data); free(g);
    }
    c->num_gates = 0;
}

void quantum_circuit_destroy(quantum_circuit_t* c) {
    if (!c) return;
    free_gates(c);
    free(c->gates);
    free(c);                    /* layers / graph / state / nodes are never populated by this layer */
}

void quantum_circuit_reset(quantum_circuit_t* c) {
    if (!c) return;
    free_gates(c);
    c->is_compiled = false;
}

static qgt_error_t push_gate(quantum_circuit_t* c, gate_type_t type, size_t q0, size_t q1, int nq, const double* angle) {
    if (!c || q0 >= c->num_qubits || (nq == 2 && (q1 >= c->num_qubits || q0 == q1))) return QGT_ERROR_INVALID_ARGUMENT;
    quantum_gate_t* g = (quantum_gate_t*)calloc(1, sizeof *g);
    if (!g) return QGT_ERROR_MEMORY_ALLOCATION;
    g->type = type;
    g->num_qubits = (size_t)nq;
    g->qubits = (size_t*)malloc((size_t)nq * sizeof(size_t));
    if (angle) g->parameters = (double*)malloc(sizeof(double));
    if (!g->qubits || (angle && !g->parameters)) { free(g->qubits); free(g->parameters); free(g); return QGT_ERROR_MEMORY_ALLOCATION; }
    g->qubits[0] = q0;
    if (nq == 2) g->qubits[1] = q1;
    if (angle) { g->parameters[0] = *angle; g->num_parameters = 1; }
    if (c->num_gates >= c->max_gates) {
        const size_t nm = c->max_gates * 2;
        quantum_gate_t** ng = (quantum_gate_t**)realloc(c->gates, nm * sizeof *ng);
        if (!ng) { free(g->qubits); free(g->parameters); free(g); return QGT_ERROR_MEMORY_ALLOCATION; }
        c->gates = ng; c->max_gates = nm;
    }
    c->gates[c->num_gates++] = g;
    return QGT_SUCCESS;
}

qgt_error_t quantum_circuit_hadamard(quantum_circuit_t* c, size_t q) { return push_gate(c, GATE_TYPE_H, q, 0, 1, NULL); }
qgt_error_t quantum_circuit_pauli_x(quantum_circuit_t* c, size_t q) { return push_gate(c, GATE_TYPE_X, q, 0, 1, NULL); }
qgt_error_t quantum_circuit_pauli_y(quantum_circuit_t* c, size_t q) { return push_gate(c, GATE_TYPE_Y, q, 0, 1, NULL); }
qgt_error_t quantum_circuit_pauli_z(quantum_circuit_t* c, size_t q) { return push_gate(c, GATE_TYPE_Z, q, 0, 1, NULL); }
qgt_error_t quantum_circuit_phase(quantum_circuit_t* c, size_t q, double angle) { return push_gate(c, GATE_TYPE_S, q, 0, 1, &angle); }
qgt_error_t quantum_circuit_rotation(quantum_circuit_t* c, size_t q, double angle, pauli_type axis) {
    if (axis != PAULI_X && axis != PAULI_Y && axis != PAULI_Z) return c && q < c->num_qubits ? QGT_ERROR_INVALID_PARAMETER : QGT_ERROR_INVALID_ARGUMENT;
    return push_gate(c, axis == PAULI_X ? GATE_TYPE_RX : axis == PAULI_Y ? GATE_TYPE_RY : GATE_TYPE_RZ, q, 0, 1, &angle);
}
qgt_error_t quantum_circuit_cnot(quantum_circuit_t* c, size_t control, size_t target) { return push_gate(c, GATE_TYPE_CNOT, control, target, 2, NULL); }
qgt_error_t quantum_circuit_cz(quantum_circuit_t* c, size_t control, size_t target) { return push_gate(c, GATE_TYPE_CZ, control, target, 2, NULL); }
qgt_error_t quantum_circuit_swap(quantum_circuit_t* c, size_t a, size_t b) { return push_gate(c, GATE_TYPE_SWAP, a, b, 2, NULL); }

/* the flat list in the device library's terms, with this path's phase conventions */
static qgt_error_t lower(const quantum_circuit_t* c, qgt_b200_gate** out, size_t* n_out) {
    qgt_b200_gate* gs = (qgt_b200_gate*)calloc(c->num_gates ? c->num_gates : 1, sizeof *gs);
    if (!gs) return QGT_ERROR_MEMORY_ALLOCATION;
    for (size_t i = 0; i < c->num_gates; i++) {
        const quantum_gate_t* g = c->gates[i];
        if (!g || !g->qubits || g->num_qubits == 0 || g->qubits[0] >= c->num_qubits) { free(gs); return QGT_ERROR_INVALID_ARGUMENT; }
        qgt_b200_gate* o = &gs[i];
        o->target = (int32_t)g->qubits[0]; o->control = -1; o->param = -1; o->scale = 1.0; o->angle = 0.0;
        const bool has_angle = g->parameters && g->num_parameters >= 1;
        switch (g->type) {
        case GATE_TYPE_H: o->kind = QGT_B200_GATE_H; break;
        case GATE_TYPE_X: o->kind = QGT_B200_GATE_RX; o->angle = M_PI; break;
        case GATE_TYPE_Y: o->kind = QGT_B200_GATE_RY; o->angle = M_PI; break;
        case GATE_TYPE_Z: o->kind = QGT_B200_GATE_RZ; o->angle = M_PI; break;
        case GATE_TYPE_S: o->kind = QGT_B200_GATE_RZ; o->angle = has_angle ? g->parameters[0] : M_PI_2; break;
        case GATE_TYPE_RX: case GATE_TYPE_RY: case GATE_TYPE_RZ:
            if (!has_angle) { free(gs); return QGT_ERROR_INVALID_ARGUMENT; }
            o->kind = (int32_t)g->type; o->angle = g->parameters[0];
            break;
        case GATE_TYPE_CNOT: case GATE_TYPE_CZ: case GATE_TYPE_SWAP:
            if (g->num_qubits < 2 || g->qubits[1] >= c->num_qubits) { free(gs); return QGT_ERROR_INVALID_ARGUMENT; }
            o->kind = (int32_t)g->type;
            o->control = (int32_t)g->qubits[0]; o->target = (int32_t)g->qubits[1];      /* qubits = {control, target} (:1176-1186) */
            break;
        default:
            free(gs);
            return QGT_ERROR_INVALID_OPERATOR;
        }
    }
    *out = gs; *n_out = c->num_gates;
    return QGT_SUCCESS;
}

static qgt_b200_state* upload(const quantum_state* s, qgt_error_t* err) {
    qgt_b200_ctx* ctx = qgt_compat_ctx();
    if (!ctx) { *err = QGT_ERROR_HARDWARE_FAILURE; return NULL; }
    qgt_b200_state* d = NULL;
    int rc = qgt_b200_state_create(ctx, (int)s->num_qubits, &d);
    if (!rc) rc = qgt_b200_state_upload_c64(d, (const float*)s->amplitudes);
    if (rc) { qgt_compat_set_error("quantum_circuit: state upload", rc); qgt_b200_state_destroy(d); *err = rc == QGT_B200_ERR_NO_MEMORY ? QGT_ERROR_MEMORY_ALLOCATION : QGT_ERROR_HARDWARE_FAILURE; return NULL; }
    return d;
}

qgt_error_t quantum_circuit_execute(quantum_circuit_t* c, quantum_state* s) {
    if (!c || !s) return QGT_ERROR_INVALID_ARGUMENT;
    if (c->num_qubits != s->num_qubits) return QGT_ERROR_INCOMPATIBLE;
    if (!s->amplitudes) return QGT_ERROR_INVALID_ARGUMENT;
    if (c->num_gates == 0) return QGT_SUCCESS;
    qgt_b200_gate* gs = NULL;
    size_t ng = 0;
    qgt_error_t err = lower(c, &gs, &ng);
    if (err) return err;
    qgt_b200_state* d = upload(s, &err);
    if (!d) { free(gs); return err; }
    qgt_b200_circuit qc;
    memset(&qc, 0, sizeof qc);
    qc.num_qubits = (int32_t)c->num_qubits; qc.gates = gs; qc.num_gates = ng;
    int rc = qgt_b200_apply_circuit(d, &qc, NULL);
    if (!rc) rc = qgt_b200_state_download_c64(d, (float*)s->amplitudes);
    if (rc) qgt_compat_set_error("quantum_circuit_execute", rc);
    qgt_b200_state_destroy(d);
    free(gs);
    return rc ? QGT_ERROR_HARDWARE_FAILURE : QGT_SUCCESS;
}

/* quantum_measure_qubit (:245-288) per qubit: probabilities, collapse and renormalisation run on the device; the outcome is
 * drawn from rand() like the reference does (outcome 0 when r < P(0)) */
static qgt_error_t measure_all(quantum_circuit_t* c, quantum_state* s, size_t* results) {
    if (!c || !s || !results) return QGT_ERROR_INVALID_ARGUMENT;
    if (c->num_qubits != s->num_qubits) return QGT_ERROR_INCOMPATIBLE;
    qgt_error_t err = QGT_SUCCESS;
    qgt_b200_state* d = upload(s, &err);
    if (!d) return err;
    int rc = 0;
    for (size_t q = 0; q < c->num_qubits && !rc; q++) {
        const double r = (double)rand() / RAND_MAX;
        int outcome = 0;
        rc = qgt_b200_state_measure(d, (int)q, 1.0 - r, 0.0, &outcome, NULL);
        results[q] = (size_t)outcome;
    }
    if (!rc) rc = qgt_b200_state_download_c64(d, (float*)s->amplitudes);
    if (rc) qgt_compat_set_error("quantum_circuit_measure", rc);
    qgt_b200_state_destroy(d);
    return rc ? QGT_ERROR_HARDWARE_FAILURE : QGT_SUCCESS;
}

qgt_error_t quantum_circuit_measure(quantum_circuit_t* c, quantum_state* s, size_t* results) { return measure_all(c, s, results); }
qgt_error_t quantum_circuit_measure_all(quantum_circuit_t* c, quantum_state* s, size_t* results) { return measure_all(c, s, results); }

qgt_error_t quantum_circuit_optimize(quantum_circuit_t* c, int level) {
    if (!c || level < 0) return QGT_ERROR_INVALID_ARGUMENT;
    c->optimization_level = level;          /* fusion into sweeps happens at execution time, whatever the level */
    return QGT_SUCCESS;
}

qgt_error_t quantum_circuit_validate(quantum_circuit_t* c) {
    if (!c) return QGT_ERROR_INVALID_ARGUMENT;
    for (size_t i = 0; i < c->num_gates; i++) {
        const quantum_gate_t* g = c->gates[i];
        if (!g || !g->qubits) return QGT_ERROR_INVALID_ARGUMENT;
        for (size_t j = 0; j < g->num_qubits; j++) if (g->qubits[j] >= c->num_qubits) return QGT_ERROR_INVALID_ARGUMENT;
        if ((g->type == GATE_TYPE_RX || g->type == GATE_TYPE_RY || g->type == GATE_TYPE_RZ || g->type == GATE_TYPE_S) &&
            (g->num_parameters != 1 || !g->parameters)) return QGT_ERROR_INVALID_ARGUMENT;       /* :1300-1312 */
    }
    return QGT_SUCCESS;
}

size_t quantum_circuit_depth(const quantum_circuit_t* c) {
    if (!c) return 0;
    size_t* last = (size_t*)calloc(c->num_qubits, sizeof(size_t));
    if (!last) return 0;
    size_t depth = 0;
    for (size_t i = 0; i < c->num_gates; i++) {
        const quantum_gate_t* g = c->gates[i];
        size_t m = 0;
        for (size_t j = 0; j < g->num_qubits; j++) if (last[g->qubits[j]] > m) m = last[g->qubits[j]];
        for (size_t j = 0; j < g->num_qubits; j++) last[g->qubits[j]] = m + 1;
        if (m + 1 > depth) depth = m + 1;
    }
    free(last);
    return depth;
}

size_t quantum_circuit_gate_count(const quantum_circuit_t* c) { return c ? c->num_gates : 0; }

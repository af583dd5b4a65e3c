/*
 * ref_driver.c — glue that drives the UNMODIFIED reference sources.   TEST INFRASTRUCTURE ONLY.
 *
 * Compiled by oracle/Makefile together with, in place from /root/reference,
 *   src/quantum_geometric/hardware/quantum_simulator.c        (oracle A: sim_* complex-double simulator)
 *   src/quantum_geometric/distributed/differential_geometry.c (oracle D: diffgeo_compute_fubini_study/_berry_curvature)
 * into oracle/_ref/libqgt_ref.so.  No reference source is copied into this repository.
 *
 * The functions here only translate the repo's POD circuit description into calls of the
 * reference's public API (sim_create_circuit / sim_add_gate / sim_execute_circuit) and form
 * derivative columns the way BASELINE.md §3 prescribes: run the reference simulator on the
 * prefix, apply the Pauli generator with the reference's own GATE_X/Y/Z, scale by -i/2, run the
 * suffix.  Only the 13 gate kinds of apply_gate_by_type (quantum_simulator.c:188-283) exist there.
 */
#include "quantum_geometric/hardware/quantum_simulator.h"
#include "quantum_geometric/distributed/differential_geometry.h"
#include <complex.h>
#include <stdarg.h>
#include <stdlib.h>
#include <string.h>

#include "../include/qgt_b200.h"

/* the two logging symbols quantum_simulator.c imports (core/quantum_geometric_logging.h:46-50) */
void geometric_log_error(const char* fmt, ...) { (void)fmt; }
void geometric_log_warning(const char* fmt, ...) { (void)fmt; }

static int ref_supported(int kind) {
    switch (kind) {
    case GATE_I: case GATE_X: case GATE_Y: case GATE_Z: case GATE_H: case GATE_S: case GATE_T:
    case GATE_RX: case GATE_RY: case GATE_RZ: case GATE_CNOT: case GATE_CZ: case GATE_SWAP: return 1;
    default: return 0;
    }
}

static double ref_angle(const qgt_b200_gate* g, const double* theta) {
    return g->param >= 0 ? g->scale * theta[g->param] + g->angle : g->angle;
}

/* run gates [lo, hi) of the description through sim_execute_circuit on `st` */
static int ref_run_range(SimulatorState* st, const qgt_b200_circuit* c, const double* theta, size_t lo, size_t hi) {
    SimulatorCircuit* circ = sim_create_circuit((uint32_t)c->num_qubits, 0);
    if (!circ) return -2;
    int rc = 0;
    for (size_t k = lo; k < hi && rc == 0; k++) {
        const qgt_b200_gate* g = &c->gates[k];
        if (!ref_supported(g->kind)) { rc = -7; break; }
        double p[4] = { ref_angle(g, theta), 0, 0, 0 };
        int rot = (g->kind == GATE_RX || g->kind == GATE_RY || g->kind == GATE_RZ);
        if (g->kind == GATE_SWAP) {
            /* sim_add_gate stores a control qubit only for CNOT/CZ (quantum_simulator.c:461-476), so a
             * SWAP added through it always pairs the target with qubit 0 at execution (:509-510) — a
             * reference defect.  Drive the reference's own decomposition
             * SWAP = CNOT(a,b) CNOT(b,a) CNOT(a,b) (:269-275) through its CNOT path instead. */
            uint32_t a = (uint32_t)g->control, b = (uint32_t)g->target;
            if (!sim_add_gate(circ, GATE_CNOT, b, a, NULL) || !sim_add_gate(circ, GATE_CNOT, a, b, NULL) ||
                !sim_add_gate(circ, GATE_CNOT, b, a, NULL)) rc = -2;
            continue;
        }
        if (!sim_add_gate(circ, (gate_type_t)g->kind, (uint32_t)g->target,
                          (uint32_t)(g->control < 0 ? 0 : g->control), rot ? p : NULL)) rc = -2;
    }
    if (rc == 0 && !sim_execute_circuit(st, circ)) rc = -15;
    sim_cleanup_circuit(circ);
    return rc;
}

int ref_apply_circuit(double* amps, const qgt_b200_circuit* c, const double* theta) {
    size_t dim = (size_t)1 << c->num_qubits;
    SimulatorState* st = sim_init((uint32_t)c->num_qubits, 0, NULL);
    if (!st) return -2;
    memcpy(st->amplitudes, amps, dim * sizeof(double complex));
    int rc = ref_run_range(st, c, theta, 0, c->num_gates);
    if (rc == 0) {
        double complex* sv = sim_get_statevector(st);
        if (sv) { memcpy(amps, sv, dim * sizeof(double complex)); free(sv); } else rc = -2;
    }
    sim_cleanup(st);
    return rc;
}

int ref_derivative(double* out, const qgt_b200_circuit* c, const double* theta, int mu) {
    size_t dim = (size_t)1 << c->num_qubits;
    double complex* acc = (double complex*)out;
    memset(acc, 0, dim * sizeof(double complex));
    for (size_t k = 0; k < c->num_gates; k++) {
        const qgt_b200_gate* g = &c->gates[k];
        if (g->param != mu) continue;
        int pauli = g->kind == GATE_RX ? GATE_X : g->kind == GATE_RY ? GATE_Y : g->kind == GATE_RZ ? GATE_Z : -1;
        if (pauli < 0) return -7;
        SimulatorState* st = sim_init((uint32_t)c->num_qubits, 0, NULL); /* |0...0> */
        if (!st) return -2;
        int rc = ref_run_range(st, c, theta, 0, k + 1);
        if (rc == 0) {
            qgt_b200_gate pg = { pauli, g->target, -1, -1, 0.0, 1.0 };
            qgt_b200_circuit one = *c;
            one.gates = &pg; one.num_gates = 1;
            rc = ref_run_range(st, &one, theta, 0, 1);
        }
        if (rc == 0) {
            double complex f = -0.5 * I * g->scale;
            for (size_t i = 0; i < dim; i++) st->amplitudes[i] *= f;
            rc = ref_run_range(st, c, theta, k + 1, c->num_gates);
        }
        if (rc == 0) for (size_t i = 0; i < dim; i++) acc[i] += st->amplitudes[i];
        sim_cleanup(st);
        if (rc) return rc;
    }
    return 0;
}

/* metric = g (Re Q); curvature = F = -2 Im Q exactly as the reference returns it */
int ref_fubini_berry(const double* psi, const double* dpsi, size_t dim, size_t P, double* metric, double* curvature) {
    diffgeo_engine_t* e = diffgeo_engine_create();
    if (!e) return -2;
    int rc = 0;
    if (metric && !diffgeo_compute_fubini_study(e, (const ComplexDouble*)psi, dim, (const ComplexDouble*)dpsi, P, metric)) rc = -15;
    if (curvature && !diffgeo_compute_berry_curvature(e, (const ComplexDouble*)psi, dim, (const ComplexDouble*)dpsi, P, curvature)) rc = -15;
    diffgeo_engine_destroy(e);
    return rc;
}

/* whole reference CPU path for one QGT evaluation (used as the cpu_baseline / --impl reference).
 * `max_cols` < P evaluates only the first max_cols columns (a bounded sample); outputs are
 * max_cols x max_cols then. */
int ref_qgt(const qgt_b200_circuit* c, const double* theta, size_t max_cols, double* metric, double* curvature) {
    if (c->initial_state != QGT_B200_INIT_ZERO) return -7;
    size_t dim = (size_t)1 << c->num_qubits, P = (size_t)c->num_params;
    if (max_cols && max_cols < P) P = max_cols;
    double* psi = (double*)calloc(dim, 16);
    double* J = (double*)malloc(P * dim * 16);
    if (!psi || !J) { free(psi); free(J); return -2; }
    psi[0] = 1.0;
    int rc = ref_apply_circuit(psi, c, theta);
    for (size_t mu = 0; mu < P && rc == 0; mu++) rc = ref_derivative(J + 2 * mu * dim, c, theta, (int)mu);
    if (rc == 0) rc = ref_fubini_berry(psi, J, dim, P, metric, curvature);
    free(psi); free(J);
    return rc;
}

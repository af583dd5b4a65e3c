/*
 * sim_compat.c — the reference's simulator entry points served by the CUDA library.
 *
 *   init_simulator_state / cpu_sim_* / simulate_circuit_cpu   hardware/quantum_simulator_cpu.c:98,761-926
 *   sim_init / sim_add_gate / sim_execute_circuit / ...        hardware/quantum_simulator.c:57-112,418-560,677
 *
 * Host buffers in and out exactly like the reference (double complex[2^n]); the wrappers stage them through
 * qgt_b200_simulate_host.  Known reference defects are NOT reproduced: 1-qubit
 * gates act correctly on every target, CRX/CRY/CRZ are real gates, SWAP uses both of its qubits.
 */
#include <math.h>
#include <pthread.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "compat_common.h"

/* One context per calling thread (its own stream and device buffers), created on first use and destroyed when the thread
 * ends: the reference's entry points are re-entrant on distinct objects, and a context's cached buffers are not. */
static int g_device = -1;
static int g_any_ctx = 0;
static pthread_mutex_t g_lock = PTHREAD_MUTEX_INITIALIZER;
static pthread_key_t g_ctx_key;
static pthread_once_t g_ctx_once = PTHREAD_ONCE_INIT;
static __thread char g_err[512];

static void ctx_destructor(void* p) { if (p) qgt_b200_destroy((qgt_b200_ctx*)p); }
static void ctx_key_init(void) { pthread_key_create(&g_ctx_key, ctx_destructor); }

void qgt_compat_set_error(const char* where, int status) {
    snprintf(g_err, sizeof g_err, "%s: %s (%s)", where, qgt_b200_error_string(status), qgt_b200_last_error());
}

const char* qgt_compat_last_error(void) { return g_err; }

int qgt_compat_set_device(int device) {
    pthread_mutex_lock(&g_lock);
    int rc = 0;
    if (g_any_ctx) rc = -17;        /* QGT_ERROR_ALREADY_INITIALIZED */
    else g_device = device;
    pthread_mutex_unlock(&g_lock);
    return rc;
}

qgt_b200_ctx* qgt_compat_ctx(void) {
    pthread_once(&g_ctx_once, ctx_key_init);
    qgt_b200_ctx* c = (qgt_b200_ctx*)pthread_getspecific(g_ctx_key);
    if (c) return c;
    pthread_mutex_lock(&g_lock);
    int dev = g_device;
    if (dev < 0) { const char* e = getenv("QGT_B200_DEVICE"); dev = e ? atoi(e) : 0; }
    int rc = qgt_b200_create(&c, dev);
    if (rc) { c = NULL; qgt_compat_set_error("qgt_b200_create", rc); }
    else g_any_ctx = 1;
    pthread_mutex_unlock(&g_lock);
    if (c) pthread_setspecific(g_ctx_key, c);
    return c;
}

int qgt_compat_convert_gate(gate_type_t type, uint32_t target, uint32_t control, double angle, int param, qgt_b200_gate* out) {
    out->kind = (int32_t)type; out->target = (int32_t)target; out->control = -1; out->param = param;
    out->angle = param >= 0 ? 0.0 : angle; out->scale = 1.0;
    switch (type) {
    case GATE_TYPE_I: case GATE_TYPE_X: case GATE_TYPE_Y: case GATE_TYPE_Z: case GATE_TYPE_H: case GATE_TYPE_S:
    case GATE_TYPE_T: case GATE_TYPE_SDG: case GATE_TYPE_TDG: case GATE_TYPE_SX:
        out->param = -1; out->angle = 0.0; return 1;
    case GATE_TYPE_RX: case GATE_TYPE_RY: case GATE_TYPE_RZ: case GATE_TYPE_U1: case GATE_TYPE_PHASE:
        return 1;
    case GATE_TYPE_CNOT: case GATE_TYPE_CY: case GATE_TYPE_CZ: case GATE_TYPE_CH: case GATE_TYPE_SWAP:
        out->control = (int32_t)control; out->param = -1; out->angle = 0.0; return 1;
    case GATE_TYPE_CRX: case GATE_TYPE_CRY: case GATE_TYPE_CRZ: case GATE_TYPE_ZZ:
        out->control = (int32_t)control; return 1;
    default:
        return 0;       /* MEASURE, BARRIER, RESET, multi-qubit kinds outside the statevector hot path */
    }
}

/* ---- hardware/quantum_simulator_cpu.h ------------------------------------------------------------------ */
void init_simulator_state(double complex* state, size_t n) {        /* n = number of amplitudes (…_cpu.c:98-101) */
    if (!state || n == 0) return;
    memset(state, 0, n * sizeof(double complex));
    state[0] = 1.0;
}

CPUSimCircuit* cpu_sim_create_circuit(size_t max_gates) {
    QuantumCircuit* c = (QuantumCircuit*)calloc(1, sizeof *c);
    if (!c) return NULL;
    c->gates = (HardwareGate*)calloc(max_gates ? max_gates : 1, sizeof(HardwareGate));
    if (!c->gates) { free(c); return NULL; }
    c->capacity = max_gates;
    return c;
}

static int two_qubit_kind(gate_type_t t) {
    return t == GATE_TYPE_CNOT || t == GATE_TYPE_CZ || t == GATE_TYPE_SWAP || t == GATE_TYPE_CRZ ||
           t == GATE_TYPE_CRX || t == GATE_TYPE_CRY || t == GATE_TYPE_CY || t == GATE_TYPE_CH || t == GATE_TYPE_ZZ;
}

void cpu_sim_add_gate(CPUSimCircuit* circuit, const QuantumGate* gate) {
    if (!circuit || !gate || circuit->num_gates >= circuit->capacity) return;     /* silent, like the reference */
    HardwareGate* g = &circuit->gates[circuit->num_gates++];
    memset(g, 0, sizeof *g);
    g->type = gate->type; g->target = gate->target_qubit; g->control = gate->control_qubit; g->parameter = gate->parameter;
    size_t top = gate->target_qubit;
    if (two_qubit_kind(gate->type) && gate->control_qubit > top) top = gate->control_qubit;
    if (top + 1 > circuit->num_qubits) circuit->num_qubits = top + 1;
}

void cpu_sim_cleanup_circuit(CPUSimCircuit* circuit) {
    if (!circuit) return;
    free(circuit->gates);
    free(circuit);
}

void cpu_sim_get_error_statistics(const CPUSimCircuit* circuit, double* avg_error_rate, double* max_error_rate) {
    if (!circuit || !avg_error_rate || !max_error_rate) return;
    /* the reference's fixed per-kind estimates (…_cpu.c:872-908): 0.5 % two-qubit, 1 % measurement, 0.1 % otherwise */
    double total = 0.0, mx = 0.0;
    for (size_t i = 0; i < circuit->num_gates; i++) {
        const gate_type_t t = circuit->gates[i].type;
        const double e = two_qubit_kind(t) ? 0.005 : (t == GATE_TYPE_MEASURE ? 0.01 : 0.001);
        total += e;
        if (e > mx) mx = e;
    }
    *max_error_rate = mx;
    *avg_error_rate = circuit->num_gates ? total / (double)circuit->num_gates : 0.0;
}

void configure_circuit_optimization(CPUSimCircuit* circuit, bool use_error_correction, bool use_tensor_networks, size_t cache_line_size) {
    (void)circuit; (void)use_error_correction; (void)use_tensor_networks; (void)cache_line_size;   /* no-op there too (:913-926) */
}

void simulate_circuit_cpu(double complex* state, const CPUSimCircuit* circuit, size_t n_qubits) {
    if (!state || !circuit || n_qubits == 0) return;
    qgt_b200_ctx* ctx = qgt_compat_ctx();
    if (!ctx) { fprintf(stderr, "simulate_circuit_cpu: %s\n", qgt_compat_last_error()); return; }
    qgt_b200_gate* gates = (qgt_b200_gate*)malloc((circuit->num_gates ? circuit->num_gates : 1) * sizeof *gates);
    if (!gates) return;
    size_t ng = 0;
    for (size_t i = 0; i < circuit->num_gates; i++) {
        const HardwareGate* h = &circuit->gates[i];
        if (qgt_compat_convert_gate(h->type, h->target, h->control, h->parameter, -1, &gates[ng])) { ng++; continue; }
        if (h->type == GATE_TYPE_BARRIER) continue;
        /* MEASURE / RESET / U2 / U3 / ISWAP / CCX / XX / YY / CUSTOM have no statevector-sweep form here: refuse the
         * whole circuit (state untouched) rather than return a plausible but wrong state */
        snprintf(g_err, sizeof g_err, "simulate_circuit_cpu: gate %zu has unsupported type %d; circuit not applied", i, (int)h->type);
        fprintf(stderr, "%s\n", g_err);
        free(gates);
        return;
    }
    qgt_b200_circuit c;
    memset(&c, 0, sizeof c);
    c.num_qubits = (int32_t)n_qubits; c.gates = gates; c.num_gates = ng;
    int rc = qgt_b200_simulate_host(ctx, (double*)state, (int)n_qubits, &c, NULL);
    if (rc) { qgt_compat_set_error("simulate_circuit_cpu", rc); fprintf(stderr, "%s\n", qgt_compat_last_error()); }
    free(gates);
}

/* ---- hardware/quantum_simulator.h ------------------------------------------------------------------------- */
struct qgt_sim_circuit {
    uint32_t num_qubits, num_classical_bits;
    qgt_b200_gate* gates;
    size_t num_gates, capacity;
};

SimulatorState* sim_init(uint32_t num_qubits, uint32_t num_classical_bits, const struct SimulatorConfig* config) {
    if (num_qubits == 0 || num_qubits > 32) return NULL;      /* MAX_QUBITS, quantum_simulator.h:16 */
    SimulatorState* s = (SimulatorState*)calloc(1, sizeof *s);
    if (!s) return NULL;
    s->num_qubits = num_qubits; s->num_classical_bits = num_classical_bits;
    s->amplitudes = (double complex*)calloc((size_t)1 << num_qubits, sizeof(double complex));
    if (!s->amplitudes) { free(s); return NULL; }
    s->amplitudes[0] = 1.0;
    if (num_classical_bits) s->classical_bits = (bool*)calloc(num_classical_bits, sizeof(bool));
    s->fidelity = 1.0;
    s->active_noise.type = NOISE_NONE;
    if (config && config->noise_model) {            /* [gate error, measurement error, decoherence rate] (quantum_simulator.c:100-109) */
        s->active_noise.gate_error_rate = config->noise_model[0];
        s->active_noise.measurement_error_rate = config->noise_model[1];
        s->active_noise.decoherence_rate = config->noise_model[2];
        if (config->noise_model[0] > 0) s->active_noise.type = NOISE_DEPOLARIZING;
    }
    return s;
}

void sim_reset_state(SimulatorState* s) {
    if (!s || !s->amplitudes) return;
    memset(s->amplitudes, 0, ((size_t)1 << s->num_qubits) * sizeof(double complex));
    s->amplitudes[0] = 1.0;
    if (s->classical_bits) memset(s->classical_bits, 0, s->num_classical_bits * sizeof(bool));
    s->fidelity = 1.0; s->error_rate = 0.0;
}

void sim_cleanup(SimulatorState* s) {
    if (!s) return;
    free(s->amplitudes); free(s->classical_bits); free(s->active_noise.custom_parameters); free(s->custom_state);
    free(s);
}

SimulatorCircuit* sim_create_circuit(uint32_t num_qubits, uint32_t num_classical_bits) {
    SimulatorCircuit* c = (SimulatorCircuit*)calloc(1, sizeof *c);
    if (!c) return NULL;
    c->num_qubits = num_qubits; c->num_classical_bits = num_classical_bits; c->capacity = 64;
    c->gates = (qgt_b200_gate*)calloc(c->capacity, sizeof *c->gates);
    if (!c->gates) { free(c); return NULL; }
    return c;
}

bool sim_add_gate(SimulatorCircuit* c, gate_type_t type, uint32_t target, uint32_t control, double* parameters) {
    if (!c) return false;
    if (c->num_gates >= c->capacity) {
        qgt_b200_gate* ng = (qgt_b200_gate*)realloc(c->gates, 2 * c->capacity * sizeof *ng);
        if (!ng) return false;
        c->gates = ng; c->capacity *= 2;
    }
    qgt_b200_gate g;
    const bool rot = type == GATE_TYPE_RX || type == GATE_TYPE_RY || type == GATE_TYPE_RZ || type == GATE_TYPE_U1 ||
                     type == GATE_TYPE_PHASE || type == GATE_TYPE_CRX || type == GATE_TYPE_CRY || type == GATE_TYPE_CRZ;
    if (rot && !parameters) return false;            /* apply_gate_by_type returns false there (:231,:241,:251) */
    if (!qgt_compat_convert_gate(type, target, control, parameters ? parameters[0] : 0.0, -1, &g)) return false;
    c->gates[c->num_gates++] = g;
    return true;
}

bool sim_add_controlled_gate(SimulatorCircuit* c, gate_type_t type, uint32_t target, uint32_t control, uint32_t control2, double* parameters) {
    (void)control2;
    return sim_add_gate(c, type, target, control, parameters);
}

static double compat_uniform(void);

bool sim_execute_circuit(SimulatorState* s, const SimulatorCircuit* c) {
    if (!s || !c) return false;
    qgt_b200_ctx* ctx = qgt_compat_ctx();
    if (!ctx) return false;
    qgt_b200_circuit qc;
    memset(&qc, 0, sizeof qc);
    qc.num_qubits = (int32_t)s->num_qubits; qc.gates = c->gates; qc.num_gates = c->num_gates;
    qgt_b200_gate* noisy = NULL;
    if (s->active_noise.type != NOISE_NONE && s->active_noise.gate_error_rate > 0 && c->num_gates) {
        /* depolarizing noise after every gate (quantum_simulator.c:290-311, 525-529): with probability gate_error_rate a
           Pauli drawn from {X, Y, Z, I} hits the gate's target.  The draws happen here, in the reference's order (one
           for the event, one for the Pauli), and the chosen Paulis join the gate list, so one noisy trajectory is still
           a single fused device circuit. */
        noisy = (qgt_b200_gate*)malloc(2 * c->num_gates * sizeof *noisy);
        if (!noisy) return false;
        size_t k = 0;
        for (size_t i = 0; i < c->num_gates; i++) {
            noisy[k++] = c->gates[i];
            if (compat_uniform() < s->active_noise.gate_error_rate) {
                const double pick = compat_uniform();
                const int kind = pick < 0.25 ? QGT_B200_GATE_X : pick < 0.5 ? QGT_B200_GATE_Y : pick < 0.75 ? QGT_B200_GATE_Z : -1;
                if (kind >= 0) {
                    qgt_b200_gate e = {kind, c->gates[i].target, -1, -1, 0.0, 1.0};
                    noisy[k++] = e;
                }
            }
        }
        qc.gates = noisy; qc.num_gates = k;
    }
    int rc = qgt_b200_simulate_host(ctx, (double*)s->amplitudes, (int)s->num_qubits, &qc, NULL);
    free(noisy);
    if (rc) { qgt_compat_set_error("sim_execute_circuit", rc); return false; }
    return true;
}

double complex* sim_get_statevector(const SimulatorState* s) {
    if (!s) return NULL;
    const size_t dim = (size_t)1 << s->num_qubits;
    double complex* sv = (double complex*)malloc(dim * sizeof *sv);
    if (sv) memcpy(sv, s->amplitudes, dim * sizeof *sv);
    return sv;                                      /* caller frees, as in the reference */
}

void sim_cleanup_circuit(SimulatorCircuit* c) {
    if (!c) return;
    free(c->gates);
    free(c);
}

/* ---- measurement, expectation values, sampling (quantum_simulator.c:563-729) -----------------------------------
 * The SimulatorState keeps its amplitudes on the host (callers read state->amplitudes); every call stages them
 * through the device, where the reductions and the collapse run (qgt_b200_state_*). */
static unsigned long long g_rng = 0;
static double compat_uniform(void) {                 /* same LCG as the reference's random_double (:47-51) */
    if (!g_rng) g_rng = 88172645463325252ull;
    g_rng = g_rng * 1103515245ull + 12345ull;
    return (double)(g_rng & 0x7fffffff) / (double)0x7fffffff;
}
void qgt_compat_seed(unsigned long long seed) { g_rng = seed ? seed : 1; }

static qgt_b200_state* stage_in(const SimulatorState* s, qgt_b200_ctx** ctx_out) {
    qgt_b200_ctx* ctx = qgt_compat_ctx();
    if (!ctx) return NULL;
    qgt_b200_state* st = NULL;
    int rc = qgt_b200_state_create(ctx, (int)s->num_qubits, &st);
    if (!rc) rc = qgt_b200_state_upload(st, (const double*)s->amplitudes);
    if (rc) { qgt_compat_set_error("state staging", rc); if (st) qgt_b200_state_destroy(st); return NULL; }
    if (ctx_out) *ctx_out = ctx;
    return st;
}

bool sim_measure_qubit(SimulatorState* s, uint32_t qubit, uint32_t classical_bit) {
    if (!s || qubit >= s->num_qubits) return false;
    if (classical_bit >= s->num_classical_bits && s->classical_bits) return false;
    qgt_b200_state* st = stage_in(s, NULL);
    if (!st) return false;
    int outcome = 0;
    int rc = qgt_b200_state_measure(st, (int)qubit, compat_uniform(), s->active_noise.measurement_error_rate, &outcome, NULL);
    if (!rc) rc = qgt_b200_state_download(st, (double*)s->amplitudes);
    qgt_b200_state_destroy(st);
    if (rc) { qgt_compat_set_error("sim_measure_qubit", rc); return false; }
    if (s->classical_bits && classical_bit < s->num_classical_bits) s->classical_bits[classical_bit] = outcome != 0;
    return true;
}

bool sim_measure_all(SimulatorState* s) {
    if (!s) return false;
    qgt_b200_state* st = stage_in(s, NULL);
    if (!st) return false;
    int rc = 0;
    for (uint32_t q = 0; q < s->num_qubits && !rc; q++) {          /* one staging for all qubits */
        int outcome = 0;
        rc = qgt_b200_state_measure(st, (int)q, compat_uniform(), s->active_noise.measurement_error_rate, &outcome, NULL);
        const uint32_t cb = q < s->num_classical_bits ? q : 0;
        if (!rc && s->classical_bits && cb < s->num_classical_bits) s->classical_bits[cb] = outcome != 0;
    }
    if (!rc) rc = qgt_b200_state_download(st, (double*)s->amplitudes);
    qgt_b200_state_destroy(st);
    if (rc) { qgt_compat_set_error("sim_measure_all", rc); return false; }
    return true;
}

bool* sim_get_measurement_results(const SimulatorState* s) {
    if (!s || !s->classical_bits) return NULL;
    bool* r = (bool*)malloc(s->num_classical_bits * sizeof(bool));
    if (r) memcpy(r, s->classical_bits, s->num_classical_bits * sizeof(bool));
    return r;                                        /* caller frees */
}

uint64_t* sim_get_measurement_counts(const SimulatorState* s, uint32_t shots) {
    if (!s) return NULL;
    const size_t dim = (size_t)1 << s->num_qubits;
    uint64_t* counts = (uint64_t*)calloc(dim, sizeof(uint64_t));
    double* u = (double*)malloc((shots ? shots : 1) * sizeof(double));
    uint64_t* idx = (uint64_t*)malloc((shots ? shots : 1) * sizeof(uint64_t));
    qgt_b200_state* st = (counts && u && idx) ? stage_in(s, NULL) : NULL;
    int rc = st ? 0 : -1;
    if (st) {
        for (uint32_t k = 0; k < shots; k++) u[k] = compat_uniform();
        rc = qgt_b200_state_sample(st, u, shots, idx);
        for (uint32_t k = 0; k < shots && !rc; k++) if (idx[k] < dim) counts[idx[k]]++;
        qgt_b200_state_destroy(st);
    }
    free(u); free(idx);
    if (rc) { free(counts); return NULL; }
    return counts;                                   /* dim entries, caller frees */
}

double sim_get_expectation_value(const SimulatorState* s, const char* observable) {
    if (!s || !observable) return 0.0;
    if (strcmp(observable, "Z") != 0) return 0.0;    /* the only observable the reference knows (:708) */
    qgt_b200_state* st = stage_in(s, NULL);
    if (!st) return 0.0;
    double v = 0.0;
    const uint64_t all = s->num_qubits >= 64 ? ~0ull : (((uint64_t)1 << s->num_qubits) - 1);
    int rc = qgt_b200_state_expectation_z(st, all, &v);
    qgt_b200_state_destroy(st);
    if (rc) { qgt_compat_set_error("sim_get_expectation_value", rc); return 0.0; }
    return v;
}

/* ---- circuit text format (quantum_simulator.c:962-1051): "QGT_CIRCUIT v1", qubits / classical_bits / gates, then one
 * line per gate "type target control [p0 p1 p2 p3]" ------------------------------------------------------------ */
bool sim_save_circuit(const SimulatorCircuit* c, const char* filename) {
    if (!c || !filename) return false;
    FILE* f = fopen(filename, "w");
    if (!f) return false;
    fprintf(f, "QGT_CIRCUIT v1\nqubits %u\nclassical_bits %u\ngates %zu\n", c->num_qubits, c->num_classical_bits, c->num_gates);
    for (size_t i = 0; i < c->num_gates; i++) {
        const qgt_b200_gate* g = &c->gates[i];
        fprintf(f, "%d %d %d", (int)g->kind, (int)g->target, g->control < 0 ? 0 : (int)g->control);
        const int rot = g->kind == GATE_TYPE_RX || g->kind == GATE_TYPE_RY || g->kind == GATE_TYPE_RZ || g->kind == GATE_TYPE_U1 ||
                        g->kind == GATE_TYPE_PHASE || g->kind == GATE_TYPE_CRX || g->kind == GATE_TYPE_CRY || g->kind == GATE_TYPE_CRZ ||
                        g->kind == GATE_TYPE_ZZ;
        if (rot) fprintf(f, " %.17g 0 0 0", g->angle);       /* full precision (the reference prints %g) */
        fprintf(f, "\n");
    }
    fclose(f);
    return true;
}

SimulatorCircuit* sim_load_circuit(const char* filename) {
    if (!filename) return NULL;
    FILE* f = fopen(filename, "r");
    if (!f) return NULL;
    char header[32];
    if (fscanf(f, "%31s", header) != 1 || strcmp(header, "QGT_CIRCUIT") != 0) { fclose(f); return NULL; }
    if (fscanf(f, "%*s") != 0) { /* version */ }
    uint32_t nq = 0, ncb = 0;
    size_t ng = 0;
    if (fscanf(f, " qubits %u", &nq) != 1 || fscanf(f, " classical_bits %u", &ncb) != 1 || fscanf(f, " gates %zu", &ng) != 1) { fclose(f); return NULL; }
    SimulatorCircuit* c = sim_create_circuit(nq, ncb);
    if (!c) { fclose(f); return NULL; }
    /* one gate per LINE.  (The reference's loader reads "type target control" and then greedily tries four more numbers
     * with fscanf, which swallows the next gate's line after every parameter-free gate: quantum_simulator.c:1033-1041,
     * BASELINE.md section 4 #18.  Not reproduced.) */
    char line[256];
    size_t got = 0;
    while (got < ng && fgets(line, sizeof line, f)) {
        int type;
        uint32_t target, control;
        double params[4] = {0, 0, 0, 0};
        const int nr = sscanf(line, "%d %u %u %lf %lf %lf %lf", &type, &target, &control, &params[0], &params[1], &params[2], &params[3]);
        if (nr < 3) continue;                        /* blank line (e.g. the rest of the header line) */
        sim_add_gate(c, (gate_type_t)type, target, control, nr > 3 ? params : NULL);
        got++;
    }
    fclose(f);
    return c;
}

/*
 * qaoa_compat.c — the reference's QAOA driver (include/quantum_geometric/algorithms/qaoa.h, src/.../algorithms/qaoa.c)
 * on the device library: graph helpers :20-118, qaoa_init :139-232, layers :344-449, expectation :455-487, gradient
 * :489-558, optimisation loop :560-667, sampling :669-718.
 *
 * The state lives on the GPU as complex double; one QAOA circuit (p cost layers exp(-i gamma E_z) + p mixer layers
 * RX(2 beta) on every qubit) is a single qgt_b200_apply_circuit call (fused sweeps; E_z comes from per-tile energy tables
 * built from the edge list, so the 2^n-entry Hamiltonian table of qaoa.c:236-294 is never materialised:
 * cost_hamiltonian->matrix stays NULL).  state->qstate->amplitudes is refreshed (ComplexFloat, as the reference keeps it)
 * by the public qaoa_apply_circuit / qaoa_apply_layer / qaoa_prepare_initial_state calls; the gradient and optimisation
 * loops run device-resident and refresh it once at the end.
 *
 * Kept from the reference on purpose: qaoa_compute_gradient is its finite-shift formula [E(+pi/2) - E(-pi/2)] / 2 per
 * parameter (4p circuits; exact only for generators with spectrum +-1/2, which neither QAOA layer has).  The exact gradient
 * (adjoint method, one forward + one backward pass) is qgt_b200_qaoa_exact_gradient.  Not kept: best_cost starts at
 * +INFINITY in the reference, so `exp_val > best_cost` never fires and best_gamma/best_beta stay zero;
 * here it starts at -INFINITY and the best parameters are tracked.
 */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

#include "compat_common.h"

typedef struct {
    qgt_b200_state* dev;       /* |psi> on the device */
    qgt_b200_edge* edges;      /* the graph in the device library's terms */
    size_t num_edges;
    bool host_stale;
} qaoa_private;

/* ---- graphs (host) ------------------------------------------------------------------------------------------------ */
qaoa_graph_t* qaoa_create_graph(size_t num_vertices) {
    if (num_vertices == 0) return NULL;
    qaoa_graph_t* g = (qaoa_graph_t*)calloc(1, sizeof *g);
    if (!g) return NULL;
    g->num_vertices = num_vertices;
    g->vertex_weights = (double*)calloc(num_vertices, sizeof(double));
    if (!g->vertex_weights) { free(g); return NULL; }
    return g;
}

qgt_error_t qaoa_add_edge(qaoa_graph_t* g, size_t i, size_t j, double weight) {
    if (!g || i >= g->num_vertices || j >= g->num_vertices || i == j) return QGT_ERROR_INVALID_ARGUMENT;
    qaoa_edge_t* ne = (qaoa_edge_t*)realloc(g->edges, (g->num_edges + 1) * sizeof *ne);
    if (!ne) return QGT_ERROR_MEMORY_ALLOCATION;
    g->edges = ne;
    g->edges[g->num_edges].i = i; g->edges[g->num_edges].j = j; g->edges[g->num_edges].weight = weight;
    g->num_edges++;
    return QGT_SUCCESS;
}

qgt_error_t qaoa_set_vertex_weight(qaoa_graph_t* g, size_t vertex, double weight) {
    if (!g || vertex >= g->num_vertices) return QGT_ERROR_INVALID_ARGUMENT;
    g->vertex_weights[vertex] = weight;
    return QGT_SUCCESS;
}

qaoa_graph_t* qaoa_create_random_graph(size_t num_vertices, double edge_probability, double min_weight, double max_weight) {
    qaoa_graph_t* g = qaoa_create_graph(num_vertices);
    if (!g) return NULL;
    srand((unsigned int)time(NULL));                     /* as the reference: seeded from the clock */
    for (size_t i = 0; i < num_vertices; i++)
        for (size_t j = i + 1; j < num_vertices; j++)
            if ((double)rand() / RAND_MAX < edge_probability)
                qaoa_add_edge(g, i, j, min_weight + (max_weight - min_weight) * ((double)rand() / RAND_MAX));
    return g;
}

qaoa_graph_t* qaoa_create_from_adjacency(const double* adjacency, size_t n) {
    if (!adjacency || n == 0) return NULL;
    qaoa_graph_t* g = qaoa_create_graph(n);
    if (!g) return NULL;
    for (size_t i = 0; i < n; i++)
        for (size_t j = i + 1; j < n; j++)
            if (adjacency[i * n + j] != 0.0) qaoa_add_edge(g, i, j, adjacency[i * n + j]);
    return g;
}

void qaoa_destroy_graph(qaoa_graph_t* g) {
    if (!g) return;
    free(g->edges); free(g->vertex_weights); free(g);
}

qaoa_config_t qaoa_default_config(size_t p) {
    qaoa_config_t c;
    memset(&c, 0, sizeof c);
    c.p = p; c.problem_type = QAOA_PROBLEM_MAXCUT; c.mixer_type = QAOA_MIXER_X; c.optimizer_type = QAOA_OPTIMIZER_COBYLA;
    c.max_iterations = 1000; c.tolerance = 1e-6; c.learning_rate = 0.01; c.num_shots = 1024; c.use_expectation = true;
    return c;
}

/* ---- state ---------------------------------------------------------------------------------------------------------- */
static qaoa_private* priv(const qaoa_state_t* s) {
    return s && s->cost_hamiltonian ? (qaoa_private*)s->cost_hamiltonian->device_data : NULL;
}

void qaoa_destroy(qaoa_state_t* s) {
    if (!s) return;
    qaoa_private* pv = priv(s);
    if (pv) { qgt_b200_state_destroy(pv->dev); free(pv->edges); free(pv); }
    qaoa_destroy_graph(s->graph);
    if (s->cost_hamiltonian) { free(s->cost_hamiltonian->matrix); free(s->cost_hamiltonian); }
    free(s->mixer_hamiltonian);
    if (s->qstate) { free(s->qstate->amplitudes); free(s->qstate->workspace); free(s->qstate); }
    free(s->gamma); free(s->beta); free(s->best_gamma); free(s->best_beta); free(s->best_solution); free(s->solution_probabilities);
    free(s);
}

qgt_error_t qaoa_construct_cost_hamiltonian(qaoa_state_t* s) {
    if (!s || !s->graph) return QGT_ERROR_INVALID_ARGUMENT;
    if (s->cost_hamiltonian) return QGT_SUCCESS;
    quantum_operator_t* h = (quantum_operator_t*)calloc(1, sizeof *h);
    if (!h) return QGT_ERROR_MEMORY_ALLOCATION;
    h->type = QUANTUM_OPERATOR_HERMITIAN; h->dimension = (size_t)1 << s->num_qubits; h->is_hermitian = true;
    h->matrix = NULL;                       /* diagonal E_z is evaluated on the device from the edge list */
    h->device_type = HARDWARE_TYPE_CUDA;
    s->cost_hamiltonian = h;
    return QGT_SUCCESS;
}

qgt_error_t qaoa_construct_mixer_hamiltonian(qaoa_state_t* s) {
    if (!s) return QGT_ERROR_INVALID_ARGUMENT;
    if (s->mixer_hamiltonian) return QGT_SUCCESS;
    quantum_operator_t* h = (quantum_operator_t*)calloc(1, sizeof *h);
    if (!h) return QGT_ERROR_MEMORY_ALLOCATION;
    h->type = QUANTUM_OPERATOR_HERMITIAN; h->dimension = (size_t)1 << s->num_qubits; h->is_hermitian = true;
    s->mixer_hamiltonian = h;               /* applied as RX(2 beta) on every qubit, no matrix (qaoa.c:296-312) */
    return QGT_SUCCESS;
}

qaoa_state_t* qaoa_init(const qaoa_graph_t* graph, const qaoa_config_t* config) {
    if (!graph || !config || config->p == 0 || graph->num_vertices == 0 || graph->num_vertices > 40) return NULL;
    qaoa_state_t* s = (qaoa_state_t*)calloc(1, sizeof *s);
    if (!s) return NULL;
    s->config = *config; s->p = config->p; s->num_qubits = graph->num_vertices;
    s->graph = qaoa_create_graph(graph->num_vertices);
    if (!s->graph) { free(s); return NULL; }
    for (size_t e = 0; e < graph->num_edges; e++)
        if (qaoa_add_edge(s->graph, graph->edges[e].i, graph->edges[e].j, graph->edges[e].weight) != QGT_SUCCESS) { qaoa_destroy(s); return NULL; }
    if (graph->vertex_weights) memcpy(s->graph->vertex_weights, graph->vertex_weights, graph->num_vertices * sizeof(double));
    s->gamma = (double*)calloc(s->p, sizeof(double)); s->beta = (double*)calloc(s->p, sizeof(double));
    s->best_gamma = (double*)calloc(s->p, sizeof(double)); s->best_beta = (double*)calloc(s->p, sizeof(double));
    s->best_solution = (int*)calloc(s->num_qubits, sizeof(int));
    s->qstate = (QuantumState*)calloc(1, sizeof(QuantumState));
    if (!s->gamma || !s->beta || !s->best_gamma || !s->best_beta || !s->best_solution || !s->qstate) { qaoa_destroy(s); return NULL; }
    if (config->initial_gamma && config->initial_beta) {
        memcpy(s->gamma, config->initial_gamma, s->p * sizeof(double));
        memcpy(s->beta, config->initial_beta, s->p * sizeof(double));
    } else {
        srand((unsigned int)time(NULL));
        for (size_t i = 0; i < s->p; i++) {
            s->gamma[i] = 2.0 * M_PI * ((double)rand() / RAND_MAX);
            s->beta[i] = M_PI * ((double)rand() / RAND_MAX);
        }
    }
    const size_t dim = (size_t)1 << s->num_qubits;
    s->qstate->num_qubits = s->num_qubits; s->qstate->dimension = dim;
    s->qstate->amplitudes = (ComplexFloat*)calloc(dim, sizeof(ComplexFloat));
    if (!s->qstate->amplitudes) { qaoa_destroy(s); return NULL; }
    /* solution_probabilities (2^n doubles in the reference, never written there) is left NULL */
    if (qaoa_construct_cost_hamiltonian(s) != QGT_SUCCESS || qaoa_construct_mixer_hamiltonian(s) != QGT_SUCCESS) { qaoa_destroy(s); return NULL; }
    qaoa_private* pv = (qaoa_private*)calloc(1, sizeof *pv);
    if (!pv) { qaoa_destroy(s); return NULL; }
    s->cost_hamiltonian->device_data = pv;
    pv->num_edges = s->graph->num_edges;
    pv->edges = (qgt_b200_edge*)calloc(pv->num_edges ? pv->num_edges : 1, sizeof(qgt_b200_edge));
    if (!pv->edges) { qaoa_destroy(s); return NULL; }
    for (size_t e = 0; e < pv->num_edges; e++) {
        pv->edges[e].i = (int32_t)s->graph->edges[e].i; pv->edges[e].j = (int32_t)s->graph->edges[e].j; pv->edges[e].weight = s->graph->edges[e].weight;
    }
    qgt_b200_ctx* ctx = qgt_compat_ctx();
    if (!ctx || qgt_b200_state_create(ctx, (int)s->num_qubits, &pv->dev) != QGT_B200_OK) {
        fprintf(stderr, "qaoa_init: %s\n", ctx ? qgt_b200_last_error() : qgt_compat_last_error());      /* no CPU fallback */
        qaoa_destroy(s);
        return NULL;
    }
    s->best_cost = -INFINITY;
    return s;
}

/* layers [first, last) of the QAOA circuit as one device circuit; parameters (gamma_1, beta_1, ..., gamma_p, beta_p) */
static void fill_circuit(const qaoa_state_t* s, size_t first, size_t last, qgt_b200_gate* gates, qgt_b200_circuit* c) {
    const size_t n = s->num_qubits;
    size_t k = 0;
    for (size_t l = first; l < last; l++) {
        qgt_b200_gate cg = {QGT_B200_GATE_COST, 0, -1, (int32_t)(2 * l), 0.0, 1.0};
        gates[k++] = cg;
        for (size_t q = 0; q < n; q++) {
            qgt_b200_gate rx = {QGT_B200_GATE_RX, (int32_t)q, -1, (int32_t)(2 * l + 1), 0.0, 2.0};
            gates[k++] = rx;
        }
    }
    const qaoa_private* pv = priv(s);
    memset(c, 0, sizeof *c);
    c->num_qubits = (int32_t)n; c->num_params = (int32_t)(2 * s->p); c->gates = gates; c->num_gates = k;
    c->edges = pv->edges; c->num_edges = pv->num_edges; c->vertex_weights = s->graph->vertex_weights;
    c->initial_state = QGT_B200_INIT_PLUS;
}

static int run_layers(qaoa_state_t* s, const double* gamma, const double* beta, size_t first, size_t last, bool init) {
    qaoa_private* pv = priv(s);
    if (!pv || !pv->dev) return QGT_B200_ERR_NOT_INIT;
    qgt_b200_gate* gates = (qgt_b200_gate*)malloc(((last - first) * (s->num_qubits + 1) + 1) * sizeof *gates);
    double* theta = (double*)malloc(2 * s->p * sizeof(double));
    if (!gates || !theta) { free(gates); free(theta); return QGT_B200_ERR_NO_MEMORY; }
    for (size_t l = 0; l < s->p; l++) { theta[2 * l] = gamma[l]; theta[2 * l + 1] = beta[l]; }
    qgt_b200_circuit c;
    fill_circuit(s, first, last, gates, &c);
    int rc = init ? qgt_b200_state_init(pv->dev, QGT_B200_INIT_PLUS) : QGT_B200_OK;
    if (!rc && c.num_gates) rc = qgt_b200_apply_circuit(pv->dev, &c, theta);
    pv->host_stale = true;
    free(gates); free(theta);
    if (rc) qgt_compat_set_error("qaoa circuit", rc);
    return rc;
}

static qgt_error_t sync_host(qaoa_state_t* s) {
    qaoa_private* pv = priv(s);
    if (!pv || !pv->host_stale) return QGT_SUCCESS;
    int rc = qgt_b200_state_download_c64(pv->dev, (float*)s->qstate->amplitudes);
    if (rc) { qgt_compat_set_error("qaoa state download", rc); return QGT_ERROR_HARDWARE_FAILURE; }
    pv->host_stale = false;
    s->qstate->is_normalized = true;
    return QGT_SUCCESS;
}

qgt_error_t qaoa_prepare_initial_state(qaoa_state_t* s) {
    if (!s || !s->qstate || !priv(s)) return QGT_ERROR_INVALID_ARGUMENT;
    int rc = qgt_b200_state_init(priv(s)->dev, QGT_B200_INIT_PLUS);
    if (rc) return QGT_ERROR_HARDWARE_FAILURE;
    priv(s)->host_stale = true;
    return sync_host(s);
}

qgt_error_t qaoa_apply_layer(qaoa_state_t* s, size_t layer_idx) {
    if (!s || layer_idx >= s->p || !priv(s)) return QGT_ERROR_INVALID_ARGUMENT;
    if (run_layers(s, s->gamma, s->beta, layer_idx, layer_idx + 1, false)) return QGT_ERROR_HARDWARE_FAILURE;
    return sync_host(s);
}

qgt_error_t qaoa_apply_circuit(qaoa_state_t* s, const double* gamma, const double* beta) {
    if (!s || !gamma || !beta || !priv(s)) return QGT_ERROR_INVALID_ARGUMENT;
    if (gamma != s->gamma) memcpy(s->gamma, gamma, s->p * sizeof(double));
    if (beta != s->beta) memcpy(s->beta, beta, s->p * sizeof(double));
    if (run_layers(s, s->gamma, s->beta, 0, s->p, true)) return QGT_ERROR_HARDWARE_FAILURE;
    return sync_host(s);
}

/* <psi|H_C|psi> of the device state */
static int device_expectation(qaoa_state_t* s, double* out) {
    qaoa_private* pv = priv(s);
    qgt_b200_gate none;
    qgt_b200_circuit c;
    fill_circuit(s, 0, 0, &none, &c);
    return qgt_b200_state_cost_expectation(pv->dev, &c, out);
}

static void note_cost(qaoa_state_t* s, double v) {
    s->current_cost = v;
    if (v > s->best_cost) {
        s->best_cost = v;
        memcpy(s->best_gamma, s->gamma, s->p * sizeof(double));
        memcpy(s->best_beta, s->beta, s->p * sizeof(double));
    }
}

qgt_error_t qaoa_compute_expectation(qaoa_state_t* s, double* expectation) {
    if (!s || !s->qstate || !s->cost_hamiltonian || !expectation || !priv(s)) return QGT_ERROR_INVALID_ARGUMENT;
    double v = 0.0;
    if (device_expectation(s, &v)) return QGT_ERROR_HARDWARE_FAILURE;
    *expectation = v;
    note_cost(s, v);
    return QGT_SUCCESS;
}

/* the reference's formula (qaoa.c:489-558): [E(theta_k + pi/2) - E(theta_k - pi/2)] / 2, every circuit on the device */
qgt_error_t qaoa_compute_gradient(qaoa_state_t* s, double* gamma_grad, double* beta_grad) {
    if (!s || !gamma_grad || !beta_grad || !priv(s)) return QGT_ERROR_INVALID_ARGUMENT;
    const double SHIFT = M_PI / 2.0;
    double* g = (double*)malloc(s->p * sizeof(double));
    double* b = (double*)malloc(s->p * sizeof(double));
    if (!g || !b) { free(g); free(b); return QGT_ERROR_MEMORY_ALLOCATION; }
    memcpy(g, s->gamma, s->p * sizeof(double)); memcpy(b, s->beta, s->p * sizeof(double));
    int rc = 0;
    for (int which = 0; which < 2 && !rc; which++) {
        double* v = which ? b : g;
        const double* base = which ? s->beta : s->gamma;
        double* out = which ? beta_grad : gamma_grad;
        for (size_t i = 0; i < s->p && !rc; i++) {
            double ep = 0.0, em = 0.0;
            v[i] = base[i] + SHIFT;
            rc = run_layers(s, g, b, 0, s->p, true);
            if (!rc) rc = device_expectation(s, &ep);
            v[i] = base[i] - SHIFT;
            if (!rc) rc = run_layers(s, g, b, 0, s->p, true);
            if (!rc) rc = device_expectation(s, &em);
            out[i] = (ep - em) / 2.0;
            v[i] = base[i];
            if (!rc) { s->current_cost = em; }          /* the reference leaves the last evaluated expectation here */
        }
    }
    free(g); free(b);
    if (rc) return QGT_ERROR_HARDWARE_FAILURE;
    return sync_host(s);                                 /* qstate holds the last shifted circuit, as in the reference */
}

/* exact dE/dgamma_l, dE/dbeta_l by the adjoint method (one forward circuit + one backward pass): not in the reference */
qgt_error_t qgt_b200_qaoa_exact_gradient(qaoa_state_t* s, double* energy, double* gamma_grad, double* beta_grad) {
    if (!s || !gamma_grad || !beta_grad || !priv(s)) return QGT_ERROR_INVALID_ARGUMENT;
    qgt_b200_ctx* ctx = qgt_compat_ctx();
    if (!ctx) return QGT_ERROR_HARDWARE_FAILURE;
    qgt_b200_gate* gates = (qgt_b200_gate*)malloc((s->p * (s->num_qubits + 1) + 1) * sizeof *gates);
    double* theta = (double*)malloc(4 * s->p * sizeof(double));
    if (!gates || !theta) { free(gates); free(theta); return QGT_ERROR_MEMORY_ALLOCATION; }
    double* grad = theta + 2 * s->p;
    for (size_t l = 0; l < s->p; l++) { theta[2 * l] = s->gamma[l]; theta[2 * l + 1] = s->beta[l]; }
    qgt_b200_circuit c;
    fill_circuit(s, 0, s->p, gates, &c);
    double e = 0.0;
    int rc = qgt_b200_expectation_gradient(ctx, &c, theta, &e, grad);
    if (!rc) for (size_t l = 0; l < s->p; l++) { gamma_grad[l] = grad[2 * l]; beta_grad[l] = grad[2 * l + 1]; }
    if (!rc && energy) *energy = e;
    if (rc) qgt_compat_set_error("qgt_b200_expectation_gradient", rc);
    free(gates); free(theta);
    return rc ? QGT_ERROR_HARDWARE_FAILURE : QGT_SUCCESS;
}

static void bits_of(uint64_t z, size_t n, int* out) { for (size_t i = 0; i < n; i++) out[i] = (int)((z >> i) & 1); }

qaoa_result_t* qaoa_optimize(qaoa_state_t* s) {
    if (!s || !priv(s)) return NULL;
    qaoa_result_t* r = (qaoa_result_t*)calloc(1, sizeof *r);
    if (!r) return NULL;
    const size_t iters = s->config.max_iterations;
    r->optimal_gamma = (double*)calloc(s->p, sizeof(double)); r->optimal_beta = (double*)calloc(s->p, sizeof(double));
    r->optimal_solution = (int*)calloc(s->num_qubits, sizeof(int));
    r->cost_history = (double*)calloc(iters ? iters : 1, sizeof(double));
    double* gg = (double*)malloc(s->p * sizeof(double));
    double* bg = (double*)malloc(s->p * sizeof(double));
    if (!r->optimal_gamma || !r->optimal_beta || !r->optimal_solution || !r->cost_history || !gg || !bg) { free(gg); free(bg); qaoa_destroy_result(r); return NULL; }
    const clock_t t0 = clock();
    double prev = -INFINITY;
    const double lr = s->config.learning_rate;
    for (size_t it = 0; it < iters; it++) {           /* gradient ascent on the cut value (qaoa.c:595-628) */
        s->iteration = it;
        if (run_layers(s, s->gamma, s->beta, 0, s->p, true)) break;
        double cost = 0.0;
        if (device_expectation(s, &cost)) break;
        note_cost(s, cost);
        r->cost_history[it] = cost; r->history_length = it + 1;
        if (fabs(cost - prev) < s->config.tolerance) break;
        prev = cost;
        if (qaoa_compute_gradient(s, gg, bg) != QGT_SUCCESS) break;
        for (size_t i = 0; i < s->p; i++) { s->gamma[i] += lr * gg[i]; s->beta[i] += lr * bg[i]; }
    }
    free(gg); free(bg);
    memcpy(r->optimal_gamma, s->best_gamma, s->p * sizeof(double));
    memcpy(r->optimal_beta, s->best_beta, s->p * sizeof(double));
    r->optimal_cost = s->best_cost;
    r->num_iterations = s->iteration + 1;
    /* most probable bit string of the best circuit */
    if (!run_layers(s, s->best_gamma, s->best_beta, 0, s->p, true)) {
        uint64_t best_z = 0;
        double pmax = 0.0;
        if (!qgt_b200_state_argmax(priv(s)->dev, &best_z, &pmax)) {
            bits_of(best_z, s->num_qubits, r->optimal_solution);
            memcpy(s->best_solution, r->optimal_solution, s->num_qubits * sizeof(int));
        }
        sync_host(s);
    }
    r->execution_time = (double)(clock() - t0) / CLOCKS_PER_SEC;
    return r;
}

qgt_error_t qaoa_sample(qaoa_state_t* s, int** samples, size_t num_samples) {
    if (!s || !samples || num_samples == 0 || !priv(s)) return QGT_ERROR_INVALID_ARGUMENT;
    double* u = (double*)malloc(num_samples * sizeof(double));
    uint64_t* z = (uint64_t*)malloc(num_samples * sizeof(uint64_t));
    if (!u || !z) { free(u); free(z); return QGT_ERROR_MEMORY_ALLOCATION; }
    for (size_t k = 0; k < num_samples; k++) u[k] = (double)rand() / RAND_MAX;      /* the reference's generator (:695) */
    int rc = qgt_b200_state_sample(priv(s)->dev, u, num_samples, z);                  /* inverse-CDF walk on the device */
    if (rc) { qgt_compat_set_error("qaoa_sample", rc); free(u); free(z); return QGT_ERROR_HARDWARE_FAILURE; }
    for (size_t k = 0; k < num_samples; k++) {
        samples[k] = (int*)malloc(s->num_qubits * sizeof(int));
        if (!samples[k]) { for (size_t i = 0; i < k; i++) free(samples[i]); free(u); free(z); return QGT_ERROR_MEMORY_ALLOCATION; }
        bits_of(z[k], s->num_qubits, samples[k]);
    }
    free(u); free(z);
    return QGT_SUCCESS;
}

double qaoa_evaluate_solution(const qaoa_graph_t* g, const int* solution) {
    if (!g || !solution) return 0.0;
    double cost = 0.0;
    for (size_t e = 0; e < g->num_edges; e++)
        if (solution[g->edges[e].i] != solution[g->edges[e].j]) cost += g->edges[e].weight;
    return cost;
}

void qaoa_destroy_result(qaoa_result_t* r) {
    if (!r) return;
    free(r->optimal_gamma); free(r->optimal_beta); free(r->optimal_solution); free(r->cost_history); free(r);
}

size_t qaoa_estimate_optimal_p(size_t num_vertices, size_t num_edges) {
    if (num_vertices < 2) return 1;
    const double density = (double)num_edges / (double)(num_vertices * (num_vertices - 1) / 2);
    const size_t base_p = (size_t)(log2((double)num_vertices) + 1);
    return base_p + (size_t)(density * (double)base_p);
}

double qaoa_approximation_ratio(const qaoa_graph_t* g, const int* solution, double optimal_cost) {
    if (!g || !solution || optimal_cost == 0.0) return 0.0;
    return qaoa_evaluate_solution(g, solution) / optimal_cost;
}

void qaoa_print_state(const qaoa_state_t* s) {
    if (!s) return;
    printf("QAOA State:\n  Qubits: %zu\n  Layers (p): %zu\n  Edges: %zu\n  Current cost: %.6f\n  Best cost: %.6f\n  Iteration: %zu\n",
           s->num_qubits, s->p, s->graph ? s->graph->num_edges : (size_t)0, s->current_cost, s->best_cost, s->iteration);
}

void qaoa_print_result(const qaoa_result_t* r) {
    if (!r) return;
    printf("QAOA Result:\n  Optimal cost: %.6f\n  Iterations: %zu\n  Execution time: %.3f seconds\n  Approximation ratio: %.4f\n",
           r->optimal_cost, r->num_iterations, r->execution_time, r->approximation_ratio);
}

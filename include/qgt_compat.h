/*
 * qgt_compat.h — the reference's C API for the QGT hot path, served by the sm_100a library.
 *
 * One consolidated header; the files under include/quantum_geometric/ only forward to it so that
 * reference callers keep their #include lines.  Type layouts follow the reference headers they replace
 * (cited per block) so objects can cross the boundary unchanged; every function below is a thin host-C
 * wrapper (quantum_geometric_tensor_b200/csrc/compat/) over include/qgt_b200.h.  There is no CPU
 * fallback: without an sm_100 device the compute entry points report an error (and, where the reference
 * signature is `void`, print it to stderr and leave their outputs untouched).
 */
#ifndef QGT_COMPAT_H
#define QGT_COMPAT_H

#include <complex.h>
#include <stdbool.h>
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ---- core/quantum_base_types.h:32-72, 118-141, 187-198 ---------------------------------------- */
typedef enum {
    GATE_TYPE_I = 0, GATE_TYPE_X = 1, GATE_TYPE_Y = 2, GATE_TYPE_Z = 3, GATE_TYPE_H = 4, GATE_TYPE_S = 5,
    GATE_TYPE_T = 6, GATE_TYPE_RX = 7, GATE_TYPE_RY = 8, GATE_TYPE_RZ = 9, GATE_TYPE_CNOT = 10,
    GATE_TYPE_CY = 11, GATE_TYPE_CZ = 12, GATE_TYPE_SWAP = 13, GATE_TYPE_CUSTOM = 14, GATE_TYPE_U1 = 15,
    GATE_TYPE_U2 = 16, GATE_TYPE_U3 = 17, GATE_TYPE_CCX = 18, GATE_TYPE_PHASE = 19, GATE_TYPE_CSWAP = 20,
    GATE_TYPE_ISWAP = 21, GATE_TYPE_CRX = 22, GATE_TYPE_CRY = 23, GATE_TYPE_CRZ = 24, GATE_TYPE_CH = 25,
    GATE_TYPE_SDG = 26, GATE_TYPE_TDG = 27, GATE_TYPE_ECR = 28, GATE_TYPE_SX = 29, GATE_TYPE_MEASURE = 30,
    GATE_TYPE_RESET = 31, GATE_TYPE_BARRIER = 32, GATE_TYPE_XX = 33, GATE_TYPE_YY = 34, GATE_TYPE_ZZ = 35
} gate_type_t;
#define GATE_I GATE_TYPE_I
#define GATE_X GATE_TYPE_X
#define GATE_Y GATE_TYPE_Y
#define GATE_Z GATE_TYPE_Z
#define GATE_H GATE_TYPE_H
#define GATE_S GATE_TYPE_S
#define GATE_T GATE_TYPE_T
#define GATE_RX GATE_TYPE_RX
#define GATE_RY GATE_TYPE_RY
#define GATE_RZ GATE_TYPE_RZ
#define GATE_CNOT GATE_TYPE_CNOT
#define GATE_CY GATE_TYPE_CY
#define GATE_CZ GATE_TYPE_CZ
#define GATE_SWAP GATE_TYPE_SWAP
#define GATE_PHASE GATE_TYPE_PHASE
#define GATE_CRX GATE_TYPE_CRX
#define GATE_CRY GATE_TYPE_CRY
#define GATE_CRZ GATE_TYPE_CRZ
#define GATE_MEASURE GATE_TYPE_MEASURE

typedef struct { float real, imag; } ComplexFloat;
typedef struct { double real, imag; } ComplexDouble;

typedef enum {
    HARDWARE_TYPE_CPU, HARDWARE_TYPE_GPU, HARDWARE_TYPE_QPU, HARDWARE_TYPE_SIMULATOR, HARDWARE_TYPE_METAL,
    HARDWARE_TYPE_CUDA, HARDWARE_TYPE_FPGA, HARDWARE_TYPE_IBM, HARDWARE_TYPE_RIGETTI, HARDWARE_TYPE_DWAVE,
    HARDWARE_TYPE_CUSTOM
} HardwareType;

/* core/error_codes.h:17-110 (the values this layer returns) */
typedef int qgt_error_t;
#define QGT_SUCCESS 0
#define QGT_ERROR_INVALID_PARAMETER (-1)
#define QGT_ERROR_INVALID_ARGUMENT (-1)
#define QGT_ERROR_MEMORY_ALLOCATION (-2)
#define QGT_ERROR_DIMENSION_MISMATCH (-3)
#define QGT_ERROR_INVALID_STATE (-4)
#define QGT_ERROR_VALIDATION_FAILED (-5)
#define QGT_ERROR_HARDWARE_FAILURE (-6)
#define QGT_ERROR_NOT_IMPLEMENTED (-7)
#define QGT_ERROR_INTERNAL (-15)
#define QGT_ERROR_INVALID_DIMENSION (-22)

/* ---- hardware/quantum_hardware_types.h:254-262, hardware/quantum_hardware_abstraction.h:51-82 ---- */
typedef struct {
    gate_type_t type;
    uint32_t qubit;
    uint32_t control_qubit;
    uint32_t target_qubit;
    double parameter;
    double* parameters;
    size_t num_parameters;
} QuantumGate;

typedef struct HardwareGate {
    gate_type_t type;
    uint32_t target, control, target1, target2;
    double parameter;
    double* parameters;
    size_t num_params;
} HardwareGate;

typedef struct QuantumCircuit {
    HardwareGate* gates;
    size_t num_gates, capacity, num_qubits, num_classical_bits, depth;
    bool* measured;
    double max_circuit_depth;
    void* optimization_data;
    void* metadata;
} QuantumCircuit;
typedef struct QuantumCircuit CPUSimCircuit;

/* ---- hardware/quantum_simulator_cpu.h:29-63 ---------------------------------------------------- */
void init_simulator_state(double complex* state, size_t n);
CPUSimCircuit* cpu_sim_create_circuit(size_t max_gates);
void cpu_sim_add_gate(CPUSimCircuit* circuit, const QuantumGate* gate);
void simulate_circuit_cpu(double complex* state, const CPUSimCircuit* circuit, size_t n_qubits);
void cpu_sim_cleanup_circuit(CPUSimCircuit* circuit);
void cpu_sim_get_error_statistics(const CPUSimCircuit* circuit, double* avg_error_rate, double* max_error_rate);
void configure_circuit_optimization(CPUSimCircuit* circuit, bool use_error_correction, bool use_tensor_networks,
                                    size_t cache_line_size);
#define init_cpu_circuit cpu_sim_create_circuit
#define add_gate_to_circuit cpu_sim_add_gate
#define get_error_statistics cpu_sim_get_error_statistics

/* ---- hardware/quantum_simulator.h:33-140 (the statevector subset) -------------------------------- */
typedef enum noise_type_t { NOISE_NONE, NOISE_DEPOLARIZING, NOISE_AMPLITUDE_DAMPING, NOISE_PHASE_DAMPING, NOISE_THERMAL, NOISE_CUSTOM } NoiseType;
typedef enum { MITIGATION_NONE, MITIGATION_RICHARDSON, MITIGATION_ZNE, MITIGATION_PROBABILISTIC, MITIGATION_CUSTOM } MitigationType;
struct HierarchicalMatrix;
typedef struct noise_model_t {
    NoiseType type;
    double gate_error_rate, measurement_error_rate, decoherence_rate;
    double* custom_parameters;
    size_t num_custom_parameters;
    size_t size;
    struct HierarchicalMatrix* h_matrix;
} NoiseModel;
typedef struct { MitigationType type; uint32_t num_samples; double scale_factors[4]; void* custom_parameters; } MitigationParams;
typedef struct {
    double complex* amplitudes;
    uint32_t num_qubits;
    uint32_t num_classical_bits;
    bool* classical_bits;
    double fidelity;
    double error_rate;
    NoiseModel active_noise;
    MitigationParams active_mitigation;
    void* device_ptr;      /* qgt_b200_state* once the state lives on the GPU */
    void* custom_state;
} SimulatorState;
struct SimulatorConfig {                 /* hardware/quantum_backend_types.h:107-117 */
    char* backend_name;
    uint32_t max_shots, max_qubits;
    double coupling_map[64][64];
    double* noise_model;                 /* [gate error rate, measurement error rate, decoherence rate] or NULL */
    bool optimize_mapping, use_gpu;
    size_t memory_limit;
    void* backend_specific_config;
};
typedef struct qgt_sim_circuit SimulatorCircuit;   /* opaque; the reference aliases struct quantum_circuit_t */

SimulatorState* sim_init(uint32_t num_qubits, uint32_t num_classical_bits, const struct SimulatorConfig* config);
void sim_reset_state(SimulatorState* state);
void sim_cleanup(SimulatorState* state);
SimulatorCircuit* sim_create_circuit(uint32_t num_qubits, uint32_t num_classical_bits);
bool sim_add_gate(SimulatorCircuit* circuit, gate_type_t type, uint32_t target, uint32_t control, double* parameters);
bool sim_add_controlled_gate(SimulatorCircuit* circuit, gate_type_t type, uint32_t target, uint32_t control,
                             uint32_t control2, double* parameters);
bool sim_execute_circuit(SimulatorState* state, const SimulatorCircuit* circuit);
double complex* sim_get_statevector(const SimulatorState* state);
void sim_cleanup_circuit(SimulatorCircuit* circuit);
/* measurement / expectation / sampling (quantum_simulator.h:100-118) and the circuit text format (:141-142): the
 * reductions, the collapse and the inverse-CDF walk run on the device (qgt_b200_state_*) */
bool sim_measure_qubit(SimulatorState* state, uint32_t qubit, uint32_t classical_bit);
bool sim_measure_all(SimulatorState* state);
bool* sim_get_measurement_results(const SimulatorState* state);
uint64_t* sim_get_measurement_counts(const SimulatorState* state, uint32_t shots);
double sim_get_expectation_value(const SimulatorState* state, const char* observable);
bool sim_save_circuit(const SimulatorCircuit* circuit, const char* filename);
SimulatorCircuit* sim_load_circuit(const char* filename);
void qgt_compat_seed(unsigned long long seed);       /* seeds the measurement RNG (the reference seeds from the clock) */

/* ---- distributed/differential_geometry.h:237-262, 573-595 ------------------------------------------ */
typedef struct diffgeo_engine diffgeo_engine_t;
diffgeo_engine_t* diffgeo_engine_create(void);
void diffgeo_engine_destroy(diffgeo_engine_t* engine);
/* g = Re Q */
bool diffgeo_compute_fubini_study(diffgeo_engine_t* engine, const ComplexDouble* state, size_t dim,
                                  const ComplexDouble* param_derivatives, size_t num_params, double* metric_out);
/* F = -2 Im Q (the diffgeo convention, differential_geometry.c:2899-2900) */
bool diffgeo_compute_berry_curvature(diffgeo_engine_t* engine, const ComplexDouble* state, size_t dim,
                                     const ComplexDouble* param_derivatives, size_t num_params, double* curvature_out);

/* ---- core/quantum_types.h:170-192, core/quantum_geometric_tensor_network.h ---------------------------- */
typedef struct quantum_gate_t {
    gate_type_t type;
    size_t num_qubits;
    bool is_controlled;
    bool is_parameterized;
    size_t* target_qubits;
    size_t* control_qubits;
    size_t num_controls;
    size_t* qubits;
    double* parameters;
    size_t num_parameters;
    ComplexFloat* matrix;
    void* custom_data;
} quantum_gate_t;

typedef enum { QGTN_BACKEND_SIMULATOR, QGTN_BACKEND_IBM, QGTN_BACKEND_RIGETTI, QGTN_BACKEND_DWAVE, QGTN_BACKEND_IONQ,
               QGTN_BACKEND_XANADU, QGTN_BACKEND_CUSTOM } qgtn_backend_type_t;
typedef struct { qgtn_backend_type_t type; void* backend_specific; bool supports_gradients; bool supports_hybrid; } qgtn_hardware_config_t;
struct tensor_network_t;
struct quantum_circuit_t;
typedef struct quantum_geometric_tensor_network {
    struct tensor_network_t* network;      /* unused: the exact statevector path needs no tensor network */
    struct quantum_circuit_t* circuit;     /* unused */
    size_t num_qubits;
    size_t num_layers;
    bool is_distributed;
    bool use_hardware_acceleration;
    qgtn_hardware_config_t hardware_config;
    void* backend_state;                   /* the gate list and the cached Q of this layer */
} quantum_geometric_tensor_network_t;

quantum_geometric_tensor_network_t* create_quantum_geometric_tensor_network(size_t num_qubits, size_t num_layers,
                                                                             bool is_distributed, bool use_hardware_acceleration);
void destroy_quantum_geometric_tensor_network(quantum_geometric_tensor_network_t* qgtn);
/* appends the gate; a parameterised gate takes the next parameter index (order of is_parameterized gates,
 * core/quantum_parameter_shift.c:278-337) with value gate->parameters[0] */
bool apply_quantum_gate(quantum_geometric_tensor_network_t* qgtn, const quantum_gate_t* gate, const size_t* qubits, size_t num_qubits);
bool get_quantum_state(const quantum_geometric_tensor_network_t* qgtn, ComplexFloat** state_vector, size_t* dimension);
bool compute_quantum_geometric_tensor(const quantum_geometric_tensor_network_t* qgtn, size_t param_i, size_t param_j, ComplexFloat* result);
bool compute_quantum_metric(const quantum_geometric_tensor_network_t* qgtn, size_t param_i, size_t param_j, double* result);
bool compute_berry_curvature(const quantum_geometric_tensor_network_t* qgtn, size_t param_i, size_t param_j, double* result);
const char* get_quantum_geometric_tensor_network_error(void);

/* ---- core/quantum_state_types.h:20-26, core/quantum_types.h:52-61,241-271, core/quantum_circuit_operations.h:14-60 ------
 * The ComplexFloat circuit path.  Gate conventions of this path (quantum_circuit_operations.c:1147-1190): X = RX(pi),
 * Y = RY(pi), Z = RZ(pi), S = RZ(pi/2); two-qubit gates keep {control, target} in gate->qubits. */
typedef struct {
    size_t num_qubits;
    ComplexFloat* amplitudes;
    void* workspace;
    size_t dimension;
    bool is_normalized;
} QuantumState;
typedef QuantumState quantum_state;
typedef enum { PAULI_I = 0, PAULI_X = 1, PAULI_Y = 2, PAULI_Z = 3 } pauli_type_t;
typedef pauli_type_t PauliOperator;
typedef pauli_type_t pauli_type;
struct circuit_layer_t; struct computational_graph_t; struct quantum_geometric_state_t; struct quantum_compute_node_t;
struct quantum_circuit_t {
    size_t num_qubits;
    bool is_parameterized;
    struct circuit_layer_t** layers;  size_t num_layers, layers_capacity;        /* unused here: sweeps are fused at execution */
    quantum_gate_t** gates;  size_t num_gates, max_gates;                          /* the flat list the builders append to */
    int optimization_level;
    bool is_compiled;
    struct computational_graph_t* graph;
    struct quantum_geometric_state_t* state;
    struct quantum_compute_node_t** nodes;  size_t num_nodes, capacity;
};
typedef struct quantum_circuit_t quantum_circuit_t;
#define QGT_ERROR_INCOMPATIBLE (-19)
#define QGT_ERROR_INVALID_OPERATOR (-650)
quantum_state* init_quantum_state(size_t num_qubits);
void quantum_state_reset(quantum_state* state);
void quantum_state_cleanup(quantum_state* state);
quantum_circuit_t* quantum_circuit_create(size_t num_qubits);
void quantum_circuit_destroy(quantum_circuit_t* circuit);
void quantum_circuit_reset(quantum_circuit_t* circuit);
qgt_error_t quantum_circuit_hadamard(quantum_circuit_t* circuit, size_t qubit);
qgt_error_t quantum_circuit_pauli_x(quantum_circuit_t* circuit, size_t qubit);
qgt_error_t quantum_circuit_pauli_y(quantum_circuit_t* circuit, size_t qubit);
qgt_error_t quantum_circuit_pauli_z(quantum_circuit_t* circuit, size_t qubit);
qgt_error_t quantum_circuit_phase(quantum_circuit_t* circuit, size_t qubit, double angle);
qgt_error_t quantum_circuit_rotation(quantum_circuit_t* circuit, size_t qubit, double angle, pauli_type axis);
qgt_error_t quantum_circuit_cnot(quantum_circuit_t* circuit, size_t control, size_t target);
qgt_error_t quantum_circuit_cz(quantum_circuit_t* circuit, size_t control, size_t target);
qgt_error_t quantum_circuit_swap(quantum_circuit_t* circuit, size_t qubit1, size_t qubit2);
qgt_error_t quantum_circuit_execute(quantum_circuit_t* circuit, quantum_state* state);
qgt_error_t quantum_circuit_measure(quantum_circuit_t* circuit, quantum_state* state, size_t* results);
qgt_error_t quantum_circuit_measure_all(quantum_circuit_t* circuit, quantum_state* state, size_t* results);
qgt_error_t quantum_circuit_optimize(quantum_circuit_t* circuit, int optimization_level);
qgt_error_t quantum_circuit_validate(quantum_circuit_t* circuit);
size_t quantum_circuit_depth(const quantum_circuit_t* circuit);
size_t quantum_circuit_gate_count(const quantum_circuit_t* circuit);

/* ---- core/quantum_state_types.h:28-60 (operator object), algorithms/qaoa.h:29-304: the QAOA driver -----------------------
 * Served by csrc/compat/qaoa_compat.c: the state lives on the device, qstate->amplitudes is refreshed by the public
 * apply calls, cost_hamiltonian->matrix stays NULL (E_z is evaluated on the device from the edge list).
 * Parameter order wherever one vector is used (QGT, exact gradient): (gamma_1, beta_1, ..., gamma_p, beta_p). */
typedef enum quantum_operator_type_t { QUANTUM_OPERATOR_UNITARY, QUANTUM_OPERATOR_HERMITIAN, QUANTUM_OPERATOR_PROJECTOR,
                                       QUANTUM_OPERATOR_KRAUS, QUANTUM_OPERATOR_LINDBLAD, QUANTUM_OPERATOR_CUSTOM } quantum_operator_type_t;
typedef struct quantum_operator_t {
    quantum_operator_type_t type;
    size_t dimension;
    ComplexFloat* matrix;
    void* auxiliary_data;
    bool is_hermitian;
    void* device_data;
    HardwareType device_type;
} quantum_operator_t;
typedef struct { size_t i, j; double weight; } qaoa_edge_t;
typedef struct { size_t num_vertices, num_edges; qaoa_edge_t* edges; double* vertex_weights; } qaoa_graph_t;
typedef enum { QAOA_PROBLEM_MAXCUT, QAOA_PROBLEM_QUBO, QAOA_PROBLEM_MAX_INDEPENDENT_SET, QAOA_PROBLEM_VERTEX_COVER,
               QAOA_PROBLEM_GRAPH_COLORING, QAOA_PROBLEM_TSP, QAOA_PROBLEM_CUSTOM } qaoa_problem_type_t;
typedef enum { QAOA_MIXER_X, QAOA_MIXER_XY, QAOA_MIXER_GROVER, QAOA_MIXER_CUSTOM } qaoa_mixer_type_t;
typedef enum { QAOA_OPTIMIZER_COBYLA, QAOA_OPTIMIZER_NELDER_MEAD, QAOA_OPTIMIZER_POWELL, QAOA_OPTIMIZER_BFGS, QAOA_OPTIMIZER_SPSA,
               QAOA_OPTIMIZER_ADAM, QAOA_OPTIMIZER_GRADIENT_DESCENT } qaoa_optimizer_type_t;
typedef struct {
    size_t p;
    qaoa_problem_type_t problem_type;
    qaoa_mixer_type_t mixer_type;
    qaoa_optimizer_type_t optimizer_type;
    double* initial_gamma;
    double* initial_beta;
    size_t max_iterations;
    double tolerance;
    double learning_rate;
    size_t num_shots;
    bool use_expectation;
    bool use_gpu;
    void* backend;
} qaoa_config_t;
typedef struct qaoa_state {
    qaoa_graph_t* graph;
    quantum_operator_t* cost_hamiltonian;
    quantum_operator_t* mixer_hamiltonian;
    double* gamma;
    double* beta;
    size_t p;
    size_t num_qubits;
    QuantumState* qstate;
    double current_cost;
    double best_cost;
    double* best_gamma;
    double* best_beta;
    size_t iteration;
    qaoa_config_t config;
    int* best_solution;
    double* solution_probabilities;
} qaoa_state_t;
typedef struct {
    double optimal_cost;
    int* optimal_solution;
    double* optimal_gamma;
    double* optimal_beta;
    size_t num_iterations;
    double* cost_history;
    size_t history_length;
    double execution_time;
    double approximation_ratio;
} qaoa_result_t;
qaoa_graph_t* qaoa_create_graph(size_t num_vertices);
qgt_error_t qaoa_add_edge(qaoa_graph_t* graph, size_t i, size_t j, double weight);
qgt_error_t qaoa_set_vertex_weight(qaoa_graph_t* graph, size_t vertex, double weight);
qaoa_graph_t* qaoa_create_random_graph(size_t num_vertices, double edge_probability, double min_weight, double max_weight);
qaoa_graph_t* qaoa_create_from_adjacency(const double* adjacency, size_t n);
void qaoa_destroy_graph(qaoa_graph_t* graph);
qaoa_state_t* qaoa_init(const qaoa_graph_t* graph, const qaoa_config_t* config);       /* NULL without an sm_100 device */
qgt_error_t qaoa_construct_cost_hamiltonian(qaoa_state_t* state);
qgt_error_t qaoa_construct_mixer_hamiltonian(qaoa_state_t* state);
qgt_error_t qaoa_prepare_initial_state(qaoa_state_t* state);
qgt_error_t qaoa_apply_layer(qaoa_state_t* state, size_t layer_idx);
qgt_error_t qaoa_apply_circuit(qaoa_state_t* state, const double* gamma, const double* beta);
qgt_error_t qaoa_compute_expectation(qaoa_state_t* state, double* expectation);
qgt_error_t qaoa_compute_gradient(qaoa_state_t* state, double* gamma_grad, double* beta_grad);   /* the reference's finite-shift formula */
qaoa_result_t* qaoa_optimize(qaoa_state_t* state);
qgt_error_t qaoa_sample(qaoa_state_t* state, int** samples, size_t num_samples);
double qaoa_evaluate_solution(const qaoa_graph_t* graph, const int* solution);
void qaoa_destroy(qaoa_state_t* state);
void qaoa_destroy_result(qaoa_result_t* result);
qaoa_config_t qaoa_default_config(size_t p);
size_t qaoa_estimate_optimal_p(size_t num_vertices, size_t num_edges);
double qaoa_approximation_ratio(const qaoa_graph_t* graph, const int* solution, double optimal_cost);
void qaoa_print_state(const qaoa_state_t* state);
void qaoa_print_result(const qaoa_result_t* result);
/* exact dE/dgamma, dE/dbeta by the adjoint method on the device (not in the reference) */
qgt_error_t qgt_b200_qaoa_exact_gradient(qaoa_state_t* state, double* energy, double* gamma_grad, double* beta_grad);

/* ---- core/quantum_parameter_shift.h:28-142 (parameter index = order of the parameterised gates) ---------------------
 * States and gradients are malloc'd ComplexFloat[2^n] arrays the caller frees.  compute_higher_order_gradient returns the
 * exact derivative column d_mu psi (the reference's combination of shifted states is not a derivative). */
bool shift_parameter(quantum_geometric_tensor_network_t* qgtn, size_t param_idx, double shift_amount);
bool compute_shifted_states(quantum_geometric_tensor_network_t* qgtn, size_t param_idx, double shift_amount,
                            ComplexFloat** forward_state, ComplexFloat** backward_state, size_t* dimension);
bool compute_parameter_shift_gradient(const quantum_geometric_tensor_network_t* qgtn, size_t param_idx, double shift_amount,
                                      ComplexFloat** gradient, size_t* dimension);
bool compute_centered_difference_gradient(const quantum_geometric_tensor_network_t* qgtn, size_t param_idx, double step_size,
                                          ComplexFloat** gradient, size_t* dimension);
bool compute_higher_order_gradient(const quantum_geometric_tensor_network_t* qgtn, size_t param_idx, const double* shift_amounts,
                                   size_t num_shifts, ComplexFloat** gradient, size_t* dimension);
bool compute_gradient_with_error(const quantum_geometric_tensor_network_t* qgtn, size_t param_idx, ComplexFloat** gradient,
                                 double* error_estimate, size_t* dimension);

/* ---- core/quantum_gate_operations.h:25-140 ------------------------------------------------------------------------------
 * Two-qubit gates keep {control, target} in target_qubits, as the reference's constructor does; apply_quantum_gate
 * understands that form.  The ComplexFloat matrix follows the reference's tables (quantum_gate_operations.c:11-150), with
 * the real SWAP matrix instead of the reference's 2x2 identity block. */
quantum_gate_t* create_quantum_gate(gate_type_t type, const size_t* qubits, size_t num_qubits, const double* parameters, size_t num_parameters);
quantum_gate_t* copy_quantum_gate(const quantum_gate_t* gate);
bool update_gate_parameters(quantum_gate_t* gate, const double* parameters, size_t num_parameters);
bool shift_gate_parameters(quantum_gate_t* gate, size_t param_idx, double shift_amount);
quantum_gate_t* create_rx_gate(size_t qubit, double angle);
quantum_gate_t* create_ry_gate(size_t qubit, double angle);
quantum_gate_t* create_rz_gate(size_t qubit, double angle);
quantum_gate_t* create_h_gate(size_t qubit);
quantum_gate_t* create_x_gate(size_t qubit);
quantum_gate_t* create_y_gate(size_t qubit);
quantum_gate_t* create_z_gate(size_t qubit);
quantum_gate_t* create_cnot_gate(size_t control, size_t target);
quantum_gate_t* create_cz_gate(size_t control, size_t target);
void destroy_quantum_gate(quantum_gate_t* gate);

/* ---- core/numerical_backend.h:9-41,146-147: the two calls reference programs make before using the network API --------- */
typedef enum { NUMERICAL_SUCCESS, NUMERICAL_ERROR_INVALID_ARGUMENT, NUMERICAL_ERROR_MEMORY, NUMERICAL_ERROR_BACKEND,
               NUMERICAL_ERROR_COMPUTATION, NUMERICAL_ERROR_NOT_IMPLEMENTED, NUMERICAL_ERROR_INVALID_STATE } numerical_error_t;
typedef enum { NUMERICAL_BACKEND_CPU, NUMERICAL_BACKEND_ACCELERATE, NUMERICAL_BACKEND_OPENBLAS, NUMERICAL_BACKEND_MKL,
               NUMERICAL_BACKEND_CUDA, NUMERICAL_BACKEND_METAL } numerical_backend_t;
typedef struct {
    numerical_backend_t type;
    size_t max_threads;
    bool use_fma, use_avx, use_neon;
    size_t cache_size;
    void* backend_specific;
} numerical_config_t;
const char* get_numerical_error_string(numerical_error_t error);
numerical_error_t initialize_numerical_backend(const numerical_config_t* config);   /* accepts any type: the work runs on the GPU */
void shutdown_numerical_backend(void);

/* ---- core/quantum_geometric_types.h:365-459, core/quantum_geometric_metric.h, _curvature.h -------------- */
#define QGT_MAX_DIMENSIONS 16
typedef enum { GEOMETRIC_METRIC_EUCLIDEAN, GEOMETRIC_METRIC_MINKOWSKI, GEOMETRIC_METRIC_FUBINI_STUDY, GEOMETRIC_METRIC_KAHLER,
               GEOMETRIC_METRIC_CUSTOM } geometric_metric_type_t;
typedef enum { GEOMETRIC_CURVATURE_RIEMANN, GEOMETRIC_CURVATURE_RICCI, GEOMETRIC_CURVATURE_SCALAR, GEOMETRIC_CURVATURE_WEYL,
               GEOMETRIC_CURVATURE_BERRY, GEOMETRIC_CURVATURE_CUSTOM } geometric_curvature_type_t;
typedef struct quantum_geometric_metric_t {
    geometric_metric_type_t type;
    size_t dimension;
    ComplexFloat* components;
    bool is_symmetric;
    void* auxiliary_data;
    HardwareType hardware;
} quantum_geometric_metric_t;
typedef struct quantum_geometric_curvature_t {
    geometric_curvature_type_t type;
    size_t dimension;
    ComplexFloat* components;
    void* auxiliary_data;
    bool is_flat;
    HardwareType hardware;
} quantum_geometric_curvature_t;

/* legacy constructors keep the reference's dimension <= QGT_MAX_DIMENSIONS check (metric.c:14, curvature.c:14);
 * qgt_b200_alloc_* lift it for the P x P outputs of real circuits */
qgt_error_t geometric_create_metric(quantum_geometric_metric_t** metric, geometric_metric_type_t type, size_t dimension, HardwareType hardware);
void geometric_destroy_metric(quantum_geometric_metric_t* metric);
qgt_error_t geometric_create_curvature(quantum_geometric_curvature_t** curvature, geometric_curvature_type_t type, size_t dimension, HardwareType hardware);
void geometric_destroy_curvature(quantum_geometric_curvature_t* curvature);
qgt_error_t qgt_b200_alloc_metric(quantum_geometric_metric_t** metric, size_t dimension);
qgt_error_t qgt_b200_alloc_curvature(quantum_geometric_curvature_t** curvature, size_t dimension);

/* g in .real (symmetric), Omega = Im Q in .real (antisymmetric), full Hermitian Q */
qgt_error_t geometric_compute_fubini_study_metric(quantum_geometric_metric_t* metric, const quantum_geometric_tensor_network_t* qgtn, size_t num_params);
qgt_error_t geometric_compute_berry_curvature(quantum_geometric_curvature_t* curvature, const quantum_geometric_tensor_network_t* qgtn, size_t num_params);
qgt_error_t geometric_compute_berry_curvature_element(const quantum_geometric_tensor_network_t* qgtn, size_t param_mu, size_t param_nu, float* result);
qgt_error_t geometric_compose_qgt(ComplexFloat* qgt, const quantum_geometric_metric_t* metric, const quantum_geometric_curvature_t* curvature, size_t dimension);
qgt_error_t geometric_compute_full_qgt(ComplexFloat* qgt, const quantum_geometric_tensor_network_t* qgtn, size_t num_params);

/* ---- core/quantum_geometric_gradient.h:188-215 ------------------------------------------------------------ */
typedef struct {
    float regularization_param;
    float condition_threshold;
    bool use_adaptive_regularization;
    bool use_pseudoinverse_fallback;
    float singular_value_cutoff;
} natural_gradient_config_t;
natural_gradient_config_t get_default_natural_gradient_config(void);
bool compute_regularized_natural_gradient(const ComplexFloat* gradient, const ComplexFloat* metric, ComplexFloat* natural_gradient,
                                          size_t dimension, const natural_gradient_config_t* config);

/* ---- core/quantum_geometric_gpu.h:28-276: the device-memory seam ------------------------------------------- */
#define MAX_GPU_NAME_LENGTH 256
typedef enum {
    QG_GPU_SUCCESS = 0, QG_GPU_ERROR_NO_DEVICE = -1, QG_GPU_ERROR_NOT_INITIALIZED = -2, QG_GPU_ERROR_OUT_OF_MEMORY = -3,
    QG_GPU_ERROR_INVALID_DEVICE = -4, QG_GPU_ERROR_INVALID_VALUE = -5, QG_GPU_ERROR_LAUNCH_FAILED = -6,
    QG_GPU_ERROR_SYNC_FAILED = -7, QG_GPU_ERROR_INTERNAL = -8
} gpu_error_t;
#define QGT_ERROR_GPU_NOT_AVAILABLE 300
#define QGT_ERROR_GPU_OUT_OF_MEMORY 301
#define QGT_ERROR_GPU_INVALID_VALUE 302
#define QGT_ERROR_GPU_LAUNCH_FAILED 303
#define QGT_ERROR_GPU_SYNC_FAILED 304
#define QGT_ERROR_GPU_INTERNAL 305
typedef struct {
    void* device_ptr;       /* device memory (page-locked host memory when is_pinned) */
    size_t size;
    bool is_pinned;
} gpu_buffer_t;
typedef struct {
    int device_id;
    char name[MAX_GPU_NAME_LENGTH];
    size_t total_memory;
    size_t free_memory;
    int compute_capability_major;
    int compute_capability_minor;
    int max_threads_per_block;
    int max_block_dimensions[3];
    int max_grid_dimensions[3];
    int compute_units;
    int backend_type;
    bool supports_unified_memory;
} gpu_device_info_t;
/* device memory only: no host-malloc fallback (the reference falls back when built without a GPU back-end) */
qgt_error_t gpu_malloc(void** ptr, size_t size);
qgt_error_t qgt_gpu_free_buffer(void* ptr);
qgt_error_t gpu_memcpy_host_to_device(void* dst, const void* src, size_t size);
qgt_error_t gpu_memcpy_device_to_host(void* dst, const void* src, size_t size);
int qg_gpu_init(void);                          /* QG_GPU_ERROR_NO_DEVICE without an sm_100 device */
void qg_gpu_cleanup(void);
void qg_gpu_shutdown(void);
int qg_gpu_get_device_count(int* count);
int qg_gpu_get_device_info(int device_id, gpu_device_info_t* info);
int qg_gpu_set_device(int device_id);
gpu_error_t qg_gpu_get_last_error(void);
const char* qg_gpu_get_error_string(gpu_error_t error);
int qg_gpu_allocate(gpu_buffer_t* buffer, size_t size);
int qg_gpu_allocate_pinned(gpu_buffer_t* buffer, size_t size);
int qg_gpu_free(gpu_buffer_t* buffer);
int qg_gpu_memcpy_to_device(gpu_buffer_t* dst, const void* src, size_t size);
int qg_gpu_memcpy_to_host(void* dst, const gpu_buffer_t* src, size_t size);
int qg_gpu_create_stream(int* stream_id);
int qg_gpu_destroy_stream(int stream_id);
int qg_gpu_synchronize_stream(int stream_id);
int qg_gpu_synchronize(void);

/* ---- core/quantum_geometric_types.h:239-250,290-353, core/quantum_geometric_tensor.h:13-100: generic rank-N tensors ----------
 * Small host algebra (csrc/compat/tensor_compat.c): the entry points the reference's tests/test_quantum_geometric_tensor.c
 * and tests/test_quantum_geometric_tensor_init.c drive. */
typedef enum { GEOMETRIC_TENSOR_SCALAR, GEOMETRIC_TENSOR_VECTOR, GEOMETRIC_TENSOR_COVECTOR, GEOMETRIC_TENSOR_BIVECTOR,
               GEOMETRIC_TENSOR_TRIVECTOR, GEOMETRIC_TENSOR_UNITARY, GEOMETRIC_TENSOR_HERMITIAN, GEOMETRIC_TENSOR_SYMMETRIC,
               GEOMETRIC_TENSOR_CUSTOM } geometric_tensor_type_t;
typedef enum { QGT_MEM_STANDARD = 0, QGT_MEM_HUGE_PAGES, QGT_MEM_PINNED, QGT_MEM_UNIFIED } qgt_memory_type_t;
typedef struct quantum_geometric_tensor_t {
    geometric_tensor_type_t type;
    size_t* dimensions;
    size_t rank;
    size_t dimension;
    ComplexFloat* components;
    size_t num_spins;
    void* auxiliary_data;
    bool is_symmetric, is_unitary, is_hermitian;
    HardwareType hardware;
    qgt_memory_type_t mem_type;
    size_t total_elements, aligned_elements;
} quantum_geometric_tensor_t;
#define QGT_ERROR_NUMERICAL_INSTABILITY (-24)
qgt_error_t geometric_tensor_create(quantum_geometric_tensor_t** tensor, geometric_tensor_type_t type, const size_t* dimensions, size_t rank);
void geometric_tensor_destroy(quantum_geometric_tensor_t* tensor);
qgt_error_t geometric_tensor_clone(quantum_geometric_tensor_t** dest, const quantum_geometric_tensor_t* src);
qgt_error_t geometric_tensor_initialize(quantum_geometric_tensor_t* tensor, const ComplexFloat* data);
qgt_error_t geometric_tensor_initialize_zero(quantum_geometric_tensor_t* tensor);
qgt_error_t geometric_tensor_initialize_random(quantum_geometric_tensor_t* tensor, float min_val, float max_val);
qgt_error_t geometric_tensor_initialize_identity(quantum_geometric_tensor_t* tensor);
qgt_error_t geometric_tensor_add(quantum_geometric_tensor_t* result, const quantum_geometric_tensor_t* a, const quantum_geometric_tensor_t* b);
qgt_error_t geometric_tensor_subtract(quantum_geometric_tensor_t* result, const quantum_geometric_tensor_t* a, const quantum_geometric_tensor_t* b);
qgt_error_t geometric_tensor_multiply(quantum_geometric_tensor_t* result, const quantum_geometric_tensor_t* a, const quantum_geometric_tensor_t* b);
qgt_error_t geometric_tensor_contract(quantum_geometric_tensor_t* result, const quantum_geometric_tensor_t* a, const quantum_geometric_tensor_t* b,
                                      const size_t* indices_a, const size_t* indices_b, size_t num_indices);
qgt_error_t geometric_tensor_outer_product(quantum_geometric_tensor_t* result, const quantum_geometric_tensor_t* a, const quantum_geometric_tensor_t* b);
qgt_error_t geometric_tensor_inner_product(ComplexFloat* result, const quantum_geometric_tensor_t* a, const quantum_geometric_tensor_t* b);
qgt_error_t geometric_tensor_transpose(quantum_geometric_tensor_t* result, const quantum_geometric_tensor_t* tensor, const size_t* permutation);
qgt_error_t geometric_tensor_conjugate(quantum_geometric_tensor_t* result, const quantum_geometric_tensor_t* tensor);
qgt_error_t geometric_tensor_scale(quantum_geometric_tensor_t* result, const quantum_geometric_tensor_t* tensor, ComplexFloat scalar);
qgt_error_t geometric_tensor_adjoint(quantum_geometric_tensor_t* result, const quantum_geometric_tensor_t* tensor);
qgt_error_t geometric_tensor_norm(float* norm, const quantum_geometric_tensor_t* tensor);
qgt_error_t geometric_tensor_trace(ComplexFloat* trace, const quantum_geometric_tensor_t* tensor);
bool geometric_tensor_is_hermitian(const quantum_geometric_tensor_t* tensor);
bool geometric_tensor_is_unitary(const quantum_geometric_tensor_t* tensor);
qgt_error_t geometric_tensor_validate(const quantum_geometric_tensor_t* tensor);
qgt_error_t geometric_tensor_validate_dimensions(const quantum_geometric_tensor_t* tensor);
qgt_error_t geometric_tensor_validate_compatibility(const quantum_geometric_tensor_t* a, const quantum_geometric_tensor_t* b);

/* ---- hardware/quantum_geometric_tensor_gpu.h:8-81: the QGT-on-GPU seam --------------------------------------------------
 * Data convention (the reference defines none, see csrc/compat/gpu_seam_compat.c): `state` is rows x cols row-major with
 * row 0 = psi and rows 1..rows-1 = d_mu psi (P = rows - 1); the output buffer has the same size and carries, at its start,
 * the P x P metric (Re Q in .real), the P x P Berry curvature (Im Q in .real) or the P Berry connections i<psi|d_a psi>. */
typedef struct { double precision; bool use_quantum_estimation; bool use_quantum_memory; int error_correction; int optimization_level; } QGTConfig;
typedef struct {
    void* device; void* command_queue; void* library; size_t device_memory;
    qgt_error_t (*execute_metric)(void* state, void* metric, size_t rows, size_t cols);
    qgt_error_t (*execute_connection)(void* state, void* connection, size_t rows, size_t cols);
    qgt_error_t (*execute_curvature)(void* state, void* curvature, size_t rows, size_t cols);
} MetalContext;
typedef struct {
    void* stream; void* module;
    qgt_error_t (*execute_metric)(void* state, void* metric, size_t rows, size_t cols);
    qgt_error_t (*execute_connection)(void* state, void* connection, size_t rows, size_t cols);
    qgt_error_t (*execute_curvature)(void* state, void* curvature, size_t rows, size_t cols);
} CUDAContext;
typedef struct {
    bool is_available;
    qgt_error_t (*malloc)(void** ptr, size_t size);
    qgt_error_t (*free)(void* ptr);
    qgt_error_t (*memcpy_to_device)(void* dst, const void* src, size_t size);
    qgt_error_t (*memcpy_from_device)(void* dst, const void* src, size_t size);
    size_t (*get_optimal_block_size)(void);
    union { MetalContext metal; CUDAContext cuda; };
} GPUContext;
QGTConfig qgt_default_config(void);
const char* qgt_error_string(qgt_error_t error);
qgt_error_t compute_quantum_metric_gpu(GPUContext* ctx, const ComplexFloat* state, ComplexFloat* metric, size_t rows, size_t cols, const QGTConfig* config);
qgt_error_t compute_quantum_connection_gpu(GPUContext* ctx, const ComplexFloat* state, ComplexFloat* connection, size_t rows, size_t cols, const QGTConfig* config);
qgt_error_t compute_quantum_curvature_gpu(GPUContext* ctx, const ComplexFloat* state, ComplexFloat* curvature, size_t rows, size_t cols, const QGTConfig* config);
/* fills in every hook of `ctx` (device memory seam + the three execute hooks); the reference leaves them NULL */
qgt_error_t qgt_b200_context_init(GPUContext* ctx);

/* ---- this layer -------------------------------------------------------------------------------------------- */
/* Threading: every calling thread gets its own device context (stream, cached buffers) on first use, released when the
 * thread ends, so the entry points are re-entrant on distinct objects like the reference's; one object must not be used
 * from two threads at once.
 * CUDA ordinal used by the wrappers (default: $QGT_B200_DEVICE or 0); call before the first compute call */
int qgt_compat_set_device(int device);
const char* qgt_compat_last_error(void);

#ifdef __cplusplus
}
#endif
#endif /* QGT_COMPAT_H */

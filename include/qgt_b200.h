/*
 * qgt_b200.h — thin C-ABI of the sm_100a statevector + quantum-geometric-tensor library.
 *
 * This is the drop-in boundary: plain pointers and sizes, `extern "C"`, no C++ or
 * torch types.  Every reference entry point on the hot path is a short
 * C wrapper over these calls (see include/quantum_geometric/ and
 * quantum_geometric_tensor_b200/csrc/compat/).  All paths are relative to the reference
 * tree tsotchke/quantum_geometric_tensor.
 *
 * What each group replaces
 *   qgt_b200_state_*  / qgt_b200_apply_circuit
 *       -> sim_init / sim_execute_circuit / sim_get_statevector
 *          (src/quantum_geometric/hardware/quantum_simulator.c:57,499,677) and
 *          init_simulator_state / simulate_circuit_cpu
 *          (src/quantum_geometric/hardware/quantum_simulator_cpu.c:98,761)
 *   qgt_b200_qgt
 *       -> geometric_compute_full_qgt (src/quantum_geometric/core/quantum_geometric_curvature.c:289),
 *          geometric_compute_fubini_study_metric (core/quantum_geometric_metric.c:353),
 *          compute_quantum_geometric_tensor (core/quantum_geometric_tensor_network.c:1058),
 *          diffgeo_compute_fubini_study / diffgeo_compute_berry_curvature
 *          (src/quantum_geometric/distributed/differential_geometry.c:2819,2864)
 *   qgt_b200_gram
 *       -> the three 2^n-long dot loops of compute_quantum_geometric_tensor
 *          (core/quantum_geometric_tensor_network.c:1127-1175) for caller-supplied columns
 *   qgt_b200_natural_gradient
 *       -> compute_regularized_natural_gradient (core/quantum_geometric_gradient.c:2887)
 *   qgt_b200_dist_*
 *       -> init_distributed_state / sync_quantum_states
 *          (src/quantum_geometric/distributed/quantum_distributed_operations.c:260,289) and the
 *          NCCL calls of core/multi_gpu_operations.c:46-264
 *
 * Conventions
 *   - amplitudes are complex double, interleaved (re, im), index bit q = qubit q (qubit 0 = LSB),
 *     exactly the layout of `double complex[2^n]` in the reference simulator;
 *   - status codes are the reference's qgt_error_t values (core/error_codes.h:17-110);
 *   - there is NO CPU fallback: without a CUDA device every compute entry point returns
 *     QGT_B200_ERR_NO_DEVICE and writes nothing.
 */
#ifndef QGT_B200_H
#define QGT_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define QGT_B200_ABI_VERSION 3

/* ---- status (numerically equal to the reference's qgt_error_t) ------------------------- */
enum {
    QGT_B200_OK               = 0,
    QGT_B200_ERR_INVALID_ARG  = -1,   /* QGT_ERROR_INVALID_PARAMETER   */
    QGT_B200_ERR_NO_MEMORY    = -2,   /* QGT_ERROR_MEMORY_ALLOCATION   */
    QGT_B200_ERR_DIMENSION    = -3,   /* QGT_ERROR_DIMENSION_MISMATCH  */
    QGT_B200_ERR_INVALID_STATE= -4,   /* QGT_ERROR_INVALID_STATE       */
    QGT_B200_ERR_HARDWARE     = -6,   /* QGT_ERROR_HARDWARE_FAILURE    */
    QGT_B200_ERR_UNSUPPORTED  = -7,   /* QGT_ERROR_NOT_IMPLEMENTED     */
    QGT_B200_ERR_INTERNAL     = -15,  /* QGT_ERROR_INTERNAL            */
    QGT_B200_ERR_NOT_INIT     = -18,  /* QGT_ERROR_NOT_INITIALIZED     */
    QGT_B200_ERR_CIRCUIT      = -36,  /* QGT_ERROR_INVALID_CIRCUIT     */
    QGT_B200_ERR_NO_DEVICE    = -31   /* QGT_ERROR_INVALID_HARDWARE    */
};

/* ---- gate kinds: numeric values follow gate_type_t (core/quantum_base_types.h:32-72) ---- */
enum {
    QGT_B200_GATE_I = 0, QGT_B200_GATE_X = 1, QGT_B200_GATE_Y = 2, QGT_B200_GATE_Z = 3,
    QGT_B200_GATE_H = 4, QGT_B200_GATE_S = 5, QGT_B200_GATE_T = 6,
    QGT_B200_GATE_RX = 7, QGT_B200_GATE_RY = 8, QGT_B200_GATE_RZ = 9,
    QGT_B200_GATE_CNOT = 10, QGT_B200_GATE_CY = 11, QGT_B200_GATE_CZ = 12, QGT_B200_GATE_SWAP = 13,
    QGT_B200_GATE_U1 = 15, QGT_B200_GATE_PHASE = 19,
    QGT_B200_GATE_CRX = 22, QGT_B200_GATE_CRY = 23, QGT_B200_GATE_CRZ = 24, QGT_B200_GATE_CH = 25,
    QGT_B200_GATE_SDG = 26, QGT_B200_GATE_TDG = 27, QGT_B200_GATE_SX = 29,
    QGT_B200_GATE_ZZ = 35,
    /* not in gate_type_t: the QAOA cost layer exp(-i*angle*E_z) of algorithms/qaoa.c:344-372,
       E_z from the circuit's edge list / vertex weights (qaoa.c:236-294) */
    QGT_B200_GATE_COST = 100
};

/* One gate.  angle_effective = scale * theta[param] + angle   (param >= 0)
 *                            = angle                           (param <  0)
 * For two-qubit kinds `control` is the control qubit (CNOT/CY/CZ/CH/CR*) or the second qubit
 * (SWAP/ZZ); it is ignored (set -1) otherwise. */
typedef struct qgt_b200_gate {
    int32_t kind;
    int32_t target;
    int32_t control;
    int32_t param;
    double  angle;
    double  scale;
} qgt_b200_gate;

typedef struct qgt_b200_edge { int32_t i, j; double weight; } qgt_b200_edge;

enum { QGT_B200_INIT_ZERO = 0, QGT_B200_INIT_PLUS = 1 };

/* A parameterised circuit.  All pointers are HOST pointers, read during the call only. */
typedef struct qgt_b200_circuit {
    int32_t              num_qubits;
    int32_t              num_params;      /* length of theta[] */
    const qgt_b200_gate* gates;
    size_t               num_gates;
    const qgt_b200_edge* edges;           /* for QGT_B200_GATE_COST, may be NULL */
    size_t               num_edges;
    const double*        vertex_weights;  /* num_qubits doubles or NULL */
    int32_t              initial_state;   /* QGT_B200_INIT_* */
} qgt_b200_circuit;

typedef struct qgt_b200_ctx qgt_b200_ctx;      /* one per process and GPU */
typedef struct qgt_b200_state qgt_b200_state;  /* a device-resident statevector (or shard) */

/* ---- library / device -------------------------------------------------------------------- */
int         qgt_b200_abi_version(void);
int         qgt_b200_device_count(void);                 /* 0 when no usable sm_100 device */
const char* qgt_b200_last_error(void);                   /* thread-local message */
const char* qgt_b200_error_string(int status);

/* Raw device memory / streams: what the reference's device-memory seam binds (qg_gpu_allocate, gpu_malloc, ...
 * include/quantum_geometric/core/quantum_geometric_gpu.h:139-276, src/quantum_geometric/core/quantum_geometric_gpu.c).
 * Device allocations never fall back to host memory: without an sm_100 device every call returns
 * QGT_B200_ERR_NO_DEVICE. */
typedef struct qgt_b200_device_info {
    int    device;                     /* CUDA ordinal */
    char   name[256];
    size_t total_memory, free_memory;  /* bytes */
    int    cc_major, cc_minor;
    int    num_sms;
    int    max_threads_per_block;
    int    max_block_dim[3], max_grid_dim[3];
    int    unified_addressing;
} qgt_b200_device_info;
int qgt_b200_device_info_get(int device, qgt_b200_device_info* out);
int qgt_b200_set_device(int device);                                   /* for the raw-memory calls of this thread */
int qgt_b200_mem_alloc(void** ptr, size_t bytes);                      /* device memory */
int qgt_b200_mem_free(void* ptr);
int qgt_b200_mem_alloc_pinned(void** ptr, size_t bytes);               /* page-locked host memory */
int qgt_b200_mem_free_pinned(void* ptr);
int qgt_b200_memcpy_h2d(void* dst, const void* src, size_t bytes);
int qgt_b200_memcpy_d2h(void* dst, const void* src, size_t bytes);
int qgt_b200_stream_create(void** stream);                             /* a cudaStream_t behind void* */
int qgt_b200_stream_destroy(void* stream);
int qgt_b200_stream_synchronize(void* stream);
int qgt_b200_device_synchronize(void);
int qgt_b200_memcpy_async(void* dst, const void* src, size_t bytes, void* stream);   /* direction from the pointers */
int qgt_b200_memset_async(void* ptr, int value, size_t bytes, void* stream);         /* host pointers are memset on the host */
int qgt_b200_event_create(void** event);                               /* a cudaEvent_t behind void* */
int qgt_b200_event_destroy(void* event);
int qgt_b200_event_record(void* event, void* stream);
int qgt_b200_event_wait(void* stream, void* event);                    /* stream waits for the event */
int qgt_b200_event_synchronize(void* event);

int  qgt_b200_create(qgt_b200_ctx** out, int device);    /* device = CUDA ordinal */
void qgt_b200_destroy(qgt_b200_ctx* ctx);
void* qgt_b200_ctx_stream(qgt_b200_ctx* ctx);            /* the cudaStream_t the context's kernels are launched on */
/* workspace cap in bytes for derivative columns (0 = 90 % of free HBM at first use) */
int  qgt_b200_set_workspace_limit(qgt_b200_ctx* ctx, size_t bytes);
/* Tuning / test knobs (defaults are the measured best; results do not depend on them):
 *   tile_qubits (11)       amplitudes staged per CTA = 2^tile_qubits       low_qubits (4)   qubits 0..low-1 always in the tile
 *   reg_qubits (3)         qubits of a stage matrix                         batch_qubits (0) extra amplitudes per thread
 *   max_ops_per_run (160)  fusion depth cap                                 birth_cut (1)    first run executed in pieces (2 = also on small states)
 *   use_mma (1)            dense stages on the FP64 tensor pipe             double_buffer (0) two tile buffers per CTA
 *   tiles_per_item (0)     tiles per work unit, 0 = automatic               gram_tile (0)    force 32 or 64 square Gram tiles
 *   max_slots (0)          cap on statevector-sized columns (forces the blocked schedule; tests)
 *   fused (-1)             schedule: -1 automatic (Gram while every column is resident, fused otherwise), 0 Gram, 1 fused
 *   fused_traj (-1)        fused schedule: phi's stage images (-1 automatic, 0 never: phi recomputed per column, 1 whenever they fit)
 *   fused_pipeline (3)     trajectory kernel: 3 direct (3 CTAs x 8 warps), 1 persistent 16-warp, 2 lean 2 x 16-warp, 0 generic
 *   fused_overlap (1)      ranged trajectory images: phi's launch of the next tile range on a side stream
 *   cost_tables (1)        QAOA cost pass through per-launch phase tables (0 = one sincos per amplitude)
 *   parity_form (1)        X-rotation stage matrices next to a cost pass in the 4-DMMA parity form
 *   wavefront (2)          gate order of unsharded plans (virtual rank qubits of the DAG walk)
 *   profile (0)            per-category CUDA-event timing into qgt_b200_stats
 *   debug_skip (0)         timing experiments in builds with -DQGT_DEBUG_SKIP only
 * Unknown keys return QGT_B200_ERR_INVALID_ARG. */
int  qgt_b200_set_option(qgt_b200_ctx* ctx, const char* key, double value);

/* ---- statevector ------------------------------------------------------------------------- */
int  qgt_b200_state_create(qgt_b200_ctx* ctx, int num_qubits, qgt_b200_state** out);
void qgt_b200_state_destroy(qgt_b200_state* st);
int  qgt_b200_state_init(qgt_b200_state* st, int initial_state);                 /* |0..0> or |+>^n */
int  qgt_b200_state_upload(qgt_b200_state* st, const double* host_amps);          /* 2^n (re,im) */
int  qgt_b200_state_download(const qgt_b200_state* st, double* host_amps);
int  qgt_b200_state_norm2(const qgt_b200_state* st, double* out);
void* qgt_b200_state_device_ptr(qgt_b200_state* st);                             /* raw device pointer */

/* Apply `circuit` (its gates only; circuit->initial_state is ignored) to the state in place. */
int  qgt_b200_apply_circuit(qgt_b200_state* st, const qgt_b200_circuit* circuit, const double* theta);

/* Host-buffer convenience used by simulate_circuit_cpu: H2D, sweeps, D2H in one call. */
int  qgt_b200_simulate_host(qgt_b200_ctx* ctx, double* host_amps, int num_qubits,
                            const qgt_b200_circuit* circuit, const double* theta);

/* ---- quantum geometric tensor ------------------------------------------------------------- */
/* Full QGT of |psi(theta)> = U(theta)|init>:  Q = <d_mu psi|d_nu psi> - <d_mu psi|psi><psi|d_nu psi>.
 * Outputs are HOST buffers, row-major P x P, any of them may be NULL:
 *   metric  = Re Q            (Fubini-Study, core/quantum_geometric_metric.c:353)
 *   berry   = Im Q            (core/quantum_geometric_curvature.c:201; diffgeo's F = -2*berry)
 *   q_full  = Q interleaved (re,im)
 * If `psi_out` (a state with the same qubit count) is given it receives psi(theta). */
int  qgt_b200_qgt(qgt_b200_ctx* ctx, const qgt_b200_circuit* circuit, const double* theta,
                  double* metric, double* berry, double* q_full, qgt_b200_state* psi_out);

/* Same formula for caller-supplied derivative columns (device or host pointers are detected):
 * psi[dim], dpsi[P*dim] row-major as in diffgeo_compute_fubini_study. */
int  qgt_b200_gram(qgt_b200_ctx* ctx, const double* psi, const double* dpsi, size_t dim, size_t num_params,
                   double* metric, double* berry, double* q_full);

/* Derivative column d_mu psi written to `out` (device state) — for tests and callers that want J. */
int  qgt_b200_derivative(qgt_b200_ctx* ctx, const qgt_b200_circuit* circuit, const double* theta,
                         int mu, qgt_b200_state* out);

/* Regularised natural-gradient step on the real metric: out = (G + lambda I)^-1 grad with the
 * adaptive-lambda / pseudo-inverse semantics of core/quantum_geometric_gradient.c:2721-2964. */
typedef struct qgt_b200_natgrad_config {
    double regularization;        /* lambda, default 1e-4 */
    double condition_threshold;   /* 1e8 */
    int    adaptive;              /* 1 */
    int    pseudoinverse_fallback;/* 1 */
    double singular_cutoff;       /* 1e-10 */
} qgt_b200_natgrad_config;
int  qgt_b200_natural_gradient(qgt_b200_ctx* ctx, const double* metric, const double* grad, size_t num_params,
                               const qgt_b200_natgrad_config* cfg, double* out, double* lambda_used);

/* Energy E = <psi|H|psi> and its gradient dE/d theta by the adjoint method on the device: one forward circuit and one
 * backward pass over two states, whatever the number of parameters (replaces the two circuit executions per parameter
 * of algorithms/qaoa.c:489-558; feeds the natural-gradient step).  H is the diagonal observable defined by the
 * circuit's edge list and vertex weights, H = sum_edges w (1 - Z_i Z_j)/2 + sum_q v_q Z_q (E_z of qaoa.c:258-289), for
 * any circuit, with or without COST gates.  Either output may be NULL.  Collective on a sharded state (every rank makes
 * the same call and receives the same energy and gradient). */
int  qgt_b200_expectation_gradient(qgt_b200_ctx* ctx, const qgt_b200_circuit* circuit, const double* theta,
                                   double* energy, double* grad);

/* Natural-gradient descent on E(theta) (see qgt_b200_expectation_gradient for H): one step evaluates the metric
 * (qgt_b200_qgt), the adjoint gradient and the regularised solve, theta_out = theta - learning_rate * (G + lambda I)^-1 grad;
 * `energy` receives E(theta) BEFORE the step.  The loop form updates theta in place; history (optional, iterations + 1
 * doubles) receives E before every step and E of the final parameters.  Replaces the host loop around
 * compute_quantum_geometric_tensor + compute_regularized_natural_gradient (core/quantum_geometric_gradient.c:2887). */
int  qgt_b200_natural_gradient_step(qgt_b200_ctx* ctx, const qgt_b200_circuit* circuit, const double* theta, double learning_rate,
                                    const qgt_b200_natgrad_config* cfg, double* theta_out, double* energy, double* lambda_used);
int  qgt_b200_natural_gradient_descent(qgt_b200_ctx* ctx, const qgt_b200_circuit* circuit, double* theta, int iterations,
                                       double learning_rate, const qgt_b200_natgrad_config* cfg, double* history);

/* ---- statistics of the last qgt/apply call (for bench.py and the roofline) ----------------- */
typedef struct qgt_b200_stats {
    double ms_total;          /* device time of the whole call (CUDA events) */
    double ms_sweep;          /* time in gate-sweep kernels */
    double ms_gram;           /* time in Gram kernels */
    double ms_other;
    double sweep_bytes;       /* algorithmic bytes moved by sweep launches (32*D or 16*D+16*D per column pass) */
    double gram_flops;        /* algorithmic real flops of Gram launches (8*na*nb*D) */
    double gram_bytes;
    double exchange_bytes;    /* bytes this rank sent in qubit-swap exchanges (sharded states) */
    int64_t sweep_launches;
    int64_t gram_launches;
    int64_t other_launches;
    int64_t sweep_column_passes;
    int32_t num_runs;         /* fused sweeps the circuit was cut into */
    int32_t resident_columns; /* b */
    int32_t blocks;           /* number of resident blocks */
    int32_t tile_qubits;
    double ms_wall;           /* host wall-clock time of the whole call */
    double ms_host_plan;      /* of which: planning (fusion, stage matrices, schedule, derivative matrices) */
    double tensor_flops;      /* real flops the sweep / fused kernels issue on the FP64 tensor pipe (stage applications of
                                 every tile they carry plus the transition-matrix products of the fused schedule) */
    int32_t fused;            /* 1: the fused schedule ran (transition matrices inside the sweeps, no Gram pass) */
    int32_t fused_launches;
    double ms_exchange;       /* sharded states: time in qubit-swap exchanges over NVLink */
} qgt_b200_stats;
int  qgt_b200_get_stats(qgt_b200_ctx* ctx, qgt_b200_stats* out);

/* Debug/verification: textual dump of the fused-run plan and the column schedule for a circuit
 * (no device needed; theta may be NULL = all zeros).  Returns the JSON length (excluding NUL) or a negative status. */
long qgt_b200_plan_dump(const qgt_b200_circuit* circuit, const double* theta, int tile_qubits, int reg_qubits,
                        size_t column_slots, char* buf, size_t buflen);

/* ---- multi-GPU (one process per GPU; amplitudes sharded on the top log2(world) qubits) ------
 * Rank 0 creates the NCCL id, the host program broadcasts it (MPI, torch.distributed, a file ...), every rank
 * calls qgt_b200_dist_init once.  Afterwards qgt_b200_state_create allocates a SHARD (2^n / world amplitudes;
 * upload/download move the shard of this rank), and apply_circuit / qgt / state_norm2 act on the sharded
 * state collectively: every rank must make the same calls with the same circuit and theta. */
#define QGT_B200_NCCL_ID_BYTES    128
int  qgt_b200_dist_unique_id(uint8_t id[QGT_B200_NCCL_ID_BYTES]);
int  qgt_b200_dist_init(qgt_b200_ctx* ctx, int rank, int world, const uint8_t id[QGT_B200_NCCL_ID_BYTES]);
int  qgt_b200_dist_world(const qgt_b200_ctx* ctx, int* rank, int* world);
int  qgt_b200_dist_barrier(qgt_b200_ctx* ctx);
/* plan_dump for a state sharded over `world` ranks (no device needed): runs, EXCHANGE pseudo-runs and the
 * per-segment logical->physical qubit maps */
long qgt_b200_plan_dump_sharded(const qgt_b200_circuit* circuit, const double* theta, int world, int restore_identity,
                                int tile_qubits, int reg_qubits, size_t column_slots, char* buf, size_t buflen);

/* ---- reductions and element-wise passes on a statevector (device kernels; collective on sharded states) ----------
 * Replace the serial loops of sim_measure_qubit / sim_get_measurement_counts / sim_get_expectation_value
 * (src/quantum_geometric/hardware/quantum_simulator.c:563-675, 705-729) and measure_qubit_cpu
 * (hardware/quantum_simulator_cpu.c:328).  Masks are over GLOBAL amplitude-index bits (bit q = qubit q). */
/* prob = sum |a_i|^2 over i with (i & mask) == want; norm2 (optional) = sum over all i */
int  qgt_b200_state_probability(const qgt_b200_state* s, uint64_t mask, uint64_t want, double* prob, double* norm2);
/* <Z...Z> on the qubits of zmask: sum (-1)^popcount(i & zmask) |a_i|^2  ("Z" of sim_get_expectation_value = all qubits) */
int  qgt_b200_state_expectation_z(const qgt_b200_state* s, uint64_t zmask, double* out);
/* out = <a|b> (re, im) */
int  qgt_b200_state_inner_product(const qgt_b200_state* a, const qgt_b200_state* b, double out[2]);
int  qgt_b200_state_scale(qgt_b200_state* s, double re, double im);
int  qgt_b200_state_axpy(qgt_b200_state* dst, double re, double im, const qgt_b200_state* src);   /* dst += (re + i im) * src */
int  qgt_b200_state_normalize(qgt_b200_state* s, double* norm_before);
/* project qubit onto `outcome` and renormalise (no-op scaling when the outcome has probability 0, as the reference) */
int  qgt_b200_state_collapse(qgt_b200_state* s, int qubit, int outcome, double* prob);
/* sim_measure_qubit: outcome = uniform < p with p = P(1) mixed with the readout error rate; collapses the state.
 * The caller supplies the uniform random number (the library holds no RNG state). */
int  qgt_b200_state_measure(qgt_b200_state* s, int qubit, double uniform, double readout_error, int* outcome, double* prob_one);
/* sim_get_measurement_counts: shots[k] = first basis index whose cumulative probability exceeds uniforms[k]
 * (host array of `shots` numbers in [0, 1)).  Sharded state: collective - every rank passes the same uniforms and receives
 * all `shots` global indices (a shot is walked on the rank whose share of the probability mass holds it) */
int  qgt_b200_state_sample(const qgt_b200_state* s, const double* uniforms, size_t shots, uint64_t* indices);
/* index of the most probable basis state (lowest index on ties) and its probability; collective on a sharded state
 * (every rank receives the global winner) */
int  qgt_b200_state_argmax(const qgt_b200_state* s, uint64_t* index, double* probability);
/* <psi|H|psi> for the diagonal observable of `observable`'s edge list and vertex weights (its gates are ignored):
 * H = sum_edges w [z_i != z_j] + sum_q v_q (1 - 2 z_q), the E_z of algorithms/qaoa.c:258-289 (qaoa_compute_expectation :455) */
int  qgt_b200_state_cost_expectation(const qgt_b200_state* s, const qgt_b200_circuit* observable, double* out);
/* ComplexFloat boundary (core/quantum_state_types.h:20-26): interleaved float pairs, host or device pointer */
int  qgt_b200_state_upload_c64(qgt_b200_state* s, const float* src);
int  qgt_b200_state_download_c64(const qgt_b200_state* s, float* dst);

/* ---- ComplexFloat buffer operations and generic collectives -------------------------------------------------------
 * What the reference's ComputeBackendOps vtable binds (include/quantum_geometric/supercomputer/compute_backend.h:122-319,
 * CPU semantics src/quantum_geometric/supercomputer/backends/compute_cpu.c:341-466 + compute_simd.c): interleaved
 * (re, im) float buffers, host or device pointers (detected; host buffers are staged and the call returns after the
 * copy back, device buffers are used in place and the call is ordered on `stream`, NULL = the context's stream).
 *   apply_matrix   targets == NULL and mdim == dim: state <- M state with a dense row-major dim x dim matrix (the
 *                  reference's quantum_unitary); otherwise M is a 2^K x 2^K gate, K = 1..4, acting on qubits
 *                  targets[0..K-1] (matrix index bit j <-> qubit targets[j]; NULL = qubits 0..K-1)
 *   normalize      state /= ||state|| unless the norm is below 1e-10; norm_before (optional) receives the norm
 *   inner_product  out[0..1] = <a|b> = sum conj(a_i) b_i
 *   expectation_diag  out[0] = sum |state_i|^2 observable[i]   (real diagonal observable, `dim` floats)
 *   matmul         result[m x k] = a[m x n] b[n x k], row-major complex */
int  qgt_b200_c64_apply_matrix(qgt_b200_ctx* ctx, float* state, size_t dim, const float* matrix, size_t mdim,
                               const int32_t* targets, void* stream);
int  qgt_b200_c64_normalize(qgt_b200_ctx* ctx, float* state, size_t dim, float* norm_before, void* stream);
int  qgt_b200_c64_inner_product(qgt_b200_ctx* ctx, const float* a, const float* b, size_t dim, float* out, void* stream);
int  qgt_b200_c64_expectation_diag(qgt_b200_ctx* ctx, const float* state, const float* observable, size_t dim, float* out, void* stream);
int  qgt_b200_c64_matmul(qgt_b200_ctx* ctx, float* result, const float* a, const float* b, size_t m, size_t n, size_t k, void* stream);
/* Collectives over the communicator of qgt_b200_dist_init (a copy with one rank).  dtype: 0 f32, 1 f64, 2 c64, 3 c128,
 * 4 i32, 5 i64, 6 u8; op: 0 sum, 1 prod, 2 min, 3 max, 4 avg (the reference's ComputeDataType / ComputeReduceOp);
 * count = elements per rank; broadcast works in place on `recv` (send may be NULL). */
enum { QGT_B200_COLL_BROADCAST = 0, QGT_B200_COLL_ALLREDUCE = 1, QGT_B200_COLL_SCATTER = 2, QGT_B200_COLL_GATHER = 3,
       QGT_B200_COLL_ALLGATHER = 4, QGT_B200_COLL_REDUCE_SCATTER = 5 };
int  qgt_b200_dist_collective(qgt_b200_ctx* ctx, int kind, const void* send, void* recv, size_t count, int dtype, int op, int root);

/* Q from caller-supplied ComplexFloat columns (host or device): psi[dim], dpsi[P*dim] row-major; the columns are widened
 * on the device and contracted by the same Gram kernel as qgt_b200_gram.  Outputs are host arrays of P*P doubles. */
int  qgt_b200_gram_c64(qgt_b200_ctx* ctx, const float* psi, const float* dpsi, size_t dim, size_t num_params,
                       double* metric, double* berry, double* q_full);

/* Roofline denominators measured on this device: sustained FP64 tensor-pipe throughput (mma.sync m8n8k4 f64, TFLOP/s)
 * and device-to-device copy bandwidth (read + write bytes, GB/s).  Either output may be NULL. */
int  qgt_b200_measure_peaks(qgt_b200_ctx* ctx, double* dmma_tflops, double* copy_gbs);

/* plan_dump with the FUSED column schedule (transition matrices contracted inside the sweeps, no Gram pass; see
 * plan.hpp) for a state on `world` ranks (1 = unsharded).  Returns QGT_B200_ERR_UNSUPPORTED when the plan does not
 * qualify (tiles below 8 qubits, a sub-pass off the tensor path, a cost layer). */
long qgt_b200_plan_dump_fused(const qgt_b200_circuit* circuit, const double* theta, int world, int tile_qubits, int reg_qubits,
                              size_t column_slots, char* buf, size_t buflen);

/* plan_dump of the INVERSE circuit with the adjoint-gradient program of qgt_b200_expectation_gradient (no device needed):
 * fused != 0 -> the fused program (QGT_B200_ERR_UNSUPPORTED when the plan does not qualify), else the generic per-run
 * programs with `scratch_slots` scratch columns.  Slots: 0 = chi, 1 = its twin, 2 = Lambda, 3.. = scratch. */
long qgt_b200_plan_dump_gradient(const qgt_b200_circuit* circuit, const double* theta, int fused, int scratch_slots, int tile_qubits,
                                 int reg_qubits, int world, char* buf, size_t buflen);     /* world > 1: the sharded plan with its exchanges */

#ifdef __cplusplus
}
#endif
#endif /* QGT_B200_H */

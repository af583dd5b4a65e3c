/*
 * qgt_compute_backend.h — the reference's pluggable compute-backend seam, served by the sm_100a library.
 *
 * SURVEY.md §8(b) "Primary" seam: the ComputeBackendOps vtable of
 *   include/quantum_geometric/supercomputer/compute_backend.h:18-412   (vtable, registry API, registration macro :523-534)
 *   include/quantum_geometric/supercomputer/compute_types.h:71-350     (enums, config, op descriptor, plan, metrics)
 * Type layouts, member order and enumerator values restate those headers (ABI compatibility is the point: the table
 * is consumed by the reference's own registry, supercomputer/compute_backend.c:41-166); the two files under
 * include/quantum_geometric/supercomputer/ forward here.  The backend itself is
 * quantum_geometric_tensor_b200/csrc/compat/compute_b200.c, built into libqgt_b200_compat.so:
 *
 *   - it registers itself as COMPUTE_BACKEND_CUDA, priority 100, from an ELF constructor when the process already
 *     carries the reference's registry (compute_register_backend is a weak reference); otherwise call
 *     qgt_b200_register_compute_backend() after loading both libraries, or take the table directly with
 *     qgt_b200_compute_backend_ops();
 *   - buffers are interleaved complex float, host or device (pointers from alloc(..., COMPUTE_MEM_DEVICE) are used in
 *     place, host pointers are staged: the reference's CUDA backend stages every call, compute_cuda.cu:703-731);
 *   - ComputeStream* is a cudaStream_t, NULL = the backend's default stream; ComputeEvent* is a cudaEvent_t;
 *   - execute(QUANTUM_OP_UNITARY) with op->num_targets > 0 applies op->parameters (a 2^k x 2^k matrix, k = num_targets
 *     <= 4, matrix index bit j <-> target_qubits[j]) to op->output_data in place — the gate-level extension §8(b)
 *     describes; without targets it is the reference's dense state_size x state_size product;
 *   - there is no CPU fallback: probe() is false without an sm_100 device and init() returns NULL.
 */
#ifndef QGT_COMPUTE_BACKEND_H
#define QGT_COMPUTE_BACKEND_H

#include <stdbool.h>
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ---- compute_types.h:71-148 ---------------------------------------------------------------------------------- */
typedef enum { COMPUTE_BACKEND_AUTO = -1, COMPUTE_BACKEND_CPU = 0, COMPUTE_BACKEND_CUDA, COMPUTE_BACKEND_METAL,
               COMPUTE_BACKEND_OPENCL, COMPUTE_BACKEND_COUNT } ComputeBackendType;
typedef enum { COMPUTE_MEM_HOST = 0, COMPUTE_MEM_DEVICE, COMPUTE_MEM_UNIFIED, COMPUTE_MEM_PINNED } ComputeMemType;
typedef enum { COMPUTE_DTYPE_FLOAT32 = 0, COMPUTE_DTYPE_FLOAT64, COMPUTE_DTYPE_COMPLEX64, COMPUTE_DTYPE_COMPLEX128,
               COMPUTE_DTYPE_INT32, COMPUTE_DTYPE_INT64, COMPUTE_DTYPE_UINT8 } ComputeDataType;
typedef enum { COMPUTE_REDUCE_SUM = 0, COMPUTE_REDUCE_PROD, COMPUTE_REDUCE_MIN, COMPUTE_REDUCE_MAX, COMPUTE_REDUCE_AVG } ComputeReduceOp;
typedef enum { COMPUTE_SUCCESS = 0, COMPUTE_ERROR_INVALID_ARGUMENT, COMPUTE_ERROR_OUT_OF_MEMORY, COMPUTE_ERROR_DEVICE_NOT_FOUND,
               COMPUTE_ERROR_BACKEND_NOT_AVAILABLE, COMPUTE_ERROR_COMMUNICATION_FAILED, COMPUTE_ERROR_SYNCHRONIZATION_FAILED,
               COMPUTE_ERROR_KERNEL_FAILED, COMPUTE_ERROR_NOT_IMPLEMENTED, COMPUTE_ERROR_INTERNAL } ComputeResult;
typedef enum { QUANTUM_OP_UNITARY = 0, QUANTUM_OP_MEASUREMENT, QUANTUM_OP_TENSOR_CONTRACT, QUANTUM_OP_GRADIENT, QUANTUM_OP_NORMALIZE,
               QUANTUM_OP_INNER_PRODUCT, QUANTUM_OP_EXPECTATION, QUANTUM_OP_DENSITY_MATRIX } QuantumOpType;
typedef enum { COMPUTE_TOPO_SINGLE = 0, COMPUTE_TOPO_RING, COMPUTE_TOPO_TREE, COMPUTE_TOPO_MESH, COMPUTE_TOPO_TORUS,
               COMPUTE_TOPO_FULLY_CONNECTED, COMPUTE_TOPO_HYBRID } ComputeTopologyType;

/* ---- compute_types.h:170-229 ---------------------------------------------------------------------------------- */
typedef struct {
    bool enable_timing, enable_memory_tracking, enable_bandwidth_tracking, enable_power_monitoring;
    size_t sample_interval_ms, history_size;
} ComputeMonitorConfig;
typedef struct {
    int node_id, num_devices, num_cores;
    size_t memory_per_device, host_memory;
    double network_bandwidth, network_latency;
} ComputeNodeConfig;
typedef struct ComputeBackendConfig ComputeBackendConfig;
typedef struct {
    int num_nodes, devices_per_node;
    ComputeTopologyType topology;
    int node_rank, local_rank, local_size;           /* local_rank selects the CUDA device */
    size_t device_buffer_size, host_buffer_size, comm_buffer_size;
    bool use_nccl, use_rdma, use_compression;
    ComputeBackendType preferred_backend;
    bool allow_fallback;
    int num_streams, num_threads_per_node;
    bool enable_async;
    ComputeMonitorConfig* monitor_config;
    ComputeBackendConfig* backend_config;            /* this backend: NULL or a qgt_b200_backend_config* */
} ComputeDistributedConfig;

/* ---- compute_types.h:235-291 ---------------------------------------------------------------------------------- */
typedef struct {
    QuantumOpType type;
    void* input_data;   size_t input_size;   ComputeDataType input_dtype;
    void* output_data;  size_t output_size;  ComputeDataType output_dtype;
    void* parameters;   size_t param_size;
    size_t* dims;       size_t num_dims;
    size_t num_qubits;  size_t* target_qubits;  size_t num_targets;
    float* parameter_gradients;  size_t num_parameters;
} ComputeQuantumOp;
typedef struct {
    int* node_assignments;  size_t num_partitions, partition_size;
    size_t* offsets;  size_t* sizes;
    int* send_targets;  int* recv_sources;  size_t num_comm_ops;
    void* workspace;  size_t workspace_size;
    int priority;  bool requires_sync;
} ComputeExecutionPlan;

typedef struct ComputeBackend ComputeBackend;
typedef struct ComputeStream ComputeStream;
typedef struct ComputeEvent ComputeEvent;
typedef struct ComputeEngine ComputeEngine;
typedef struct ComputeBuffer ComputeBuffer;
typedef struct ComputeKernel ComputeKernel;

/* ---- compute_types.h:316-350 ---------------------------------------------------------------------------------- */
typedef struct {
    double total_time_ms, compute_time_ms, communication_time_ms, synchronization_time_ms, execution_time;
    double operations_per_second, flops, bandwidth_gbps;
    size_t peak_memory_bytes, current_memory_bytes, memory_used;
    double memory_efficiency;
    size_t bytes_sent, bytes_received, num_messages;
    double avg_latency_us;
    double gate_fidelity, state_fidelity, error_rate;
} ComputeMetrics;

static inline size_t compute_dtype_size(ComputeDataType t) {
    static const size_t sz[] = {4, 8, 8, 16, 4, 8, 1};
    return ((unsigned)t < 7u) ? sz[t] : 0;
}

/* ---- compute_backend.h:18-412: the vtable (slot order is the ABI) -------------------------------------------------- */
typedef struct ComputeBackendOps {
    ComputeBackend* (*init)(const ComputeDistributedConfig* config);
    void (*cleanup)(ComputeBackend* backend);
    bool (*probe)(void);
    ComputeResult (*get_capabilities)(ComputeBackend* backend, int* num_devices, size_t* total_memory);
    void* (*alloc)(ComputeBackend* backend, size_t size, ComputeMemType mem_type);
    void (*free)(ComputeBackend* backend, void* ptr, ComputeMemType mem_type);
    ComputeResult (*memcpy)(ComputeBackend* backend, void* dst, ComputeMemType dst_type, const void* src, ComputeMemType src_type,
                            size_t size, ComputeStream* stream);
    ComputeResult (*memset)(ComputeBackend* backend, void* ptr, int value, size_t size, ComputeStream* stream);
    ComputeStream* (*create_stream)(ComputeBackend* backend);
    void (*destroy_stream)(ComputeBackend* backend, ComputeStream* stream);
    ComputeResult (*synchronize_stream)(ComputeBackend* backend, ComputeStream* stream);
    ComputeEvent* (*create_event)(ComputeBackend* backend);
    void (*destroy_event)(ComputeBackend* backend, ComputeEvent* event);
    ComputeResult (*record_event)(ComputeBackend* backend, ComputeEvent* event, ComputeStream* stream);
    ComputeResult (*wait_event)(ComputeBackend* backend, ComputeStream* stream, ComputeEvent* event);
    /* state <- U state; state_size complex elements, unitary state_size x state_size row-major */
    ComputeResult (*quantum_unitary)(ComputeBackend* backend, float* state, size_t state_size, const float* unitary,
                                     size_t unitary_size, ComputeStream* stream);
    ComputeResult (*quantum_normalize)(ComputeBackend* backend, float* state, size_t size, ComputeStream* stream);
    /* result[m x k] = a[m x n] b[n x k] */
    ComputeResult (*quantum_tensor_contract)(ComputeBackend* backend, float* result, const float* a, const float* b,
                                             size_t m, size_t n, size_t k, ComputeStream* stream);
    /* gradients[0..1] = <backward|forward> */
    ComputeResult (*quantum_gradient)(ComputeBackend* backend, float* gradients, const float* forward_state,
                                      const float* backward_state, size_t size, ComputeStream* stream);
    ComputeResult (*quantum_inner_product)(ComputeBackend* backend, float* result, const float* state_a, const float* state_b,
                                           size_t size, ComputeStream* stream);
    /* result[0] = sum |state_i|^2 observable[i] */
    ComputeResult (*quantum_expectation)(ComputeBackend* backend, float* result, const float* state, const float* observable,
                                         size_t size, ComputeStream* stream);
    ComputeResult (*barrier)(ComputeBackend* backend);
    ComputeResult (*broadcast)(ComputeBackend* backend, void* data, size_t size, ComputeDataType dtype, int root);
    ComputeResult (*allreduce)(ComputeBackend* backend, const void* send_data, void* recv_data, size_t count,
                               ComputeDataType dtype, ComputeReduceOp op);
    ComputeResult (*scatter)(ComputeBackend* backend, const void* send_data, void* recv_data, size_t count, ComputeDataType dtype, int root);
    ComputeResult (*gather)(ComputeBackend* backend, const void* send_data, void* recv_data, size_t count, ComputeDataType dtype, int root);
    ComputeResult (*allgather)(ComputeBackend* backend, const void* send_data, void* recv_data, size_t count, ComputeDataType dtype);
    ComputeResult (*reduce_scatter)(ComputeBackend* backend, const void* send_data, void* recv_data, size_t count,
                                    ComputeDataType dtype, ComputeReduceOp op);
    ComputeResult (*execute)(ComputeBackend* backend, const ComputeQuantumOp* op, const ComputeExecutionPlan* plan, ComputeStream* stream);
    ComputeExecutionPlan* (*create_plan)(ComputeBackend* backend, const ComputeQuantumOp* op);
    void (*destroy_plan)(ComputeBackend* backend, ComputeExecutionPlan* plan);
    ComputeResult (*get_metrics)(ComputeBackend* backend, ComputeMetrics* metrics);
    ComputeResult (*reset_metrics)(ComputeBackend* backend);
} ComputeBackendOps;

typedef struct {
    ComputeBackendType type;
    const char* name;
    const char* version;
    int priority;
    const ComputeBackendOps* ops;
} ComputeBackendInfo;

/* ---- compute_backend.h:440-515: the registry / engine API, implemented by the REFERENCE (compute_backend.c) ------ */
ComputeResult compute_register_backend(const ComputeBackendInfo* info);
int compute_get_backend_count(void);
const ComputeBackendInfo* compute_get_backend_info(int index);
const ComputeBackendInfo* compute_get_backend_by_type(ComputeBackendType type);
const ComputeBackendInfo* compute_select_backend(ComputeBackendType preferred, bool allow_fallback);
bool compute_backend_available(ComputeBackendType type);
ComputeEngine* compute_engine_init(const ComputeDistributedConfig* config);
void compute_engine_cleanup(ComputeEngine* engine);
ComputeBackendType compute_engine_get_backend_type(const ComputeEngine* engine);
const ComputeBackendOps* compute_engine_get_ops(const ComputeEngine* engine);
ComputeBackend* compute_engine_get_backend(const ComputeEngine* engine);

/* compute_backend.h:523-534 */
#define COMPUTE_REGISTER_BACKEND(backend_type, backend_name, backend_version, backend_priority, backend_ops)      \
    __attribute__((constructor)) static void register_##backend_type##_backend(void) {                            \
        static const ComputeBackendInfo info = {.type = backend_type, .name = backend_name,                       \
                                                .version = backend_version, .priority = backend_priority,         \
                                                .ops = &backend_ops};                                             \
        compute_register_backend(&info);                                                                           \
    }

/* ---- this backend ------------------------------------------------------------------------------------------------- */
/* Optional ComputeDistributedConfig.backend_config: multi-GPU rendezvous (one process per GPU).  nccl_id is the
 * 128-byte id of qgt_b200_dist_unique_id() made on rank 0 and distributed by the host program. */
typedef struct qgt_b200_backend_config {
    int rank, world;
    const uint8_t* nccl_id;
} qgt_b200_backend_config;

const ComputeBackendOps* qgt_b200_compute_backend_ops(void);
const ComputeBackendInfo* qgt_b200_compute_backend_info(void);
/* registers the table with the reference's registry when it is present in the process; COMPUTE_ERROR_BACKEND_NOT_AVAILABLE otherwise */
ComputeResult qgt_b200_register_compute_backend(void);
/* the qgt_b200_ctx behind a backend instance (for callers that mix the vtable with include/qgt_b200.h) */
struct qgt_b200_ctx;
struct qgt_b200_ctx* qgt_b200_compute_backend_ctx(ComputeBackend* backend);

#ifdef __cplusplus
}
#endif
#endif /* QGT_COMPUTE_BACKEND_H */

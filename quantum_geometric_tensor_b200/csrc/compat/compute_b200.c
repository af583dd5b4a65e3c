/*
 * compute_b200.c — a ComputeBackendOps table over the sm_100a library: the file a maintainer adds next to the
 * reference's src/quantum_geometric/supercomputer/backends/compute_{cpu,cuda,metal,opencl}.c.
 *
 * Seam: include/quantum_geometric/supercomputer/compute_backend.h:18-412 (vtable), :523-534 (registration),
 * consumed by supercomputer/compute_backend.c:41-166 (registry, compute_engine_init -> probe() -> init()).
 * Semantics of every slot follow the reference's CPU backend (backends/compute_cpu.c) — same argument checks and
 * result codes — with device work done by include/qgt_b200.h.  Differences, all deliberate:
 *   - no CPU fallback (the reference's CUDA backend runs sizes < 1024 on the host, compute_cuda.cu:700-760);
 *   - pointers from alloc(COMPUTE_MEM_DEVICE) are used in place instead of being re-staged on every call;
 *   - reductions accumulate in double;
 *   - execute(QUANTUM_OP_UNITARY) understands target qubits (gate-level application, no 2^n x 2^n matrix).
 * One thread per backend instance, as the reference assumes (its registry is unsynchronised global state).
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "qgt_b200.h"
#include "qgt_compute_backend.h"

typedef struct B200Backend {
    qgt_b200_ctx* ctx;
    int device;
    int rank, world;
    size_t bytes_now, bytes_peak;
    ComputeMetrics metrics;
} B200Backend;

/* NULL = the backend's default stream = the stream of its context (everything stays ordered on one stream) */
#define B200_STREAM(b, s) ((s) ? (void*)(s) : qgt_b200_ctx_stream((b)->ctx))

static ComputeResult map_status(int rc) {
    switch (rc) {
    case QGT_B200_OK: return COMPUTE_SUCCESS;
    case QGT_B200_ERR_INVALID_ARG: case QGT_B200_ERR_DIMENSION: return COMPUTE_ERROR_INVALID_ARGUMENT;
    case QGT_B200_ERR_NO_MEMORY: return COMPUTE_ERROR_OUT_OF_MEMORY;
    case QGT_B200_ERR_NO_DEVICE: return COMPUTE_ERROR_DEVICE_NOT_FOUND;
    case QGT_B200_ERR_UNSUPPORTED: return COMPUTE_ERROR_NOT_IMPLEMENTED;
    case QGT_B200_ERR_NOT_INIT: return COMPUTE_ERROR_BACKEND_NOT_AVAILABLE;
    case QGT_B200_ERR_HARDWARE: return COMPUTE_ERROR_KERNEL_FAILED;
    default: return COMPUTE_ERROR_INTERNAL;
    }
}

/* ---- lifecycle -------------------------------------------------------------------------------------------------- */
static bool b200_probe(void) { return qgt_b200_device_count() > 0; }

static ComputeBackend* b200_init(const ComputeDistributedConfig* config) {
    if (!config || !b200_probe()) return NULL;
    B200Backend* b = (B200Backend*)calloc(1, sizeof *b);
    if (!b) return NULL;
    const int ndev = qgt_b200_device_count();
    b->device = config->local_rank >= 0 ? config->local_rank % ndev : 0;
    b->world = 1;
    if (qgt_b200_create(&b->ctx, b->device) != QGT_B200_OK) { free(b); return NULL; }
    const qgt_b200_backend_config* bc = (const qgt_b200_backend_config*)config->backend_config;
    if (bc && bc->world > 1) {
        if (!bc->nccl_id || qgt_b200_dist_init(b->ctx, bc->rank, bc->world, bc->nccl_id) != QGT_B200_OK) {
            fprintf(stderr, "compute_b200: communicator setup failed: %s\n", qgt_b200_last_error());
            qgt_b200_destroy(b->ctx); free(b); return NULL;
        }
        b->rank = bc->rank; b->world = bc->world;
    }
    return (ComputeBackend*)b;
}

static void b200_cleanup(ComputeBackend* be) {
    B200Backend* b = (B200Backend*)be;
    if (!b) return;
    qgt_b200_destroy(b->ctx);
    free(b);
}

static ComputeResult b200_get_capabilities(ComputeBackend* be, int* num_devices, size_t* total_memory) {
    B200Backend* b = (B200Backend*)be;
    if (!b) return COMPUTE_ERROR_INVALID_ARGUMENT;
    qgt_b200_device_info info;
    int rc = qgt_b200_device_info_get(b->device, &info);
    if (rc) return map_status(rc);
    if (num_devices) *num_devices = qgt_b200_device_count();
    if (total_memory) *total_memory = info.total_memory;
    return COMPUTE_SUCCESS;
}

/* ---- memory: a small header-less ledger is enough for the metrics (sizes are the caller's) ---------------------- */
static void* b200_alloc(ComputeBackend* be, size_t size, ComputeMemType mem_type) {
    B200Backend* b = (B200Backend*)be;
    if (!b || size == 0) return NULL;
    void* p = NULL;
    qgt_b200_set_device(b->device);
    int rc;
    switch (mem_type) {
    case COMPUTE_MEM_HOST: p = malloc(size); rc = p ? 0 : QGT_B200_ERR_NO_MEMORY; break;
    case COMPUTE_MEM_PINNED: rc = qgt_b200_mem_alloc_pinned(&p, size); break;
    case COMPUTE_MEM_DEVICE: case COMPUTE_MEM_UNIFIED: rc = qgt_b200_mem_alloc(&p, size); break;   /* never host memory */
    default: return NULL;
    }
    if (rc) return NULL;
    b->bytes_now += size;
    if (b->bytes_now > b->bytes_peak) b->bytes_peak = b->bytes_now;
    return p;
}

static void b200_free(ComputeBackend* be, void* ptr, ComputeMemType mem_type) {
    B200Backend* b = (B200Backend*)be;
    if (!b || !ptr) return;
    qgt_b200_set_device(b->device);
    switch (mem_type) {
    case COMPUTE_MEM_HOST: free(ptr); break;
    case COMPUTE_MEM_PINNED: qgt_b200_mem_free_pinned(ptr); break;
    default: qgt_b200_mem_free(ptr); break;
    }
}

static ComputeResult b200_memcpy(ComputeBackend* be, void* dst, ComputeMemType dst_type, const void* src, ComputeMemType src_type,
                                 size_t size, ComputeStream* stream) {
    B200Backend* b = (B200Backend*)be;
    (void)dst_type; (void)src_type;     /* direction is detected from the pointers */
    if (!b || !dst || !src) return COMPUTE_ERROR_INVALID_ARGUMENT;
    if (size == 0) return COMPUTE_SUCCESS;
    qgt_b200_set_device(b->device);
    int rc = qgt_b200_memcpy_async(dst, src, size, B200_STREAM(b, stream));
    if (!rc && !stream) rc = qgt_b200_stream_synchronize(B200_STREAM(b, stream));     /* blocking without a stream, as cudaMemcpy */
    return map_status(rc);
}

static ComputeResult b200_memset(ComputeBackend* be, void* ptr, int value, size_t size, ComputeStream* stream) {
    B200Backend* b = (B200Backend*)be;
    if (!b || !ptr) return COMPUTE_ERROR_INVALID_ARGUMENT;
    qgt_b200_set_device(b->device);
    return map_status(qgt_b200_memset_async(ptr, value, size, B200_STREAM(b, stream)));
}

/* ---- streams and events ---------------------------------------------------------------------------------------- */
static ComputeStream* b200_create_stream(ComputeBackend* be) {
    B200Backend* b = (B200Backend*)be;
    if (!b) return NULL;
    void* s = NULL;
    qgt_b200_set_device(b->device);
    return qgt_b200_stream_create(&s) == QGT_B200_OK ? (ComputeStream*)s : NULL;
}
static void b200_destroy_stream(ComputeBackend* be, ComputeStream* s) { (void)be; if (s) qgt_b200_stream_destroy((void*)s); }
static ComputeResult b200_synchronize_stream(ComputeBackend* be, ComputeStream* s) {
    B200Backend* b = (B200Backend*)be;
    if (!b) return COMPUTE_ERROR_INVALID_ARGUMENT;
    qgt_b200_set_device(b->device);
    int rc = s ? qgt_b200_stream_synchronize((void*)s) : qgt_b200_device_synchronize();      /* NULL: everything */
    return rc ? COMPUTE_ERROR_SYNCHRONIZATION_FAILED : COMPUTE_SUCCESS;
}
static ComputeEvent* b200_create_event(ComputeBackend* be) {
    B200Backend* b = (B200Backend*)be;
    if (!b) return NULL;
    void* e = NULL;
    qgt_b200_set_device(b->device);
    return qgt_b200_event_create(&e) == QGT_B200_OK ? (ComputeEvent*)e : NULL;
}
static void b200_destroy_event(ComputeBackend* be, ComputeEvent* e) { (void)be; if (e) qgt_b200_event_destroy((void*)e); }
static ComputeResult b200_record_event(ComputeBackend* be, ComputeEvent* e, ComputeStream* s) {
    if (!be || !e) return COMPUTE_ERROR_INVALID_ARGUMENT;
    return qgt_b200_event_record((void*)e, B200_STREAM((B200Backend*)be, s)) ? COMPUTE_ERROR_SYNCHRONIZATION_FAILED : COMPUTE_SUCCESS;
}
static ComputeResult b200_wait_event(ComputeBackend* be, ComputeStream* s, ComputeEvent* e) {
    if (!be || !e) return COMPUTE_ERROR_INVALID_ARGUMENT;
    return qgt_b200_event_wait(B200_STREAM((B200Backend*)be, s), (void*)e) ? COMPUTE_ERROR_SYNCHRONIZATION_FAILED : COMPUTE_SUCCESS;
}

/* ---- quantum operations (argument checks as backends/compute_cpu.c:341-466) ------------------------------------ */
static ComputeResult b200_quantum_unitary(ComputeBackend* be, float* state, size_t state_size, const float* unitary,
                                          size_t unitary_size, ComputeStream* stream) {
    B200Backend* b = (B200Backend*)be;
    if (!b || !state || !unitary || state_size == 0 || unitary_size == 0) return COMPUTE_ERROR_INVALID_ARGUMENT;
    /* unitary_size == state_size: the reference's dense product; a smaller power of two: a gate on the low qubits */
    return map_status(qgt_b200_c64_apply_matrix(b->ctx, state, state_size, unitary, unitary_size, NULL, (void*)stream));
}

static ComputeResult b200_quantum_normalize(ComputeBackend* be, float* state, size_t size, ComputeStream* stream) {
    B200Backend* b = (B200Backend*)be;
    if (!b || !state || size == 0) return COMPUTE_ERROR_INVALID_ARGUMENT;
    return map_status(qgt_b200_c64_normalize(b->ctx, state, size, NULL, (void*)stream));
}

static ComputeResult b200_quantum_tensor_contract(ComputeBackend* be, float* result, const float* a, const float* bm,
                                                  size_t m, size_t n, size_t k, ComputeStream* stream) {
    B200Backend* b = (B200Backend*)be;
    if (!b || !result || !a || !bm) return COMPUTE_ERROR_INVALID_ARGUMENT;
    if (m == 0 || n == 0 || k == 0) return COMPUTE_SUCCESS;          /* the reference's loops do nothing */
    return map_status(qgt_b200_c64_matmul(b->ctx, result, a, bm, m, n, k, (void*)stream));
}

static ComputeResult b200_quantum_inner_product(ComputeBackend* be, float* result, const float* sa, const float* sb,
                                                size_t size, ComputeStream* stream) {
    B200Backend* b = (B200Backend*)be;
    if (!b || !result || !sa || !sb || size == 0) return COMPUTE_ERROR_INVALID_ARGUMENT;
    return map_status(qgt_b200_c64_inner_product(b->ctx, sa, sb, size, result, (void*)stream));
}

/* gradients[0..1] = <backward|forward>  (compute_cpu.c:404-427) */
static ComputeResult b200_quantum_gradient(ComputeBackend* be, float* gradients, const float* forward_state,
                                           const float* backward_state, size_t size, ComputeStream* stream) {
    B200Backend* b = (B200Backend*)be;
    if (!b || !gradients || !forward_state || !backward_state || size == 0) return COMPUTE_ERROR_INVALID_ARGUMENT;
    return map_status(qgt_b200_c64_inner_product(b->ctx, backward_state, forward_state, size, gradients, (void*)stream));
}

static ComputeResult b200_quantum_expectation(ComputeBackend* be, float* result, const float* state, const float* observable,
                                              size_t size, ComputeStream* stream) {
    B200Backend* b = (B200Backend*)be;
    if (!b || !result || !state || !observable || size == 0) return COMPUTE_ERROR_INVALID_ARGUMENT;
    return map_status(qgt_b200_c64_expectation_diag(b->ctx, state, observable, size, result, (void*)stream));
}

/* ---- collectives ------------------------------------------------------------------------------------------------ */
static ComputeResult coll(ComputeBackend* be, int kind, const void* s, void* r, size_t count, ComputeDataType dt, int op, int root) {
    B200Backend* b = (B200Backend*)be;
    if (!b) return COMPUTE_ERROR_INVALID_ARGUMENT;
    int rc = qgt_b200_dist_collective(b->ctx, kind, s, r, count, (int)dt, op, root);
    if (rc == QGT_B200_OK && b->world > 1) {
        b->metrics.num_messages += 1;
        b->metrics.bytes_sent += count * compute_dtype_size(dt);
        b->metrics.bytes_received += count * compute_dtype_size(dt);
    }
    return rc == QGT_B200_ERR_HARDWARE ? COMPUTE_ERROR_COMMUNICATION_FAILED : map_status(rc);
}
static ComputeResult b200_barrier(ComputeBackend* be) {
    B200Backend* b = (B200Backend*)be;
    if (!b) return COMPUTE_ERROR_INVALID_ARGUMENT;
    return qgt_b200_dist_barrier(b->ctx) ? COMPUTE_ERROR_COMMUNICATION_FAILED : COMPUTE_SUCCESS;
}
static ComputeResult b200_broadcast(ComputeBackend* be, void* data, size_t size, ComputeDataType dt, int root) {
    return coll(be, QGT_B200_COLL_BROADCAST, data, data, size, dt, 0, root);
}
static ComputeResult b200_allreduce(ComputeBackend* be, const void* s, void* r, size_t count, ComputeDataType dt, ComputeReduceOp op) {
    return coll(be, QGT_B200_COLL_ALLREDUCE, s, r, count, dt, (int)op, 0);
}
static ComputeResult b200_scatter(ComputeBackend* be, const void* s, void* r, size_t count, ComputeDataType dt, int root) {
    return coll(be, QGT_B200_COLL_SCATTER, s, r, count, dt, 0, root);
}
static ComputeResult b200_gather(ComputeBackend* be, const void* s, void* r, size_t count, ComputeDataType dt, int root) {
    return coll(be, QGT_B200_COLL_GATHER, s, r, count, dt, 0, root);
}
static ComputeResult b200_allgather(ComputeBackend* be, const void* s, void* r, size_t count, ComputeDataType dt) {
    return coll(be, QGT_B200_COLL_ALLGATHER, s, r, count, dt, 0, 0);
}
static ComputeResult b200_reduce_scatter(ComputeBackend* be, const void* s, void* r, size_t count, ComputeDataType dt, ComputeReduceOp op) {
    return coll(be, QGT_B200_COLL_REDUCE_SCATTER, s, r, count, dt, (int)op, 0);
}

/* ---- execute / plans (dispatch as compute_cpu.c:664-740) --------------------------------------------------------- */
static ComputeResult b200_execute(ComputeBackend* be, const ComputeQuantumOp* op, const ComputeExecutionPlan* plan, ComputeStream* stream) {
    B200Backend* b = (B200Backend*)be;
    (void)plan;
    if (!b || !op) return COMPUTE_ERROR_INVALID_ARGUMENT;
    ComputeResult r = COMPUTE_SUCCESS;
    switch (op->type) {
    case QUANTUM_OP_UNITARY:
        if (op->num_targets > 0 && op->target_qubits) {
            /* gate-level: parameters = 2^k x 2^k complex float matrix on target_qubits, state = output_data */
            if (!op->output_data || !op->parameters || op->output_size == 0 || op->num_targets > 4) {
                r = op->num_targets > 4 ? COMPUTE_ERROR_NOT_IMPLEMENTED : COMPUTE_ERROR_INVALID_ARGUMENT;
                break;
            }
            int32_t tg[4];
            for (size_t j = 0; j < op->num_targets; j++) tg[j] = (int32_t)op->target_qubits[j];
            r = map_status(qgt_b200_c64_apply_matrix(b->ctx, (float*)op->output_data, op->output_size, (const float*)op->parameters,
                                                     (size_t)1 << op->num_targets, tg, (void*)stream));
        } else {
            r = b200_quantum_unitary(be, (float*)op->output_data, op->output_size, (const float*)op->parameters, op->param_size, stream);
        }
        break;
    case QUANTUM_OP_NORMALIZE:
        r = b200_quantum_normalize(be, (float*)op->output_data, op->output_size, stream);
        break;
    case QUANTUM_OP_TENSOR_CONTRACT:
        if (op->num_dims >= 3 && op->dims)
            r = b200_quantum_tensor_contract(be, (float*)op->output_data, (const float*)op->input_data, (const float*)op->parameters,
                                             op->dims[0], op->dims[1], op->dims[2], stream);
        break;
    case QUANTUM_OP_GRADIENT:
        r = b200_quantum_gradient(be, (float*)op->output_data, (const float*)op->input_data, (const float*)op->parameters, op->input_size, stream);
        break;
    case QUANTUM_OP_INNER_PRODUCT:
        r = b200_quantum_inner_product(be, (float*)op->output_data, (const float*)op->input_data, (const float*)op->parameters, op->input_size, stream);
        break;
    case QUANTUM_OP_EXPECTATION:
        r = b200_quantum_expectation(be, (float*)op->output_data, (const float*)op->input_data, (const float*)op->parameters, op->input_size, stream);
        break;
    default:
        r = COMPUTE_ERROR_NOT_IMPLEMENTED;
        break;
    }
    b->metrics.operations_per_second += 1.0;       /* the reference counts executed ops in this field */
    return r;
}

static ComputeExecutionPlan* b200_create_plan(ComputeBackend* be, const ComputeQuantumOp* op) {
    B200Backend* b = (B200Backend*)be;
    if (!b || !op) return NULL;
    ComputeExecutionPlan* p = (ComputeExecutionPlan*)calloc(1, sizeof *p);
    if (!p) return NULL;
    const size_t parts = (size_t)b->world;         /* one partition per rank: the top-qubit sharding of the library */
    p->num_partitions = parts;
    p->partition_size = op->input_size / parts;
    p->node_assignments = (int*)calloc(parts, sizeof(int));
    p->offsets = (size_t*)calloc(parts, sizeof(size_t));
    p->sizes = (size_t*)calloc(parts, sizeof(size_t));
    if (!p->node_assignments || !p->offsets || !p->sizes) {
        free(p->node_assignments); free(p->offsets); free(p->sizes); free(p);
        return NULL;
    }
    for (size_t i = 0; i < parts; i++) {
        p->node_assignments[i] = (int)i;
        p->offsets[i] = i * p->partition_size;
        p->sizes[i] = i + 1 == parts ? op->input_size - i * p->partition_size : p->partition_size;
    }
    return p;
}

static void b200_destroy_plan(ComputeBackend* be, ComputeExecutionPlan* p) {
    (void)be;
    if (!p) return;
    free(p->node_assignments); free(p->offsets); free(p->sizes); free(p->send_targets); free(p->recv_sources); free(p->workspace);
    free(p);
}

static ComputeResult b200_get_metrics(ComputeBackend* be, ComputeMetrics* m) {
    B200Backend* b = (B200Backend*)be;
    if (!b || !m) return COMPUTE_ERROR_INVALID_ARGUMENT;
    *m = b->metrics;
    m->peak_memory_bytes = b->bytes_peak;
    m->current_memory_bytes = b->bytes_now;
    m->memory_used = b->bytes_now;
    return COMPUTE_SUCCESS;
}
static ComputeResult b200_reset_metrics(ComputeBackend* be) {
    B200Backend* b = (B200Backend*)be;
    if (!b) return COMPUTE_ERROR_INVALID_ARGUMENT;
    memset(&b->metrics, 0, sizeof b->metrics);
    return COMPUTE_SUCCESS;
}

static const ComputeBackendOps b200_ops = {
    .init = b200_init, .cleanup = b200_cleanup, .probe = b200_probe, .get_capabilities = b200_get_capabilities,
    .alloc = b200_alloc, .free = b200_free, .memcpy = b200_memcpy, .memset = b200_memset,
    .create_stream = b200_create_stream, .destroy_stream = b200_destroy_stream, .synchronize_stream = b200_synchronize_stream,
    .create_event = b200_create_event, .destroy_event = b200_destroy_event, .record_event = b200_record_event, .wait_event = b200_wait_event,
    .quantum_unitary = b200_quantum_unitary, .quantum_normalize = b200_quantum_normalize,
    .quantum_tensor_contract = b200_quantum_tensor_contract, .quantum_gradient = b200_quantum_gradient,
    .quantum_inner_product = b200_quantum_inner_product, .quantum_expectation = b200_quantum_expectation,
    .barrier = b200_barrier, .broadcast = b200_broadcast, .allreduce = b200_allreduce, .scatter = b200_scatter,
    .gather = b200_gather, .allgather = b200_allgather, .reduce_scatter = b200_reduce_scatter,
    .execute = b200_execute, .create_plan = b200_create_plan, .destroy_plan = b200_destroy_plan,
    .get_metrics = b200_get_metrics, .reset_metrics = b200_reset_metrics,
};

static const ComputeBackendInfo b200_info = {
    .type = COMPUTE_BACKEND_CUDA, .name = "b200", .version = "sm_100a statevector + QGT", .priority = 100, .ops = &b200_ops,
};

const ComputeBackendOps* qgt_b200_compute_backend_ops(void) { return &b200_ops; }
const ComputeBackendInfo* qgt_b200_compute_backend_info(void) { return &b200_info; }
struct qgt_b200_ctx* qgt_b200_compute_backend_ctx(ComputeBackend* be) { return be ? ((B200Backend*)be)->ctx : NULL; }

/* The registry lives in the reference (supercomputer/compute_backend.c:41-64).  Weak: this library also loads into
 * processes that do not carry it. */
extern ComputeResult compute_register_backend(const ComputeBackendInfo* info) __attribute__((weak));

ComputeResult qgt_b200_register_compute_backend(void) {
    if (!compute_register_backend) return COMPUTE_ERROR_BACKEND_NOT_AVAILABLE;
    return compute_register_backend(&b200_info);
}

/* what COMPUTE_REGISTER_BACKEND(COMPUTE_BACKEND_CUDA, "b200", ..., 100, b200_ops) expands to, minus the hard link */
__attribute__((constructor)) static void register_COMPUTE_BACKEND_CUDA_backend(void) { (void)qgt_b200_register_compute_backend(); }

// capi.cu — the C-ABI of include/qgt_b200.h: context, device memory, program executor.
//
// No CPU fallback: every compute entry point needs a CUDA device and reports
// QGT_B200_ERR_NO_DEVICE otherwise (north star: "backend dispatch ... targets only the new sm_100a
// library with no CPU fallback"; contrast gpu_malloc's malloc fallback in the reference,
// src/quantum_geometric/core/quantum_geometric_gpu.c:30-61).
#include <cuda_runtime.h>

#include <algorithm>
#include <chrono>
#include <thread>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <string>
#include <vector>

#include "../../include/qgt_b200.h"
#include "ctx.hpp"
#include "kernels.cuh"
#include "plan.hpp"

using namespace qgt;

namespace qgt {
thread_local std::string g_last_error;

int fail(int status, const std::string& msg) {
    g_last_error = msg;
    return status;
}

int cuda_fail(cudaError_t e, const char* what) {
    char buf[256];
    snprintf(buf, sizeof buf, "%s: %s", what, cudaGetErrorString(e));
    g_last_error = buf;
    return e == cudaErrorMemoryAllocation ? QGT_B200_ERR_NO_MEMORY : QGT_B200_ERR_HARDWARE;
}

int DevBuf::reserve(size_t n) {
    if (n <= bytes) return QGT_B200_OK;
    if (ptr) cudaFree(ptr);
    ptr = nullptr; bytes = 0;
    cudaError_t e = cudaMalloc(&ptr, n);
    if (e != cudaSuccess) return cuda_fail(e, "cudaMalloc");
    bytes = n;
    return QGT_B200_OK;
}

void DevBuf::release() {
    if (ptr) cudaFree(ptr);
    ptr = nullptr; bytes = 0;
}

void Timer::begin(cudaStream_t st, int cat, const char* label) {
    if (!enabled) return;
    if (trace) labels.push_back(label ? label : "");
    if (used + 2 > events.size()) {
        const size_t old = events.size();
        events.resize(old + 64);
        for (size_t i = old; i < events.size(); i++) cudaEventCreate(&events[i]);
    }
    cats.push_back(cat);
    cudaEventRecord(events[used], st);
}

void Timer::end(cudaStream_t st) {
    if (!enabled) return;
    cudaEventRecord(events[used + 1], st);
    used += 2;
}

void Timer::collect(double ms[4]) {
    ms[0] = ms[1] = ms[2] = ms[3] = 0.0;
    for (size_t i = 0; i < cats.size(); i++) {
        float t = 0.f;
        cudaEventElapsedTime(&t, events[2 * i], events[2 * i + 1]);
        ms[cats[i]] += t;
        if (trace) fprintf(stderr, "[qgt_b200 trace] cat=%d ms=%.4f %s\n", cats[i], t, i < labels.size() ? labels[i].c_str() : "");
    }
    labels.clear();
    cats.clear();
    used = 0;
}

Timer::~Timer() {
    for (cudaEvent_t e : events) cudaEventDestroy(e);
}
}  // namespace qgt

// ------------------------------------------------------------------------------------------------
extern "C" {

int qgt_b200_abi_version(void) { return QGT_B200_ABI_VERSION; }

int qgt_b200_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
    int ok = 0;
    for (int d = 0; d < n; d++) {
        cudaDeviceProp p;
        if (cudaGetDeviceProperties(&p, d) == cudaSuccess && p.major == 10) ok++;
    }
    return ok;
}

const char* qgt_b200_last_error(void) { return g_last_error.c_str(); }

// ---- raw device memory / streams (the reference's qg_gpu_* seam) -----------------------------------------------
static int raw_need_device(const char* what) {
    if (qgt_b200_device_count() <= 0) return fail(QGT_B200_ERR_NO_DEVICE, std::string(what) + ": no sm_100 device (there is no host fallback)");
    return QGT_B200_OK;
}

int qgt_b200_device_info_get(int device, qgt_b200_device_info* out) {
    if (!out) return fail(QGT_B200_ERR_INVALID_ARG, "device_info: out is NULL");
    int rc = raw_need_device("device_info");
    if (rc) return rc;
    cudaDeviceProp p;
    if (cudaGetDeviceProperties(&p, device) != cudaSuccess) { cudaGetLastError(); return fail(QGT_B200_ERR_INVALID_ARG, "device_info: bad device ordinal"); }
    std::memset(out, 0, sizeof *out);
    out->device = device;
    std::snprintf(out->name, sizeof out->name, "%s", p.name);
    out->total_memory = p.totalGlobalMem;
    int cur = 0;
    cudaGetDevice(&cur);
    size_t fr = 0, tot = 0;
    if (cudaSetDevice(device) == cudaSuccess && cudaMemGetInfo(&fr, &tot) == cudaSuccess) out->free_memory = fr;
    cudaSetDevice(cur);
    out->cc_major = p.major; out->cc_minor = p.minor;
    out->num_sms = p.multiProcessorCount;
    out->max_threads_per_block = p.maxThreadsPerBlock;
    for (int i = 0; i < 3; i++) { out->max_block_dim[i] = p.maxThreadsDim[i]; out->max_grid_dim[i] = p.maxGridSize[i]; }
    out->unified_addressing = p.unifiedAddressing;
    return QGT_B200_OK;
}

int qgt_b200_set_device(int device) {
    int rc = raw_need_device("set_device");
    if (rc) return rc;
    if (cudaSetDevice(device) != cudaSuccess) { cudaGetLastError(); return fail(QGT_B200_ERR_INVALID_ARG, "set_device: bad device ordinal"); }
    return QGT_B200_OK;
}

int qgt_b200_mem_alloc(void** ptr, size_t bytes) {
    if (!ptr || bytes == 0) return fail(QGT_B200_ERR_INVALID_ARG, "mem_alloc: ptr is NULL or size is 0");
    *ptr = nullptr;
    int rc = raw_need_device("mem_alloc");
    if (rc) return rc;
    cudaError_t e = cudaMalloc(ptr, bytes);
    if (e == cudaErrorMemoryAllocation) { cudaGetLastError(); return fail(QGT_B200_ERR_NO_MEMORY, "mem_alloc: out of device memory"); }
    if (e != cudaSuccess) return cuda_fail(e, "mem_alloc");
    return QGT_B200_OK;
}

int qgt_b200_mem_free(void* ptr) {
    if (!ptr) return QGT_B200_OK;
    cudaError_t e = cudaFree(ptr);
    return e == cudaSuccess ? QGT_B200_OK : cuda_fail(e, "mem_free");
}

int qgt_b200_mem_alloc_pinned(void** ptr, size_t bytes) {
    if (!ptr || bytes == 0) return fail(QGT_B200_ERR_INVALID_ARG, "mem_alloc_pinned: ptr is NULL or size is 0");
    *ptr = nullptr;
    int rc = raw_need_device("mem_alloc_pinned");
    if (rc) return rc;
    cudaError_t e = cudaMallocHost(ptr, bytes);
    if (e != cudaSuccess) { cudaGetLastError(); return fail(QGT_B200_ERR_NO_MEMORY, "mem_alloc_pinned: cudaMallocHost failed"); }
    return QGT_B200_OK;
}

int qgt_b200_mem_free_pinned(void* ptr) {
    if (!ptr) return QGT_B200_OK;
    cudaError_t e = cudaFreeHost(ptr);
    return e == cudaSuccess ? QGT_B200_OK : cuda_fail(e, "mem_free_pinned");
}

int qgt_b200_memcpy_h2d(void* dst, const void* src, size_t bytes) {
    if (!dst || !src) return fail(QGT_B200_ERR_INVALID_ARG, "memcpy_h2d: NULL pointer");
    cudaError_t e = cudaMemcpy(dst, src, bytes, cudaMemcpyHostToDevice);
    return e == cudaSuccess ? QGT_B200_OK : cuda_fail(e, "memcpy_h2d");
}

int qgt_b200_memcpy_d2h(void* dst, const void* src, size_t bytes) {
    if (!dst || !src) return fail(QGT_B200_ERR_INVALID_ARG, "memcpy_d2h: NULL pointer");
    cudaError_t e = cudaMemcpy(dst, src, bytes, cudaMemcpyDeviceToHost);
    return e == cudaSuccess ? QGT_B200_OK : cuda_fail(e, "memcpy_d2h");
}

int qgt_b200_stream_create(void** stream) {
    if (!stream) return fail(QGT_B200_ERR_INVALID_ARG, "stream_create: stream is NULL");
    *stream = nullptr;
    int rc = raw_need_device("stream_create");
    if (rc) return rc;
    cudaStream_t s;
    cudaError_t e = cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking);
    if (e != cudaSuccess) return cuda_fail(e, "stream_create");
    *stream = (void*)s;
    return QGT_B200_OK;
}

int qgt_b200_stream_destroy(void* stream) {
    if (!stream) return QGT_B200_OK;
    cudaError_t e = cudaStreamDestroy((cudaStream_t)stream);
    return e == cudaSuccess ? QGT_B200_OK : cuda_fail(e, "stream_destroy");
}

int qgt_b200_stream_synchronize(void* stream) {
    cudaError_t e = cudaStreamSynchronize((cudaStream_t)stream);
    return e == cudaSuccess ? QGT_B200_OK : cuda_fail(e, "stream_synchronize");
}

int qgt_b200_device_synchronize(void) {
    int rc = raw_need_device("device_synchronize");
    if (rc) return rc;
    cudaError_t e = cudaDeviceSynchronize();
    return e == cudaSuccess ? QGT_B200_OK : cuda_fail(e, "device_synchronize");
}

int qgt_b200_memcpy_async(void* dst, const void* src, size_t bytes, void* stream) {
    if (!dst || !src) return fail(QGT_B200_ERR_INVALID_ARG, "memcpy_async: NULL pointer");
    cudaError_t e = cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDefault, (cudaStream_t)stream);
    return e == cudaSuccess ? QGT_B200_OK : cuda_fail(e, "memcpy_async");
}

int qgt_b200_memset_async(void* ptr, int value, size_t bytes, void* stream) {
    if (!ptr) return fail(QGT_B200_ERR_INVALID_ARG, "memset_async: NULL pointer");
    cudaPointerAttributes a;
    if (cudaPointerGetAttributes(&a, ptr) != cudaSuccess || a.type == cudaMemoryTypeUnregistered || a.type == cudaMemoryTypeHost) {
        cudaGetLastError();
        std::memset(ptr, value, bytes);            // plain or pinned host memory
        return QGT_B200_OK;
    }
    cudaError_t e = cudaMemsetAsync(ptr, value, bytes, (cudaStream_t)stream);
    return e == cudaSuccess ? QGT_B200_OK : cuda_fail(e, "memset_async");
}

int qgt_b200_event_create(void** event) {
    if (!event) return fail(QGT_B200_ERR_INVALID_ARG, "event_create: event is NULL");
    *event = nullptr;
    int rc = raw_need_device("event_create");
    if (rc) return rc;
    cudaEvent_t ev;
    cudaError_t e = cudaEventCreateWithFlags(&ev, cudaEventDisableTiming);
    if (e != cudaSuccess) return cuda_fail(e, "event_create");
    *event = (void*)ev;
    return QGT_B200_OK;
}

int qgt_b200_event_destroy(void* event) {
    if (!event) return QGT_B200_OK;
    cudaError_t e = cudaEventDestroy((cudaEvent_t)event);
    return e == cudaSuccess ? QGT_B200_OK : cuda_fail(e, "event_destroy");
}

int qgt_b200_event_record(void* event, void* stream) {
    if (!event) return fail(QGT_B200_ERR_INVALID_ARG, "event_record: event is NULL");
    cudaError_t e = cudaEventRecord((cudaEvent_t)event, (cudaStream_t)stream);
    return e == cudaSuccess ? QGT_B200_OK : cuda_fail(e, "event_record");
}

int qgt_b200_event_wait(void* stream, void* event) {
    if (!event) return fail(QGT_B200_ERR_INVALID_ARG, "event_wait: event is NULL");
    cudaError_t e = cudaStreamWaitEvent((cudaStream_t)stream, (cudaEvent_t)event, 0);
    return e == cudaSuccess ? QGT_B200_OK : cuda_fail(e, "event_wait");
}

int qgt_b200_event_synchronize(void* event) {
    if (!event) return fail(QGT_B200_ERR_INVALID_ARG, "event_synchronize: event is NULL");
    cudaError_t e = cudaEventSynchronize((cudaEvent_t)event);
    return e == cudaSuccess ? QGT_B200_OK : cuda_fail(e, "event_synchronize");
}

void* qgt_b200_ctx_stream(qgt_b200_ctx* c) { return c ? (void*)c->stream : nullptr; }

const char* qgt_b200_error_string(int s) {
    switch (s) {
    case QGT_B200_OK: return "success";
    case QGT_B200_ERR_INVALID_ARG: return "invalid argument";
    case QGT_B200_ERR_NO_MEMORY: return "out of memory";
    case QGT_B200_ERR_DIMENSION: return "dimension mismatch";
    case QGT_B200_ERR_INVALID_STATE: return "invalid state";
    case QGT_B200_ERR_HARDWARE: return "CUDA failure";
    case QGT_B200_ERR_UNSUPPORTED: return "not implemented";
    case QGT_B200_ERR_INTERNAL: return "internal error";
    case QGT_B200_ERR_NOT_INIT: return "not initialized";
    case QGT_B200_ERR_CIRCUIT: return "invalid circuit";
    case QGT_B200_ERR_NO_DEVICE: return "no sm_100 CUDA device (this library has no CPU fallback)";
    default: return "unknown status";
    }
}

int qgt_b200_create(qgt_b200_ctx** out, int device) {
    if (!out) return fail(QGT_B200_ERR_INVALID_ARG, "out is NULL");
    *out = nullptr;
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess || n == 0) {
        cudaGetLastError();
        return fail(QGT_B200_ERR_NO_DEVICE, "no CUDA device visible; qgt_b200 has no CPU fallback");
    }
    if (device < 0 || device >= n) return fail(QGT_B200_ERR_INVALID_ARG, "device ordinal out of range");
    cudaDeviceProp p;
    cudaError_t e = cudaGetDeviceProperties(&p, device);
    if (e != cudaSuccess) return cuda_fail(e, "cudaGetDeviceProperties");
    if (p.major != 10) {
        char buf[160];
        snprintf(buf, sizeof buf, "device %d is sm_%d%d; this library ships sm_100a code only", device, p.major, p.minor);
        return fail(QGT_B200_ERR_NO_DEVICE, buf);
    }
    e = cudaSetDevice(device);
    if (e != cudaSuccess) return cuda_fail(e, "cudaSetDevice");
    qgt_b200_ctx* c = new qgt_b200_ctx();
    c->device = device;
    c->num_sms = p.multiProcessorCount;
    e = cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking);
    if (e != cudaSuccess) { delete c; return cuda_fail(e, "cudaStreamCreate"); }
    cudaEventCreate(&c->ev0);
    cudaEventCreate(&c->ev1);
    {
        int lo = 0, hi = 0;
        cudaDeviceGetStreamPriorityRange(&lo, &hi);           // hi = numerically lowest = greatest priority
        e = cudaStreamCreateWithPriority(&c->side_stream, cudaStreamNonBlocking, hi);
        if (e != cudaSuccess) { cudaStreamDestroy(c->stream); delete c; return cuda_fail(e, "cudaStreamCreateWithPriority"); }
        for (int k = 0; k < 2; k++) {
            cudaEventCreateWithFlags(&c->ev_side[k], cudaEventDisableTiming);
            cudaEventCreateWithFlags(&c->ev_main[k], cudaEventDisableTiming);
        }
        cudaEventCreateWithFlags(&c->ev_group, cudaEventDisableTiming);
    }
    if (const char* tr = getenv("QGT_B200_TRACE")) c->timer.trace = (tr[0] == '1');
    *out = c;
    return QGT_B200_OK;
}

void qgt_b200_destroy(qgt_b200_ctx* c) {
    if (!c) return;
    cudaSetDevice(c->device);
    cudaStreamSynchronize(c->stream);
    qgt::dist_shutdown(c);
    c->arena.release(); c->img_runs.release(); c->img_subs.release(); c->img_stages.release();
    c->img_tdiags.release(); c->img_costs.release(); c->img_pool.release(); c->ovr_pool.release();
    c->items.release(); c->aux.release(); c->partial.release(); c->cmat.release(); c->outbuf.release();
    c->edges.release(); c->vweights.release(); c->scratch.release();
    c->fx_pool.release(); c->fx_tab.release(); c->rho.release(); c->rho_self.release(); c->amat.release();
    if (c->pinned) cudaFreeHost(c->pinned);
    cudaEventDestroy(c->ev0); cudaEventDestroy(c->ev1);
    c->partial_side.release(); c->rho_side.release(); c->cost_phase.release();
    for (int k = 0; k < 2; k++) { cudaEventDestroy(c->ev_side[k]); cudaEventDestroy(c->ev_main[k]); }
    cudaEventDestroy(c->ev_group);
    cudaStreamSynchronize(c->side_stream);
    cudaStreamDestroy(c->side_stream);
    cudaStreamDestroy(c->stream);
    delete c;
}

int qgt_b200_set_workspace_limit(qgt_b200_ctx* c, size_t bytes) {
    if (!c) return fail(QGT_B200_ERR_INVALID_ARG, "ctx is NULL");
    c->ws_limit = bytes;
    return QGT_B200_OK;
}

int qgt_b200_set_option(qgt_b200_ctx* c, const char* key, double value) {
    if (!c || !key) return fail(QGT_B200_ERR_INVALID_ARG, "ctx/key is NULL");
    const std::string k(key);
    if (k == "tile_qubits") { if (value) c->opt.tile_qubits = (int)value; }
    else if (k == "low_qubits") c->opt.low_qubits = (int)value;
    else if (k == "birth_cut") c->opt.birth_cut = (int)value;
    else if (k == "parity_form") c->opt.parity_form = (int)value;
    else if (k == "wavefront") c->opt.wavefront = (int)value;
    else if (k == "max_ops_per_run") c->opt.max_ops_per_run = (int)value;
    else if (k == "reg_qubits") c->opt.reg_qubits = (int)value;
    else if (k == "batch_qubits") c->opt.batch_qubits = (int)value;
    else if (k == "use_mma") c->use_mma = value != 0;
    else if (k == "double_buffer") c->double_buffer = value != 0;
    else if (k == "debug_skip") c->debug_skip = (int)value;
    else if (k == "tiles_per_item") c->tiles_per_item = (int)value;
    else if (k == "gram_tile") set_gram_tile_override((int)value);
    else if (k == "profile") c->timer.enabled = value != 0;
    else if (k == "max_slots") c->max_slots = (size_t)value;
    else if (k == "fused") c->fused_mode = (int)value;
    else if (k == "fused_traj") c->fused_traj = (int)value;
    else if (k == "fused_overlap") c->fused_overlap = (int)value;
    else if (k == "cost_tables") c->cost_tables = (int)value;
    else if (k == "fused_debug") c->fused_debug = (int)value;
    else if (k == "fused_pipeline") c->fused_pipeline = (int)value;
    else return fail(QGT_B200_ERR_INVALID_ARG, "unknown option " + k);
    return QGT_B200_OK;
}

int qgt_b200_get_stats(qgt_b200_ctx* c, qgt_b200_stats* out) {
    if (!c || !out) return fail(QGT_B200_ERR_INVALID_ARG, "ctx/out is NULL");
    *out = c->stats;
    return QGT_B200_OK;
}

// ---- statevector ---------------------------------------------------------------------------------
int qgt_b200_state_create(qgt_b200_ctx* c, int n, qgt_b200_state** out) {
    if (!c || !out) return fail(QGT_B200_ERR_INVALID_ARG, "ctx/out is NULL");
    *out = nullptr;
    if (n < 1 || n > QGT_MAX_QUBITS) return fail(QGT_B200_ERR_INVALID_ARG, "num_qubits out of range");
    int gbits = 0;
    while ((1 << gbits) < c->world) gbits++;
    if (n - gbits < 1) return fail(QGT_B200_ERR_INVALID_ARG, "state too small for this many ranks");
    cudaSetDevice(c->device);
    qgt_b200_state* s = new qgt_b200_state();
    s->ctx = c; s->n = n; s->nloc = n - gbits; s->D = (uint64_t)1 << s->nloc;
    cudaError_t e = cudaMalloc((void**)&s->d, s->D * sizeof(cplx));
    if (e != cudaSuccess) { delete s; return cuda_fail(e, "cudaMalloc(state)"); }
    s->owns = true;
    *out = s;
    return qgt_b200_state_init(s, QGT_B200_INIT_ZERO);
}

void qgt_b200_state_destroy(qgt_b200_state* s) {
    if (!s) return;
    cudaSetDevice(s->ctx->device);
    cudaStreamSynchronize(s->ctx->stream);
    if (s->owns && s->d) cudaFree(s->d);
    delete s;
}

int qgt_b200_state_init(qgt_b200_state* s, int initial_state) {
    if (!s) return fail(QGT_B200_ERR_INVALID_ARG, "state is NULL");
    if (initial_state != QGT_B200_INIT_ZERO && initial_state != QGT_B200_INIT_PLUS)
        return fail(QGT_B200_ERR_INVALID_ARG, "unknown initial state");
    qgt_b200_ctx* c = s->ctx;
    cudaSetDevice(c->device);
    const double amp = std::pow(2.0, -0.5 * s->n);
    cudaError_t e = launch_init_state(s->d, s->D, initial_state, amp, (uint64_t)c->rank * s->D, c->stream);
    if (e != cudaSuccess) return cuda_fail(e, "init kernel");
    e = cudaStreamSynchronize(c->stream);
    if (e != cudaSuccess) return cuda_fail(e, "init sync");
    return QGT_B200_OK;
}

int qgt_b200_state_upload(qgt_b200_state* s, const double* host) {
    if (!s || !host) return fail(QGT_B200_ERR_INVALID_ARG, "state/host is NULL");
    cudaSetDevice(s->ctx->device);
    cudaError_t e = cudaMemcpyAsync(s->d, host, s->D * sizeof(cplx), cudaMemcpyHostToDevice, s->ctx->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(s->ctx->stream);
    if (e != cudaSuccess) return cuda_fail(e, "state upload");
    return QGT_B200_OK;
}

int qgt_b200_state_download(const qgt_b200_state* s, double* host) {
    if (!s || !host) return fail(QGT_B200_ERR_INVALID_ARG, "state/host is NULL");
    cudaSetDevice(s->ctx->device);
    cudaError_t e = cudaMemcpyAsync(host, s->d, s->D * sizeof(cplx), cudaMemcpyDeviceToHost, s->ctx->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(s->ctx->stream);
    if (e != cudaSuccess) return cuda_fail(e, "state download");
    return QGT_B200_OK;
}

int qgt_b200_state_norm2(const qgt_b200_state* s, double* out) {
    if (!s || !out) return fail(QGT_B200_ERR_INVALID_ARG, "state/out is NULL");
    qgt_b200_ctx* c = s->ctx;
    cudaSetDevice(c->device);
    int rc = c->scratch.reserve(256);
    if (rc) return rc;
    cudaError_t e = launch_norm2(s->d, s->D, (double*)c->scratch.ptr, c->stream);
    if (e != cudaSuccess) return cuda_fail(e, "norm kernel");
    double v = 0;
    e = cudaMemcpyAsync(&v, c->scratch.ptr, sizeof v, cudaMemcpyDeviceToHost, c->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(c->stream);
    if (e != cudaSuccess) return cuda_fail(e, "norm readback");
    rc = qgt::dist_allreduce_host(c, &v, 1);
    if (rc) return rc;
    *out = v;
    return QGT_B200_OK;
}

void* qgt_b200_state_device_ptr(qgt_b200_state* s) { return s ? (void*)s->d : nullptr; }

}  // extern "C"

// ------------------------------------------------------------------------------------------------
// executor
// ------------------------------------------------------------------------------------------------
namespace qgt {

static int check_circuit(const qgt_b200_circuit* circ, const double* theta) {
    if (!circ) return fail(QGT_B200_ERR_INVALID_ARG, "circuit is NULL");
    if (circ->num_params > 0 && !theta) return fail(QGT_B200_ERR_INVALID_ARG, "theta is NULL");
    if (circ->num_params < 0) return fail(QGT_B200_ERR_INVALID_ARG, "num_params < 0");
    return QGT_B200_OK;
}

// the circuit's diagonal cost observable (edge list, vertex weights) on the device: c->cost
int upload_cost_table(qgt_b200_ctx* c, const qgt_b200_circuit& circ) {
    int rc;
    cudaError_t e = cudaSuccess;
    c->seg_cost.clear();
    c->cost.edges = nullptr; c->cost.num_edges = 0; c->cost.vertex_weights = nullptr; c->cost.n = circ.num_qubits;
    if (circ.num_edges > QGT_COST_MAX_EDGES) return fail(QGT_B200_ERR_UNSUPPORTED, "more than 1024 cost-layer edges");
    if (circ.num_edges && circ.edges) {
        std::vector<QgtDevEdge> ed(circ.num_edges);
        for (size_t k = 0; k < circ.num_edges; k++) {
            if (circ.edges[k].i < 0 || circ.edges[k].i >= circ.num_qubits || circ.edges[k].j < 0 || circ.edges[k].j >= circ.num_qubits)
                return fail(QGT_B200_ERR_CIRCUIT, "edge endpoint out of range");
            ed[k].i = circ.edges[k].i; ed[k].j = circ.edges[k].j; ed[k].w = circ.edges[k].weight;
        }
        if ((rc = c->edges.reserve(ed.size() * sizeof(QgtDevEdge)))) return rc;
        e = cudaMemcpyAsync(c->edges.ptr, ed.data(), ed.size() * sizeof(QgtDevEdge), cudaMemcpyHostToDevice, c->stream);
        if (e == cudaSuccess) e = cudaStreamSynchronize(c->stream);   // `ed` is a local
        if (e != cudaSuccess) return cuda_fail(e, "edge upload");
        c->cost.edges = (const QgtDevEdge*)c->edges.ptr; c->cost.num_edges = (int)ed.size();
    }
    if (circ.vertex_weights) {
        if ((rc = c->vweights.reserve(circ.num_qubits * sizeof(double)))) return rc;
        e = cudaMemcpyAsync(c->vweights.ptr, circ.vertex_weights, circ.num_qubits * sizeof(double), cudaMemcpyHostToDevice, c->stream);
        if (e == cudaSuccess) e = cudaStreamSynchronize(c->stream);
        if (e != cudaSuccess) return cuda_fail(e, "vertex weight upload");
        c->cost.vertex_weights = (const double*)c->vweights.ptr;
    }
    return QGT_B200_OK;
}

// upload the device image of a plan and the circuit's cost table
int upload_plan(qgt_b200_ctx* c, const qgt_b200_circuit& circ, const CircuitPlan& plan, PlanImage& img) {
    build_image(plan, img);
    c->img_stage_form.clear(); c->img_run_stage_off.clear();
    for (const QgtDevStage& st : img.stages) c->img_stage_form.push_back(st.form);
    for (const QgtDevRun& r : img.runs) c->img_run_stage_off.push_back(r.stage_off);
    int rc;
    struct Up { DevBuf* buf; const void* src; size_t bytes; };
    const Up ups[] = {
        {&c->img_runs, img.runs.data(), img.runs.size() * sizeof(QgtDevRun)},
        {&c->img_subs, img.subs.data(), img.subs.size() * sizeof(QgtDevSubPass)},
        {&c->img_stages, img.stages.data(), img.stages.size() * sizeof(QgtDevStage)},
        {&c->img_tdiags, img.tdiags.data(), img.tdiags.size() * sizeof(QgtDevThrDiag)},
        {&c->img_costs, img.costs.data(), img.costs.size() * sizeof(QgtDevCost)},
        {&c->img_pool, img.pool.data(), img.pool.size() * sizeof(double)},
    };
    cudaError_t e = cudaSuccess;
    for (const Up& u : ups) {
        if ((rc = u.buf->reserve(std::max<size_t>(16, u.bytes)))) return rc;
        if (u.bytes && e == cudaSuccess) e = cudaMemcpyAsync(u.buf->ptr, u.src, u.bytes, cudaMemcpyHostToDevice, c->stream);
    }
    if (e != cudaSuccess) return cuda_fail(e, "plan upload");
    return upload_cost_table(c, circ);
}

struct SweepBatch { int run; size_t item_off; int nitems; int out_of_place; };

// one fused sweep launch over `cols` (slot pointers resolved by the caller)
static int do_sweep(qgt_b200_ctx* c, const CircuitPlan& plan, int run, const QgtSweepItem* d_items, int nitems,
                    uint64_t shard_tiles) {
    SweepLaunch a;
    a.runs = (const QgtDevRun*)c->img_runs.ptr;
    a.subs = (const QgtDevSubPass*)c->img_subs.ptr;
    a.stages = (const QgtDevStage*)c->img_stages.ptr;
    a.tdiags = (const QgtDevThrDiag*)c->img_tdiags.ptr;
    a.costs = (const QgtDevCost*)c->img_costs.ptr;
    a.pool = (const cplx*)c->img_pool.ptr;
    a.run_idx = run;
    a.items = d_items;
    a.nitems = nitems;
    a.ntiles = shard_tiles;
    a.use_mma = c->use_mma;
    a.double_buffer = c->double_buffer;
    a.debug_skip = c->debug_skip;
    a.tiles_per_item = c->tiles_per_item;
    a.mma_only = c->use_mma && plan.R == 3 && plan.B == 0;
    int has_cost = 0, ncost = 0;
    for (const SubPass& sp : plan.runs[run].subs) {
        if (sp.is_cost) { has_cost = 1; ncost++; }
        if (!sp.is_cost && !sp.mma_ok) a.mma_only = 0;     // cost passes have their own code in the tensor-only kernel
    }
    a.gprefix = (uint64_t)c->rank << plan.nloc;
    a.ct = c->seg_cost.empty() ? c->cost : c->seg_cost[plan.runs[run].segment];
    const int K = plan.runs[run].K;
    const int R = plan.R;
    int mat_count = 0;
    for (const SubPass& sp : plan.runs[run].subs)
        for (const Stage& stg : sp.stages) mat_count += QGT_VARIANT_STRIDE(1 << R) << stg.vqubits.size();
    char label[160];
    if (c->timer.trace) {
        int nst = 0;
        for (const SubPass& sp : plan.runs[run].subs) nst += (int)sp.stages.size();
        snprintf(label, sizeof label, "sweep run=%d items=%d subs=%d stages=%d ops=%d tiles=%llu mats=%d", run, nitems,
                 (int)plan.runs[run].subs.size(), nst, (int)plan.runs[run].ops.size(), (unsigned long long)shard_tiles, mat_count);
    }
    a.cost_phase = nullptr; a.cost_ein = nullptr;
    if (has_cost && c->cost_tables) {
        int rc = c->cost_phase.reserve(((size_t)ncost << K) * sizeof(cplx) + ((size_t)1 << K) * sizeof(double));
        if (rc) return rc;
        a.cost_phase = (const cplx*)c->cost_phase.ptr;
        a.cost_ein = (const double*)(a.cost_phase + ((size_t)ncost << K));
    }
    c->timer.begin(c->stream, 0, label);
    cudaError_t e = a.cost_phase ? launch_cost_phase_tables(a, K, ncost, (cplx*)c->cost_phase.ptr, c->stream) : cudaSuccess;
    if (e == cudaSuccess) e = launch_sweep(a, K, R, plan.B, mat_count, (int)plan.runs[run].subs.size(), has_cost, c->num_sms, c->stream);
    c->timer.end(c->stream);
    if (e != cudaSuccess) return cuda_fail(e, "sweep launch");
    c->stats.sweep_launches++;
    c->stats.sweep_column_passes += nitems;
    return QGT_B200_OK;
}

// Apply every run of `plan` to the column at `d` in place.
int apply_plan_inplace(qgt_b200_ctx* c, const CircuitPlan& plan, cplx* d, uint64_t D) {
    if (plan.runs.empty()) return QGT_B200_OK;
    int rc = c->items.reserve(sizeof(QgtSweepItem));
    if (rc) return rc;
    QgtSweepItem it;
    std::memset(&it, 0, sizeof it);
    it.src = d; it.dst = d; it.ovr_kind = 0; it.ovr_index = -1; it.accumulate = 0;
    cudaError_t e = cudaMemcpyAsync(c->items.ptr, &it, sizeof it, cudaMemcpyHostToDevice, c->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(c->stream);
    if (e != cudaSuccess) return cuda_fail(e, "item upload");
    for (size_t r = 0; r < plan.runs.size(); r++) {
        if (plan.runs[r].exchange_gbit >= 0) {
            if ((rc = dist_exchange(c, d, D, plan.runs[r].exchange_mask))) return rc;
            continue;
        }
        const uint64_t ntiles = D >> plan.runs[r].K;
        if ((rc = do_sweep(c, plan, (int)r, (const QgtSweepItem*)c->items.ptr, 1, ntiles))) return rc;
        c->stats.sweep_bytes += 32.0 * (double)D;
    }
    return QGT_B200_OK;
}

// executes a Program on `nslots` columns of D amplitudes each living at arena + slot*D
int run_program(qgt_b200_ctx* c, const qgt_b200_circuit& circ, const CircuitPlan& plan, const Program& prog,
                cplx* arena, uint64_t D, cplx* cmat /* (P+1)^2 device */) {
    const int P = plan.P;
    const auto t_pack0 = std::chrono::steady_clock::now();
    // ---- pack every launch's arguments and upload them once ------------------------------------
    std::vector<QgtSweepItem> items;
    std::vector<double> ovr_pool;                // derivative matrices of the spawn items
    struct OvrTask { int run, sub, stage; std::vector<int> dops; size_t item, off; };
    std::vector<OvrTask> tasks;
    size_t ovr_doubles = 0;
    std::vector<const cplx*> ptrs;
    std::vector<int> ids;
    struct GramRec { size_t a_off, b_off, aid_off, bid_off; int na, nb; bool symmetric; };
    std::vector<size_t> item_off(prog.instrs.size(), 0);
    std::vector<GramRec> grec(prog.instrs.size());
    size_t partial_bytes = 0;
    // Program slots -> arena columns.  Identity, except that a grouped exchange receives every column into the spare
    // column (index num_slots) and the column's old place becomes the next spare: the map is rotated statically here.
    std::vector<int> phys((size_t)prog.num_slots);
    for (int i = 0; i < prog.num_slots; i++) phys[i] = i;
    int spare = prog.num_slots;
    struct Resolved { int a = -1, b = -1; std::vector<std::pair<int, int>> moves; int traj[QGT_MAX_TRAJ]; };
    std::vector<Resolved> res(prog.instrs.size());
    for (size_t i = 0; i < prog.instrs.size(); i++) {
        const Instr& in = prog.instrs[i];
        if (in.kind == INSTR_INIT) res[i].b = phys[in.dst];
        if (in.kind == INSTR_COPY) { res[i].a = phys[in.src]; res[i].b = phys[in.dst]; }
        if (in.kind == INSTR_FUSED) {
            res[i].a = phys[in.phi];
            for (int t = 0; t < QGT_MAX_TRAJ; t++) res[i].traj[t] = t < (int)prog.traj_slots.size() ? phys[prog.traj_slots[t]] : -1;
        }
        if (in.kind == INSTR_SWEEP && plan.runs[in.run].exchange_gbit >= 0) {
            for (const SweepCol& sc : in.cols) {
                const int from = phys[sc.dst];
                res[i].moves.push_back({from, spare});
                phys[sc.dst] = spare;
                spare = from;
            }
            continue;
        }
        if (in.kind == INSTR_SWEEP || in.kind == INSTR_FUSED) {
            item_off[i] = items.size();
            const Run& run = plan.runs[in.run];
            for (const SweepCol& sc : in.cols) {
                QgtSweepItem it;
                std::memset(&it, 0, sizeof it);
                it.src = arena + (size_t)phys[sc.src] * D;
                it.dst = arena + (size_t)phys[sc.dst] * D;
                it.accumulate = sc.accumulate ? 1u : 0u;
                it.ovr_kind = 0; it.ovr_index = -1;
                it.self = sc.self ? 1 : 0;
                it.rho_from = sc.rho_from;
                it.phi_dst = sc.phi_dst >= 0 ? (void*)(arena + (size_t)phys[sc.phi_dst] * D) : nullptr;
                if (sc.ovr_op >= 0) {
                    const OpLocation loc = locate_op(run, sc.ovr_op);
                    if (loc.kind == 0) return fail(QGT_B200_ERR_INTERNAL, "override op outside every sub-pass");
                    it.ovr_kind = loc.kind;
                    it.ovr_index = loc.index;
                    const SubPass& sp = run.subs[loc.sub];
                    if (loc.kind == 1) {
                        int first = 0;
                        for (int s2 = 0; s2 < loc.sub; s2++) first += (int)run.subs[s2].stages.size();
                        std::vector<int> dops(1, sc.ovr_op);
                        dops.insert(dops.end(), sc.ovr_extra.begin(), sc.ovr_extra.end());
                        // the matrices themselves are computed below, in parallel: only the layout is fixed here
                        const Stage& stg = sp.stages[loc.index - first];
                        OvrTask task;
                        task.run = in.run; task.sub = loc.sub; task.stage = loc.index - first; task.dops = std::move(dops);
                        task.item = items.size(); task.off = ovr_doubles;
                        ovr_doubles += ((size_t)QGT_VARIANT_STRIDE(1 << (int)sp.reg_local.size()) << stg.vqubits.size()) * 2;
                        tasks.push_back(std::move(task));
                    } else if (loc.kind == 2) {
                        it.ovr_tdiag = make_tdiag(run.ops[sc.ovr_op], true);
                    } else {
                        it.ovr_cost = make_cost(run.ops[sc.ovr_op], true);
                    }
                }
                items.push_back(it);
            }
        } else if (in.kind == INSTR_GRAM) {
            GramRec& g = grec[i];
            g.na = (int)in.a_slots.size(); g.nb = (int)in.b_slots.size();
            g.a_off = ptrs.size();
            for (int s : in.a_slots) ptrs.push_back(arena + (size_t)phys[s] * D);
            g.b_off = ptrs.size();
            for (int s : in.b_slots) ptrs.push_back(arena + (size_t)phys[s] * D);
            g.aid_off = ids.size();
            for (int v : in.a_ids) ids.push_back(v);
            g.bid_off = ids.size();
            for (int v : in.b_ids) ids.push_back(v);
            g.symmetric = g.nb >= g.na && std::equal(in.a_slots.begin(), in.a_slots.end(), in.b_slots.begin());
            const GramShape shp = gram_shape(g.na, g.nb);
            GramLaunch gl; gl.na = g.na; gl.nb = g.nb; gl.D = D; gl.symmetric = g.symmetric ? 1 : 0;
            const size_t per_split = gram_configure(gl, shp);
            const int ks = gram_choose_ksplit(gl, shp, c->num_sms);
            partial_bytes = std::max(partial_bytes, (size_t)ks * per_split * sizeof(cplx));
        }
    }
    // ---- fused schedule: evolved generators (X pool), contraction tables, layout of the self transition matrices ----
    struct FusedRec { size_t group_off = 0; int ngroups = 0; int self_item = -1; int tpc = 0, tg = 0; };
    std::vector<FusedRec> frec(prog.instrs.size());
    std::vector<QgtContractEntry> centries;
    std::vector<QgtContractGroup> cgroups;
    std::vector<double> xpool;
    std::vector<std::vector<int>> xbase(plan.runs.size());     // per run and stage (run-relative): first X block, -1 = none
    std::vector<char> xdone(plan.runs.size(), 0);
    size_t rho_doubles_max = 0, side_partial_bytes = 0;
    FusedHost& fh = c->fused_host;
    if (prog.fused) {
        fh = FusedHost();
        fh.self_off.assign(plan.runs.size(), (size_t)-1);
        std::vector<double> gens;
        for (size_t i = 0; i < prog.instrs.size(); i++) {
            const Instr& in = prog.instrs[i];
            if (in.kind != INSTR_FUSED) continue;
            const Run& run = plan.runs[in.run];
            if (!xdone[in.run]) {
                xdone[in.run] = 1;
                for (const SubPass& sp : run.subs)
                    for (const Stage& st : sp.stages) {
                        if (st.params.empty()) { xbase[in.run].push_back(-1); continue; }
                        xbase[in.run].push_back((int)(xpool.size() / 128));
                        stage_generators(run, sp, st, gens);
                        const int nv = 1 << (int)st.vqubits.size();
                        for (size_t kk = 0; kk < st.params.size(); kk++)
                            for (int v = 0; v < nv; v++) {
                                const double* G = &gens[(kk * (size_t)nv + v) * 128];
                                const size_t base = xpool.size();
                                xpool.resize(base + 128);
                                for (int a = 0; a < 8; a++)
                                    for (int cc = 0; cc < 8; cc++) {          // X(c, a) = Gt[a][c], stored in rho's element order
                                        xpool[base + rho_index(cc, a, 0)] = G[2 * (a * 8 + cc)];
                                        xpool[base + rho_index(cc, a, 1)] = G[2 * (a * 8 + cc) + 1];
                                    }
                            }
                        FusedHost::StageGen sg;
                        sg.run = in.run; sg.rho_off = st.rho_off; sg.nvar = (int)st.vqubits.size(); sg.params = st.params; sg.gens = gens;
                        fh.stages.push_back(std::move(sg));
                    }
            }
            FusedRec& fr = frec[i];
            std::map<int, std::vector<QgtContractEntry>> byout;
            for (size_t j = 0; j < in.cols.size(); j++) {
                const SweepCol& sc = in.cols[j];
                if (sc.self) {
                    if (sc.rho_from == 0) {
                        fr.self_item = (int)j;
                        if (fh.self_off[in.run] == (size_t)-1) { fh.self_off[in.run] = fh.self_doubles; fh.self_doubles += (size_t)run.rho_blocks * 128; }
                    }
                    continue;
                }
                int sidx = 0;
                for (const SubPass& sp : run.subs)
                    for (const Stage& st : sp.stages) {
                        if (st.rho_off >= 0 && sidx >= sc.rho_from) {
                            const int nv = 1 << (int)st.vqubits.size();
                            for (size_t kk = 0; kk < st.params.size(); kk++) {
                                QgtContractEntry en;
                                en.item = (int)j; en.rho_off = st.rho_off; en.nblocks = nv;
                                en.x_off = xbase[in.run][sidx] + (int)kk * nv;
                                byout[sc.id * P + st.params[kk]].push_back(en);
                            }
                        }
                        sidx++;
                    }
            }
            fr.group_off = cgroups.size();
            for (auto& kv : byout) {
                QgtContractGroup g;
                g.begin = (int)centries.size();
                centries.insert(centries.end(), kv.second.begin(), kv.second.end());
                g.end = (int)centries.size(); g.out = kv.first; g.pad = 0;
                cgroups.push_back(g);
            }
            fr.ngroups = (int)(cgroups.size() - fr.group_off);
            {
                int mc = 0, ns = 0;
                for (const SubPass& sp : run.subs)
                    for (const Stage& stg : sp.stages) { mc += QGT_VARIANT_STRIDE(8) << stg.vqubits.size(); ns++; }
                const bool pipe = fused_uses_pipe(run.K, in.traj ? 1 : 0, c->fused_pipeline, mc, (int)run.subs.size(), run.rho_blocks, ns);
                int simple = 1;
                for (const SubPass& sp : run.subs) if (sp.is_cost || sp.stages.size() != 1 || !sp.tdiags.empty()) simple = 0;
                const bool direct = fused_uses_direct(run.K, in.traj ? 1 : 0, c->fused_pipeline, simple, mc, (int)run.subs.size(), run.rho_blocks);
                const int nranges = in.traj && prog.traj_ranges > 1 ? prog.traj_ranges : 1;
                fused_geometry((D >> run.K) / (uint64_t)nranges, (int)in.cols.size(), c->num_sms, direct ? 2 : pipe ? 1 : 0, in.traj, &fr.tpc, &fr.tg);
            }
            const size_t per_item = (size_t)run.rho_blocks * 128;
            rho_doubles_max = std::max(rho_doubles_max, in.cols.size() * per_item);
            partial_bytes = std::max(partial_bytes, (size_t)fr.tg * in.cols.size() * per_item * sizeof(double));
            if (in.cols.size() == 1) side_partial_bytes = std::max(side_partial_bytes, (size_t)fr.tg * per_item * sizeof(double));
        }
    }
    // derivative (product-rule) stage matrices: independent of each other, a few microseconds each, hundreds per
    // evaluation -> a handful of host threads
    ovr_pool.assign(ovr_doubles, 0.0);
    {
        auto work = [&](size_t lo, size_t hi) {
            std::vector<double> mats;
            for (size_t k = lo; k < hi; k++) {
                const OvrTask& t = tasks[k];
                const Run& run = plan.runs[t.run];
                const SubPass& sp = run.subs[t.sub];
                items[t.item].ovr_form = stage_matrices_sum(run, sp, sp.stages[t.stage], t.dops, mats);
                std::memcpy(&ovr_pool[t.off], mats.data(), mats.size() * sizeof(double));
            }
        };
        // threads only for large programs: at a few hundred matrices (< 1 ms serial) a descheduled helper thread on a busy
        // host costs more than it saves
        const size_t nthreads = tasks.size() >= 1024 ? std::min<size_t>(4, std::max<unsigned>(1u, std::thread::hardware_concurrency())) : 1;
        if (nthreads <= 1) work(0, tasks.size());
        else {
            std::vector<std::thread> pool;
            const size_t per = (tasks.size() + nthreads - 1) / nthreads;
            size_t started = 1;                  // chunk 0 runs on this thread
            try {
                for (; started < nthreads; started++)
                    pool.emplace_back(work, std::min(tasks.size(), started * per), std::min(tasks.size(), (started + 1) * per));
            } catch (...) {}                     // no thread available: the remaining chunks run here
            work(0, std::min(tasks.size(), per));
            for (size_t t = started; t < nthreads; t++) work(std::min(tasks.size(), t * per), std::min(tasks.size(), (t + 1) * per));
            for (std::thread& th : pool) th.join();
        }
    }
    c->ms_prog_pack = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t_pack0).count();
    int rc;
    if ((rc = c->ovr_pool.reserve(std::max<size_t>(16, ovr_pool.size() * sizeof(double))))) return rc;
    for (const OvrTask& t : tasks)
        items[t.item].ovr_mat = (const char*)c->ovr_pool.ptr + t.off * sizeof(double);
    if ((rc = c->items.reserve(std::max<size_t>(1, items.size()) * sizeof(QgtSweepItem)))) return rc;
    const size_t ptr_bytes = ptrs.size() * sizeof(cplx*), id_bytes = ids.size() * sizeof(int);
    if ((rc = c->aux.reserve(std::max<size_t>(16, ptr_bytes + id_bytes)))) return rc;
    if ((rc = c->partial.reserve(std::max<size_t>(16, partial_bytes)))) return rc;
    cudaError_t e = cudaSuccess;
    if (!items.empty()) e = cudaMemcpyAsync(c->items.ptr, items.data(), items.size() * sizeof(QgtSweepItem), cudaMemcpyHostToDevice, c->stream);
    if (e == cudaSuccess && !ovr_pool.empty()) e = cudaMemcpyAsync(c->ovr_pool.ptr, ovr_pool.data(), ovr_pool.size() * sizeof(double), cudaMemcpyHostToDevice, c->stream);
    if (e == cudaSuccess && ptr_bytes) e = cudaMemcpyAsync(c->aux.ptr, ptrs.data(), ptr_bytes, cudaMemcpyHostToDevice, c->stream);
    if (e == cudaSuccess && id_bytes) e = cudaMemcpyAsync((char*)c->aux.ptr + ptr_bytes, ids.data(), id_bytes, cudaMemcpyHostToDevice, c->stream);
    const size_t entry_bytes = centries.size() * sizeof(QgtContractEntry), group_bytes = cgroups.size() * sizeof(QgtContractGroup);
    if (prog.fused) {
        if ((rc = c->fx_pool.reserve(std::max<size_t>(16, xpool.size() * sizeof(double))))) return rc;
        if ((rc = c->fx_tab.reserve(std::max<size_t>(16, entry_bytes + group_bytes)))) return rc;
        if ((rc = c->rho.reserve(std::max<size_t>(16, rho_doubles_max * sizeof(double))))) return rc;
        if ((rc = c->rho_side.reserve(std::max<size_t>(16, rho_doubles_max * sizeof(double))))) return rc;
        if ((rc = c->partial_side.reserve(std::max<size_t>(16, side_partial_bytes)))) return rc;
        if ((rc = c->rho_self.reserve(std::max<size_t>(16, fh.self_doubles * sizeof(double))))) return rc;
        if ((rc = c->amat.reserve(std::max<size_t>(16, (size_t)P * P * sizeof(cplx))))) return rc;
        if (e == cudaSuccess && !xpool.empty()) e = cudaMemcpyAsync(c->fx_pool.ptr, xpool.data(), xpool.size() * sizeof(double), cudaMemcpyHostToDevice, c->stream);
        if (e == cudaSuccess && entry_bytes) e = cudaMemcpyAsync(c->fx_tab.ptr, centries.data(), entry_bytes, cudaMemcpyHostToDevice, c->stream);
        if (e == cudaSuccess && group_bytes) e = cudaMemcpyAsync((char*)c->fx_tab.ptr + entry_bytes, cgroups.data(), group_bytes, cudaMemcpyHostToDevice, c->stream);
        if (e == cudaSuccess) e = cudaMemsetAsync(c->amat.ptr, 0, std::max<size_t>(16, (size_t)P * P * sizeof(cplx)), c->stream);
        if (e == cudaSuccess) e = cudaMemsetAsync(c->rho_self.ptr, 0, std::max<size_t>(16, fh.self_doubles * sizeof(double)), c->stream);
    }
    if (e == cudaSuccess) e = cudaStreamSynchronize(c->stream);   // host vectors are locals
    if (e != cudaSuccess) return cuda_fail(e, "program upload");
    const QgtContractEntry* d_entries = (const QgtContractEntry*)c->fx_tab.ptr;
    const QgtContractGroup* d_groups = (const QgtContractGroup*)((const char*)c->fx_tab.ptr + entry_bytes);
    c->stats.fused = prog.fused ? 1 : 0;
    const cplx* const* d_ptrs = (const cplx* const*)c->aux.ptr;
    const int* d_ids = (const int*)((char*)c->aux.ptr + ptr_bytes);

    c->psi_phys_slot = prog.psi_slot < (int)phys.size() ? phys[prog.psi_slot] : prog.psi_slot;
    // ---- execute ----------------------------------------------------------------------------------
    // one fused launch: instruction `i`, tile range `rg` of `nrg` (nrg > 1: ranged trajectory mode, Program::traj_ranges)
    // `side`: on the side stream with its own partial / transition-matrix buffers (untimed); `ib`: which half of the image column
    auto fused_launch = [&](size_t i, int rg, int nrg, bool side, int ib) -> int {
            cudaStream_t st = side ? c->side_stream : c->stream;
            double* partial = (double*)(side ? c->partial_side.ptr : c->partial.ptr);
            double* rho = (double*)(side ? c->rho_side.ptr : c->rho.ptr);
            const Instr& in = prog.instrs[i];
            const Run& run = plan.runs[in.run];
            const FusedRec& fr = frec[i];
            const double Dr = (double)D / nrg;
            const int nitems = (int)in.cols.size();
            const size_t per_item = (size_t)run.rho_blocks * 128;
            FusedLaunch a;
            a.runs = (const QgtDevRun*)c->img_runs.ptr;
            a.subs = (const QgtDevSubPass*)c->img_subs.ptr;
            a.stages = (const QgtDevStage*)c->img_stages.ptr;
            a.tdiags = (const QgtDevThrDiag*)c->img_tdiags.ptr;
            a.pool = (const cplx*)c->img_pool.ptr;
            a.run_idx = in.run;
            a.items = (const QgtSweepItem*)c->items.ptr + item_off[i];
            a.nitems = nitems;
            a.phi = arena + (size_t)res[i].a * D;
            a.ntiles = (D >> run.K) / (uint64_t)nrg;
            a.tile_off = (uint64_t)rg * a.ntiles;
            a.tiles_per_cta = fr.tpc; a.tile_groups = fr.tg;
            a.gprefix = (uint64_t)c->rank << plan.nloc;
            a.rho_partial = partial;
            a.use_traj = in.traj ? 1 : 0;
            a.debug = c->fused_debug;
            a.pipeline = c->fused_pipeline;
            a.all_simple = 1;
            for (const SubPass& sp : run.subs) if (sp.is_cost || sp.stages.size() != 1 || !sp.tdiags.empty()) a.all_simple = 0;
            for (int t = 0; t < QGT_MAX_TRAJ; t++)
                a.traj[t] = nrg > 1 ? arena + (size_t)res[i].traj[0] * D + ((size_t)ib * (size_t)(nrg / 2) + (size_t)t) * (D / (uint64_t)nrg)
                                    : res[i].traj[t] >= 0 ? arena + (size_t)res[i].traj[t] * D : nullptr;
            int mat_count = 0, nstage_rho = 0, nstages = 0;
            double flops_ab = 0.0;                  // per amplitude: one tile through every stage
            for (const SubPass& sp : run.subs)
                for (const Stage& stg : sp.stages) {
                    mat_count += QGT_VARIANT_STRIDE(8) << stg.vqubits.size();
                    // DMMA flops per amplitude of one stage application: 8 complex MACs as 4 real products (64), or the
                    // diagonal-real form's 2 real products (32)
                    flops_ab += c->img_stage_form[(size_t)c->img_run_stage_off[in.run] + nstages] == QGT_FORM_DIAG_REAL ? 32.0 : 64.0;
                    nstages++;
                    if (stg.rho_off >= 0) nstage_rho++;
                }
            char label[160];
            if (c->timer.trace)
                snprintf(label, sizeof label, "fused run=%d items=%d subs=%d rho_stages=%d rho_blocks=%d tiles=%llu chunks=%d", in.run, nitems,
                         (int)run.subs.size(), nstage_rho, run.rho_blocks, (unsigned long long)a.ntiles, fr.tg);
            if (!side) c->timer.begin(st, 0, label);
            e = launch_fused(a, run.K, mat_count, (int)run.subs.size(), run.rho_blocks, nstages, st);
            if (!side) c->timer.end(st);
            if (e != cudaSuccess) return cuda_fail(e, "fused launch");
            c->stats.sweep_launches++; c->stats.fused_launches++;
            if (rg == 0) c->stats.sweep_column_passes += nitems;
            for (const SweepCol& sc : in.cols) {
                c->stats.sweep_bytes += (sc.accumulate ? 48.0 : 32.0) * Dr;
                int sidx = 0, nrho = 0;
                for (const SubPass& sp : run.subs)
                    for (const Stage& stg : sp.stages) { if (stg.rho_off >= 0 && sidx >= sc.rho_from) nrho++; sidx++; }
                const bool use_b = !in.traj && !sc.self && sc.rho_from <= run.last_rho_stage;
                c->stats.tensor_flops += Dr * (flops_ab * (use_b ? 2.0 : 1.0) + 48.0 * nrho);      // rho: 3M complex products
                if (in.traj) c->stats.sweep_bytes += 16.0 * Dr * nrho;     // phi's tile images: written by the self item, fetched (L2) by the others
            }
            if (per_item > 0) {
                if (!side) c->timer.begin(st, 1, "rho reduce + contract");
                e = launch_rho_reduce(partial, fr.tg, nitems, (int)per_item, rho, st);
                if (e == cudaSuccess && fr.self_item >= 0) {
                    double* self_dst = (double*)c->rho_self.ptr + fh.self_off[in.run];
                    const double* self_src = rho + (size_t)fr.self_item * per_item;
                    if (rg == 0) e = cudaMemcpyAsync(self_dst, self_src, per_item * sizeof(double), cudaMemcpyDeviceToDevice, st);
                    else e = launch_add_doubles(self_dst, self_src, per_item, st);       // the ranges' shares add up
                }
                if (e == cudaSuccess && fr.ngroups > 0)
                    e = launch_rho_contract(rho, (int)per_item, (const double*)c->fx_pool.ptr, d_groups + fr.group_off, fr.ngroups,
                                            d_entries, (cplx*)c->amat.ptr, st);
                if (!side) c->timer.end(st);
                if (e != cudaSuccess) return cuda_fail(e, "rho reduce/contract launch");
                c->stats.other_launches += 2;
            }
            return QGT_B200_OK;
    };
    const double plus_amp = std::pow(2.0, -0.5 * plan.n);
    for (size_t i = 0; i < prog.instrs.size(); i++) {
        const Instr& in = prog.instrs[i];
        switch (in.kind) {
        case INSTR_INIT: {
            c->timer.begin(c->stream, 2);
            e = launch_init_state(arena + (size_t)res[i].b * D, D, circ.initial_state, plus_amp, (uint64_t)c->rank * D, c->stream);
            c->timer.end(c->stream);
            if (e != cudaSuccess) return cuda_fail(e, "init launch");
            c->stats.other_launches++;
            break; }
        case INSTR_COPY: {
            c->timer.begin(c->stream, 2);
            e = cudaMemcpyAsync(arena + (size_t)res[i].b * D, arena + (size_t)res[i].a * D, D * sizeof(cplx), cudaMemcpyDeviceToDevice, c->stream);
            c->timer.end(c->stream);
            if (e != cudaSuccess) return cuda_fail(e, "column copy");
            c->stats.other_launches++;
            break; }
        case INSTR_SWEEP: {
            if (plan.runs[in.run].exchange_gbit >= 0) {      // sharded state: one grouped all-to-all per column, into the spare column
                for (const std::pair<int, int>& mv : res[i].moves) {
                    c->timer.begin(c->stream, 3, "exchange");
                    rc = dist_exchange_multi(c, arena + (size_t)mv.first * D, arena + (size_t)mv.second * D, D, plan.runs[in.run].exchange_mask);
                    c->timer.end(c->stream);
                    if (rc) return rc;
                    c->stats.other_launches++;
                }
                break;
            }
            const uint64_t ntiles = D >> plan.runs[in.run].K;
            if ((rc = do_sweep(c, plan, in.run, (const QgtSweepItem*)c->items.ptr + item_off[i], (int)in.cols.size(), ntiles))) return rc;
            for (const SweepCol& sc : in.cols)
                c->stats.sweep_bytes += (sc.accumulate ? 48.0 : 32.0) * (double)D;
            break; }
        case INSTR_FUSED: {
            // ranged trajectory mode: the launches of one run (phi's, the columns', the accumulating items') form a group that
            // is walked range by range, so that phi's images of a range are consumed before the next range overwrites them
            const int nrg = in.traj && prog.traj_ranges > 1 ? prog.traj_ranges : 1;
            size_t last = i;
            if (nrg > 1)
                while (last + 1 < prog.instrs.size() && prog.instrs[last + 1].kind == INSTR_FUSED && prog.instrs[last + 1].run == in.run &&
                       prog.instrs[last + 1].traj) last++;
            // phi's launch is bound by HBM (it writes an image per stage), the columns' launches by the tensor pipe: phi's launch
            // of range r + 1 runs on the (high-priority) side stream next to the columns' launches of range r.  The image column
            // then holds two ranges (halves, alternating), which needs 2 x stages <= ranges.
            const bool overlap = nrg > 1 && last > i && in.cols.size() == 1 && in.cols[0].self && c->fused_overlap != 0 &&
                                 2 * plan.runs[in.run].rho_stages <= nrg;
            if (!overlap) {
                for (int rg = 0; rg < nrg; rg++)
                    for (size_t k = i; k <= last; k++)
                        if ((rc = fused_launch(k, rg, nrg, false, 0))) return rc;
            } else {
                e = cudaEventRecord(c->ev_group, c->stream);               // phi (copies, exchanges) and the previous run's readers
                if (e == cudaSuccess) e = cudaStreamWaitEvent(c->side_stream, c->ev_group, 0);
                if (e != cudaSuccess) return cuda_fail(e, "side stream hand-over");
                if ((rc = fused_launch(i, 0, nrg, true, 0))) return rc;
                e = cudaEventRecord(c->ev_side[0], c->side_stream);
                for (int rg = 0; rg < nrg && e == cudaSuccess; rg++) {
                    if (rg + 1 < nrg) {
                        const int nb = (rg + 1) & 1;
                        if (rg >= 1) e = cudaStreamWaitEvent(c->side_stream, c->ev_main[nb], 0);      // range rg - 1 has consumed that half
                        if (e != cudaSuccess) break;
                        if ((rc = fused_launch(i, rg + 1, nrg, true, nb))) return rc;
                        e = cudaEventRecord(c->ev_side[nb], c->side_stream);
                        if (e != cudaSuccess) break;
                    }
                    e = cudaStreamWaitEvent(c->stream, c->ev_side[rg & 1], 0);
                    if (e != cudaSuccess) break;
                    for (size_t k = i + 1; k <= last; k++)
                        if ((rc = fused_launch(k, rg, nrg, false, rg & 1))) return rc;
                    e = cudaEventRecord(c->ev_main[rg & 1], c->stream);
                }
                if (e != cudaSuccess) return cuda_fail(e, "side stream ordering");
            }
            i = last;
            break; }
        case INSTR_GRAM: {
            const GramRec& g = grec[i];
            GramLaunch gl;
            gl.a_ptrs = d_ptrs + g.a_off; gl.b_ptrs = d_ptrs + g.b_off;
            gl.na = g.na; gl.nb = g.nb; gl.D = D;
            const GramShape shp = gram_shape(g.na, g.nb);
            gl.symmetric = g.symmetric ? 1 : 0;
            gram_configure(gl, shp);
            gl.ksplit = gram_choose_ksplit(gl, shp, c->num_sms);
            gl.partial = (cplx*)c->partial.ptr;
            c->timer.begin(c->stream, 1);
            e = launch_gram(gl, shp, c->stream);
            if (e == cudaSuccess) e = launch_gram_reduce(gl, shp, d_ids + g.aid_off, d_ids + g.bid_off, cmat, P + 1, c->stream);
            c->timer.end(c->stream);
            if (e != cudaSuccess) return cuda_fail(e, "gram launch");
            c->stats.gram_launches++;
            const double pairs = g.symmetric ? 0.5 * g.na * (g.na + 1) + (g.nb - g.na) * (double)g.na : (double)g.na * g.nb;
            c->stats.gram_flops += 8.0 * pairs * (double)D;
            c->stats.gram_bytes += 16.0 * (double)D * (g.symmetric ? g.nb : g.na + g.nb);
            break; }
        }
    }
    return QGT_B200_OK;
}

bool choose_fused(const qgt_b200_ctx* c, const CircuitPlan& plan, size_t slots) {
    if (c->fused_mode == 0 || !plan_supports_fused(plan)) return false;
    size_t Pa = 0;
    for (int f : plan.first_run) if (f >= 0) Pa++;
    if (Pa == 0) return false;
    if (c->fused_mode == 1) return true;
    return Pa + 2 > slots;       // automatic: the Gram schedule while every column is resident, the fused one otherwise
}

int fused_finish(qgt_b200_ctx* c, const CircuitPlan& plan, double* metric, double* berry, double* q_full) {
    const int P = plan.P;
    const FusedHost& fh = c->fused_host;
    int rc;
    if (c->world > 1) {
        if ((rc = dist_allreduce_device(c, (double*)c->amat.ptr, (size_t)P * P * 2))) return rc;
        if (fh.self_doubles && (rc = dist_allreduce_device(c, (double*)c->rho_self.ptr, fh.self_doubles))) return rc;
    }
    std::vector<double> A((size_t)P * P * 2), rs(fh.self_doubles);
    cudaError_t e = cudaSuccess;
    if (P > 0) e = cudaMemcpyAsync(A.data(), c->amat.ptr, A.size() * sizeof(double), cudaMemcpyDeviceToHost, c->stream);
    if (e == cudaSuccess && !rs.empty()) e = cudaMemcpyAsync(rs.data(), c->rho_self.ptr, rs.size() * sizeof(double), cudaMemcpyDeviceToHost, c->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(c->stream);
    if (e != cudaSuccess) return cuda_fail(e, "fused result download");
    // C = A + A^H + D,  v_mu = <d_mu psi|psi>;  D (pairs inside one stage, the diagonal) and v from rho of phi with itself
    std::vector<double> Cr((size_t)P * P), Ci((size_t)P * P), vr(P, 0.0), vi(P, 0.0);
    for (int m = 0; m < P; m++)
        for (int n = 0; n < P; n++) {
            Cr[(size_t)m * P + n] = A[2 * ((size_t)m * P + n)] + A[2 * ((size_t)n * P + m)];
            Ci[(size_t)m * P + n] = A[2 * ((size_t)m * P + n) + 1] - A[2 * ((size_t)n * P + m) + 1];
        }
    for (const FusedHost::StageGen& sg : fh.stages) {
        if (fh.self_off[sg.run] == (size_t)-1) continue;
        const int nv = 1 << sg.nvar, np = (int)sg.params.size();
        for (int v = 0; v < nv; v++) {
            const double* rho = &rs[fh.self_off[sg.run] + (size_t)(sg.rho_off + v) * 128];
            // W_k[i][a] = sum_c G_k[i][c] rho[c][a]   (G_k rho), then  D[mu][nu] = sum_{i,a} conj(G_mu[i][a]) W_nu[i][a],  v_mu = sum conj(G_mu[i][a]) rho[i][a]
            std::vector<double> W((size_t)np * 128, 0.0);
            for (int k = 0; k < np; k++) {
                const double* G = &sg.gens[((size_t)k * nv + v) * 128];
                for (int i = 0; i < 8; i++)
                    for (int a = 0; a < 8; a++) {
                        double wr = 0.0, wi = 0.0;
                        for (int cc = 0; cc < 8; cc++) {
                            const double gr = G[2 * (i * 8 + cc)], gi = G[2 * (i * 8 + cc) + 1];
                            const double rr = rho[rho_index(cc, a, 0)], ri = rho[rho_index(cc, a, 1)];
                            wr += gr * rr - gi * ri; wi += gr * ri + gi * rr;
                        }
                        W[((size_t)k * 64 + i * 8 + a) * 2] = wr; W[((size_t)k * 64 + i * 8 + a) * 2 + 1] = wi;
                    }
            }
            for (int km = 0; km < np; km++) {
                const double* Gm = &sg.gens[((size_t)km * nv + v) * 128];
                const int mu = sg.params[km];
                double sr = 0.0, si = 0.0;
                for (int i = 0; i < 8; i++)
                    for (int a = 0; a < 8; a++) {
                        const double gr = Gm[2 * (i * 8 + a)], gi = -Gm[2 * (i * 8 + a) + 1];
                        const double rr = rho[rho_index(i, a, 0)], ri = rho[rho_index(i, a, 1)];
                        sr += gr * rr - gi * ri; si += gr * ri + gi * rr;
                    }
                vr[mu] += sr; vi[mu] += si;
                for (int kn = 0; kn < np; kn++) {
                    const double* Wn = &W[(size_t)kn * 128];
                    double dr = 0.0, di = 0.0;
                    for (int x = 0; x < 64; x++) {
                        const double gr = Gm[2 * x], gi = -Gm[2 * x + 1];
                        dr += gr * Wn[2 * x] - gi * Wn[2 * x + 1]; di += gr * Wn[2 * x + 1] + gi * Wn[2 * x];
                    }
                    Cr[(size_t)mu * P + sg.params[kn]] += dr; Ci[(size_t)mu * P + sg.params[kn]] += di;
                }
            }
        }
    }
    for (int m = 0; m < P; m++)
        for (int n = 0; n < P; n++) {
            // Q = C - v_mu conj(v_nu)
            const double qr = Cr[(size_t)m * P + n] - (vr[m] * vr[n] + vi[m] * vi[n]);
            const double qi = Ci[(size_t)m * P + n] - (vi[m] * vr[n] - vr[m] * vi[n]);
            if (metric) metric[(size_t)m * P + n] = qr;
            if (berry) berry[(size_t)m * P + n] = qi;
            if (q_full) { q_full[2 * ((size_t)m * P + n)] = qr; q_full[2 * ((size_t)m * P + n) + 1] = qi; }
        }
    return QGT_B200_OK;
}

size_t workspace_slots(qgt_b200_ctx* c, uint64_t D, size_t reserve_bytes) {
    size_t free_b = 0, total_b = 0;
    cudaMemGetInfo(&free_b, &total_b);
    size_t avail = free_b + c->arena.bytes;          // the cached arena is ours to reuse
    avail = avail > reserve_bytes ? avail - reserve_bytes : 0;
    size_t limit = c->ws_limit ? std::min(c->ws_limit, avail) : (size_t)(0.90 * (double)avail);
    size_t slots = limit / (D * sizeof(cplx));
    if (c->max_slots && slots > c->max_slots) slots = c->max_slots;
    return slots;
}

void stats_begin(qgt_b200_ctx* c) {
    std::memset(&c->stats, 0, sizeof c->stats);
    cudaEventRecord(c->ev0, c->stream);
}

int stats_end(qgt_b200_ctx* c) {
    cudaEventRecord(c->ev1, c->stream);
    cudaError_t e = cudaStreamSynchronize(c->stream);
    if (e != cudaSuccess) return cuda_fail(e, "stream sync");
    float ms = 0.f;
    cudaEventElapsedTime(&ms, c->ev0, c->ev1);
    double cat[4];
    c->timer.collect(cat);
    c->stats.ms_exchange = cat[3];
    c->stats.ms_total = ms;
    c->stats.ms_sweep = cat[0]; c->stats.ms_gram = cat[1]; c->stats.ms_other = cat[2];
    return QGT_B200_OK;
}

}  // namespace qgt

extern "C" {

int qgt_b200_apply_circuit(qgt_b200_state* s, const qgt_b200_circuit* circ, const double* theta) {
    if (!s) return fail(QGT_B200_ERR_INVALID_ARG, "state is NULL");
    int rc = check_circuit(circ, theta);
    if (rc) return rc;
    if (circ->num_qubits != s->n) return fail(QGT_B200_ERR_DIMENSION, "circuit and state qubit counts differ");
    qgt_b200_ctx* c = s->ctx;
    cudaSetDevice(c->device);
    if (c->world > 1) return qgt::dist_apply_circuit(s, circ, theta);
    CircuitPlan plan;
    std::string err;
    if ((rc = build_plan(*circ, theta, c->opt, plan, err))) return fail(rc, err);
    PlanImage img;
    stats_begin(c);
    if ((rc = upload_plan(c, *circ, plan, img))) return rc;
    if ((rc = apply_plan_inplace(c, plan, s->d, s->D))) return rc;
    if ((rc = stats_end(c))) return rc;
    c->stats.num_runs = (int)plan.runs.size();
    c->stats.tile_qubits = plan.runs.empty() ? 0 : plan.runs[0].K;
    return QGT_B200_OK;
}

int qgt_b200_simulate_host(qgt_b200_ctx* c, double* host, int n, const qgt_b200_circuit* circ, const double* theta) {
    if (!c || !host) return fail(QGT_B200_ERR_INVALID_ARG, "ctx/host is NULL");
    qgt_b200_state* s = nullptr;
    int rc = qgt_b200_state_create(c, n, &s);
    if (rc) return rc;
    rc = qgt_b200_state_upload(s, host);
    if (!rc) rc = qgt_b200_apply_circuit(s, circ, theta);
    if (!rc) rc = qgt_b200_state_download(s, host);
    qgt_b200_state_destroy(s);
    return rc;
}

int qgt_b200_qgt(qgt_b200_ctx* c, const qgt_b200_circuit* circ, const double* theta,
                 double* metric, double* berry, double* q_full, qgt_b200_state* psi_out) {
    if (!c) return fail(QGT_B200_ERR_INVALID_ARG, "ctx is NULL");
    int rc = check_circuit(circ, theta);
    if (rc) return rc;
    cudaSetDevice(c->device);
    if (psi_out && psi_out->n != circ->num_qubits) return fail(QGT_B200_ERR_DIMENSION, "psi_out has a different qubit count");
    if (c->world > 1) return qgt::dist_qgt(c, circ, theta, metric, berry, q_full, psi_out);
    const auto t_wall0 = std::chrono::steady_clock::now();
    const int n = circ->num_qubits, P = circ->num_params;
    const uint64_t D = (uint64_t)1 << n;
    CircuitPlan plan;
    std::string err;
    if ((rc = build_plan(*circ, theta, c->opt, plan, err))) return fail(rc, err);
    Program prog;
    const size_t slots = workspace_slots(c, D, (size_t)256 << 20);
    const bool use_fused = choose_fused(c, plan, slots);
    if ((rc = use_fused ? build_fused_program(plan, slots, psi_out != nullptr, prog, err, c->fused_traj)
                        : build_qgt_program(plan, slots, psi_out != nullptr, prog, err))) return fail(rc, err);
    const double ms_plan0 = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t_wall0).count();
    if ((rc = c->arena.reserve((size_t)prog.num_slots * D * sizeof(cplx)))) return rc;
    const size_t cm = (size_t)(P + 1) * (P + 1);
    if ((rc = c->cmat.reserve(std::max<size_t>(16, cm * sizeof(cplx))))) return rc;
    if ((rc = c->outbuf.reserve(std::max<size_t>(16, (size_t)P * P * 4 * sizeof(double))))) return rc;
    PlanImage img;
    stats_begin(c);
    if ((rc = upload_plan(c, *circ, plan, img))) return rc;
    cudaError_t e = cudaMemsetAsync(c->cmat.ptr, 0, cm * sizeof(cplx), c->stream);
    if (e != cudaSuccess) return cuda_fail(e, "memset");
    if ((rc = run_program(c, *circ, plan, prog, (cplx*)c->arena.ptr, D, (cplx*)c->cmat.ptr))) return rc;
    double* d_metric = (double*)c->outbuf.ptr;
    double* d_berry = d_metric + (size_t)P * P;
    cplx* d_q = (cplx*)(d_berry + (size_t)P * P);
    if (prog.fused) {
        if ((rc = fused_finish(c, plan, metric, berry, q_full))) return rc;
    } else {
        c->timer.begin(c->stream, 2);
        e = launch_finalize((const cplx*)c->cmat.ptr, P, d_metric, d_berry, d_q, c->stream);
        c->timer.end(c->stream);
        if (e != cudaSuccess) return cuda_fail(e, "finalize launch");
        c->stats.other_launches++;
    }
    const size_t pp = (size_t)P * P;
    if (P > 0 && !prog.fused) {
        if (metric && e == cudaSuccess) e = cudaMemcpyAsync(metric, d_metric, pp * sizeof(double), cudaMemcpyDeviceToHost, c->stream);
        if (berry && e == cudaSuccess) e = cudaMemcpyAsync(berry, d_berry, pp * sizeof(double), cudaMemcpyDeviceToHost, c->stream);
        if (q_full && e == cudaSuccess) e = cudaMemcpyAsync(q_full, d_q, pp * sizeof(cplx), cudaMemcpyDeviceToHost, c->stream);
    }
    if (psi_out && e == cudaSuccess)
        e = cudaMemcpyAsync(psi_out->d, (cplx*)c->arena.ptr + (size_t)c->psi_phys_slot * D, D * sizeof(cplx), cudaMemcpyDeviceToDevice, c->stream);
    if (e != cudaSuccess) return cuda_fail(e, "result copy");
    if ((rc = stats_end(c))) return rc;
    c->stats.num_runs = (int)plan.runs.size();
    c->stats.resident_columns = prog.resident;
    c->stats.blocks = prog.blocks;
    c->stats.tile_qubits = plan.runs.empty() ? 0 : plan.runs[0].K;
    c->stats.ms_wall = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t_wall0).count();
    c->stats.ms_host_plan = ms_plan0 + c->ms_prog_pack;
    return QGT_B200_OK;
}

int qgt_b200_derivative(qgt_b200_ctx* c, const qgt_b200_circuit* circ, const double* theta, int mu, qgt_b200_state* out) {
    if (!c || !out) return fail(QGT_B200_ERR_INVALID_ARG, "ctx/out is NULL");
    int rc = check_circuit(circ, theta);
    if (rc) return rc;
    if (mu < 0 || mu >= circ->num_params) return fail(QGT_B200_ERR_INVALID_ARG, "parameter index out of range");
    if (out->n != circ->num_qubits) return fail(QGT_B200_ERR_DIMENSION, "out has a different qubit count");
    if (c->world > 1) return fail(QGT_B200_ERR_UNSUPPORTED, "derivative columns are single-GPU only");
    cudaSetDevice(c->device);
    const uint64_t D = out->D;
    CircuitPlan plan;
    std::string err;
    if ((rc = build_plan(*circ, theta, c->opt, plan, err))) return fail(rc, err);
    // a two-slot program: slot 0 = phi, slot 1 = the column
    Program prog;
    prog.num_slots = 2;
    { Instr in; in.kind = INSTR_INIT; in.dst = 0; prog.instrs.push_back(in); }
    bool alive = false;
    for (int r = 0; r < (int)plan.runs.size(); r++) {
        if (alive) { Instr in; in.kind = INSTR_SWEEP; in.run = r; in.cols.push_back({1, 1, -1, false}); prog.instrs.push_back(in); }
        for (const ParamOcc& oc : plan.runs[r].occ) {
            if (oc.param != mu) continue;
            Instr in; in.kind = INSTR_SWEEP; in.run = r; in.cols.push_back({0, 1, oc.op, alive});
            prog.instrs.push_back(in);
            alive = true;
        }
        Instr in; in.kind = INSTR_SWEEP; in.run = r; in.cols.push_back({0, 0, -1, false}); prog.instrs.push_back(in);
    }
    if ((rc = c->arena.reserve(2 * D * sizeof(cplx)))) return rc;
    if ((rc = c->cmat.reserve(16 * sizeof(cplx)))) return rc;
    PlanImage img;
    stats_begin(c);
    if ((rc = upload_plan(c, *circ, plan, img))) return rc;
    cudaError_t e = cudaMemsetAsync((cplx*)c->arena.ptr + D, 0, D * sizeof(cplx), c->stream);
    if (e != cudaSuccess) return cuda_fail(e, "memset");
    if ((rc = run_program(c, *circ, plan, prog, (cplx*)c->arena.ptr, D, (cplx*)c->cmat.ptr))) return rc;
    e = cudaMemcpyAsync(out->d, (cplx*)c->arena.ptr + D, D * sizeof(cplx), cudaMemcpyDeviceToDevice, c->stream);
    if (e != cudaSuccess) return cuda_fail(e, "column copy");
    return stats_end(c);
}

int qgt_b200_gram(qgt_b200_ctx* c, const double* psi, const double* dpsi, size_t dim, size_t num_params,
                  double* metric, double* berry, double* q_full) {
    if (!c || !psi || !dpsi) return fail(QGT_B200_ERR_INVALID_ARG, "ctx/psi/dpsi is NULL");
    if (dim == 0 || num_params == 0) return fail(QGT_B200_ERR_INVALID_ARG, "empty problem");
    if (c->world > 1) return fail(QGT_B200_ERR_UNSUPPORTED, "qgt_b200_gram is single-GPU only");
    cudaSetDevice(c->device);
    const int P = (int)num_params;
    const uint64_t D = dim;
    auto is_device = [](const void* p) {
        cudaPointerAttributes a;
        if (cudaPointerGetAttributes(&a, p) != cudaSuccess) { cudaGetLastError(); return false; }
        return a.type == cudaMemoryTypeDevice || a.type == cudaMemoryTypeManaged;
    };
    int rc;
    stats_begin(c);
    const cplx* d_psi = (const cplx*)psi;
    const cplx* d_cols = (const cplx*)dpsi;
    size_t need = 0;
    const bool up_psi = !is_device(psi), up_cols = !is_device(dpsi);
    if (up_psi) need += D;
    if (up_cols) need += (size_t)P * D;
    if (need) {
        if ((rc = c->arena.reserve(need * sizeof(cplx)))) return rc;
        cplx* a = (cplx*)c->arena.ptr;
        cudaError_t e = cudaSuccess;
        if (up_cols) { e = cudaMemcpyAsync(a, dpsi, (size_t)P * D * sizeof(cplx), cudaMemcpyHostToDevice, c->stream); d_cols = a; a += (size_t)P * D; }
        if (up_psi && e == cudaSuccess) { e = cudaMemcpyAsync(a, psi, D * sizeof(cplx), cudaMemcpyHostToDevice, c->stream); d_psi = a; }
        if (e != cudaSuccess) return cuda_fail(e, "column upload");
    }
    std::vector<const cplx*> ptrs;
    std::vector<int> ids;
    for (int i = 0; i < P; i++) ptrs.push_back(d_cols + (size_t)i * D);      // A list
    for (int i = 0; i < P; i++) ptrs.push_back(d_cols + (size_t)i * D);      // B list = A list + psi
    ptrs.push_back(d_psi);
    for (int i = 0; i < P; i++) ids.push_back(i);
    for (int i = 0; i <= P; i++) ids.push_back(i);
    const size_t ptr_bytes = ptrs.size() * sizeof(cplx*), id_bytes = ids.size() * sizeof(int);
    if ((rc = c->aux.reserve(ptr_bytes + id_bytes))) return rc;
    const size_t cm = (size_t)(P + 1) * (P + 1);
    if ((rc = c->cmat.reserve(cm * sizeof(cplx)))) return rc;
    if ((rc = c->outbuf.reserve((size_t)P * P * 4 * sizeof(double)))) return rc;
    GramLaunch gl;
    const GramShape shp = gram_shape(P, P + 1);
    gl.na = P; gl.nb = P + 1; gl.D = D;
    gl.symmetric = 1;
    const size_t per_split = gram_configure(gl, shp);
    gl.ksplit = gram_choose_ksplit(gl, shp, c->num_sms);
    if ((rc = c->partial.reserve((size_t)gl.ksplit * per_split * sizeof(cplx)))) return rc;
    gl.partial = (cplx*)c->partial.ptr;
    cudaError_t e = cudaMemcpyAsync(c->aux.ptr, ptrs.data(), ptr_bytes, cudaMemcpyHostToDevice, c->stream);
    if (e == cudaSuccess) e = cudaMemcpyAsync((char*)c->aux.ptr + ptr_bytes, ids.data(), id_bytes, cudaMemcpyHostToDevice, c->stream);
    if (e == cudaSuccess) e = cudaMemsetAsync(c->cmat.ptr, 0, cm * sizeof(cplx), c->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(c->stream);
    if (e != cudaSuccess) return cuda_fail(e, "gram setup");
    gl.a_ptrs = (const cplx* const*)c->aux.ptr;
    gl.b_ptrs = gl.a_ptrs + P;
    const int* d_ids = (const int*)((char*)c->aux.ptr + ptr_bytes);
    c->timer.begin(c->stream, 1);
    e = launch_gram(gl, shp, c->stream);
    if (e == cudaSuccess) e = launch_gram_reduce(gl, shp, d_ids, d_ids + P, (cplx*)c->cmat.ptr, P + 1, c->stream);
    c->timer.end(c->stream);
    if (e != cudaSuccess) return cuda_fail(e, "gram launch");
    c->stats.gram_launches = 1;
    c->stats.gram_flops = 8.0 * (0.5 * P * (P + 1) + P) * (double)D;
    c->stats.gram_bytes = 16.0 * (double)D * (P + 1);
    double* d_metric = (double*)c->outbuf.ptr;
    double* d_berry = d_metric + (size_t)P * P;
    cplx* d_q = (cplx*)(d_berry + (size_t)P * P);
    e = launch_finalize((const cplx*)c->cmat.ptr, P, d_metric, d_berry, d_q, c->stream);
    const size_t pp = (size_t)P * P;
    if (metric && e == cudaSuccess) e = cudaMemcpyAsync(metric, d_metric, pp * sizeof(double), cudaMemcpyDeviceToHost, c->stream);
    if (berry && e == cudaSuccess) e = cudaMemcpyAsync(berry, d_berry, pp * sizeof(double), cudaMemcpyDeviceToHost, c->stream);
    if (q_full && e == cudaSuccess) e = cudaMemcpyAsync(q_full, d_q, pp * sizeof(cplx), cudaMemcpyDeviceToHost, c->stream);
    if (e != cudaSuccess) return cuda_fail(e, "gram result copy");
    return stats_end(c);
}

long qgt_b200_plan_dump(const qgt_b200_circuit* circ, const double* theta, int tile_qubits, int reg_qubits,
                        size_t column_slots, char* buf, size_t buflen) {
    if (!circ) return fail(QGT_B200_ERR_INVALID_ARG, "circuit is NULL");
    PlanOptions opt;
    if (tile_qubits) opt.tile_qubits = tile_qubits;
    if (reg_qubits) opt.reg_qubits = reg_qubits;
    std::vector<double> zeros((size_t)std::max(1, circ->num_params), 0.0);
    CircuitPlan plan;
    std::string err;
    int rc = build_plan(*circ, theta ? theta : zeros.data(), opt, plan, err);
    if (rc) return fail(rc, err);
    Program prog;
    const Program* pp = nullptr;
    if (column_slots) {
        if ((rc = build_qgt_program(plan, column_slots, true, prog, err))) return fail(rc, err);
        pp = &prog;
    }
    const std::string js = dump_json(*circ, plan, pp);
    if (buf && buflen > js.size()) std::memcpy(buf, js.c_str(), js.size() + 1);
    return (long)js.size();
}

long qgt_b200_plan_dump_sharded(const qgt_b200_circuit* circ, const double* theta, int world, int restore_identity,
                                int tile_qubits, int reg_qubits, size_t column_slots, char* buf, size_t buflen) {
    if (!circ) return fail(QGT_B200_ERR_INVALID_ARG, "circuit is NULL");
    if (world < 1 || (world & (world - 1))) return fail(QGT_B200_ERR_INVALID_ARG, "world must be a power of two");
    int gbits = 0;
    while ((1 << gbits) < world) gbits++;
    PlanOptions opt;
    if (tile_qubits) opt.tile_qubits = tile_qubits;
    if (reg_qubits) opt.reg_qubits = reg_qubits;
    std::vector<double> zeros((size_t)std::max(1, circ->num_params), 0.0);
    CircuitPlan plan;
    std::vector<MappedSegment> segs;
    std::string err;
    int rc = build_plan_sharded(*circ, theta ? theta : zeros.data(), opt, circ->num_qubits - gbits, restore_identity != 0, plan, segs, err);
    if (rc) return fail(rc, err);
    Program prog;
    const Program* pp = nullptr;
    if (column_slots) {
        if ((rc = build_qgt_program(plan, column_slots, restore_identity != 0, prog, err))) return fail(rc, err);
        pp = &prog;
    }
    const std::string js = dump_json(*circ, plan, pp, &segs);
    if (buf && buflen > js.size()) std::memcpy(buf, js.c_str(), js.size() + 1);
    return (long)js.size();
}

int qgt_b200_measure_peaks(qgt_b200_ctx* c, double* dmma_tflops, double* copy_gbs) {
    if (!c) return fail(QGT_B200_ERR_INVALID_ARG, "ctx is NULL");
    cudaSetDevice(c->device);
    int rc;
    if (dmma_tflops) {
        if ((rc = c->scratch.reserve((size_t)c->num_sms * 1024 * sizeof(double)))) return rc;
        cudaError_t e = measure_dmma_peak(c->num_sms, (double*)c->scratch.ptr, c->stream, dmma_tflops);
        if (e != cudaSuccess) return cuda_fail(e, "DMMA peak kernel");
    }
    if (copy_gbs) {
        const size_t bytes = (size_t)1 << 30;
        void *a = nullptr, *b = nullptr;
        cudaError_t e = cudaMalloc(&a, bytes);
        if (e == cudaSuccess) e = cudaMalloc(&b, bytes);
        cudaEvent_t e0, e1;
        cudaEventCreate(&e0); cudaEventCreate(&e1);
        float best = 1e30f;
        if (e == cudaSuccess) e = cudaMemsetAsync(a, 1, bytes, c->stream);
        for (int r = 0; r < 8 && e == cudaSuccess; r++) {
            cudaEventRecord(e0, c->stream);
            e = cudaMemcpyAsync(b, a, bytes, cudaMemcpyDeviceToDevice, c->stream);
            cudaEventRecord(e1, c->stream);
            if (e == cudaSuccess) e = cudaEventSynchronize(e1);
            float ms = 0.f;
            cudaEventElapsedTime(&ms, e0, e1);
            if (r > 0 && ms < best) best = ms;
        }
        cudaEventDestroy(e0); cudaEventDestroy(e1);
        if (a) cudaFree(a);
        if (b) cudaFree(b);
        if (e != cudaSuccess) return cuda_fail(e, "copy bandwidth measurement");
        *copy_gbs = 2.0 * (double)bytes / best * 1e-6;
    }
    return QGT_B200_OK;
}

long qgt_b200_plan_dump_fused(const qgt_b200_circuit* circ, const double* theta, int world, int tile_qubits, int reg_qubits,
                              size_t column_slots, char* buf, size_t buflen) {
    if (!circ) return fail(QGT_B200_ERR_INVALID_ARG, "circuit is NULL");
    if (world < 1 || (world & (world - 1))) return fail(QGT_B200_ERR_INVALID_ARG, "world must be a power of two");
    int gbits = 0;
    while ((1 << gbits) < world) gbits++;
    PlanOptions opt;
    if (tile_qubits) opt.tile_qubits = tile_qubits;
    if (reg_qubits) opt.reg_qubits = reg_qubits;
    std::vector<double> zeros((size_t)std::max(1, circ->num_params), 0.0);
    CircuitPlan plan;
    std::vector<MappedSegment> segs;
    std::string err;
    int rc = world > 1 ? build_plan_sharded(*circ, theta ? theta : zeros.data(), opt, circ->num_qubits - gbits, true, plan, segs, err)
                       : build_plan(*circ, theta ? theta : zeros.data(), opt, plan, err);
    if (rc) return fail(rc, err);
    if (!plan_supports_fused(plan)) return fail(QGT_B200_ERR_UNSUPPORTED, "plan does not qualify for the fused schedule");
    Program prog;
    int traj_mode = -1;
    if (const char* e = std::getenv("QGT_B200_FUSED_TRAJ")) traj_mode = std::atoi(e);      // test hook
    if ((rc = build_fused_program(plan, column_slots, true, prog, err, traj_mode))) return fail(rc, err);
    const std::string js = dump_json(*circ, plan, &prog, world > 1 ? &segs : nullptr);
    if (buf && buflen > js.size()) std::memcpy(buf, js.c_str(), js.size() + 1);
    return (long)js.size();
}

// Plan of the inverse circuit with the adjoint-gradient program (adjoint.cu): the fused one when the plan qualifies and
// `fused` is non-zero, else the per-run programs of the generic path concatenated (scratch_slots scratch columns).
long qgt_b200_plan_dump_gradient(const qgt_b200_circuit* circ, const double* theta, int fused, int scratch_slots, int tile_qubits,
                                 int reg_qubits, int world, char* buf, size_t buflen) {
    if (!circ) return fail(QGT_B200_ERR_INVALID_ARG, "circuit is NULL");
    if (world < 1 || (world & (world - 1))) return fail(QGT_B200_ERR_INVALID_ARG, "world must be a power of two");
    int gbits = 0;
    while ((1 << gbits) < world) gbits++;
    PlanOptions opt;
    if (tile_qubits) opt.tile_qubits = tile_qubits;
    if (reg_qubits) opt.reg_qubits = reg_qubits;
    std::vector<double> zeros((size_t)std::max(1, circ->num_params), 0.0);
    std::vector<qgt_b200_gate> inv_gates;
    invert_circuit(*circ, inv_gates);
    qgt_b200_circuit icirc = *circ;
    icirc.gates = inv_gates.data(); icirc.num_gates = inv_gates.size();
    CircuitPlan plan;
    std::vector<MappedSegment> segs;
    std::string err;
    int rc = world > 1 ? build_plan_sharded(icirc, theta ? theta : zeros.data(), opt, circ->num_qubits - gbits, false, plan, segs, err)
                       : build_plan(icirc, theta ? theta : zeros.data(), opt, plan, err);
    if (rc) return fail(rc, err);
    Program prog;
    if (fused) {
        if (!plan_supports_fused(plan)) return fail(QGT_B200_ERR_UNSUPPORTED, "plan does not qualify for the fused schedule");
        if ((rc = build_gradient_fused_program(plan, prog))) return fail(rc, "gradient program");
    } else {
        prog.num_slots = 3 + std::max(1, scratch_slots);
        std::vector<Program> progs;
        for (int r = 0; r < (int)plan.runs.size(); r++) {
            if ((rc = build_gradient_run_programs(plan, r, std::max(1, scratch_slots), progs))) return fail(rc, "gradient program");
            for (const Program& g : progs) prog.instrs.insert(prog.instrs.end(), g.instrs.begin(), g.instrs.end());
        }
    }
    const std::string js = dump_json(icirc, plan, &prog, world > 1 ? &segs : nullptr);
    if (buf && buflen > js.size()) std::memcpy(buf, js.c_str(), js.size() + 1);
    return (long)js.size();
}

}  // extern "C"

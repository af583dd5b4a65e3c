/*
 * gpu_compat.c — the reference's device-memory seam served by the sm_100a library.
 *
 *   gpu_malloc / qgt_gpu_free_buffer / gpu_memcpy_*          core/quantum_geometric_gpu.c:30-165
 *   qg_gpu_init / cleanup / get_device_count / get_device_info / set_device   :212-456
 *   qg_gpu_allocate / allocate_pinned / free / memcpy_to_device / memcpy_to_host   :458-611
 *   qg_gpu_create_stream / destroy_stream / synchronize_stream / synchronize   :715-860
 *   qg_gpu_get_error_string / qg_gpu_get_last_error   :897-920
 *
 * Same names, argument meaning, validation order and gpu_error_t values.  Differences on purpose: device buffers
 * are never backed by host malloc (the reference falls back to malloc when it has no GPU back-end compiled in,
 * :41-60, :480-497) — without an sm_100 device qg_gpu_init returns QG_GPU_ERROR_NO_DEVICE and the allocators
 * fail; the reference's process-wide memory pool is not reproduced (cudaMalloc is called directly).
 */
#include <pthread.h>
#include <string.h>

#include "compat_common.h"

#define QGT_COMPAT_MAX_STREAMS 64

static struct {
    int initialized;
    int num_devices;
    gpu_error_t last_error;
    void* streams[QGT_COMPAT_MAX_STREAMS];
} g_gpu;
static pthread_mutex_t g_gpu_lock = PTHREAD_MUTEX_INITIALIZER;

static int set_err(gpu_error_t e) { g_gpu.last_error = e; return (int)e; }

static gpu_error_t map_status(int rc) {
    switch (rc) {
    case QGT_B200_OK: return QG_GPU_SUCCESS;
    case QGT_B200_ERR_NO_DEVICE: return QG_GPU_ERROR_NO_DEVICE;
    case QGT_B200_ERR_NO_MEMORY: return QG_GPU_ERROR_OUT_OF_MEMORY;
    case QGT_B200_ERR_INVALID_ARG: return QG_GPU_ERROR_INVALID_VALUE;
    default: return QG_GPU_ERROR_INTERNAL;
    }
}

/* ---- low level (qgt_error_t returns) ---------------------------------------------------------------------- */
qgt_error_t gpu_malloc(void** ptr, size_t size) {
    if (!ptr || size == 0) return QGT_ERROR_INVALID_PARAMETER;
    int rc = qgt_b200_mem_alloc(ptr, size);
    if (rc == QGT_B200_ERR_NO_DEVICE) return QGT_ERROR_GPU_NOT_AVAILABLE;
    if (rc == QGT_B200_ERR_NO_MEMORY) return QGT_ERROR_GPU_OUT_OF_MEMORY;
    return rc == QGT_B200_OK ? QGT_SUCCESS : QGT_ERROR_GPU_INTERNAL;
}

qgt_error_t qgt_gpu_free_buffer(void* ptr) {
    if (!ptr) return QGT_ERROR_INVALID_PARAMETER;
    return qgt_b200_mem_free(ptr) == QGT_B200_OK ? QGT_SUCCESS : QGT_ERROR_GPU_INTERNAL;
}

qgt_error_t gpu_memcpy_host_to_device(void* dst, const void* src, size_t size) {
    if (!dst || !src) return QGT_ERROR_INVALID_PARAMETER;
    if (size == 0) return QGT_SUCCESS;
    return qgt_b200_memcpy_h2d(dst, src, size) == QGT_B200_OK ? QGT_SUCCESS : QGT_ERROR_GPU_INTERNAL;
}

qgt_error_t gpu_memcpy_device_to_host(void* dst, const void* src, size_t size) {
    if (!dst || !src) return QGT_ERROR_INVALID_PARAMETER;
    if (size == 0) return QGT_SUCCESS;
    return qgt_b200_memcpy_d2h(dst, src, size) == QGT_B200_OK ? QGT_SUCCESS : QGT_ERROR_GPU_INTERNAL;
}

/* ---- system ---------------------------------------------------------------------------------------------- */
int qg_gpu_init(void) {
    pthread_mutex_lock(&g_gpu_lock);
    int rc = QG_GPU_SUCCESS;
    if (!g_gpu.initialized) {
        const int n = qgt_b200_device_count();
        if (n <= 0) rc = set_err(QG_GPU_ERROR_NO_DEVICE);
        else { g_gpu.num_devices = n; g_gpu.initialized = 1; g_gpu.last_error = QG_GPU_SUCCESS; }
    }
    pthread_mutex_unlock(&g_gpu_lock);
    return rc;
}

void qg_gpu_cleanup(void) {
    pthread_mutex_lock(&g_gpu_lock);
    if (g_gpu.initialized) {
        qgt_b200_device_synchronize();
        for (int i = 0; i < QGT_COMPAT_MAX_STREAMS; i++)
            if (g_gpu.streams[i]) { qgt_b200_stream_destroy(g_gpu.streams[i]); g_gpu.streams[i] = NULL; }
        g_gpu.initialized = 0;
        g_gpu.num_devices = 0;
    }
    pthread_mutex_unlock(&g_gpu_lock);
}

void qg_gpu_shutdown(void) { qg_gpu_cleanup(); }

int qg_gpu_get_device_count(int* count) {
    if (!g_gpu.initialized || !count) return set_err(QG_GPU_ERROR_NOT_INITIALIZED);
    *count = g_gpu.num_devices;
    return QG_GPU_SUCCESS;
}

int qg_gpu_get_device_info(int device_id, gpu_device_info_t* info) {
    if (!g_gpu.initialized || !info) return set_err(QG_GPU_ERROR_NOT_INITIALIZED);
    if (device_id < 0 || device_id >= g_gpu.num_devices) return set_err(QG_GPU_ERROR_INVALID_DEVICE);
    qgt_b200_device_info di;
    if (qgt_b200_device_info_get(device_id, &di) != QGT_B200_OK) return set_err(QG_GPU_ERROR_INVALID_DEVICE);
    memset(info, 0, sizeof *info);
    info->device_id = di.device;
    strncpy(info->name, di.name, MAX_GPU_NAME_LENGTH - 1);
    info->total_memory = di.total_memory;
    info->free_memory = di.free_memory;
    info->compute_capability_major = di.cc_major;
    info->compute_capability_minor = di.cc_minor;
    info->max_threads_per_block = di.max_threads_per_block;
    for (int i = 0; i < 3; i++) { info->max_block_dimensions[i] = di.max_block_dim[i]; info->max_grid_dimensions[i] = di.max_grid_dim[i]; }
    info->compute_units = di.num_sms;
    info->backend_type = 1;                     /* CUDA */
    info->supports_unified_memory = di.unified_addressing != 0;
    return QG_GPU_SUCCESS;
}

int qg_gpu_set_device(int device_id) {
    if (!g_gpu.initialized) return set_err(QG_GPU_ERROR_NOT_INITIALIZED);
    if (device_id < 0 || device_id >= g_gpu.num_devices) return set_err(QG_GPU_ERROR_INVALID_DEVICE);
    if (qgt_b200_set_device(device_id) != QGT_B200_OK) return set_err(QG_GPU_ERROR_INVALID_DEVICE);
    return QG_GPU_SUCCESS;
}

gpu_error_t qg_gpu_get_last_error(void) { return g_gpu.last_error; }

const char* qg_gpu_get_error_string(gpu_error_t error) {
    switch (error) {
    case QG_GPU_SUCCESS: return "Success";
    case QG_GPU_ERROR_NO_DEVICE: return "No GPU device available";
    case QG_GPU_ERROR_OUT_OF_MEMORY: return "Out of memory";
    case QG_GPU_ERROR_INVALID_DEVICE: return "Invalid device";
    case QG_GPU_ERROR_LAUNCH_FAILED: return "Kernel launch failed";
    case QG_GPU_ERROR_INVALID_VALUE: return "Invalid value";
    case QG_GPU_ERROR_NOT_INITIALIZED: return "GPU not initialized";
    default: return "Unknown error";
    }
}

/* ---- buffers ---------------------------------------------------------------------------------------------- */
int qg_gpu_allocate(gpu_buffer_t* buffer, size_t size) {
    if (!g_gpu.initialized || !buffer || size == 0) return set_err(QG_GPU_ERROR_INVALID_VALUE);
    buffer->size = size;
    buffer->is_pinned = false;
    buffer->device_ptr = NULL;
    int rc = qgt_b200_mem_alloc(&buffer->device_ptr, size);
    if (rc != QGT_B200_OK) { buffer->size = 0; return set_err(rc == QGT_B200_ERR_NO_DEVICE ? QG_GPU_ERROR_NO_DEVICE : QG_GPU_ERROR_OUT_OF_MEMORY); }
    return QG_GPU_SUCCESS;
}

int qg_gpu_allocate_pinned(gpu_buffer_t* buffer, size_t size) {
    if (!g_gpu.initialized || !buffer || size == 0) return set_err(QG_GPU_ERROR_INVALID_VALUE);
    buffer->size = size;
    buffer->is_pinned = true;
    buffer->device_ptr = NULL;
    int rc = qgt_b200_mem_alloc_pinned(&buffer->device_ptr, size);
    if (rc != QGT_B200_OK) { buffer->size = 0; return set_err(rc == QGT_B200_ERR_NO_DEVICE ? QG_GPU_ERROR_NO_DEVICE : QG_GPU_ERROR_OUT_OF_MEMORY); }
    return QG_GPU_SUCCESS;
}

int qg_gpu_free(gpu_buffer_t* buffer) {
    if (!g_gpu.initialized || !buffer || !buffer->device_ptr) return set_err(QG_GPU_ERROR_INVALID_VALUE);
    if (buffer->is_pinned) qgt_b200_mem_free_pinned(buffer->device_ptr);
    else qgt_b200_mem_free(buffer->device_ptr);
    buffer->device_ptr = NULL;
    buffer->size = 0;
    return QG_GPU_SUCCESS;
}

int qg_gpu_memcpy_to_device(gpu_buffer_t* dst, const void* src, size_t size) {
    if (!g_gpu.initialized || !dst || !src || !dst->device_ptr || size > dst->size) return set_err(QG_GPU_ERROR_INVALID_VALUE);
    if (dst->is_pinned) { memcpy(dst->device_ptr, src, size); return QG_GPU_SUCCESS; }
    if (qgt_b200_memcpy_h2d(dst->device_ptr, src, size) != QGT_B200_OK) return set_err(QG_GPU_ERROR_LAUNCH_FAILED);
    return QG_GPU_SUCCESS;
}

int qg_gpu_memcpy_to_host(void* dst, const gpu_buffer_t* src, size_t size) {
    if (!g_gpu.initialized || !dst || !src || !src->device_ptr || size > src->size) return set_err(QG_GPU_ERROR_INVALID_VALUE);
    if (src->is_pinned) { memcpy(dst, src->device_ptr, size); return QG_GPU_SUCCESS; }
    if (qgt_b200_memcpy_d2h(dst, src->device_ptr, size) != QGT_B200_OK) return set_err(QG_GPU_ERROR_LAUNCH_FAILED);
    return QG_GPU_SUCCESS;
}

/* ---- streams (small integer handles, as in the reference) -------------------------------------------------- */
int qg_gpu_create_stream(int* stream_id) {
    if (!g_gpu.initialized || !stream_id) return set_err(QG_GPU_ERROR_NOT_INITIALIZED);
    *stream_id = -1;
    void* s = NULL;
    int rc = qgt_b200_stream_create(&s);
    if (rc != QGT_B200_OK) return set_err(map_status(rc) == QG_GPU_ERROR_NO_DEVICE ? QG_GPU_ERROR_NO_DEVICE : QG_GPU_ERROR_LAUNCH_FAILED);
    pthread_mutex_lock(&g_gpu_lock);
    for (int i = 0; i < QGT_COMPAT_MAX_STREAMS; i++)
        if (!g_gpu.streams[i]) { g_gpu.streams[i] = s; *stream_id = i; break; }
    pthread_mutex_unlock(&g_gpu_lock);
    if (*stream_id < 0) { qgt_b200_stream_destroy(s); return set_err(QG_GPU_ERROR_OUT_OF_MEMORY); }
    return QG_GPU_SUCCESS;
}

int qg_gpu_destroy_stream(int stream_id) {
    if (!g_gpu.initialized) return set_err(QG_GPU_ERROR_NOT_INITIALIZED);
    if (stream_id < 0 || stream_id >= QGT_COMPAT_MAX_STREAMS || !g_gpu.streams[stream_id]) return set_err(QG_GPU_ERROR_INVALID_VALUE);
    pthread_mutex_lock(&g_gpu_lock);
    void* s = g_gpu.streams[stream_id];
    g_gpu.streams[stream_id] = NULL;
    pthread_mutex_unlock(&g_gpu_lock);
    return qgt_b200_stream_destroy(s) == QGT_B200_OK ? QG_GPU_SUCCESS : set_err(QG_GPU_ERROR_SYNC_FAILED);
}

int qg_gpu_synchronize_stream(int stream_id) {
    if (!g_gpu.initialized) return set_err(QG_GPU_ERROR_NOT_INITIALIZED);
    if (stream_id < 0 || stream_id >= QGT_COMPAT_MAX_STREAMS || !g_gpu.streams[stream_id]) return set_err(QG_GPU_ERROR_INVALID_VALUE);
    return qgt_b200_stream_synchronize(g_gpu.streams[stream_id]) == QGT_B200_OK ? QG_GPU_SUCCESS : set_err(QG_GPU_ERROR_SYNC_FAILED);
}

int qg_gpu_synchronize(void) {
    if (!g_gpu.initialized) return set_err(QG_GPU_ERROR_NOT_INITIALIZED);
    return qgt_b200_device_synchronize() == QGT_B200_OK ? QG_GPU_SUCCESS : set_err(QG_GPU_ERROR_SYNC_FAILED);
}

// c64_ops.cu — operations on interleaved complex-FLOAT buffers: what the quantum_* slots of the reference's
// ComputeBackendOps vtable bind (include/quantum_geometric/supercomputer/compute_backend.h:122-211 of the reference;
// CPU semantics in src/quantum_geometric/supercomputer/compute_simd.c:46-57, 117-128, 163-216, 284-293 and
// backends/compute_cpu.c:341-466).  Buffers may be host or device pointers (cudaPointerGetAttributes): host buffers
// are staged through the context's scratch buffer, device buffers are used in place and the call is stream ordered.
//
//   * gate on target qubits: every thread owns one group of 2^K amplitudes (K <= 4), the matrix sits in shared memory;
//     one read and one write of the state = 16 * D bytes, HBM-bound;
//   * dense matrix-vector product (the reference's quantum_unitary: a state_size x state_size matrix): one warp per
//     row, bound by reading the matrix once;
//   * norm / inner product / diagonal expectation: grid-stride single pass, double accumulators, fixed-order two-stage
//     sums (the reference accumulates in float; results agree to float rounding of the inputs);
//   * complex matrix product C[m x k] = A[m x n] B[n x k] (quantum_tensor_contract): 32 x 32 shared-memory tiles.
#include <cuda_runtime.h>

#include <algorithm>
#include <cmath>
#include <cstring>
#include <vector>

#include "ctx.hpp"

namespace qgt {

static bool c64_is_device(const void* p) {
    cudaPointerAttributes a;
    if (cudaPointerGetAttributes(&a, p) != cudaSuccess) { cudaGetLastError(); return false; }
    return a.type == cudaMemoryTypeDevice || a.type == cudaMemoryTypeManaged;
}

__device__ __forceinline__ double c64_warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

template <int NV>
__device__ __forceinline__ void c64_block_store(double (&v)[NV], double* partial) {
    __shared__ double ws[8][NV];
#pragma unroll
    for (int j = 0; j < NV; ++j) v[j] = c64_warp_sum(v[j]);
    if ((threadIdx.x & 31) == 0)
#pragma unroll
        for (int j = 0; j < NV; ++j) ws[threadIdx.x >> 5][j] = v[j];
    __syncthreads();
    if (threadIdx.x < NV) {
        double s = 0.0;
        for (int w = 0; w < (int)(blockDim.x >> 5); ++w) s += ws[w][threadIdx.x];
        partial[(size_t)blockIdx.x * NV + threadIdx.x] = s;
    }
}

// second stage: out_f[j] (float) and out_d[j] (double) = sum of the per-CTA partials, fixed order
__global__ void c64_final_kernel(const double* partial, int nblocks, int nv, float* out_f, double* out_d, int sqrt_first) {
    const int j = threadIdx.x;
    if (j >= nv) return;
    double s = 0.0;
    for (int b = 0; b < nblocks; ++b) s += partial[(size_t)b * nv + j];
    if (sqrt_first && j == 0) s = sqrt(s);
    if (out_d) out_d[j] = s;
    if (out_f) out_f[j] = (float)s;
}

__global__ void __launch_bounds__(256) c64_norm2_kernel(const float2* a, uint64_t n, double* partial) {
    double v[1] = {0.0};
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        const float2 z = a[i];
        v[0] += (double)z.x * z.x + (double)z.y * z.y;
    }
    c64_block_store<1>(v, partial);
}

// a_i *= 1/sqrt(norm2[0]) unless the norm is below the reference's 1e-10 cut (compute_cpu.c:378-381)
__global__ void __launch_bounds__(256) c64_scale_by_norm_kernel(float2* a, uint64_t n, const double* norm) {
    const double nv = norm[0];
    if (!(nv > 1e-10)) return;
    const float s = (float)(1.0 / nv);
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        float2 z = a[i];
        z.x *= s; z.y *= s;
        a[i] = z;
    }
}

// <a|b> = sum conj(a_i) b_i   (compute_simd.c:117-128)
__global__ void __launch_bounds__(256) c64_inner_kernel(const float2* a, const float2* b, uint64_t n, double* partial) {
    double v[2] = {0.0, 0.0};
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        const float2 x = a[i], y = b[i];
        v[0] += (double)x.x * y.x + (double)x.y * y.y;
        v[1] += (double)x.x * y.y - (double)x.y * y.x;
    }
    c64_block_store<2>(v, partial);
}

// sum |a_i|^2 o_i for a real diagonal observable (compute_simd.c:284-293)
__global__ void __launch_bounds__(256) c64_expect_kernel(const float2* a, const float* obs, uint64_t n, double* partial) {
    double v[1] = {0.0};
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        const float2 z = a[i];
        v[0] += ((double)z.x * z.x + (double)z.y * z.y) * (double)obs[i];
    }
    c64_block_store<1>(v, partial);
}

// out = M x with M row-major dim x dim: one warp per row (compute_simd.c:163-179)
__global__ void __launch_bounds__(256) c64_matvec_kernel(float2* out, const float2* M, const float2* x, uint64_t rows, uint64_t cols) {
    const uint64_t row = (uint64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (row >= rows) return;
    const int lane = threadIdx.x & 31;
    const float2* mr = M + row * cols;
    double sr = 0.0, si = 0.0;
    for (uint64_t j = lane; j < cols; j += 32) {
        const float2 a = mr[j], b = x[j];
        sr += (double)a.x * b.x - (double)a.y * b.y;
        si += (double)a.x * b.y + (double)a.y * b.x;
    }
    sr = c64_warp_sum(sr); si = c64_warp_sum(si);
    if (lane == 0) out[row] = make_float2((float)sr, (float)si);
}

struct C64Targets { int bit[4]; int sorted[4]; };

// K-qubit gate: matrix index bit j <-> qubit tg.bit[j]; group base = thread id with zero bits inserted at the sorted targets
template <int K>
__global__ void __launch_bounds__(256) c64_gate_kernel(float2* st, uint64_t groups, const float2* M, C64Targets tg) {
    constexpr int N = 1 << K;
    __shared__ float2 sm[N * N];
    for (int i = threadIdx.x; i < N * N; i += blockDim.x) sm[i] = M[i];
    __syncthreads();
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t g = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; g < groups; g += stride) {
        uint64_t base = g;
#pragma unroll
        for (int j = 0; j < K; ++j) {
            const uint64_t low = base & (((uint64_t)1 << tg.sorted[j]) - 1);
            base = ((base >> tg.sorted[j]) << (tg.sorted[j] + 1)) | low;
        }
        float2 v[N];
#pragma unroll
        for (int a = 0; a < N; ++a) {
            uint64_t idx = base;
#pragma unroll
            for (int j = 0; j < K; ++j) if (a >> j & 1) idx |= (uint64_t)1 << tg.bit[j];
            v[a] = st[idx];
        }
#pragma unroll
        for (int r = 0; r < N; ++r) {
            float sr = 0.f, si = 0.f;
#pragma unroll
            for (int a = 0; a < N; ++a) {
                const float2 m = sm[r * N + a];
                sr = fmaf(m.x, v[a].x, sr); sr = fmaf(-m.y, v[a].y, sr);
                si = fmaf(m.x, v[a].y, si); si = fmaf(m.y, v[a].x, si);
            }
            uint64_t idx = base;
#pragma unroll
            for (int j = 0; j < K; ++j) if (r >> j & 1) idx |= (uint64_t)1 << tg.bit[j];
            st[idx] = make_float2(sr, si);
        }
    }
}

// C[m x k] = A[m x n] B[n x k], row-major complex float, 32 x 32 tiles, 4 outputs per thread (compute_simd.c:195-216)
__global__ void __launch_bounds__(256) c64_matmul_kernel(float2* C, const float2* A, const float2* B, int m, int n, int k) {
    __shared__ float2 sa[32][33], sb[32][33];
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;      // 8 rows of threads, each owning 4 rows of the tile
    const int row0 = blockIdx.y * 32, col0 = blockIdx.x * 32;
    float2 acc[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) acc[q] = make_float2(0.f, 0.f);
    for (int l0 = 0; l0 < n; l0 += 32) {
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const int r = ty + 8 * q;
            sa[r][tx] = (row0 + r < m && l0 + tx < n) ? A[(size_t)(row0 + r) * n + l0 + tx] : make_float2(0.f, 0.f);
            sb[r][tx] = (l0 + r < n && col0 + tx < k) ? B[(size_t)(l0 + r) * k + col0 + tx] : make_float2(0.f, 0.f);
        }
        __syncthreads();
#pragma unroll 8
        for (int l = 0; l < 32; ++l) {
            const float2 b = sb[l][tx];
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const float2 a = sa[ty + 8 * q][l];
                acc[q].x = fmaf(a.x, b.x, acc[q].x); acc[q].x = fmaf(-a.y, b.y, acc[q].x);
                acc[q].y = fmaf(a.x, b.y, acc[q].y); acc[q].y = fmaf(a.y, b.x, acc[q].y);
            }
        }
        __syncthreads();
    }
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        const int r = row0 + ty + 8 * q;
        if (r < m && col0 + tx < k) C[(size_t)r * k + col0 + tx] = acc[q];
    }
}

__global__ void __launch_bounds__(256) c64_widen_kernel(cplx* dst, const float2* src, uint64_t n) {
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        const float2 f = src[i];
        cplx z; z.x = (double)f.x; z.y = (double)f.y;
        dst[i] = z;
    }
}

static unsigned c64_grid(const qgt_b200_ctx* c, uint64_t n) {
    const uint64_t want = (n + 255) / 256;
    const uint64_t cap = (uint64_t)c->num_sms * 8;
    return (unsigned)std::max<uint64_t>(1, std::min(want, cap));
}

// Staging of host buffers: a bump allocator over c->scratch (one reserve per call, so pointers stay valid)
struct C64Stage {
    qgt_b200_ctx* c;
    cudaStream_t st;
    char* base = nullptr;
    size_t off = 0;
    bool any_host = false;
    int begin(size_t bytes) {
        int rc = c->scratch.reserve(bytes + 1024);
        base = (char*)c->scratch.ptr;
        return rc;
    }
    void* take(size_t bytes) { void* p = base + off; off += (bytes + 255) & ~(size_t)255; return p; }
    // device view of a read-only input
    int in(const void* p, size_t bytes, const void** dev) {
        if (c64_is_device(p)) { *dev = p; return QGT_B200_OK; }
        void* d = take(bytes);
        cudaError_t e = cudaMemcpyAsync(d, p, bytes, cudaMemcpyHostToDevice, st);
        if (e != cudaSuccess) return cuda_fail(e, "c64 staging H2D");
        any_host = true;
        *dev = d;
        return QGT_B200_OK;
    }
    // device view of an output (copied back by out_done when the caller's pointer is host memory)
    void* out(void* p, size_t bytes, bool load, int* rc) {
        *rc = QGT_B200_OK;
        if (c64_is_device(p)) return p;
        void* d = take(bytes);
        any_host = true;
        if (load) {
            cudaError_t e = cudaMemcpyAsync(d, p, bytes, cudaMemcpyHostToDevice, st);
            if (e != cudaSuccess) *rc = cuda_fail(e, "c64 staging H2D");
        }
        return d;
    }
    int out_done(void* p, const void* d, size_t bytes) {
        if (p == d) return QGT_B200_OK;
        cudaError_t e = cudaMemcpyAsync(p, d, bytes, cudaMemcpyDeviceToHost, st);
        return e == cudaSuccess ? QGT_B200_OK : cuda_fail(e, "c64 staging D2H");
    }
    int finish() {
        cudaError_t e = cudaGetLastError();
        if (e == cudaSuccess && any_host) e = cudaStreamSynchronize(st);
        return e == cudaSuccess ? QGT_B200_OK : cuda_fail(e, "c64 operation");
    }
};

static size_t pad256(size_t b) { return (b + 255) & ~(size_t)255; }

}  // namespace qgt

using namespace qgt;

extern "C" {

int qgt_b200_c64_apply_matrix(qgt_b200_ctx* c, float* state, size_t dim, const float* matrix, size_t mdim,
                              const int32_t* targets, void* stream) {
    if (!c || !state || !matrix) return fail(QGT_B200_ERR_INVALID_ARG, "ctx/state/matrix is NULL");
    if (dim == 0 || mdim == 0 || mdim > dim) return fail(QGT_B200_ERR_DIMENSION, "matrix larger than the state or empty");
    cudaSetDevice(c->device);
    cudaStream_t st = stream ? (cudaStream_t)stream : c->stream;
    C64Stage sg{c, st};
    const size_t sbytes = dim * sizeof(float2), mbytes = mdim * mdim * sizeof(float2);
    int rc;
    if (mdim == dim && !targets) {
        // the reference's quantum_unitary: a dense dim x dim matrix (compute_cpu.c:341-366)
        if ((rc = sg.begin(2 * pad256(sbytes) + pad256(mbytes)))) return rc;
        const void *dm, *dx;
        if ((rc = sg.in(matrix, mbytes, &dm))) return rc;
        float2* y = (float2*)sg.take(sbytes);
        if ((rc = sg.in(state, sbytes, &dx))) return rc;
        c64_matvec_kernel<<<(unsigned)((dim + 7) / 8), 256, 0, st>>>(y, (const float2*)dm, (const float2*)dx, dim, dim);
        cudaError_t e = cudaMemcpyAsync(state, y, sbytes, dx == state ? cudaMemcpyDeviceToDevice : cudaMemcpyDeviceToHost, st);
        if (e != cudaSuccess) return cuda_fail(e, "matvec result copy");
        return sg.finish();
    }
    int K = 0;
    while (((size_t)1 << K) < mdim) K++;
    if (((size_t)1 << K) != mdim || (dim & (dim - 1))) return fail(QGT_B200_ERR_DIMENSION, "gate and state dimensions must be powers of two");
    if (K < 1 || K > 4) return fail(QGT_B200_ERR_UNSUPPORTED, "gates on 1 to 4 target qubits are supported");
    int nq = 0;
    while (((size_t)1 << nq) < dim) nq++;
    C64Targets tg;
    for (int j = 0; j < 4; j++) tg.bit[j] = tg.sorted[j] = 0;
    for (int j = 0; j < K; j++) {
        const int t = targets ? targets[j] : j;
        if (t < 0 || t >= nq) return fail(QGT_B200_ERR_INVALID_ARG, "target qubit out of range");
        for (int i = 0; i < j; i++) if (tg.bit[i] == t) return fail(QGT_B200_ERR_INVALID_ARG, "duplicate target qubit");
        tg.bit[j] = tg.sorted[j] = t;
    }
    std::sort(tg.sorted, tg.sorted + K);
    if ((rc = sg.begin(pad256(sbytes) + pad256(mbytes)))) return rc;
    const void* dm;
    if ((rc = sg.in(matrix, mbytes, &dm))) return rc;
    float2* ds = (float2*)sg.out(state, sbytes, true, &rc);
    if (rc) return rc;
    const uint64_t groups = dim >> K;
    const unsigned grid = c64_grid(c, groups);
    switch (K) {
        case 1: c64_gate_kernel<1><<<grid, 256, 0, st>>>(ds, groups, (const float2*)dm, tg); break;
        case 2: c64_gate_kernel<2><<<grid, 256, 0, st>>>(ds, groups, (const float2*)dm, tg); break;
        case 3: c64_gate_kernel<3><<<grid, 256, 0, st>>>(ds, groups, (const float2*)dm, tg); break;
        default: c64_gate_kernel<4><<<grid, 256, 0, st>>>(ds, groups, (const float2*)dm, tg); break;
    }
    if ((rc = sg.out_done(state, ds, sbytes))) return rc;
    return sg.finish();
}

int qgt_b200_c64_normalize(qgt_b200_ctx* c, float* state, size_t dim, float* norm_before, void* stream) {
    if (!c || !state || dim == 0) return fail(QGT_B200_ERR_INVALID_ARG, "ctx/state is NULL or the state is empty");
    cudaSetDevice(c->device);
    cudaStream_t st = stream ? (cudaStream_t)stream : c->stream;
    C64Stage sg{c, st};
    const size_t sbytes = dim * sizeof(float2);
    const unsigned grid = c64_grid(c, dim);
    int rc;
    if ((rc = sg.begin(pad256(sbytes) + pad256((grid + 8) * sizeof(double)) + 512))) return rc;
    float2* ds = (float2*)sg.out(state, sbytes, true, &rc);
    if (rc) return rc;
    double* partial = (double*)sg.take((grid + 8) * sizeof(double));
    double* nrm = partial + grid;
    float* nrm_f = (float*)sg.take(256);
    c64_norm2_kernel<<<grid, 256, 0, st>>>(ds, dim, partial);
    c64_final_kernel<<<1, 32, 0, st>>>(partial, (int)grid, 1, nrm_f, nrm, 1);
    c64_scale_by_norm_kernel<<<grid, 256, 0, st>>>(ds, dim, nrm);
    if ((rc = sg.out_done(state, ds, sbytes))) return rc;
    if (norm_before) {
        cudaError_t e = cudaMemcpyAsync(norm_before, nrm_f, sizeof(float), c64_is_device(norm_before) ? cudaMemcpyDeviceToDevice : cudaMemcpyDeviceToHost, st);
        if (e != cudaSuccess) return cuda_fail(e, "norm copy");
        if (!c64_is_device(norm_before)) sg.any_host = true;
    }
    return sg.finish();
}

static int c64_reduce2(qgt_b200_ctx* c, const float* a, const float* b, size_t dim, int nv, int kind, float* out, void* stream) {
    cudaSetDevice(c->device);
    cudaStream_t st = stream ? (cudaStream_t)stream : c->stream;
    C64Stage sg{c, st};
    const size_t abytes = dim * sizeof(float2), bbytes = kind == 1 ? dim * sizeof(float) : abytes;
    const unsigned grid = c64_grid(c, dim);
    int rc;
    if ((rc = sg.begin(pad256(abytes) + pad256(bbytes) + pad256(((size_t)grid * 2 + 8) * sizeof(double)) + 512))) return rc;
    const void *da, *db;
    if ((rc = sg.in(a, abytes, &da))) return rc;
    if (b == a) db = da;
    else if ((rc = sg.in(b, bbytes, &db))) return rc;
    double* partial = (double*)sg.take(((size_t)grid * 2 + 8) * sizeof(double));
    float* res = (float*)sg.take(256);
    if (kind == 0) c64_inner_kernel<<<grid, 256, 0, st>>>((const float2*)da, (const float2*)db, dim, partial);
    else c64_expect_kernel<<<grid, 256, 0, st>>>((const float2*)da, (const float*)db, dim, partial);
    c64_final_kernel<<<1, 32, 0, st>>>(partial, (int)grid, nv, res, nullptr, 0);
    const bool dev_out = c64_is_device(out);
    cudaError_t e = cudaMemcpyAsync(out, res, nv * sizeof(float), dev_out ? cudaMemcpyDeviceToDevice : cudaMemcpyDeviceToHost, st);
    if (e != cudaSuccess) return cuda_fail(e, "reduction result copy");
    if (!dev_out) sg.any_host = true;
    return sg.finish();
}

int qgt_b200_c64_inner_product(qgt_b200_ctx* c, const float* a, const float* b, size_t dim, float* out, void* stream) {
    if (!c || !a || !b || !out || dim == 0) return fail(QGT_B200_ERR_INVALID_ARG, "ctx/a/b/out is NULL or the states are empty");
    return c64_reduce2(c, a, b, dim, 2, 0, out, stream);
}

int qgt_b200_c64_expectation_diag(qgt_b200_ctx* c, const float* state, const float* observable, size_t dim, float* out, void* stream) {
    if (!c || !state || !observable || !out || dim == 0) return fail(QGT_B200_ERR_INVALID_ARG, "ctx/state/observable/out is NULL or the state is empty");
    return c64_reduce2(c, state, observable, dim, 1, 1, out, stream);
}

int qgt_b200_c64_matmul(qgt_b200_ctx* c, float* result, const float* a, const float* b, size_t m, size_t n, size_t k, void* stream) {
    if (!c || !result || !a || !b) return fail(QGT_B200_ERR_INVALID_ARG, "ctx/result/a/b is NULL");
    if (m == 0 || n == 0 || k == 0) return fail(QGT_B200_ERR_DIMENSION, "empty matrix product");
    if (m > 0x7fffffff || n > 0x7fffffff || k > 0x7fffffff) return fail(QGT_B200_ERR_DIMENSION, "matrix dimension too large");
    cudaSetDevice(c->device);
    cudaStream_t st = stream ? (cudaStream_t)stream : c->stream;
    C64Stage sg{c, st};
    const size_t ab = m * n * sizeof(float2), bb = n * k * sizeof(float2), cb = m * k * sizeof(float2);
    int rc;
    if ((rc = sg.begin(pad256(ab) + pad256(bb) + pad256(cb)))) return rc;
    const void *da, *db;
    if ((rc = sg.in(a, ab, &da))) return rc;
    if ((rc = sg.in(b, bb, &db))) return rc;
    float2* dc = (float2*)sg.out(result, cb, false, &rc);
    if (rc) return rc;
    dim3 grid((unsigned)((k + 31) / 32), (unsigned)((m + 31) / 32));
    c64_matmul_kernel<<<grid, 256, 0, st>>>(dc, (const float2*)da, (const float2*)db, (int)m, (int)n, (int)k);
    if ((rc = sg.out_done(result, dc, cb))) return rc;
    return sg.finish();
}

int qgt_b200_gram_c64(qgt_b200_ctx* c, const float* psi, const float* dpsi, size_t dim, size_t num_params,
                      double* metric, double* berry, double* q_full) {
    if (!c || !psi || !dpsi) return fail(QGT_B200_ERR_INVALID_ARG, "ctx/psi/dpsi is NULL");
    if (dim == 0 || num_params == 0) return fail(QGT_B200_ERR_INVALID_ARG, "empty problem");
    cudaSetDevice(c->device);
    const size_t total = (num_params + 1) * dim;
    int rc;
    if ((rc = c->arena.reserve(total * sizeof(cplx)))) return rc;
    C64Stage sg{c, c->stream};
    if ((rc = sg.begin(pad256(dim * sizeof(float2)) + pad256(num_params * dim * sizeof(float2))))) return rc;
    const void *dp, *dc;
    if ((rc = sg.in(psi, dim * sizeof(float2), &dp))) return rc;
    if ((rc = sg.in(dpsi, num_params * dim * sizeof(float2), &dc))) return rc;
    cplx* wide = (cplx*)c->arena.ptr;                       // [dpsi rows | psi]
    c64_widen_kernel<<<c64_grid(c, num_params * dim), 256, 0, c->stream>>>(wide, (const float2*)dc, num_params * dim);
    c64_widen_kernel<<<c64_grid(c, dim), 256, 0, c->stream>>>(wide + num_params * dim, (const float2*)dp, dim);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return cuda_fail(e, "widen kernel");
    return qgt_b200_gram(c, (const double*)(wide + num_params * dim), (const double*)wide, dim, num_params, metric, berry, q_full);
}

}  // extern "C"

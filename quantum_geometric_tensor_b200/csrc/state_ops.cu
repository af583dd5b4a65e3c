// state_ops.cu — device reductions and element-wise passes on a statevector (SURVEY.md section 8 f4 and the float
// boundary of the ComplexFloat APIs): measurement probabilities, collapse, Pauli-Z parity expectation, inner products,
// scaling, inverse-CDF sampling, complex-float <-> complex-double conversion.
//
// Replaces the serial loops of sim_measure_qubit / sim_get_measurement_counts / sim_get_expectation_value
// (reference src/quantum_geometric/hardware/quantum_simulator.c:563-675, 705-729) and measure_qubit_cpu
// (hardware/quantum_simulator_cpu.c:328).  All HBM-bound single passes: 128-bit loads, grid = a multiple of the SM
// count, fixed-order two-stage sums (per-CTA partials, then one CTA) so results are reproducible.
#include <cuda_runtime.h>

#include <algorithm>
#include <cmath>
#include <cstring>
#include <vector>

#include "ctx.hpp"

namespace qgt {

__device__ __forceinline__ double so_warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// block-level sum of NV values per thread -> partial[blockIdx][NV]
template <int NV>
__device__ __forceinline__ void so_block_store(double (&v)[NV], double* partial) {
    __shared__ double ws[8][NV];
#pragma unroll
    for (int j = 0; j < NV; ++j) v[j] = so_warp_sum(v[j]);
    if ((threadIdx.x & 31) == 0)
#pragma unroll
        for (int j = 0; j < NV; ++j) ws[threadIdx.x >> 5][j] = v[j];
    __syncthreads();
    if (threadIdx.x < NV) {
        double s = 0.0;
        for (int w = 0; w < (int)(blockDim.x >> 5); ++w) s += ws[w][threadIdx.x];      // fixed order
        partial[(size_t)blockIdx.x * NV + threadIdx.x] = s;
    }
}

__global__ void so_final_sum_kernel(const double* partial, int nblocks, int nv, double* out) {
    const int j = threadIdx.x;
    if (j >= nv) return;
    double s = 0.0;
    for (int b = 0; b < nblocks; ++b) s += partial[(size_t)b * nv + j];                // fixed order
    out[j] = s;
}

// out[0] = sum |a_i|^2 over indices with (g & mask) == want, out[1] = sum over all indices; g = goff + i
__global__ void __launch_bounds__(256) so_masked_prob_kernel(const cplx* src, uint64_t D, uint64_t goff, uint64_t mask, uint64_t want, double* partial) {
    double v[2] = {0.0, 0.0};
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < D; i += stride) {
        const cplx z = src[i];
        const double p = z.x * z.x + z.y * z.y;
        v[1] += p;
        if (((goff + i) & mask) == want) v[0] += p;
    }
    so_block_store<2>(v, partial);
}

// out[0] = sum (-1)^popcount(g & zmask) |a_i|^2
__global__ void __launch_bounds__(256) so_parity_kernel(const cplx* src, uint64_t D, uint64_t goff, uint64_t zmask, double* partial) {
    double v[1] = {0.0};
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < D; i += stride) {
        const cplx z = src[i];
        const double p = z.x * z.x + z.y * z.y;
        v[0] += (__popcll((goff + i) & zmask) & 1) ? -p : p;
    }
    so_block_store<1>(v, partial);
}

// <a|b> = sum conj(a_i) b_i
__global__ void __launch_bounds__(256) so_inner_kernel(const cplx* a, const cplx* b, uint64_t D, double* partial) {
    double v[2] = {0.0, 0.0};
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < D; i += stride) {
        const cplx x = a[i], y = b[i];
        v[0] += x.x * y.x + x.y * y.y;
        v[1] += x.x * y.y - x.y * y.x;
    }
    so_block_store<2>(v, partial);
}

// a_i <- (keep ? a_i * s : 0) with keep = ((g & mask) == want); mask = 0 scales everything
__global__ void __launch_bounds__(256) so_scale_kernel(cplx* dst, uint64_t D, uint64_t goff, uint64_t mask, uint64_t want, double sr, double si) {
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < D; i += stride) {
        cplx z = dst[i];
        if (((goff + i) & mask) == want) { const cplx q = z; z.x = q.x * sr - q.y * si; z.y = q.x * si + q.y * sr; }
        else { z.x = 0.0; z.y = 0.0; }
        dst[i] = z;
    }
}

// dst_i += a * src_i
__global__ void __launch_bounds__(256) so_axpy_kernel(cplx* dst, const cplx* src, uint64_t D, double ar, double ai) {
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < D; i += stride) {
        const cplx x = src[i];
        cplx z = dst[i];
        z.x += ar * x.x - ai * x.y;
        z.y += ar * x.y + ai * x.x;
        dst[i] = z;
    }
}

__global__ void __launch_bounds__(256) so_widen_kernel(cplx* dst, const float2* src, uint64_t D) {
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < D; i += stride) {
        const float2 f = src[i];
        cplx z; z.x = (double)f.x; z.y = (double)f.y;
        dst[i] = z;
    }
}

__global__ void __launch_bounds__(256) so_narrow_kernel(float2* dst, const cplx* src, uint64_t D) {
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < D; i += stride) {
        const cplx z = src[i];
        dst[i] = make_float2((float)z.x, (float)z.y);
    }
}

// sampling, level 1: probability mass of every chunk of 2^CH amplitudes (one CTA per chunk, fixed-order sum)
constexpr int SO_CHUNK_LOG2 = 14;
__global__ void __launch_bounds__(256) so_chunk_mass_kernel(const cplx* src, uint64_t D, double* mass) {
    const uint64_t c0 = (uint64_t)blockIdx.x << SO_CHUNK_LOG2;
    const uint64_t c1 = c0 + ((uint64_t)1 << SO_CHUNK_LOG2) < D ? c0 + ((uint64_t)1 << SO_CHUNK_LOG2) : D;
    double v[1] = {0.0};
    for (uint64_t i = c0 + threadIdx.x; i < c1; i += blockDim.x) { const cplx z = src[i]; v[0] += z.x * z.x + z.y * z.y; }
    so_block_store<1>(v, mass);
}

// sampling, level 2: shot s falls into chunk[s] with residual r[s] (0 <= r < mass of the chunk): the first index whose
// running sum inside the chunk exceeds r.  One CTA per shot; threads own consecutive segments, a scan over the 256
// segment sums finds the segment, its owner walks it.
__global__ void __launch_bounds__(256) so_sample_kernel(const cplx* src, uint64_t D, const uint32_t* chunk, const double* resid, uint64_t* out) {
    __shared__ double seg[256];
    __shared__ int which;
    const uint64_t c0 = (uint64_t)chunk[blockIdx.x] << SO_CHUNK_LOG2;
    const uint64_t c1 = c0 + ((uint64_t)1 << SO_CHUNK_LOG2) < D ? c0 + ((uint64_t)1 << SO_CHUNK_LOG2) : D;
    const uint64_t per = ((c1 - c0) + 255) / 256;
    const uint64_t s0 = c0 + per * threadIdx.x < c1 ? c0 + per * threadIdx.x : c1;
    const uint64_t s1 = s0 + per < c1 ? s0 + per : c1;
    double m = 0.0;
    for (uint64_t i = s0; i < s1; ++i) { const cplx z = src[i]; m += z.x * z.x + z.y * z.y; }
    seg[threadIdx.x] = m;
    if (threadIdx.x == 0) which = 255;
    __syncthreads();
    const double r = resid[blockIdx.x];
    if (threadIdx.x == 0) {
        double run = 0.0;
        int w = 255;
        for (int t = 0; t < 256; ++t) { if (r < run + seg[t]) { w = t; break; } run += seg[t]; }
        which = w;
        seg[0] = run;                                  // mass before the chosen segment
    }
    __syncthreads();
    if ((int)threadIdx.x == which) {
        double run = seg[0];
        uint64_t pick = s1 > s0 ? s1 - 1 : (c1 > c0 ? c1 - 1 : c0);
        for (uint64_t i = s0; i < s1; ++i) {
            const cplx z = src[i];
            run += z.x * z.x + z.y * z.y;
            if (r < run) { pick = i; break; }
        }
        out[blockIdx.x] = pick;
    }
}

// per-CTA maximum of |a_i|^2 (first index wins ties): partial[b] = (probability, index)
__global__ void __launch_bounds__(256) so_argmax_kernel(const cplx* src, uint64_t D, double* pmax, unsigned long long* imax) {
    double best = -1.0;
    unsigned long long bi = 0;
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < D; i += stride) {
        const cplx z = src[i];
        const double p = z.x * z.x + z.y * z.y;
        if (p > best) { best = p; bi = i; }
    }
    __shared__ double sp[256];
    __shared__ unsigned long long si[256];
    sp[threadIdx.x] = best; si[threadIdx.x] = bi;
    __syncthreads();
    for (int o = 128; o > 0; o >>= 1) {
        if ((int)threadIdx.x < o) {
            const double q = sp[threadIdx.x + o];
            const unsigned long long qi = si[threadIdx.x + o];
            if (q > sp[threadIdx.x] || (q == sp[threadIdx.x] && qi < si[threadIdx.x])) { sp[threadIdx.x] = q; si[threadIdx.x] = qi; }
        }
        __syncthreads();
    }
    if (threadIdx.x == 0) { pmax[blockIdx.x] = sp[0]; imax[blockIdx.x] = si[0]; }
}

static unsigned so_grid(const qgt_b200_ctx* c, uint64_t D) {
    const uint64_t want = (D + 255) / 256;
    const uint64_t cap = (uint64_t)c->num_sms * 8;
    return (unsigned)std::max<uint64_t>(1, std::min(want, cap));
}

// runs a partial-sum kernel launch (already issued into c->scratch as partials) to its final NV doubles on the host,
// summed over ranks
static int so_finish(qgt_b200_ctx* c, int nblocks, int nv, double* host_out) {
    double* partial = (double*)c->scratch.ptr;
    double* out = partial + (size_t)nblocks * nv;
    so_final_sum_kernel<<<1, 32, 0, c->stream>>>(partial, nblocks, nv, out);
    cudaError_t e = cudaGetLastError();
    if (e == cudaSuccess) e = cudaMemcpyAsync(host_out, out, nv * sizeof(double), cudaMemcpyDeviceToHost, c->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(c->stream);
    if (e != cudaSuccess) return cuda_fail(e, "state reduction");
    return dist_allreduce_host(c, host_out, nv);
}

}  // namespace qgt

using namespace qgt;

extern "C" {

int qgt_b200_state_probability(const qgt_b200_state* s, uint64_t mask, uint64_t want, double* prob, double* norm2) {
    if (!s || !prob) return fail(QGT_B200_ERR_INVALID_ARG, "state/prob is NULL");
    qgt_b200_ctx* c = s->ctx;
    cudaSetDevice(c->device);
    const unsigned grid = so_grid(c, s->D);
    int rc = c->scratch.reserve(((size_t)grid * 2 + 8) * sizeof(double));
    if (rc) return rc;
    so_masked_prob_kernel<<<grid, 256, 0, c->stream>>>(s->d, s->D, (uint64_t)c->rank * s->D, mask, want, (double*)c->scratch.ptr);
    double h[2];
    if ((rc = so_finish(c, (int)grid, 2, h))) return rc;
    *prob = h[0];
    if (norm2) *norm2 = h[1];
    return QGT_B200_OK;
}

int qgt_b200_state_expectation_z(const qgt_b200_state* s, uint64_t zmask, double* out) {
    if (!s || !out) return fail(QGT_B200_ERR_INVALID_ARG, "state/out is NULL");
    qgt_b200_ctx* c = s->ctx;
    cudaSetDevice(c->device);
    const unsigned grid = so_grid(c, s->D);
    int rc = c->scratch.reserve(((size_t)grid + 8) * sizeof(double));
    if (rc) return rc;
    so_parity_kernel<<<grid, 256, 0, c->stream>>>(s->d, s->D, (uint64_t)c->rank * s->D, zmask, (double*)c->scratch.ptr);
    return so_finish(c, (int)grid, 1, out);
}

int qgt_b200_state_inner_product(const qgt_b200_state* a, const qgt_b200_state* b, double out[2]) {
    if (!a || !b || !out) return fail(QGT_B200_ERR_INVALID_ARG, "state/out is NULL");
    if (a->ctx != b->ctx || a->n != b->n) return fail(QGT_B200_ERR_DIMENSION, "states differ in context or size");
    qgt_b200_ctx* c = a->ctx;
    cudaSetDevice(c->device);
    const unsigned grid = so_grid(c, a->D);
    int rc = c->scratch.reserve(((size_t)grid * 2 + 8) * sizeof(double));
    if (rc) return rc;
    so_inner_kernel<<<grid, 256, 0, c->stream>>>(a->d, b->d, a->D, (double*)c->scratch.ptr);
    return so_finish(c, (int)grid, 2, out);
}

int qgt_b200_state_scale(qgt_b200_state* s, double re, double im) {
    if (!s) return fail(QGT_B200_ERR_INVALID_ARG, "state is NULL");
    qgt_b200_ctx* c = s->ctx;
    cudaSetDevice(c->device);
    so_scale_kernel<<<so_grid(c, s->D), 256, 0, c->stream>>>(s->d, s->D, 0, 0, 0, re, im);
    cudaError_t e = cudaGetLastError();
    if (e == cudaSuccess) e = cudaStreamSynchronize(c->stream);
    return e == cudaSuccess ? QGT_B200_OK : cuda_fail(e, "scale kernel");
}

int qgt_b200_state_argmax(const qgt_b200_state* s, uint64_t* index, double* probability) {
    if (!s || !index) return fail(QGT_B200_ERR_INVALID_ARG, "state/index is NULL");
    qgt_b200_ctx* c = s->ctx;
    cudaSetDevice(c->device);
    const unsigned grid = so_grid(c, s->D);
    int rc = c->scratch.reserve((size_t)grid * 16 + 64);
    if (rc) return rc;
    double* d_p = (double*)c->scratch.ptr;
    unsigned long long* d_i = (unsigned long long*)(d_p + grid);
    so_argmax_kernel<<<grid, 256, 0, c->stream>>>(s->d, s->D, d_p, d_i);
    std::vector<double> hp(grid);
    std::vector<unsigned long long> hi(grid);
    cudaError_t e = cudaGetLastError();
    if (e == cudaSuccess) e = cudaMemcpyAsync(hp.data(), d_p, grid * sizeof(double), cudaMemcpyDeviceToHost, c->stream);
    if (e == cudaSuccess) e = cudaMemcpyAsync(hi.data(), d_i, grid * sizeof(unsigned long long), cudaMemcpyDeviceToHost, c->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(c->stream);
    if (e != cudaSuccess) return cuda_fail(e, "argmax kernel");
    size_t b = 0;
    for (size_t k = 1; k < grid; k++) if (hp[k] > hp[b] || (hp[k] == hp[b] && hi[k] < hi[b])) b = k;
    double best_p = hp[b];
    uint64_t best_i = ((uint64_t)c->rank * s->D) | hi[b];
    if (c->world > 1) {
        // sharded state: every rank contributes (probability, global index) in its own slot of a summed table, so all ranks
        // see all candidates and pick the same one (indices below 2^53 are exact as doubles)
        std::vector<double> tab((size_t)c->world * 2, 0.0);
        tab[(size_t)c->rank * 2] = best_p;
        tab[(size_t)c->rank * 2 + 1] = (double)best_i;
        if ((rc = dist_allreduce_host(c, tab.data(), c->world * 2))) return rc;
        int br = 0;
        for (int r = 1; r < c->world; r++) if (tab[(size_t)r * 2] > tab[(size_t)br * 2]) br = r;     // ties: the lower rank = the lower index
        best_p = tab[(size_t)br * 2];
        best_i = (uint64_t)tab[(size_t)br * 2 + 1];
    }
    *index = best_i;
    if (probability) *probability = best_p;
    return QGT_B200_OK;
}

int qgt_b200_state_cost_expectation(const qgt_b200_state* s, const qgt_b200_circuit* observable, double* out) {
    if (!s || !observable || !out) return fail(QGT_B200_ERR_INVALID_ARG, "state/observable/out is NULL");
    if (observable->num_qubits != s->n) return fail(QGT_B200_ERR_DIMENSION, "observable has a different qubit count");
    qgt_b200_ctx* c = s->ctx;
    cudaSetDevice(c->device);
    int rc = upload_cost_table(c, *observable);
    if (!rc) rc = c->scratch.reserve(256);
    if (rc) return rc;
    double h[2] = {0.0, 0.0};
    cudaError_t e = launch_cost_dot(s->d, s->d, s->D, c->cost, (uint64_t)c->rank * s->D, (double*)c->scratch.ptr, c->stream);
    if (e == cudaSuccess) e = cudaMemcpyAsync(h, c->scratch.ptr, sizeof h, cudaMemcpyDeviceToHost, c->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(c->stream);
    if (e != cudaSuccess) return cuda_fail(e, "cost expectation");
    if ((rc = dist_allreduce_host(c, h, 1))) return rc;
    *out = h[0];
    return QGT_B200_OK;
}

int qgt_b200_state_axpy(qgt_b200_state* dst, double re, double im, const qgt_b200_state* src) {
    if (!dst || !src) return fail(QGT_B200_ERR_INVALID_ARG, "state is NULL");
    if (dst->ctx != src->ctx || dst->n != src->n) return fail(QGT_B200_ERR_DIMENSION, "states differ in context or size");
    qgt_b200_ctx* c = dst->ctx;
    cudaSetDevice(c->device);
    so_axpy_kernel<<<so_grid(c, dst->D), 256, 0, c->stream>>>(dst->d, src->d, dst->D, re, im);
    cudaError_t e = cudaGetLastError();
    if (e == cudaSuccess) e = cudaStreamSynchronize(c->stream);
    return e == cudaSuccess ? QGT_B200_OK : cuda_fail(e, "axpy kernel");
}

int qgt_b200_state_normalize(qgt_b200_state* s, double* norm_before) {
    double n2 = 0.0;
    int rc = qgt_b200_state_norm2(s, &n2);
    if (rc) return rc;
    if (norm_before) *norm_before = std::sqrt(n2);
    if (!(n2 > 0.0)) return fail(QGT_B200_ERR_INVALID_STATE, "cannot normalise the zero vector");
    return qgt_b200_state_scale(s, 1.0 / std::sqrt(n2), 0.0);
}

int qgt_b200_state_collapse(qgt_b200_state* s, int qubit, int outcome, double* prob) {
    if (!s) return fail(QGT_B200_ERR_INVALID_ARG, "state is NULL");
    if (qubit < 0 || qubit >= s->n) return fail(QGT_B200_ERR_INVALID_ARG, "qubit out of range");
    const uint64_t mask = (uint64_t)1 << qubit, want = outcome ? mask : 0;
    double p = 0.0;
    int rc = qgt_b200_state_probability(s, mask, want, &p, nullptr);
    if (rc) return rc;
    if (prob) *prob = p;
    qgt_b200_ctx* c = s->ctx;
    // amplitudes of the other outcome become 0; the rest is rescaled when anything is left (quantum_simulator.c:589-603)
    const double sc = p > 0.0 ? 1.0 / std::sqrt(p) : 1.0;
    so_scale_kernel<<<so_grid(c, s->D), 256, 0, c->stream>>>(s->d, s->D, (uint64_t)c->rank * s->D, mask, want, sc, 0.0);
    cudaError_t e = cudaGetLastError();
    if (e == cudaSuccess) e = cudaStreamSynchronize(c->stream);
    return e == cudaSuccess ? QGT_B200_OK : cuda_fail(e, "collapse kernel");
}

int qgt_b200_state_measure(qgt_b200_state* s, int qubit, double uniform, double readout_error, int* outcome, double* prob_one) {
    if (!s || !outcome) return fail(QGT_B200_ERR_INVALID_ARG, "state/outcome is NULL");
    if (qubit < 0 || qubit >= s->n) return fail(QGT_B200_ERR_INVALID_ARG, "qubit out of range");
    double p1 = 0.0;
    int rc = qgt_b200_state_probability(s, (uint64_t)1 << qubit, (uint64_t)1 << qubit, &p1, nullptr);
    if (rc) return rc;
    double p = p1;
    if (readout_error > 0.0) p = p * (1.0 - readout_error) + (1.0 - p) * readout_error;      // quantum_simulator.c:577-582
    *outcome = uniform < p ? 1 : 0;
    if (prob_one) *prob_one = p1;
    return qgt_b200_state_collapse(s, qubit, *outcome, nullptr);
}

int qgt_b200_state_sample(const qgt_b200_state* s, const double* uniforms, size_t shots, uint64_t* indices) {
    if (!s || !uniforms || !indices) return fail(QGT_B200_ERR_INVALID_ARG, "state/uniforms/indices is NULL");
    qgt_b200_ctx* c = s->ctx;
    if (shots == 0) return QGT_B200_OK;
    cudaSetDevice(c->device);
    const uint64_t D = s->D;
    const size_t nch = (size_t)((D + ((uint64_t)1 << SO_CHUNK_LOG2) - 1) >> SO_CHUNK_LOG2);
    const size_t bytes = nch * sizeof(double) + shots * (sizeof(uint32_t) + sizeof(double) + sizeof(uint64_t)) + 64;
    int rc = c->scratch.reserve(bytes);
    if (rc) return rc;
    double* d_mass = (double*)c->scratch.ptr;
    so_chunk_mass_kernel<<<(unsigned)nch, 256, 0, c->stream>>>(s->d, D, d_mass);
    std::vector<double> mass(nch);
    cudaError_t e = cudaGetLastError();
    if (e == cudaSuccess) e = cudaMemcpyAsync(mass.data(), d_mass, nch * sizeof(double), cudaMemcpyDeviceToHost, c->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(c->stream);
    if (e != cudaSuccess) return cuda_fail(e, "chunk masses");
    // inverse CDF over the chunk masses on the host (2^(n-14) entries), the walk inside a chunk on the device
    std::vector<double> cum(nch + 1, 0.0);
    for (size_t i = 0; i < nch; i++) cum[i + 1] = cum[i] + mass[i];
    // sharded state (collective call, the same uniforms on every rank): the ranks' masses are laid end to end in rank order
    // = global index order; a shot belongs to the rank whose interval holds it and is walked there with the interval's start
    // taken off; every rank computes the same interval ends from the same summed table, so exactly one rank claims a shot
    double base = 0.0, top = cum[nch];
    if (c->world > 1) {
        std::vector<double> rm((size_t)c->world, 0.0);
        rm[(size_t)c->rank] = cum[nch];
        if ((rc = dist_allreduce_host(c, rm.data(), c->world))) return rc;
        for (int r = 0; r < c->rank; r++) base += rm[(size_t)r];
        top = base + rm[(size_t)c->rank];
    }
    std::vector<size_t> own;
    std::vector<uint32_t> chunk;
    std::vector<double> resid;
    for (size_t k = 0; k < shots; k++) {
        const double u = uniforms[k];
        if (c->world > 1 && !((c->rank == 0 || u >= base) && (c->rank == c->world - 1 || u < top))) continue;
        const double r = u - base;
        size_t ci = (size_t)(std::upper_bound(cum.begin() + 1, cum.end(), r) - (cum.begin() + 1));
        if (ci >= nch) ci = nch - 1;
        own.push_back(k);
        chunk.push_back((uint32_t)ci);
        resid.push_back(r - cum[ci]);
    }
    const size_t nown = own.size();
    std::vector<uint64_t> local(nown);
    if (nown > 0) {
        double* d_resid = d_mass + nch;
        uint64_t* d_out = (uint64_t*)(d_resid + shots);
        uint32_t* d_chunk = (uint32_t*)(d_out + shots);
        e = cudaMemcpyAsync(d_resid, resid.data(), nown * sizeof(double), cudaMemcpyHostToDevice, c->stream);
        if (e == cudaSuccess) e = cudaMemcpyAsync(d_chunk, chunk.data(), nown * sizeof(uint32_t), cudaMemcpyHostToDevice, c->stream);
        if (e != cudaSuccess) return cuda_fail(e, "sample staging");
        so_sample_kernel<<<(unsigned)nown, 256, 0, c->stream>>>(s->d, D, d_chunk, d_resid, d_out);
        e = cudaGetLastError();
        if (e == cudaSuccess) e = cudaMemcpyAsync(local.data(), d_out, nown * sizeof(uint64_t), cudaMemcpyDeviceToHost, c->stream);
        if (e == cudaSuccess) e = cudaStreamSynchronize(c->stream);
        if (e != cudaSuccess) return cuda_fail(e, "sample kernel");
    }
    if (c->world == 1) {
        for (size_t j = 0; j < nown; j++) indices[own[j]] = local[j];
        return QGT_B200_OK;
    }
    // every rank receives every shot: owned entries as global indices (exact as doubles below 2^53), zeros elsewhere, summed
    std::vector<double> all(shots, 0.0);
    for (size_t j = 0; j < nown; j++) all[own[j]] = (double)(((uint64_t)c->rank * D) | local[j]);
    for (size_t off = 0; off < shots; off += (size_t)1 << 20) {
        const size_t cnt = std::min(shots - off, (size_t)1 << 20);
        if ((rc = dist_allreduce_host(c, all.data() + off, (int)cnt))) return rc;
    }
    for (size_t k = 0; k < shots; k++) indices[k] = (uint64_t)all[k];
    return QGT_B200_OK;
}

static bool so_is_device(const void* p) {
    cudaPointerAttributes a;
    if (cudaPointerGetAttributes(&a, p) != cudaSuccess) { cudaGetLastError(); return false; }
    return a.type == cudaMemoryTypeDevice || a.type == cudaMemoryTypeManaged;
}

int qgt_b200_state_upload_c64(qgt_b200_state* s, const float* src) {
    if (!s || !src) return fail(QGT_B200_ERR_INVALID_ARG, "state/src is NULL");
    qgt_b200_ctx* c = s->ctx;
    cudaSetDevice(c->device);
    const float2* d_src = (const float2*)src;
    if (!so_is_device(src)) {
        int rc = c->scratch.reserve(s->D * sizeof(float2));
        if (rc) return rc;
        cudaError_t e = cudaMemcpyAsync(c->scratch.ptr, src, s->D * sizeof(float2), cudaMemcpyHostToDevice, c->stream);
        if (e != cudaSuccess) return cuda_fail(e, "c64 upload");
        d_src = (const float2*)c->scratch.ptr;
    }
    so_widen_kernel<<<so_grid(c, s->D), 256, 0, c->stream>>>(s->d, d_src, s->D);
    cudaError_t e = cudaGetLastError();
    if (e == cudaSuccess) e = cudaStreamSynchronize(c->stream);
    return e == cudaSuccess ? QGT_B200_OK : cuda_fail(e, "widen kernel");
}

int qgt_b200_state_download_c64(const qgt_b200_state* s, float* dst) {
    if (!s || !dst) return fail(QGT_B200_ERR_INVALID_ARG, "state/dst is NULL");
    qgt_b200_ctx* c = s->ctx;
    cudaSetDevice(c->device);
    const bool dev = so_is_device(dst);
    float2* d_dst = (float2*)dst;
    if (!dev) {
        int rc = c->scratch.reserve(s->D * sizeof(float2));
        if (rc) return rc;
        d_dst = (float2*)c->scratch.ptr;
    }
    so_narrow_kernel<<<so_grid(c, s->D), 256, 0, c->stream>>>(d_dst, s->d, s->D);
    cudaError_t e = cudaGetLastError();
    if (e == cudaSuccess && !dev) e = cudaMemcpyAsync(dst, d_dst, s->D * sizeof(float2), cudaMemcpyDeviceToHost, c->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(c->stream);
    return e == cudaSuccess ? QGT_B200_OK : cuda_fail(e, "narrow kernel");
}

}  // extern "C"

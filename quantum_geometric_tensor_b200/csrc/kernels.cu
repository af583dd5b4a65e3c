// kernels.cu — hand-written sm_100a kernels of the QGT hot path.
//
//   qgt_sweep_kernel   fused gate sweep (HBM-bound): stages a 2^K-amplitude tile in shared memory with
//                      128-bit loads, applies every op of a run with 2^R amplitudes per thread in
//                      registers, writes the tile back.  Batched over columns; a column may replace one
//                      op by its derivative (generator folded in) and may accumulate into its destination.
//                      Replaces apply_single_gate / apply_controlled_gate
//                      (reference hardware/quantum_simulator.c:147-185) and the kernels of
//                      src/cuda/quantum_geometric_cuda.cu:196-217.
//   qgt_gram_kernel    C = A^H B for column blocks, complex double on the FP64 tensor pipe
//                      (mma.sync m8n8k4 f64 -> SASS DMMA.8x8x4; tcgen05 has no f64 kind), split-K with a
//                      deterministic second-stage reduction.  Replaces the dot loops of
//                      compute_quantum_geometric_tensor (core/quantum_geometric_tensor_network.c:1127-1175)
//                      and diffgeo_compute_fubini_study (distributed/differential_geometry.c:2819-2862).
//   qgt_finalize       Q = C - v v^H  ->  metric (Re), Berry curvature (Im), full Q.
#include "kernels.cuh"

#include <cstdio>

namespace qgt {

// ------------------------------------------------------------------------------------------------
// gate sweep
// ------------------------------------------------------------------------------------------------
template <int R>
__global__ void __launch_bounds__(256, 2) qgt_sweep_kernel(SweepLaunch a) {
    extern __shared__ __align__(16) unsigned char qgt_smem_raw[];
    cplx* tile = reinterpret_cast<cplx*>(qgt_smem_raw);
    __shared__ QgtDevRun run;
    if (threadIdx.x < sizeof(QgtDevRun) / 4) {
        reinterpret_cast<uint32_t*>(&run)[threadIdx.x] =
            reinterpret_cast<const uint32_t*>(&a.runs[a.run_idx])[threadIdx.x];
    }
    __syncthreads();
    const QgtDevOp* ops = a.ops + run.ops_off;
    const QgtDevSubPass* subs = a.subs + run.sub_off;
    const int tid = threadIdx.x;
    const int T = blockDim.x;
    const QgtIoMap<R> io = qgt_make_iomap<R>(run, tid);
    const uint64_t total = a.ntiles * (uint64_t)a.nitems;
    for (uint64_t w = blockIdx.x; w < total; w += gridDim.x) {
        const int item = (int)(w % (uint64_t)a.nitems);
        const uint64_t tau = w / (uint64_t)a.nitems;
        const QgtSweepItem& it = a.items[item];
        const uint64_t tilebase = qgt_tile_base(run, tau);
        qgt_phase_load<R>(io, tile, reinterpret_cast<const cplx*>(it.src), tilebase, tid, T);
        __syncthreads();
        const int ovr = it.ovr_op;
        for (int s = 0; s < run.nsub; ++s) {
            qgt_phase_subpass<R>(run, subs[s], ops, ovr, it.ovr, tile, tilebase, tid, a.ct);
            __syncthreads();
        }
        qgt_phase_store<R>(io, tile, reinterpret_cast<cplx*>(it.dst), tilebase, tid, T, it.accumulate != 0);
        __syncthreads();
    }
}

// small states (fewer than 2^R * 32 amplitudes per tile) use the same code with one thread per 2^R
// amplitudes; sizeof(QgtDevRun)/4 threads may not exist there, so the run header is copied in a loop.
template <int R>
__global__ void __launch_bounds__(256) qgt_sweep_small_kernel(SweepLaunch a) {
    extern __shared__ __align__(16) unsigned char qgt_smem_raw[];
    cplx* tile = reinterpret_cast<cplx*>(qgt_smem_raw);
    __shared__ QgtDevRun run;
    for (int i = threadIdx.x; i < (int)(sizeof(QgtDevRun) / 4); i += blockDim.x)
        reinterpret_cast<uint32_t*>(&run)[i] = reinterpret_cast<const uint32_t*>(&a.runs[a.run_idx])[i];
    __syncthreads();
    const QgtDevOp* ops = a.ops + run.ops_off;
    const QgtDevSubPass* subs = a.subs + run.sub_off;
    const int tid = threadIdx.x;
    const int T = blockDim.x;
    const QgtIoMap<R> io = qgt_make_iomap<R>(run, tid);
    const uint64_t total = a.ntiles * (uint64_t)a.nitems;
    for (uint64_t w = blockIdx.x; w < total; w += gridDim.x) {
        const int item = (int)(w % (uint64_t)a.nitems);
        const uint64_t tau = w / (uint64_t)a.nitems;
        const QgtSweepItem& it = a.items[item];
        const uint64_t tilebase = qgt_tile_base(run, tau);
        qgt_phase_load<R>(io, tile, reinterpret_cast<const cplx*>(it.src), tilebase, tid, T);
        __syncthreads();
        for (int s = 0; s < run.nsub; ++s) {
            qgt_phase_subpass<R>(run, subs[s], ops, it.ovr_op, it.ovr, tile, tilebase, tid, a.ct);
            __syncthreads();
        }
        qgt_phase_store<R>(io, tile, reinterpret_cast<cplx*>(it.dst), tilebase, tid, T, it.accumulate != 0);
        __syncthreads();
    }
}

template <int R>
static cudaError_t launch_sweep_r(const SweepLaunch& a, int K, int num_sms, cudaStream_t st) {
    const int T = 1 << (K - R);
    const size_t smem = sizeof(cplx) << K;
    const uint64_t total = a.ntiles * (uint64_t)a.nitems;
    if (total == 0) return cudaSuccess;
    if (T >= 32) {
        static bool attr_set = false;
        if (!attr_set) {
            cudaError_t e = cudaFuncSetAttribute(qgt_sweep_kernel<R>, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
            if (e != cudaSuccess) return e;
            attr_set = true;
        }
        const uint64_t cap = (uint64_t)num_sms * 2 * 4;
        const unsigned grid = (unsigned)(total < cap ? total : cap);
        qgt_sweep_kernel<R><<<grid, T, smem, st>>>(a);
    } else {
        const uint64_t cap = (uint64_t)num_sms * 8;
        const unsigned grid = (unsigned)(total < cap ? total : cap);
        qgt_sweep_small_kernel<R><<<grid, T, smem, st>>>(a);
    }
    return cudaGetLastError();
}

cudaError_t launch_sweep(const SweepLaunch& a, int K, int R, int num_sms, cudaStream_t st) {
    if (K - R > 8 || K > QGT_MAX_TILE_QUBITS) return cudaErrorInvalidValue;
    switch (R) {
    case 1: return launch_sweep_r<1>(a, K, num_sms, st);
    case 2: return launch_sweep_r<2>(a, K, num_sms, st);
    case 3: return launch_sweep_r<3>(a, K, num_sms, st);
    default: return cudaErrorInvalidValue;
    }
}

// ------------------------------------------------------------------------------------------------
// Gram  C = A^H B  on the FP64 tensor pipe
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void dmma884(double& c0, double& c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                 : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

// CTA tile: MT = WM*BM*8 rows (columns of A) x NT = WN*BN*8 cols (columns of B); WM*WN warps, each
// warp owns BM x BN blocks of 8x8.  K (the 2^n amplitude axis) is consumed in chunks of KC staged in
// shared memory as split re / im planes with row stride KC+4 doubles (conflict-free fragment loads).
template <int WM, int WN, int BM, int BN>
__global__ void __launch_bounds__(WM * WN * 32) qgt_gram_kernel(GramLaunch g) {
    constexpr int MT = WM * BM * 8, NT = WN * BN * 8, KC = 16, S = KC + 4, NTHR = WM * WN * 32;
    constexpr int ELEMS = (MT + NT) * KC;
    static_assert(ELEMS % NTHR == 0, "tile/threads mismatch");
    constexpr int PER = ELEMS / NTHR;
    __shared__ double sAr[MT * S], sAi[MT * S], sBr[NT * S], sBi[NT * S];

    const int tiles = g.mtiles * g.ntiles;
    const int ks = blockIdx.x / tiles;
    const int mt = (blockIdx.x % tiles) / g.ntiles;
    const int nt = (blockIdx.x % tiles) % g.ntiles;
    if (g.symmetric && (nt + 1) * NT <= mt * MT) return;   // whole tile below the diagonal: mirrored later

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int wm = warp / WN, wn = warp % WN;
    uint64_t per = (g.D + (uint64_t)g.ksplit - 1) / (uint64_t)g.ksplit;
    per = (per + KC - 1) / KC * KC;
    const uint64_t k0 = (uint64_t)ks * per;
    const uint64_t k1 = (k0 + per < g.D) ? k0 + per : g.D;

    // this thread's share of a chunk: element e -> (column e / KC, offset e % KC)
    const cplx* colptr[PER];
    int soff[PER];
    bool isA[PER];
#pragma unroll
    for (int j = 0; j < PER; ++j) {
        const int e = tid + j * NTHR;
        const int col = e / KC, k = e % KC;
        if (col < MT) {
            const int gc = mt * MT + col;
            colptr[j] = gc < g.na ? g.a_ptrs[gc] + k : nullptr;
            soff[j] = col * S + k; isA[j] = true;
        } else {
            const int gc = nt * NT + (col - MT);
            colptr[j] = gc < g.nb ? g.b_ptrs[gc] + k : nullptr;
            soff[j] = (col - MT) * S + k; isA[j] = false;
        }
    }
    const int koff = tid % KC;   // == e % KC for every j because NTHR % KC == 0

    double cre[BM][BN][2], cim[BM][BN][2];
#pragma unroll
    for (int i = 0; i < BM; ++i)
#pragma unroll
        for (int j = 0; j < BN; ++j) { cre[i][j][0] = cre[i][j][1] = cim[i][j][0] = cim[i][j][1] = 0.0; }

    // blocks entirely in the padding do no tensor work
    bool rowok[BM], colok[BN];
#pragma unroll
    for (int i = 0; i < BM; ++i) rowok[i] = (mt * MT + (wm * BM + i) * 8) < g.na;
#pragma unroll
    for (int j = 0; j < BN; ++j) colok[j] = (nt * NT + (wn * BN + j) * 8) < g.nb;

    cplx pre[PER];
    auto fetch = [&](uint64_t kb) {
#pragma unroll
        for (int j = 0; j < PER; ++j) {
            cplx z; z.x = 0.0; z.y = 0.0;
            if (colptr[j] != nullptr && kb + (uint64_t)koff < k1) z = colptr[j][kb];
            pre[j] = z;
        }
    };
    if (k0 < k1) fetch(k0);
    for (uint64_t kb = k0; kb < k1; kb += KC) {
#pragma unroll
        for (int j = 0; j < PER; ++j) {
            if (isA[j]) { sAr[soff[j]] = pre[j].x; sAi[soff[j]] = pre[j].y; }
            else        { sBr[soff[j]] = pre[j].x; sBi[soff[j]] = pre[j].y; }
        }
        __syncthreads();
        if (kb + KC < k1) fetch(kb + KC);
        const int fr = lane >> 2, fk = lane & 3;
#pragma unroll
        for (int kk = 0; kk < KC; kk += 4) {
            double ar[BM], ai[BM], br[BN], bi[BN];
#pragma unroll
            for (int i = 0; i < BM; ++i) {
                const int o = ((wm * BM + i) * 8 + fr) * S + kk + fk;
                ar[i] = sAr[o]; ai[i] = sAi[o];
            }
#pragma unroll
            for (int j = 0; j < BN; ++j) {
                const int o = ((wn * BN + j) * 8 + fr) * S + kk + fk;
                br[j] = sBr[o]; bi[j] = sBi[o];
            }
#pragma unroll
            for (int i = 0; i < BM; ++i) {
                if (!rowok[i]) continue;
                const double nai = -ai[i];
#pragma unroll
                for (int j = 0; j < BN; ++j) {
                    if (!colok[j]) continue;
                    // conj(a) * b = (ar*br + ai*bi) + i (ar*bi - ai*br)
                    dmma884(cre[i][j][0], cre[i][j][1], ar[i], br[j]);
                    dmma884(cre[i][j][0], cre[i][j][1], ai[i], bi[j]);
                    dmma884(cim[i][j][0], cim[i][j][1], ar[i], bi[j]);
                    dmma884(cim[i][j][0], cim[i][j][1], nai, br[j]);
                }
            }
        }
        __syncthreads();
    }
    const int Npad = g.ntiles * NT;
    const int Mpad = g.mtiles * MT;
    cplx* out = g.partial + (size_t)ks * Mpad * Npad;
#pragma unroll
    for (int i = 0; i < BM; ++i)
#pragma unroll
        for (int j = 0; j < BN; ++j) {
            const int row = mt * MT + (wm * BM + i) * 8 + (lane >> 2);
            const int col = nt * NT + (wn * BN + j) * 8 + (lane & 3) * 2;
            cplx z0, z1;
            z0.x = cre[i][j][0]; z0.y = cim[i][j][0];
            z1.x = cre[i][j][1]; z1.y = cim[i][j][1];
            out[(size_t)row * Npad + col] = z0;
            out[(size_t)row * Npad + col + 1] = z1;
        }
}

__global__ void qgt_gram_reduce_kernel(const cplx* partial, int ksplit, int Mpad, int Npad, int na, int nb,
                                       const int* a_ids, const int* b_ids, cplx* C, int ldc, int symmetric) {
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= na * nb) return;
    const int i = idx / nb, j = idx % nb;
    if (symmetric && j < i) return;
    double sr = 0.0, si = 0.0;
    for (int ks = 0; ks < ksplit; ++ks) {           // fixed order: deterministic
        const cplx z = partial[((size_t)ks * Mpad + i) * Npad + j];
        sr += z.x; si += z.y;
    }
    const int ai = a_ids[i], bj = b_ids[j];
    cplx z; z.x = sr; z.y = si;
    C[(size_t)ai * ldc + bj] = z;
    if (ai != bj) { z.y = -si; C[(size_t)bj * ldc + ai] = z; }
}

__global__ void qgt_finalize_kernel(const cplx* C, int P, double* metric, double* berry, cplx* q_full) {
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= P * P) return;
    const int mu = idx / P, nu = idx % P;
    const int ld = P + 1;
    const cplx c = C[(size_t)mu * ld + nu];
    const cplx vm = C[(size_t)mu * ld + P];   // <d_mu psi | psi>
    const cplx vn = C[(size_t)nu * ld + P];
    // Q = c - vm * conj(vn)
    const double qr = c.x - (vm.x * vn.x + vm.y * vn.y);
    const double qi = c.y - (vm.y * vn.x - vm.x * vn.y);
    if (metric) metric[idx] = qr;
    if (berry) berry[idx] = qi;
    if (q_full) { q_full[idx].x = qr; q_full[idx].y = qi; }
}

GramShape gram_shape(int na, int nb) {
    GramShape s;
    if (na > 32 || nb > 32) { s.MT = 64; s.NT = 64; }
    else { s.MT = 32; s.NT = 16; }
    return s;
}

cudaError_t launch_gram(const GramLaunch& g, GramShape shp, cudaStream_t st) {
    const unsigned grid = (unsigned)(g.mtiles * g.ntiles * g.ksplit);
    if (grid == 0) return cudaSuccess;
    if (shp.MT == 64) qgt_gram_kernel<2, 4, 4, 2><<<grid, 256, 0, st>>>(g);
    else qgt_gram_kernel<4, 2, 1, 1><<<grid, 256, 0, st>>>(g);
    return cudaGetLastError();
}

cudaError_t launch_gram_reduce(const GramLaunch& g, GramShape shp, const int* a_ids, const int* b_ids,
                               cplx* C, int ldc, cudaStream_t st) {
    const int total = g.na * g.nb;
    if (total == 0) return cudaSuccess;
    qgt_gram_reduce_kernel<<<(total + 255) / 256, 256, 0, st>>>(g.partial, g.ksplit, g.mtiles * shp.MT, g.ntiles * shp.NT,
                                                             g.na, g.nb, a_ids, b_ids, C, ldc, g.symmetric);
    return cudaGetLastError();
}

cudaError_t launch_finalize(const cplx* C, int P, double* metric, double* berry, cplx* q_full, cudaStream_t st) {
    if (P <= 0) return cudaSuccess;
    qgt_finalize_kernel<<<(P * P + 255) / 256, 256, 0, st>>>(C, P, metric, berry, q_full);
    return cudaGetLastError();
}

// ------------------------------------------------------------------------------------------------
// small helpers
// ------------------------------------------------------------------------------------------------
__global__ void qgt_init_kernel(cplx* dst, uint64_t D, int initial_state, double plus_amp, uint64_t goff) {
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < D; i += stride) {
        cplx z; z.y = 0.0;
        z.x = initial_state == 1 ? plus_amp : ((goff + i) == 0 ? 1.0 : 0.0);
        dst[i] = z;
    }
}

cudaError_t launch_init_state(cplx* dst, uint64_t D, int initial_state, double plus_amp, uint64_t goff, cudaStream_t st) {
    const uint64_t want = (D + 255) / 256;
    const unsigned grid = (unsigned)(want < 148 * 16 ? want : 148 * 16);
    qgt_init_kernel<<<grid, 256, 0, st>>>(dst, D, initial_state, plus_amp, goff);
    return cudaGetLastError();
}

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

__global__ void qgt_norm2_kernel(const cplx* src, uint64_t D, double* out) {
    double s = 0.0;
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < D; i += stride) {
        const cplx z = src[i];
        s += z.x * z.x + z.y * z.y;
    }
    s = warp_sum(s);
    __shared__ double ws[8];
    if ((threadIdx.x & 31) == 0) ws[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x < 32) {
        double t = threadIdx.x < (blockDim.x >> 5) ? ws[threadIdx.x] : 0.0;
        t = warp_sum(t);
        if (threadIdx.x == 0) atomicAdd(out, t);
    }
}

cudaError_t launch_norm2(const cplx* src, uint64_t D, double* out, cudaStream_t st) {
    cudaError_t e = cudaMemsetAsync(out, 0, sizeof(double), st);
    if (e != cudaSuccess) return e;
    const uint64_t want = (D + 255) / 256;
    const unsigned grid = (unsigned)(want < 148 * 8 ? want : 148 * 8);
    qgt_norm2_kernel<<<grid, 256, 0, st>>>(src, D, out);
    return cudaGetLastError();
}

__global__ void qgt_axpy_kernel(cplx* dst, const cplx* src, uint64_t D, double ar, double ai) {
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < D; i += stride) {
        const cplx s = src[i];
        cplx d = dst[i];
        d.x += ar * s.x - ai * s.y;
        d.y += ar * s.y + ai * s.x;
        dst[i] = d;
    }
}

cudaError_t launch_axpy(cplx* dst, const cplx* src, uint64_t D, double ar, double ai, cudaStream_t st) {
    const uint64_t want = (D + 255) / 256;
    const unsigned grid = (unsigned)(want < 148 * 16 ? want : 148 * 16);
    qgt_axpy_kernel<<<grid, 256, 0, st>>>(dst, src, D, ar, ai);
    return cudaGetLastError();
}

__global__ void qgt_cost_dot_kernel(const cplx* a, const cplx* b, uint64_t D, QgtCostTable ct, uint64_t goff, double* out2) {
    double sr = 0.0, si = 0.0;
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < D; i += stride) {
        const double e = qgt_cost_energy(ct, goff + i);
        const cplx x = a[i], y = b[i];
        sr += e * (x.x * y.x + x.y * y.y);
        si += e * (x.x * y.y - x.y * y.x);
    }
    sr = warp_sum(sr); si = warp_sum(si);
    __shared__ double wr[8], wi[8];
    if ((threadIdx.x & 31) == 0) { wr[threadIdx.x >> 5] = sr; wi[threadIdx.x >> 5] = si; }
    __syncthreads();
    if (threadIdx.x < 32) {
        double tr = threadIdx.x < (blockDim.x >> 5) ? wr[threadIdx.x] : 0.0;
        double ti = threadIdx.x < (blockDim.x >> 5) ? wi[threadIdx.x] : 0.0;
        tr = warp_sum(tr); ti = warp_sum(ti);
        if (threadIdx.x == 0) { atomicAdd(out2, tr); atomicAdd(out2 + 1, ti); }
    }
}

cudaError_t launch_cost_dot(const cplx* a, const cplx* b, uint64_t D, QgtCostTable ct, uint64_t goff, double* out2, cudaStream_t st) {
    cudaError_t e = cudaMemsetAsync(out2, 0, 2 * sizeof(double), st);
    if (e != cudaSuccess) return e;
    const uint64_t want = (D + 255) / 256;
    const unsigned grid = (unsigned)(want < 148 * 8 ? want : 148 * 8);
    qgt_cost_dot_kernel<<<grid, 256, 0, st>>>(a, b, D, ct, goff, out2);
    return cudaGetLastError();
}

}  // namespace qgt

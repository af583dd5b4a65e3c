// sweep_core.cuh — per-thread phases of the fused gate-sweep kernel.
//
// Written against plain C++ so the very same index arithmetic can be compiled for the host by the
// emulation harness in tests/emul/ (threads become a loop, __syncthreads() a phase boundary).
// The CUDA kernel in kernels.cu strings the phases together.
//
// Replaces the inner loops of apply_single_gate / apply_controlled_gate
// (reference src/quantum_geometric/hardware/quantum_simulator.c:147-185): instead of one pass over
// 2^n amplitudes per gate, a tile of 2^K amplitudes is staged on chip once per RUN and every op of
// the run is applied there; 2^R amplitudes per thread live in registers during a SUB-PASS.
#pragma once
#include <stdint.h>

#include "dev_structs.h"

#if defined(__CUDACC__)
#define QGT_HD __host__ __device__ __forceinline__
#else
#define QGT_HD inline
#include <cmath>
#endif

struct alignas(16) cplx {
    double x, y;
};

// XOR swizzle of the local amplitude index: bit p of the index lands in 16-byte bank-group bit (p mod 3)
QGT_HD uint32_t qgt_swz(uint32_t idx) {
    return idx ^ ((idx >> 3) & 7u) ^ ((idx >> 6) & 7u) ^ ((idx >> 9) & 7u);
}

QGT_HD int qgt_popc64(uint64_t v) {
#if defined(__CUDA_ARCH__)
    return __popcll(v);
#else
    return __builtin_popcountll(v);
#endif
}

// deposit the bits of the tile id into the non-tile qubit positions
QGT_HD uint64_t qgt_tile_base(const QgtDevRun& run, uint64_t tau) {
    uint64_t g = 0;
    const int m = run.n - run.K;
    for (int j = 0; j < m; j++) g |= ((tau >> j) & 1ull) << run.ntq[j];
    return g;
}

// local index (bits 0..nbits-1, starting at local position `first`) -> global offset
QGT_HD uint64_t qgt_local_to_global(const QgtDevRun& run, uint32_t idx) {
    uint64_t g = 0;
    for (int j = 0; j < run.K; j++) g |= (uint64_t)((idx >> j) & 1u) << run.tq[j];
    return g;
}

struct QgtCostTable {
    const QgtDevEdge* edges;
    int num_edges;
    const double* vertex_weights;   // n doubles or null
    int n;
};

// E_z of the cost layer: cut weight + vertex terms (reference algorithms/qaoa.c:258-289)
QGT_HD double qgt_cost_energy(const QgtCostTable& ct, uint64_t g) {
    double e = 0.0;
    for (int k = 0; k < ct.num_edges; k++) {
        const QgtDevEdge ed = ct.edges[k];
        if (((g >> ed.i) ^ (g >> ed.j)) & 1ull) e += ed.w;
    }
    if (ct.vertex_weights)
        for (int q = 0; q < ct.n; q++) e += ct.vertex_weights[q] * (double)(1 - 2 * (int)((g >> q) & 1ull));
    return e;
}

// ---- register-level op application -----------------------------------------------------------------
template <int R, int J>
QGT_HD void qgt_apply_u(cplx (&v)[1 << R], const QgtDevOp& op) {
    const double* m = op.m;
#pragma unroll
    for (int c = 0; c < (1 << R); ++c) {
        if (c & (1 << J)) continue;
        const int c1 = c | (1 << J);
        if (((uint32_t)c & op.creg) != op.creg) {
            if (op.flags & QGT_FLAG_ZERO_CTRL_FAIL) { v[c].x = v[c].y = v[c1].x = v[c1].y = 0.0; }
            continue;
        }
        const cplx a = v[c], b = v[c1];
        if (op.type == QGT_OP_PERM) {
            v[c] = b; v[c1] = a;
        } else if (op.type == QGT_OP_UREAL) {
            v[c].x = m[0] * a.x + m[2] * b.x;   v[c].y = m[0] * a.y + m[2] * b.y;
            v[c1].x = m[4] * a.x + m[6] * b.x;  v[c1].y = m[4] * a.y + m[6] * b.y;
        } else if (op.type == QGT_OP_URX) {      // real diagonal, imaginary off-diagonal
            v[c].x = m[0] * a.x - m[3] * b.y;   v[c].y = m[0] * a.y + m[3] * b.x;
            v[c1].x = m[6] * b.x - m[5] * a.y;  v[c1].y = m[6] * b.y + m[5] * a.x;
        } else {
            v[c].x = m[0] * a.x - m[1] * a.y + m[2] * b.x - m[3] * b.y;
            v[c].y = m[0] * a.y + m[1] * a.x + m[2] * b.y + m[3] * b.x;
            v[c1].x = m[4] * a.x - m[5] * a.y + m[6] * b.x - m[7] * b.y;
            v[c1].y = m[4] * a.y + m[5] * a.x + m[6] * b.y + m[7] * b.x;
        }
    }
}

template <int R>
QGT_HD void qgt_zero_all(cplx (&v)[1 << R]) {
#pragma unroll
    for (int c = 0; c < (1 << R); ++c) v[c].x = v[c].y = 0.0;
}

// apply one op to the 2^R amplitudes of a thread.  gbase = global index of combo 0 (register bits clear),
// regg[r] = global bit of register qubit r.
template <int R>
QGT_HD void qgt_apply_op(cplx (&v)[1 << R], const QgtDevOp& op, uint64_t gbase, const uint64_t* regg,
                         const QgtCostTable& ct) {
    const bool thread_ok = (gbase & op.cmask) == op.cmask;
    if (op.type <= QGT_OP_PERM) {
        if (!thread_ok) {
            if (op.flags & QGT_FLAG_ZERO_CTRL_FAIL) qgt_zero_all<R>(v);
            return;
        }
        switch (op.tbit) {
        case 0: qgt_apply_u<R, 0>(v, op); break;
        case 1: if (R > 1) qgt_apply_u<R, (R > 1 ? 1 : 0)>(v, op); break;
        case 2: if (R > 2) qgt_apply_u<R, (R > 2 ? 2 : 0)>(v, op); break;
        case 3: if (R > 3) qgt_apply_u<R, (R > 3 ? 3 : 0)>(v, op); break;
        default: break;
        }
    } else if (op.type == QGT_OP_DIAG) {
        const int par_t = qgt_popc64(gbase & op.pmask) & 1;
#pragma unroll
        for (int c = 0; c < (1 << R); ++c) {
            if (thread_ok && ((uint32_t)c & op.creg) == op.creg) {
                const int par = par_t ^ (qgt_popc64((uint64_t)((uint32_t)c & op.preg)) & 1);
                const double dr = par ? op.m[2] : op.m[0], di = par ? op.m[3] : op.m[1];
                const cplx a = v[c];
                v[c].x = dr * a.x - di * a.y;
                v[c].y = dr * a.y + di * a.x;
            } else if (op.flags & QGT_FLAG_ZERO_CTRL_FAIL) {
                v[c].x = v[c].y = 0.0;
            }
        }
    } else {   // QGT_OP_COST
#pragma unroll
        for (int c = 0; c < (1 << R); ++c) {
            uint64_t g = gbase;
#pragma unroll
            for (int r = 0; r < R; ++r) if (c & (1 << r)) g |= regg[r];
            const double e = qgt_cost_energy(ct, g);
            double sn, cs;
#if defined(__CUDA_ARCH__)
            sincos(-op.m[0] * e, &sn, &cs);
#else
            sn = std::sin(-op.m[0] * e); cs = std::cos(-op.m[0] * e);
#endif
            cplx a = v[c];
            cplx b; b.x = cs * a.x - sn * a.y; b.y = cs * a.y + sn * a.x;
            if (op.flags & QGT_FLAG_COST_DERIV) {   // times (-i * scale * E)
                const double f = op.m[1] * e;
                a.x = f * b.y; a.y = -f * b.x; b = a;
            }
            v[c] = b;
        }
    }
}

// ---- per-thread phases --------------------------------------------------------------------------------
// Per-thread constants of the load/store phases (independent of the tile): thread `tid` of T moves the
// amplitudes with local index idx = tid + i*T, i < 2^R, i.e. thread bits occupy local positions 0..K-R-1
// and i the top R positions.
template <int R>
struct QgtIoMap {
    uint64_t glo;        // global offset of local index `tid`
    uint64_t ghi[R];     // global bit of local position K-R+r
};

template <int R>
QGT_HD QgtIoMap<R> qgt_make_iomap(const QgtDevRun& run, int tid) {
    QgtIoMap<R> io;
    io.glo = qgt_local_to_global(run, (uint32_t)tid);
#pragma unroll
    for (int r = 0; r < R; ++r) io.ghi[r] = 1ull << run.tq[run.K - R + r];
    return io;
}

template <int R>
QGT_HD uint64_t qgt_io_offset(const QgtIoMap<R>& io, int i) {
    uint64_t g = io.glo;
#pragma unroll
    for (int r = 0; r < R; ++r) if (i & (1 << r)) g |= io.ghi[r];
    return g;
}

// global -> shared
template <int R>
QGT_HD void qgt_phase_load(const QgtIoMap<R>& io, cplx* tile, const cplx* src, uint64_t tilebase, int tid, int T) {
#pragma unroll
    for (int i = 0; i < (1 << R); ++i) {
        const uint32_t idx = (uint32_t)tid + (uint32_t)i * (uint32_t)T;
        tile[qgt_swz(idx)] = src[tilebase | qgt_io_offset<R>(io, i)];
    }
}

template <int R>
QGT_HD void qgt_phase_store(const QgtIoMap<R>& io, const cplx* tile, cplx* dst, uint64_t tilebase, int tid, int T, bool accumulate) {
#pragma unroll
    for (int i = 0; i < (1 << R); ++i) {
        const uint32_t idx = (uint32_t)tid + (uint32_t)i * (uint32_t)T;
        const uint64_t g = tilebase | qgt_io_offset<R>(io, i);
        cplx v = tile[qgt_swz(idx)];
        if (accumulate) { const cplx o = dst[g]; v.x += o.x; v.y += o.y; }
        dst[g] = v;
    }
}

// one sub-pass for one thread: shared -> registers, ops, registers -> shared
template <int R>
QGT_HD void qgt_phase_subpass(const QgtDevRun& run, const QgtDevSubPass& sp, const QgtDevOp* ops, int ovr_op,
                              const QgtDevOp& ovr, cplx* tile, uint64_t tilebase, int tid, const QgtCostTable& ct) {
    uint32_t lbase = 0;
    uint64_t gbase = tilebase;
    const int nthr_bits = run.K - R;
    for (int i = 0; i < nthr_bits; i++) {
        if ((tid >> i) & 1) {
            const int p = sp.tperm[i];
            lbase |= 1u << p;
            gbase |= 1ull << run.tq[p];
        }
    }
    uint32_t regl[R];
    uint64_t regg[R];
#pragma unroll
    for (int r = 0; r < R; ++r) { regl[r] = 1u << sp.regq[r]; regg[r] = 1ull << run.tq[sp.regq[r]]; }
    cplx v[1 << R];
#pragma unroll
    for (int c = 0; c < (1 << R); ++c) {
        uint32_t l = lbase;
#pragma unroll
        for (int r = 0; r < R; ++r) if (c & (1 << r)) l |= regl[r];
        v[c] = tile[qgt_swz(l)];
    }
    for (int o = sp.op_begin; o < sp.op_end; ++o) {
        if (o == ovr_op) qgt_apply_op<R>(v, ovr, gbase, regg, ct);
        else qgt_apply_op<R>(v, ops[o], gbase, regg, ct);
    }
#pragma unroll
    for (int c = 0; c < (1 << R); ++c) {
        uint32_t l = lbase;
#pragma unroll
        for (int r = 0; r < R; ++r) if (c & (1 << r)) l |= regl[r];
        tile[qgt_swz(l)] = v[c];
    }
}

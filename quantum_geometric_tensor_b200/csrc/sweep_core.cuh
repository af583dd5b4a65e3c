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

QGT_HD double qgt_fma(double a, double b, double c) {
#if defined(__CUDA_ARCH__)
    return fma(a, b, c);
#else
    return std::fma(a, b, c);
#endif
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

// ---- register-level work ---------------------------------------------------------------------------
// fold one thread-level diagonal gate into a pending scalar
QGT_HD void qgt_thread_diag(cplx& pend, const QgtDevThrDiag& t, uint64_t g) {
    if ((g & t.cmask) == t.cmask) {
        const int par = qgt_popc64(g & t.pmask) & 1;
        const double dr = par ? t.d[2] : t.d[0], di = par ? t.d[3] : t.d[1];
        const cplx a = pend;
        pend.x = dr * a.x - di * a.y;
        pend.y = dr * a.y + di * a.x;
    } else if (t.flags & QGT_FLAG_ZERO_CTRL_FAIL) {
        pend.x = 0.0; pend.y = 0.0;
    }
}

QGT_HD int qgt_variant_index(const QgtDevStage& st, uint64_t g) {
    int v = 0;
    for (int k = 0; k < st.nvar; ++k) v |= ((g & st.vmask[k]) != 0) << k;
    return v;
}

// element (i, j) of an 8x8 stage matrix stored in QGT_FORM_PARITY (dev_structs.h), for the paths that do plain arithmetic
QGT_HD cplx qgt_parity_elem(const cplx* M, int i, int j) {
    const int pi = (i ^ (i >> 1) ^ (i >> 2)) & 1, pj = (j ^ (j >> 1) ^ (j >> 2)) & 1;
    const cplx slot = M[QGT_MIDX(8, (pi << 2) | (i & 3), (pj << 2) | (j & 3))];
    cplx m; m.x = 0.0; m.y = 0.0;
    if (!pj) { if (!pi) m.x = slot.x; else m.y = slot.x; }       // even column: x = (row even ? Re : Im)
    else { if (!pi) m.y = slot.y; else m.x = slot.y; }           // odd column:  y = (row even ? Im : Re)
    return m;
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

// everything a sub-pass needs besides the tile: the run's stage / thread-diagonal tables, the matrix
// pool (staged in shared memory) and the item's override
struct QgtSubCtx {
    const QgtDevStage* stages;
    const QgtDevThrDiag* tdiags;
    const cplx* pool;            // the run's matrices (element 0 = run.mat_off of the plan-wide pool) ...
    int ovr_mat_off;             // ... followed, at this element offset, by the item's override matrices
    int ovr_kind, ovr_index;
    int ovr_form;                // QGT_FORM_* of the override matrices
    const QgtDevThrDiag* ovr_tdiag;
};

// One sub-pass for one thread.  The thread owns 2^(R+B) amplitudes: R "matrix" qubits and B "batch" qubits.
// Every gate of the sub-pass acting on the R matrix qubits was multiplied on the host into one dense
// 2^R x 2^R complex matrix per stage, so the device does no per-gate dispatch; the 2^B batch halves reuse
// each matrix element loaded from shared memory (the LSU wavefront rate, not the FP64 pipe, is the first
// limit otherwise).  Results go straight back to the thread's own tile slots, row by row.
template <int R, int B>
QGT_HD void qgt_phase_subpass(const QgtDevRun& run, const QgtDevSubPass& sp, const QgtSubCtx& cx, cplx* tile,
                              uint64_t tilebase, int tid) {
    constexpr int N = 1 << R;
    constexpr int NB = 1 << B;
    uint32_t sbase = 0;                  // swizzled slot of the thread's combo 0
    uint64_t gbase = tilebase;
    const int nthr_bits = run.K - R - B;
    for (int i = 0; i < nthr_bits; i++)
        if ((tid >> i) & 1) { sbase ^= sp.s_thr[i]; gbase |= sp.g_thr[i]; }
    uint32_t sreg[R + B];
#pragma unroll
    for (int r = 0; r < R + B; ++r) sreg[r] = sp.s_reg[r];
    // swizzled tile slot of (batch half h, matrix combo c)
    uint32_t slot[NB][N];
    uint64_t gh[NB];
#pragma unroll
    for (int h = 0; h < NB; ++h) {
        uint32_t sh = sbase;
        gh[h] = gbase;
#pragma unroll
        for (int r = 0; r < B; ++r) if (h & (1 << r)) { sh ^= sreg[R + r]; gh[h] |= sp.g_reg[R + r]; }
#pragma unroll
        for (int c = 0; c < N; ++c) {
            uint32_t l = sh;
#pragma unroll
            for (int r = 0; r < R; ++r) if (c & (1 << r)) l ^= sreg[r];
            slot[h][c] = l;
        }
    }
    cplx pend[NB];
#pragma unroll
    for (int h = 0; h < NB; ++h) { pend[h].x = 1.0; pend[h].y = 0.0; }
    for (int t = sp.tdiag_begin; t < sp.tdiag_end; ++t) {
        const QgtDevThrDiag& td = (cx.ovr_kind == 2 && t == cx.ovr_index) ? *cx.ovr_tdiag : cx.tdiags[t];
#pragma unroll
        for (int h = 0; h < NB; ++h) qgt_thread_diag(pend[h], td, gh[h]);
    }
    const int nstage = sp.stage_end - sp.stage_begin;
    for (int s = sp.stage_begin; s < sp.stage_end; ++s) {
        const QgtDevStage& st = cx.stages[s];
        const bool ovr = (cx.ovr_kind == 1 && s == cx.ovr_index);
        const int off = ovr ? cx.ovr_mat_off : st.mat_off;
        const bool diag_real = (ovr ? cx.ovr_form : (int)st.form) == QGT_FORM_DIAG_REAL;   // M = D * Rm: rows scaled afterwards
        const bool parity = (ovr ? cx.ovr_form : (int)st.form) == QGT_FORM_PARITY;
        const bool last = (s == sp.stage_end - 1);
        cplx v[NB][N];
#pragma unroll
        for (int h = 0; h < NB; ++h)
#pragma unroll
            for (int c = 0; c < N; ++c) v[h][c] = tile[slot[h][c]];
        const cplx* M[NB];
        bool same = true;
#pragma unroll
        for (int h = 0; h < NB; ++h) {
            M[h] = cx.pool + off + qgt_variant_index(st, gh[h]) * QGT_VARIANT_STRIDE(N);
            same = same && (M[h] == M[0]);
        }
        if (same) {
#pragma unroll
            for (int i = 0; i < N; ++i) {
                double xr[NB], xi[NB];
#pragma unroll
                for (int h = 0; h < NB; ++h) { xr[h] = 0.0; xi[h] = 0.0; }
#pragma unroll
                for (int j = 0; j < N; ++j) {
                    cplx m;
                    if (diag_real) { m.x = reinterpret_cast<const double*>(M[0])[QGT_MIDX(N, i, j)]; m.y = 0.0; }
                    else if (N == 8 && parity) m = qgt_parity_elem(M[0], i, j);
                    else m = M[0][QGT_MIDX(N, i, j)];
#pragma unroll
                    for (int h = 0; h < NB; ++h) {
                        xr[h] = qgt_fma(m.x, v[h][j].x, xr[h]); xr[h] = qgt_fma(-m.y, v[h][j].y, xr[h]);
                        xi[h] = qgt_fma(m.x, v[h][j].y, xi[h]); xi[h] = qgt_fma(m.y, v[h][j].x, xi[h]);
                    }
                }
#pragma unroll
                for (int h = 0; h < NB; ++h) {
                    cplx o; o.x = xr[h]; o.y = xi[h];
                    if (diag_real) { const cplx d = M[0][N * N + i]; const cplx q = o; o.x = d.x * q.x - d.y * q.y; o.y = d.x * q.y + d.y * q.x; }
                    if (last) { const cplx q = o; o.x = pend[h].x * q.x - pend[h].y * q.y; o.y = pend[h].x * q.y + pend[h].y * q.x; }
                    tile[slot[h][i]] = o;
                }
            }
        } else {
#pragma unroll
            for (int h = 0; h < NB; ++h) {
#pragma unroll 1
                for (int i = 0; i < N; ++i) {
                    double xr = 0.0, xi = 0.0;
#pragma unroll
                    for (int j = 0; j < N; ++j) {
                        cplx m;
                        if (diag_real) { m.x = reinterpret_cast<const double*>(M[h])[QGT_MIDX(N, i, j)]; m.y = 0.0; }
                        else if (N == 8 && parity) m = qgt_parity_elem(M[h], i, j);
                        else m = M[h][QGT_MIDX(N, i, j)];
                        xr = qgt_fma(m.x, v[h][j].x, xr); xr = qgt_fma(-m.y, v[h][j].y, xr);
                        xi = qgt_fma(m.x, v[h][j].y, xi); xi = qgt_fma(m.y, v[h][j].x, xi);
                    }
                    cplx o; o.x = xr; o.y = xi;
                    if (diag_real) { const cplx d = M[h][N * N + i]; const cplx q = o; o.x = d.x * q.x - d.y * q.y; o.y = d.x * q.y + d.y * q.x; }
                    if (last) { const cplx q = o; o.x = pend[h].x * q.x - pend[h].y * q.y; o.y = pend[h].x * q.y + pend[h].y * q.x; }
                    uint32_t l = sbase;   // slot[h][i] with a run-time i: recompute instead of indexing registers
#pragma unroll
                    for (int r = 0; r < R; ++r) if (i & (1 << r)) l ^= sreg[r];
#pragma unroll
                    for (int r = 0; r < B; ++r) if (h & (1 << r)) l ^= sreg[R + r];
                    tile[l] = o;
                }
            }
        }
    }
    if (nstage == 0) {          // only thread diagonals: scale the thread's amplitudes in place
#pragma unroll
        for (int h = 0; h < NB; ++h) {
            if (pend[h].x == 1.0 && pend[h].y == 0.0) continue;
#pragma unroll
            for (int c = 0; c < N; ++c) {
                const cplx q = tile[slot[h][c]];
                cplx o; o.x = pend[h].x * q.x - pend[h].y * q.y; o.y = pend[h].x * q.y + pend[h].y * q.x;
                tile[slot[h][c]] = o;
            }
        }
    }
}

// ---- cost layer ---------------------------------------------------------------------------------------
// E(z) = sum_edges w [z_i != z_j] + sum_q vw_q (1 - 2 z_q) is split by where the qubits live:
//   ein[idx]   both ends inside the tile (+ vertex terms of tile qubits): a 2^K table built once per CTA;
//   per tile   edges with both ends outside and outside vertex terms give a constant, edges with one end at
//              tile position p contribute  z_j + z_p (1 - 2 z_j): a constant plus lin[p] * z_p, folded into
//              two small tables lo[idx & 63], hi[idx >> 6].
// So the per-amplitude energy is three shared-memory reads instead of a loop over every edge.
#define QGT_COST_LO_BITS 6
#define QGT_COST_MAX_EDGES 1024
struct QgtCostSmem {
    double* ein;       // [2^K]
    double* lo;        // [64]
    double* hi;        // [2^(K-6)] (at least 1)
    double* lin;       // [K] slopes, [K] constants, [1] outside constant
    double* plo;       // [64][2]       exp(-i angle lo[x]) of the cost op being applied (re, im)
    double* phi;       // [2^(K-6)][2]  exp(-i angle hi[y])
    double* outp;      // [32] shares of the tile's outside constant (edges and vertex terms with no end in the tile)
    double* cross_w;   // cross edges (one end at a tile position) grouped by position: weights ...
    double* out_w;     // edges with both ends outside: weights ...
    int16_t* cross_start;   // [K+1] CSR offsets into cross_w / cross_q
    int8_t* cross_q;   // ... and the outside qubit of each
    int8_t* out_i;     // ... and their two qubits
    int8_t* out_j;
    int* n_out;        // [1]
};

// with_ein = false: the launch carries global tables (phases per cost op + ein), the 2^K energy table leaves shared memory
QGT_HD size_t qgt_cost_smem_doubles(int K, int num_edges, bool with_ein = true) {
    const size_t ne = (size_t)(num_edges > 0 ? num_edges : 1);
    const size_t bytes_small = 2 * (QGT_MAX_TILE_QUBITS + 2) + 3 * ne + 16;     // int16 starts + three int8 arrays + count
    return (with_ein ? ((size_t)1 << K) : 0) + 3 * (64 + ((size_t)1 << (K > QGT_COST_LO_BITS ? K - QGT_COST_LO_BITS : 0))) + 32 + 2 * QGT_MAX_TILE_QUBITS + 8 +
           2 * ne + (bytes_small + 7) / 8;
}

QGT_HD QgtCostSmem qgt_cost_smem_carve(double* base, int K, int num_edges, bool with_ein = true) {
    const size_t ne = (size_t)(num_edges > 0 ? num_edges : 1);
    QgtCostSmem c;
    c.ein = with_ein ? base : nullptr;
    c.lo = base + (with_ein ? ((size_t)1 << K) : 0);
    c.hi = c.lo + 64;
    const size_t nhi = (size_t)1 << (K > QGT_COST_LO_BITS ? K - QGT_COST_LO_BITS : 0);
    c.plo = c.hi + nhi;
    c.phi = c.plo + 128;
    c.outp = c.phi + 2 * nhi;
    c.lin = c.outp + 32;
    c.cross_w = c.lin + 2 * QGT_MAX_TILE_QUBITS + 8;
    c.out_w = c.cross_w + ne;
    c.cross_start = reinterpret_cast<int16_t*>(c.out_w + ne);
    c.cross_q = reinterpret_cast<int8_t*>(c.cross_start + QGT_MAX_TILE_QUBITS + 2);
    c.out_i = c.cross_q + ne;
    c.out_j = c.out_i + ne;
    c.n_out = reinterpret_cast<int*>(reinterpret_cast<uintptr_t>(c.out_j + ne + 3) & ~(uintptr_t)3);
    return c;
}

QGT_HD int qgt_local_pos(const QgtDevRun& run, int q) {
    for (int j = 0; j < run.K; j++) if (run.tq[j] == q) return j;
    return -1;
}

// once per CTA: thread `tid` of T fills its share of ein[]; thread 0 also sorts the edges into the cross
// lists (by tile position) and the outside list
QGT_HD double qgt_cost_ein(const QgtDevRun& run, const QgtCostTable& ct, uint32_t idx) {
    double e = 0.0;
    for (int k = 0; k < ct.num_edges; k++) {
        const int li = qgt_local_pos(run, ct.edges[k].i), lj = qgt_local_pos(run, ct.edges[k].j);
        if (li >= 0 && lj >= 0 && (((idx >> li) ^ (idx >> lj)) & 1u)) e += ct.edges[k].w;
    }
    if (ct.vertex_weights)
        for (int j = 0; j < run.K; j++) e += ct.vertex_weights[run.tq[j]] * (double)(1 - 2 * (int)((idx >> j) & 1u));
    return e;
}

QGT_HD void qgt_sincos(double x, double* sn, double* cs_) {
#if defined(__CUDA_ARCH__)
    sincos(x, sn, cs_);
#else
    *sn = std::sin(x); *cs_ = std::cos(x);
#endif
}

// entry idx of the per-launch phase table of one cost op: exp(-i angle ein[idx]) (filled by a small kernel before the sweep,
// read per amplitude instead of a sincos: the tile part of the phase then comes from two small per-tile tables)
QGT_HD cplx qgt_cost_phase_entry(const QgtDevRun& run, const QgtCostTable& ct, double angle, uint32_t idx) {
    cplx p;
    qgt_sincos(-angle * qgt_cost_ein(run, ct, idx), &p.y, &p.x);
    return p;
}

QGT_HD void qgt_cost_build_ein(const QgtDevRun& run, const QgtCostTable& ct, const QgtCostSmem& cs, int tid, int T) {
    const uint32_t count = 1u << run.K;
    if (cs.ein)
        for (uint32_t idx = (uint32_t)tid; idx < count; idx += (uint32_t)T) cs.ein[idx] = qgt_cost_ein(run, ct, idx);
    if (tid == 0) {
        int nc = 0, no = 0;
        for (int p = 0; p < run.K; p++) {
            cs.cross_start[p] = (int16_t)nc;
            const int qp = run.tq[p];
            for (int k = 0; k < ct.num_edges; k++) {
                const int i = ct.edges[k].i, j = ct.edges[k].j;
                const int other = (i == qp) ? j : (j == qp ? i : -1);
                if (other < 0 || qgt_local_pos(run, other) >= 0) continue;
                cs.cross_q[nc] = (int8_t)other; cs.cross_w[nc] = ct.edges[k].w; nc++;
            }
        }
        cs.cross_start[run.K] = (int16_t)nc;
        for (int k = 0; k < ct.num_edges; k++) {
            const int i = ct.edges[k].i, j = ct.edges[k].j;
            if (qgt_local_pos(run, i) >= 0 || qgt_local_pos(run, j) >= 0) continue;
            cs.out_i[no] = (int8_t)i; cs.out_j[no] = (int8_t)j; cs.out_w[no] = ct.edges[k].w; no++;
        }
        *cs.n_out = no;
    }
}

// per tile, step 1, threads 0..31: thread p < K computes lin[p] and its constant from its own cross list; every one of the
// 32 takes its share (edges / qubits p, p + 32, ...) of the outside constant - one thread walking all outside edges and
// qubits was a serial chain as long as the tile's whole HBM time
#define QGT_COST_LIN_THREADS 32
QGT_HD void qgt_cost_tile_lin(const QgtDevRun& run, const QgtCostTable& ct, const QgtCostSmem& cs, uint64_t tileg, int tid) {
    if (tid < run.K) {
        double lin = 0.0, c0 = 0.0;
        for (int k = cs.cross_start[tid]; k < cs.cross_start[tid + 1]; k++) {
            const double zj = (double)((tileg >> cs.cross_q[k]) & 1ull);
            lin += cs.cross_w[k] * (1.0 - 2.0 * zj);
            c0 += cs.cross_w[k] * zj;
        }
        cs.lin[tid] = lin;
        cs.lin[QGT_MAX_TILE_QUBITS + tid] = c0;
    }
    if (tid < QGT_COST_LIN_THREADS) {
        double e = 0.0;
        const int no = *cs.n_out;
        for (int k = tid; k < no; k += QGT_COST_LIN_THREADS)
            if (((tileg >> cs.out_i[k]) ^ (tileg >> cs.out_j[k])) & 1ull) e += cs.out_w[k];
        if (ct.vertex_weights)
            for (int q = tid; q < ct.n; q += QGT_COST_LIN_THREADS)
                if (qgt_local_pos(run, q) < 0) e += ct.vertex_weights[q] * (double)(1 - 2 * (int)((tileg >> q) & 1ull));
        cs.outp[tid] = e;
    }
}

// per tile, step 2 (after a barrier): the two small tables; the constants are folded into lo[]
// (`angle`: of the cost op being applied; the phases of the two tables go next to their energies)
QGT_HD void qgt_cost_tile_tables(const QgtDevRun& run, const QgtCostSmem& cs, int tid, int T, double angle) {
    const int lo_bits = run.K < QGT_COST_LO_BITS ? run.K : QGT_COST_LO_BITS;
    const int hi_bits = run.K - lo_bits;
    for (int x = tid; x < (1 << lo_bits) + (1 << hi_bits); x += T) {
        if (x < (1 << lo_bits)) {
            double e = 0.0;
            for (int p = 0; p < QGT_COST_LIN_THREADS; p++) e += cs.outp[p];          // fixed order
            for (int p = 0; p < run.K; p++) e += cs.lin[QGT_MAX_TILE_QUBITS + p];
            for (int p = 0; p < lo_bits; p++) if ((x >> p) & 1) e += cs.lin[p];
            cs.lo[x] = e;
            qgt_sincos(-angle * e, &cs.plo[2 * x + 1], &cs.plo[2 * x]);
        } else {
            const int y = x - (1 << lo_bits);
            double e = 0.0;
            for (int p = 0; p < hi_bits; p++) if ((y >> p) & 1) e += cs.lin[lo_bits + p];
            cs.hi[y] = e;
            qgt_sincos(-angle * e, &cs.phi[2 * y + 1], &cs.phi[2 * y]);
        }
    }
}

// per tile, step 3 (after a barrier): thread `tid` of T handles local indices tid, tid+T, ... in shared memory
// `ptab`: the op's phase table exp(-i angle ein[idx]) (qgt_cost_phase_entry), or null = one sincos per amplitude
// (then `etab` = ein[] in global memory, read by derivative items only)
QGT_HD void qgt_phase_cost(const QgtDevRun& run, const QgtDevCost& op, cplx* tile, const QgtCostSmem& cs, const cplx* ptab, const double* etab,
                           int tid, int T) {
    const uint32_t count = 1u << run.K;
    const int lo_bits = run.K < QGT_COST_LO_BITS ? run.K : QGT_COST_LO_BITS;
    const bool need_e = !ptab || (op.flags & QGT_FLAG_COST_DERIV);
    for (uint32_t idx = (uint32_t)tid; idx < count; idx += (uint32_t)T) {
        const uint32_t xl = idx & ((1u << lo_bits) - 1u), xh = idx >> lo_bits;
        const double e = need_e ? (cs.ein ? cs.ein[idx] : etab[idx]) + cs.lo[xl] + cs.hi[xh] : 0.0;
        double sn, cs_;
        if (ptab) {
            // exp(-i angle e) = table entry x phase of the low tile bits x phase of the high tile bits
            const cplx p = ptab[idx];
            const double lr = cs.plo[2 * xl], li = cs.plo[2 * xl + 1], hr = cs.phi[2 * xh], hi_ = cs.phi[2 * xh + 1];
            const double tr = lr * hr - li * hi_, ti = lr * hi_ + li * hr;
            cs_ = p.x * tr - p.y * ti; sn = p.x * ti + p.y * tr;
        } else {
            qgt_sincos(-op.angle * e, &sn, &cs_);
        }
        const cplx a = tile[qgt_swz(idx)];
        cplx b; b.x = cs_ * a.x - sn * a.y; b.y = cs_ * a.y + sn * a.x;
        if (op.flags & QGT_FLAG_COST_DERIV) {   // times (-i * scale * E)
            const double f = op.dscale * e;
            cplx d; d.x = f * b.y; d.y = -f * b.x; b = d;
        }
        tile[qgt_swz(idx)] = b;
    }
}

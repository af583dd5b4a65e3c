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
#include "mma_common.cuh"

#include <cstdio>

namespace qgt {

// Timing experiments (build with -DQGT_DEBUG_SKIP; results are wrong by construction): option debug_skip
// bit 1 no tile load, 2 no store, 4 no sub-passes, 8 no barrier between sub-passes, 16 no operand
// loads / result stores inside a sub-pass, 32 no DMMAs.
#ifdef QGT_DEBUG_SKIP
#define QGT_DBG(bit) ((a.debug_skip & (bit)) != 0)
#define QGT_DBGF(bit) ((dbg & (bit)) != 0)
#else
#define QGT_DBG(bit) false
#define QGT_DBGF(bit) false
#endif

// ------------------------------------------------------------------------------------------------
// gate sweep
// ------------------------------------------------------------------------------------------------
// Streamlined tensor-pipe sub-pass for the common shape: exactly one dense stage, no thread diagonals.
// Same fragment layout as qgt_warp_subpass_mma below.  All 8 operand loads are issued before the first
// DMMA (one __syncwarp for the four groups), so the shared-memory latency is paid once per sub-pass.
__device__ __forceinline__ void qgt_warp_subpass_fast(const QgtFastSub& f, const QgtFastWarp& fw, uint32_t lt, const QgtSubCtx& cx,
                                                      cplx* tile, uint64_t tileg, int lane) {
    constexpr int N = 8;
    const uint32_t baseB = fw.s ^ (lt & 0xffffu), baseC = fw.s ^ (lt >> 16);
    const uint64_t gwarp = tileg | fw.g;
    const bool ovr = (cx.ovr_kind == 1 && (int)f.stage == cx.ovr_index);
    const int off = ovr ? cx.ovr_mat_off : (int)f.mat_off;
    const bool diag_real = (ovr ? (uint32_t)cx.ovr_form : f.form) == QGT_FORM_DIAG_REAL;
    const int var = ((gwarp & f.vm0) != 0 ? 1 : 0) | ((gwarp & f.vm1) != 0 ? 2 : 0);
    const cplx* M = cx.pool + off + var * QGT_VARIANT_STRIDE(N);
    const uint32_t gx1 = f.gx1, gx2 = f.gx2, sr2 = f.sr2, st0 = f.st0;    // locals: tile stores must not force reloads
    if (diag_real) {
        // M = D * Rm with Rm real: out = D (Rm v_re + i Rm v_im), 4 DMMAs per 8 vectors and one complex multiply
        // per result (row lane>>2 of the C fragment).  A fragments: QGT_MIDX(8, q, k) == lane, packed doubles
        cplx m0, m1;
        m0.x = reinterpret_cast<const double*>(M)[lane]; m1.x = reinterpret_cast<const double*>(M)[32 + lane];
        const cplx d = M[N * N + (lane >> 2)];
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            cplx v0[2], v1[2];
#pragma unroll
            for (int g = 0; g < 2; ++g) {
                const uint32_t gx = (g ? gx1 : 0u) ^ (h ? gx2 : 0u);
                v0[g] = tile[baseB ^ gx];
                v1[g] = tile[baseB ^ gx ^ sr2];
            }
            __syncwarp();
#pragma unroll
            for (int g = 0; g < 2; ++g) {
                const uint32_t gx = (g ? gx1 : 0u) ^ (h ? gx2 : 0u);
                double cr0 = 0.0, cr1 = 0.0, ci0 = 0.0, ci1 = 0.0;
                dmma884(cr0, cr1, m0.x, v0[g].x);
                dmma884(ci0, ci1, m0.x, v0[g].y);
                dmma884(cr0, cr1, m1.x, v1[g].x);
                dmma884(ci0, ci1, m1.x, v1[g].y);
                cplx o0, o1;
                o0.x = d.x * cr0 - d.y * ci0; o0.y = d.x * ci0 + d.y * cr0;
                o1.x = d.x * cr1 - d.y * ci1; o1.y = d.x * ci1 + d.y * cr1;
                tile[baseC ^ gx] = o0;
                tile[baseC ^ gx ^ st0] = o1;
            }
        }
        return;
    }
    const cplx m0 = M[lane], m1 = M[32 + lane];           // A fragments: QGT_MIDX(8, q, k) == lane
    if ((ovr ? (uint32_t)cx.ovr_form : f.form) == QGT_FORM_PARITY) {
        // real part on even, imaginary part on odd index distance (dev_structs.h): inputs and results are taken in the
        // (parity, bit 1, bit 0) order of the component index - which only changes the slots this lane reads and writes -
        // and 4 DMMAs give [y_e.re; y_o.im] and [y_e.im; y_o.re]
        const uint32_t pb = (__popc(lane & 3) & 1) ? sr2 : 0u;             // B operand: the even-parity component of column pair k first
        const uint32_t pc = (__popc((lane >> 2) & 3) & 1) ? sr2 : 0u;      // C rows: (p, b1, b0) -> component bit 2 = p ^ b1 ^ b0
        const bool odd_row = lane >= 16;
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            cplx v0[2], v1[2];
#pragma unroll
            for (int g = 0; g < 2; ++g) {
                const uint32_t gx = (g ? gx1 : 0u) ^ (h ? gx2 : 0u);
                v0[g] = tile[baseB ^ gx ^ pb];
                v1[g] = tile[baseB ^ gx ^ pb ^ sr2];
            }
            __syncwarp();
#pragma unroll
            for (int g = 0; g < 2; ++g) {
                const uint32_t gx = (g ? gx1 : 0u) ^ (h ? gx2 : 0u);
                double x0 = 0.0, x1 = 0.0, y0 = 0.0, y1 = 0.0;
                dmma884(x0, x1, m0.x, v0[g].x);
                dmma884(y0, y1, m0.y, v0[g].y);
                dmma884(x0, x1, m1.x, v1[g].y);
                dmma884(y0, y1, m1.y, v1[g].x);
                cplx o0, o1;
                o0.x = odd_row ? y0 : x0; o0.y = odd_row ? x0 : y0;
                o1.x = odd_row ? y1 : x1; o1.y = odd_row ? x1 : y1;
                tile[baseC ^ gx ^ pc] = o0;
                tile[baseC ^ gx ^ pc ^ st0] = o1;
            }
        }
        return;
    }
    const double nm0y = -m0.y, nm1y = -m1.y;
#pragma unroll
    for (int h = 0; h < 2; ++h) {     // two groups of 8 vectors at a time: operand loads first, then the DMMAs
        cplx v0[2], v1[2];
#pragma unroll
        for (int g = 0; g < 2; ++g) {
            const uint32_t gx = (g ? gx1 : 0u) ^ (h ? gx2 : 0u);
            v0[g] = tile[baseB ^ gx];
            v1[g] = tile[baseB ^ gx ^ sr2];
        }
        __syncwarp();                 // every lane has read its slots before any is overwritten
#pragma unroll
        for (int g = 0; g < 2; ++g) {
            const uint32_t gx = (g ? gx1 : 0u) ^ (h ? gx2 : 0u);
            double cr0 = 0.0, cr1 = 0.0, ci0 = 0.0, ci1 = 0.0;
            dmma884(cr0, cr1, m0.x, v0[g].x);
            dmma884(ci0, ci1, m0.x, v0[g].y);
            dmma884(cr0, cr1, m1.x, v1[g].x);
            dmma884(ci0, ci1, m1.x, v1[g].y);
            dmma884(cr0, cr1, nm0y, v0[g].y);
            dmma884(ci0, ci1, m0.y, v0[g].x);
            dmma884(cr0, cr1, nm1y, v1[g].y);
            dmma884(ci0, ci1, m1.y, v1[g].x);
            cplx o0, o1;
            o0.x = cr0; o0.y = ci0; o1.x = cr1; o1.y = ci1;
            tile[baseC ^ gx] = o0;
            tile[baseC ^ gx ^ st0] = o1;
        }
    }
}

// Tensor-pipe version of a sub-pass with 3 matrix qubits (DMMA.8x8x4).  A "vector" is the 8 amplitudes one
// thread of the register path would own; a warp owns 32 vectors = 4 groups of 8.  Per stage the 8x8 complex
// matrix M sits in A fragments (lane (r, k) holds M[r][k] and M[r][4+k]), the 8 vectors of a group form the
// B operand (lane (n, k) holds amplitudes k and 4+k of vector n) and
//     out_re = Mre Vre - Mim Vim,   out_im = Mre Vim + Mim Vre
// takes 8 DMMAs per group.  The C fragment (lane (r, k') holds row r of vectors 2k', 2k'+1) goes straight
// back to the vectors' tile slots.  Needs every stage's variant to be uniform over the warp (sp.mma_ok).
__device__ __forceinline__ void qgt_warp_subpass_mma(const QgtDevRun& run, const QgtDevSubPass& sp, const QgtSubCtx& cx,
                                                     cplx* tile, uint64_t tilebase, int warp, int lane) {
    constexpr int N = 8;
    const int q = lane >> 2, k = lane & 3;
    const int nthr_bits = run.K - 3;
    // virtual thread (= vector) index t = warp*32 + g*8 + v: bits 0..2 <- v, bits 3..4 <- g, bits 5.. <- warp.
    // Slots are XORs of the host-precomputed per-bit swizzled contributions.
    uint32_t st5[5];
#pragma unroll
    for (int i = 0; i < 5; ++i) st5[i] = sp.s_thr[i];
    const uint32_t sr0 = sp.s_reg[0], sr1 = sp.s_reg[1], sr2 = sp.s_reg[2];
    uint32_t swarp = 0;
    uint64_t gwarp = tilebase;
    for (int i = 5; i < nthr_bits; ++i)
        if ((warp >> (i - 5)) & 1) { swarp ^= sp.s_thr[i]; gwarp |= sp.g_thr[i]; }
    const uint32_t s_q = ((q & 1) ? st5[0] : 0u) ^ ((q & 2) ? st5[1] : 0u) ^ ((q & 4) ? st5[2] : 0u);    // vector q
    const uint32_t s_k = ((k & 1) ? st5[1] : 0u) ^ ((k & 2) ? st5[2] : 0u);                               // vector 2k
    const uint32_t c_k = ((k & 1) ? sr0 : 0u) ^ ((k & 2) ? sr1 : 0u);                                     // combo k
    const uint32_t c_q = ((q & 1) ? sr0 : 0u) ^ ((q & 2) ? sr1 : 0u) ^ ((q & 4) ? sr2 : 0u);              // combo q
    // slots of this lane: B layout (vector g*8+q, combos k and 4+k), C layout (vectors g*8+2k, +1, combo q)
    uint32_t sB0[4], sC0[4];
    const bool has_tdiag = sp.tdiag_end > sp.tdiag_begin;
    cplx pend0[4], pend1[4];
#pragma unroll
    for (int g = 0; g < 4; ++g) {
        const uint32_t sg = swarp ^ ((g & 1) ? st5[3] : 0u) ^ ((g & 2) ? st5[4] : 0u);
        sB0[g] = sg ^ s_q ^ c_k;          // combo 4+k: ^ sr2
        sC0[g] = sg ^ s_k ^ c_q;          // vector 2k+1: ^ st5[0]
        if (has_tdiag) {
            pend0[g].x = 1.0; pend0[g].y = 0.0; pend1[g] = pend0[g];
            uint64_t gc0 = gwarp;
            const int t0 = g * 8 + 2 * k;
#pragma unroll
            for (int i = 0; i < 5; ++i) if ((t0 >> i) & 1) gc0 |= sp.g_thr[i];
            const uint64_t gc1 = gc0 | sp.g_thr[0];
            for (int t = sp.tdiag_begin; t < sp.tdiag_end; ++t) {
                const QgtDevThrDiag& td = (cx.ovr_kind == 2 && t == cx.ovr_index) ? *cx.ovr_tdiag : cx.tdiags[t];
                qgt_thread_diag(pend0[g], td, gc0);
                qgt_thread_diag(pend1[g], td, gc1);
            }
        }
    }
    const int nstage = sp.stage_end - sp.stage_begin;
    for (int s = sp.stage_begin; s < sp.stage_end; ++s) {
        const QgtDevStage& st = cx.stages[s];
        const bool ovr = (cx.ovr_kind == 1 && s == cx.ovr_index);
        const int off = ovr ? cx.ovr_mat_off : st.mat_off;
        const bool diag_real = (ovr ? cx.ovr_form : (int)st.form) == QGT_FORM_DIAG_REAL;
        const cplx* M = cx.pool + off + qgt_variant_index(st, gwarp) * QGT_VARIANT_STRIDE(N);
        cplx m0, m1, dq;
        dq.x = 1.0; dq.y = 0.0;
        if (diag_real) {                                  // M = D * Rm: real fragments (packed doubles), row q scaled afterwards
            m0.x = reinterpret_cast<const double*>(M)[QGT_MIDX(N, q, k)]; m1.x = reinterpret_cast<const double*>(M)[QGT_MIDX(N, q, 4 + k)];
            m0.y = 0.0; m1.y = 0.0;
            dq = M[N * N + q];
        } else if ((ovr ? cx.ovr_form : (int)st.form) == QGT_FORM_PARITY) {
            m0 = qgt_parity_elem(M, q, k); m1 = qgt_parity_elem(M, q, 4 + k);      // (multi-stage sub-passes: plain 8-DMMA form)
        } else {
            m0 = M[QGT_MIDX(N, q, k)]; m1 = M[QGT_MIDX(N, q, 4 + k)];
        }
        const double nm0y = -m0.y, nm1y = -m1.y;
        const bool last = (s == sp.stage_end - 1);
#pragma unroll
        for (int g = 0; g < 4; ++g) {
            const cplx v0 = tile[sB0[g]], v1 = tile[sB0[g] ^ sr2];
            double cr0 = 0.0, cr1 = 0.0, ci0 = 0.0, ci1 = 0.0;
            dmma884(cr0, cr1, m0.x, v0.x);
            dmma884(ci0, ci1, m0.x, v0.y);
            dmma884(cr0, cr1, m1.x, v1.x);
            dmma884(ci0, ci1, m1.x, v1.y);
            dmma884(cr0, cr1, nm0y, v0.y);
            dmma884(ci0, ci1, m0.y, v0.x);
            dmma884(cr0, cr1, nm1y, v1.y);
            dmma884(ci0, ci1, m1.y, v1.x);
            cplx o0, o1;
            o0.x = cr0; o0.y = ci0; o1.x = cr1; o1.y = ci1;
            if (diag_real) {
                cplx t0 = o0, t1 = o1;
                o0.x = dq.x * t0.x - dq.y * t0.y; o0.y = dq.x * t0.y + dq.y * t0.x;
                o1.x = dq.x * t1.x - dq.y * t1.y; o1.y = dq.x * t1.y + dq.y * t1.x;
            }
            if (last && has_tdiag) {
                cplx t0 = o0, t1 = o1;
                o0.x = pend0[g].x * t0.x - pend0[g].y * t0.y; o0.y = pend0[g].x * t0.y + pend0[g].y * t0.x;
                o1.x = pend1[g].x * t1.x - pend1[g].y * t1.y; o1.y = pend1[g].x * t1.y + pend1[g].y * t1.x;
            }
            __syncwarp();                 // every lane has read the group's slots before they are overwritten
            tile[sC0[g]] = o0;
            tile[sC0[g] ^ st5[0]] = o1;
        }
        __syncwarp();                     // the next stage reads slots other lanes just wrote
    }
    if (nstage == 0 && has_tdiag) {       // only thread diagonals
#pragma unroll
        for (int g = 0; g < 4; ++g) {
            const cplx a0 = tile[sC0[g]], a1 = tile[sC0[g] ^ st5[0]];
            cplx o0, o1;
            o0.x = pend0[g].x * a0.x - pend0[g].y * a0.y; o0.y = pend0[g].x * a0.y + pend0[g].y * a0.x;
            o1.x = pend1[g].x * a1.x - pend1[g].y * a1.y; o1.y = pend1[g].x * a1.y + pend1[g].y * a1.x;
            tile[sC0[g]] = o0; tile[sC0[g] ^ st5[0]] = o1;
        }
    }
}

// Shared memory: [tile: 2^K amplitudes][matrix pool of the run][override matrices of the current item].
// The kernel is persistent over (tile, column) work items; the run header and its matrix pool are staged
// once per CTA.  R = qubits of a stage matrix, B = batch qubits: a thread owns 2^(R+B) amplitudes.
// COST (with MMA_ONLY): the run also holds cost-layer passes; their energy tables sit behind the lookup tables.
// MMA_ONLY: every sub-pass of the run takes the tensor-pipe path (no cost pass, warp-uniform variants): the
// register-FMA code is not instantiated, which halves the register count and doubles the resident warps.
// DB: two tile buffers with cp.async prefetch of the next work item.
template <int R, int B, bool MMA_ONLY, bool DB, int MAXT, int MINB, bool COST = false>
__global__ void __launch_bounds__(MAXT, MINB) qgt_sweep_kernel(SweepLaunch a) {
    constexpr int N = 1 << R;
    constexpr int OVR_ELEMS = QGT_VARIANT_STRIDE(N) << QGT_MAX_VARIANT_BITS;
    extern __shared__ __align__(16) unsigned char qgt_smem_raw[];
    cplx* tile = reinterpret_cast<cplx*>(qgt_smem_raw);
    __shared__ QgtDevRun run;
    const int tid = threadIdx.x;
    const int T = blockDim.x;
    for (int i = tid; i < (int)(sizeof(QgtDevRun) / 4); i += T)
        reinterpret_cast<uint32_t*>(&run)[i] = reinterpret_cast<const uint32_t*>(&a.runs[a.run_idx])[i];
    __syncthreads();
    // two tile buffers: the next work item's tile streams in with cp.async while this one is processed
    cplx* spool = tile + ((size_t)(DB ? 2 : 1) << run.K);
    cplx* sovr = spool + run.mat_count;
    QgtDevSubPass* subs = reinterpret_cast<QgtDevSubPass*>(sovr + OVR_ELEMS);
    // [sub-pass descriptors][cost tables (runs with a cost pass)][lookup tables of the tensor-only kernel]
    const bool ein_smem = a.cost_phase == nullptr;         // no global tables: the energy table lives in shared memory
    QgtCostSmem cost_sm = qgt_cost_smem_carve(reinterpret_cast<double*>(subs + run.nsub), run.K, a.ct.num_edges, ein_smem);
    QgtFastSub* fast = reinterpret_cast<QgtFastSub*>(reinterpret_cast<double*>(subs + run.nsub) +
                                                     ((!MMA_ONLY || COST) && run.has_cost ? qgt_cost_smem_doubles(run.K, a.ct.num_edges, ein_smem) : 0));
    QgtFastWarp* fwarp = reinterpret_cast<QgtFastWarp*>(fast + run.nsub);
    uint32_t* flane = reinterpret_cast<uint32_t*>(fwarp + 8 * run.nsub);
    {
        const cplx* gpool = a.pool + run.mat_off;
        for (int i = tid; i < run.mat_count; i += T) spool[i] = gpool[i];
        const uint32_t* gs = reinterpret_cast<const uint32_t*>(a.subs + run.sub_off);
        uint32_t* ss = reinterpret_cast<uint32_t*>(subs);
        for (int i = tid; i < run.nsub * (int)(sizeof(QgtDevSubPass) / 4); i += T) ss[i] = gs[i];
        if ((!MMA_ONLY || COST) && run.has_cost) qgt_cost_build_ein(run, a.ct, cost_sm, tid, T);
    }
    if (MMA_ONLY) {
        __syncthreads();
        qgt_fast_build(run, subs, a.stages + run.stage_off, fast, fwarp, flane, tid, T);
    }
    QgtSubCtx cx;
    cx.stages = a.stages + run.stage_off;
    cx.tdiags = a.tdiags + run.tdiag_off;
    cx.pool = spool;
    cx.ovr_mat_off = run.mat_count;
    const QgtIoMap<R + B> io = qgt_make_iomap<R + B>(run, tid);
    // work unit = (group of `tpi` consecutive tiles, column item), items fastest: concurrently running CTAs read the
    // same source tiles (one phi spawns many columns).  The item set-up (descriptor, override matrices) is paid once
    // per unit.
    const int tpi = a.tiles_per_item;
    const uint64_t total = (a.ntiles / (uint64_t)tpi) * (uint64_t)a.nitems;
    auto prefetch = [&](uint64_t w, int t, cplx* buf) {
        const QgtSweepItem& it = a.items[(int)(w % (uint64_t)a.nitems)];
        const uint64_t tb = qgt_tile_base(run, (w / (uint64_t)a.nitems) * (uint64_t)tpi + (uint64_t)t);
        const cplx* src = reinterpret_cast<const cplx*>(it.src);
#pragma unroll
        for (int i = 0; i < (1 << (R + B)); ++i) {
            const uint32_t idx = (uint32_t)tid + (uint32_t)i * (uint32_t)T;
            cp_async16(buf + qgt_swz(idx), src + (tb | qgt_io_offset<R + B>(io, i)), 16);
        }
    };
    uint64_t w = blockIdx.x;
    int t = 0;
    if (DB) {
        if (w < total) prefetch(w, 0, tile);
        cp_async_commit();
    }
    for (int par = 0; w < total; par ^= 1) {
        cplx* cur = DB ? tile + ((size_t)par << run.K) : tile;
        const int item = (int)(w % (uint64_t)a.nitems);
        const uint64_t tau = (w / (uint64_t)a.nitems) * (uint64_t)tpi + (uint64_t)t;
        const QgtSweepItem& it = a.items[item];
        const uint64_t tilebase = qgt_tile_base(run, tau);
        const uint64_t tileg = tilebase | a.gprefix;      // global index bits incl. the rank's (sharded states)
        cx.ovr_kind = it.ovr_kind;
        cx.ovr_index = it.ovr_index;
        cx.ovr_form = it.ovr_form;
        cx.ovr_tdiag = &it.ovr_tdiag;
        if (t == 0 && it.ovr_kind == 1) {
            // every warp is past the previous unit's last sub-pass (barrier), so the override buffer is free
            const cplx* g = reinterpret_cast<const cplx*>(it.ovr_mat);
            const int cnt = QGT_VARIANT_STRIDE(N) << cx.stages[it.ovr_index].nvar;
            for (int i = tid; i < cnt && i < OVR_ELEMS; i += T) sovr[i] = g[i];
        }
        int nt = t + 1;
        uint64_t nw = w;
        if (nt == tpi) { nt = 0; nw = w + gridDim.x; }
        if (DB) {
            // the other buffer was last read by this thread itself (store phase, same slots as the load)
            if (nw < total) prefetch(nw, nt, tile + ((size_t)(par ^ 1) << run.K));
            cp_async_commit();
            cp_async_wait<1>();
        } else {
            if (!QGT_DBG(1)) prefetch(w, t, cur);
            cp_async_commit();
            cp_async_wait<0>();
        }
        __syncthreads();
        for (int s = 0; s < (QGT_DBG(4) ? 0 : run.nsub); ++s) {
            if (MMA_ONLY && COST && subs[s].nreg == 0) {
                const QgtDevCost& co = (it.ovr_kind == 3 && subs[s].cost == it.ovr_index) ? it.ovr_cost : a.costs[run.cost_off + subs[s].cost];
                for (int t2 = tid; t2 < QGT_COST_LIN_THREADS; t2 += T) qgt_cost_tile_lin(run, a.ct, cost_sm, tileg, t2);
                __syncthreads();
                qgt_cost_tile_tables(run, cost_sm, tid, T, co.angle);
                __syncthreads();
                qgt_phase_cost(run, co, cur, cost_sm, a.cost_phase ? a.cost_phase + ((size_t)subs[s].cost << run.K) : nullptr, a.cost_ein, tid, T);
            } else if (MMA_ONLY) {
                if (fast[s].simple)
                    qgt_warp_subpass_fast(fast[s], fwarp[s * 8 + (tid >> 5)], flane[s * 32 + (tid & 31)], cx, cur, tileg, tid & 31);
                else
                    qgt_warp_subpass_mma(run, subs[s], cx, cur, tileg, tid >> 5, tid & 31);
            } else if (subs[s].nreg == 0) {
                const QgtDevCost& co = (it.ovr_kind == 3 && subs[s].cost == it.ovr_index) ? it.ovr_cost : a.costs[run.cost_off + subs[s].cost];
                for (int t2 = tid; t2 < QGT_COST_LIN_THREADS; t2 += T) qgt_cost_tile_lin(run, a.ct, cost_sm, tileg, t2);
                __syncthreads();
                qgt_cost_tile_tables(run, cost_sm, tid, T, co.angle);
                __syncthreads();
                qgt_phase_cost(run, co, cur, cost_sm, a.cost_phase ? a.cost_phase + ((size_t)subs[s].cost << run.K) : nullptr, a.cost_ein, tid, T);
            } else if (R == 3 && B == 0 && a.use_mma && subs[s].mma_ok && T >= 32) {
                qgt_warp_subpass_mma(run, subs[s], cx, cur, tileg, tid >> 5, tid & 31);
            } else {
                qgt_phase_subpass<R, B>(run, subs[s], cx, cur, tileg, tid);
            }
            if (!QGT_DBG(8)) __syncthreads();
        }
        // no barrier after the store: the next load of this buffer writes, per thread, exactly the slots the thread
        // has just read (load and store use the same index map), and the next sub-pass sits behind the load barrier
        if (!QGT_DBG(2)) qgt_phase_store<R + B>(io, cur, reinterpret_cast<cplx*>(it.dst), tilebase, tid, T, it.accumulate != 0);
        t = nt; w = nw;
    }
    cp_async_wait<0>();
}

template <int R, int B, bool MMA_ONLY, bool DB, int MAXT, int MINB, bool COST = false>
static cudaError_t launch_sweep_cfg(const SweepLaunch& a_in, int T, size_t smem, int num_sms, cudaStream_t st) {
    auto kern = qgt_sweep_kernel<R, B, MMA_ONLY, DB, MAXT, MINB, COST>;
    static int ctas_per_sm = 0, threads_seen = 0;
    static size_t smem_seen = 0;
    if (!ctas_per_sm || smem != smem_seen || T != threads_seen) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
        if (e != cudaSuccess) return e;
        int occ = 0;
        e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, T, smem);
        if (e != cudaSuccess) return e;
        ctas_per_sm = occ > 0 ? occ : 1;
        smem_seen = smem;
        threads_seen = T;
    }
    SweepLaunch a = a_in;
    const uint64_t resident = (uint64_t)num_sms * ctas_per_sm;
    // tiles per item set-up: 1 unless asked for (more tiles per unit measured no faster, and fewer units per launch
    // leave less room for the multi-wave grid below)
    int tpi = 1;
    if (a.tiles_per_item > 0)
        while (tpi * 2 <= a.tiles_per_item && a.ntiles % (uint64_t)(tpi * 2) == 0) tpi *= 2;
    a.tiles_per_item = tpi;
    const uint64_t total = (a.ntiles / (uint64_t)tpi) * (uint64_t)a.nitems;
    // grid: several waves of resident CTAs (CTAs that finish early are replaced, which evens out the tail; measured
    // best at 8-16 waves, 2 waves +3 %), as long as every CTA still gets ~8 work units to amortise its set-up
    uint64_t waves = total / (resident * 8);
    waves = waves < 2 ? 2 : (waves > 16 ? 16 : waves);
    if (COST || a.ct.num_edges > 0) waves = 2;          // cost runs build their energy table once per CTA
    const uint64_t cap = resident * waves;
    const unsigned grid = (unsigned)(total < cap ? total : cap);
    kern<<<grid, T, smem, st>>>(a);
    return cudaGetLastError();
}

template <int R, int B>
static cudaError_t launch_sweep_rb(const SweepLaunch& a, int K, int mat_count, int nsub, int has_cost, int num_sms, cudaStream_t st) {
    constexpr int N = 1 << R;
    const int T = 1 << (K - R - B);
    if (a.ntiles * (uint64_t)a.nitems == 0) return cudaSuccess;
    const size_t fixed0 = sizeof(cplx) * (size_t)mat_count + sizeof(cplx) * (size_t)(QGT_VARIANT_STRIDE(N) << QGT_MAX_VARIANT_BITS) +
                         (size_t)nsub * sizeof(QgtDevSubPass) + (has_cost ? qgt_cost_smem_doubles(K, a.ct.num_edges, a.cost_phase == nullptr) * sizeof(double) : 0);
    const size_t tile_bytes = sizeof(cplx) << K;
    const size_t fixed = fixed0;
    if (fixed + (size_t)nsub * QGT_FAST_BYTES_PER_SUB + 2 * tile_bytes > 200 * 1024 || T > 256) return cudaErrorInvalidValue;
    if (R == 3 && B == 0 && a.mma_only && T >= 32) {
        const size_t fixed = fixed0 + (size_t)nsub * QGT_FAST_BYTES_PER_SUB;
        // single tile buffer: more resident CTAs hide the load latency instead of a second buffer
        // with the global phase / energy tables the 16 KB energy table is gone from shared memory: four CTAs per SM again
        if (has_cost && a.cost_phase) return launch_sweep_cfg<3, 0, true, false, 256, 4, true>(a, T, fixed + tile_bytes, num_sms, st);
        if (has_cost) return launch_sweep_cfg<3, 0, true, false, 256, 3, true>(a, T, fixed + tile_bytes, num_sms, st);
        if (a.double_buffer) return launch_sweep_cfg<3, 0, true, true, 256, 3>(a, T, fixed + 2 * tile_bytes, num_sms, st);
        return launch_sweep_cfg<3, 0, true, false, 256, 4>(a, T, fixed + tile_bytes, num_sms, st);
    }
    if (B > 0 && T <= 128) return launch_sweep_cfg<R, B, false, true, 128, 3>(a, T, fixed + 2 * tile_bytes, num_sms, st);
    if (B > 0) return launch_sweep_cfg<R, B, false, true, 256, 1>(a, T, fixed + 2 * tile_bytes, num_sms, st);
    return launch_sweep_cfg<R, B, false, true, 256, 2>(a, T, fixed + 2 * tile_bytes, num_sms, st);
}

// per-launch phase tables of a run's cost ops: out[c][idx] = exp(-i angle_c ein[idx]) (sweep_core.cuh: qgt_cost_phase_entry)
__global__ void __launch_bounds__(256) qgt_cost_phase_kernel(const QgtDevRun* runs, int run_idx, QgtCostTable ct, const QgtDevCost* costs, cplx* out) {
    __shared__ QgtDevRun run;
    for (int i = threadIdx.x; i < (int)(sizeof(QgtDevRun) / 4); i += blockDim.x)
        reinterpret_cast<uint32_t*>(&run)[i] = reinterpret_cast<const uint32_t*>(&runs[run_idx])[i];
    __syncthreads();
    const uint32_t idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx < (1u << run.K)) {
        out[((size_t)blockIdx.y << run.K) + idx] = qgt_cost_phase_entry(run, ct, costs[run.cost_off + blockIdx.y].angle, idx);
        if (blockIdx.y == 0) reinterpret_cast<double*>(out + ((size_t)gridDim.y << run.K))[idx] = qgt_cost_ein(run, ct, idx);    // ein[] behind the phases
    }
}

cudaError_t launch_cost_phase_tables(const SweepLaunch& a, int K, int ncost, cplx* out, cudaStream_t st) {
    if (ncost <= 0) return cudaSuccess;
    const unsigned nb = (unsigned)(((1u << K) + 255u) / 256u);
    qgt_cost_phase_kernel<<<dim3(nb, (unsigned)ncost), 256, 0, st>>>(a.runs, a.run_idx, a.ct, a.costs, out);
    return cudaGetLastError();
}

cudaError_t launch_sweep(const SweepLaunch& a, int K, int R, int B, int mat_count, int nsub, int has_cost, int num_sms, cudaStream_t st) {
    if (K > QGT_MAX_TILE_QUBITS) return cudaErrorInvalidValue;
    switch (R * 2 + B) {
    case 2: return launch_sweep_rb<1, 0>(a, K, mat_count, nsub, has_cost, num_sms, st);
    case 3: return launch_sweep_rb<1, 1>(a, K, mat_count, nsub, has_cost, num_sms, st);
    case 4: return launch_sweep_rb<2, 0>(a, K, mat_count, nsub, has_cost, num_sms, st);
    case 5: return launch_sweep_rb<2, 1>(a, K, mat_count, nsub, has_cost, num_sms, st);
    case 6: return launch_sweep_rb<3, 0>(a, K, mat_count, nsub, has_cost, num_sms, st);
    case 7: return launch_sweep_rb<3, 1>(a, K, mat_count, nsub, has_cost, num_sms, st);
    default: return cudaErrorInvalidValue;
    }
}

// ------------------------------------------------------------------------------------------------
// Gram  C = A^H B  on the FP64 tensor pipe
// ------------------------------------------------------------------------------------------------

// one staged chunk of KC amplitudes: fragments from shared memory, 4 DMMAs per 8x8 block and k-step.
// conj(a) * b = (ar*br + ai*bi) + i (ar*bi - ai*br); the two updates of one accumulator are issued a whole
// block sweep apart so that dependent DMMAs never sit back to back.
template <int WM, int WN, int BM, int BN, int KC, bool ALL>
__device__ __forceinline__ void gram_chunk(const cplx* __restrict__ sA, const cplx* __restrict__ sB, int wm, int wn, int fr, int fk,
                                           const bool (&blk)[BM][BN], double (&cre)[BM][BN][2], double (&cim)[BM][BN][2]) {
    constexpr int S = KC + 4;
#pragma unroll
    for (int kk = 0; kk < KC; kk += 4) {
        cplx a[BM], b[BN];
#pragma unroll
        for (int i = 0; i < BM; ++i) a[i] = sA[((wm * BM + i) * 8 + fr) * S + kk + fk];
#pragma unroll
        for (int j = 0; j < BN; ++j) b[j] = sB[((wn * BN + j) * 8 + fr) * S + kk + fk];
#pragma unroll
        for (int i = 0; i < BM; ++i) {
#pragma unroll
            for (int j = 0; j < BN; ++j) {
                if (!ALL && !blk[i][j]) continue;
                dmma884(cre[i][j][0], cre[i][j][1], a[i].x, b[j].x);
                dmma884(cim[i][j][0], cim[i][j][1], a[i].x, b[j].y);
            }
        }
#pragma unroll
        for (int i = 0; i < BM; ++i) {
            const double nai = -a[i].y;
#pragma unroll
            for (int j = 0; j < BN; ++j) {
                if (!ALL && !blk[i][j]) continue;
                dmma884(cre[i][j][0], cre[i][j][1], a[i].y, b[j].y);
                dmma884(cim[i][j][0], cim[i][j][1], nai, b[j].x);
            }
        }
    }
}

// CTA tile: MT = WM*BM*8 rows (columns of A) x NT = WN*BN*8 cols (columns of B); WM*WN warps, each warp
// owns BM x BN blocks of 8x8.  The 2^n-long amplitude axis is consumed in chunks of KC amplitudes moved
// global -> shared with 16-byte cp.async through a STAGES-deep ring (no register staging).  Shared memory
// keeps the complex numbers interleaved with a row stride of KC+4 elements: one 128-bit load then
// delivers (re, im) of a fragment element and a quarter-warp touches 8 distinct 16-byte bank groups.
template <int WM, int WN, int BM, int BN, int KC, int STAGES>
__global__ void __launch_bounds__(WM * WN * 32, (WM == 2 && WN == 4 && BM == 2 && BN == 1) ? 3 : 2) qgt_gram_kernel(GramLaunch g) {
    constexpr int MT = WM * BM * 8, NT = WN * BN * 8, S = KC + 4, NTHR = WM * WN * 32;
    constexpr int ELEMS = (MT + NT) * KC;
    static_assert(ELEMS % NTHR == 0 && NTHR % KC == 0, "tile/threads mismatch");
    constexpr int PER = ELEMS / NTHR;
    constexpr bool STRIP = (WM == 2 && WN == 4 && BM == 2 && BN == 1);     // the 32x32 shape carries the strip
    constexpr int STAGE_ELEMS = (MT + NT + (STRIP ? 8 : 0)) * S;
    extern __shared__ __align__(16) unsigned char qgt_gram_smem[];
    cplx* sm = reinterpret_cast<cplx*>(qgt_gram_smem);

    const int tiles = g.mtiles * g.ntiles;
    const int ks = blockIdx.x / tiles;
    const int mt = (blockIdx.x % tiles) / g.ntiles;
    const int nt = (blockIdx.x % tiles) % g.ntiles;
    if (g.symmetric && (nt + 1) * NT <= mt * MT) return;   // whole tile below the diagonal: mirrored later

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    int wm = warp / WN, wn = warp % WN;
    // diagonal tile of a symmetric launch: warps (1,0) and (1,1) own only blocks below the diagonal; they take
    // the strip instead (block rows 0-1 / 2-3 of this tile times the 8 strip columns staged behind B)
    const bool strip_tile = STRIP && g.nstrip > 0 && mt == nt;
    const bool strip_warp = strip_tile && wm == 1 && wn < 2;
    if (strip_warp) { wm = wn; wn = WN; }
    uint64_t per = (g.D + (uint64_t)g.ksplit - 1) / (uint64_t)g.ksplit;
    per = (per + KC - 1) / KC * KC;
    const uint64_t k0 = (uint64_t)ks * per;
    const uint64_t k1 = (k0 + per < g.D) ? k0 + per : g.D;

    // this thread's share of a chunk: element e -> (column e / KC, offset e % KC)
    const cplx* colptr[PER];
    int soff[PER];
    const int koff = tid % KC;
#pragma unroll
    for (int j = 0; j < PER; ++j) {
        const int col = (tid + j * NTHR) / KC;
        const int gc = col < MT ? mt * MT + col : nt * NT + (col - MT);
        const bool ok = col < MT ? gc < g.na : gc < g.nb;
        colptr[j] = ok ? (col < MT ? g.a_ptrs[gc] : g.b_ptrs[gc]) + koff : nullptr;
        soff[j] = col * S + koff;
    }
    const cplx* strip_ptr = nullptr;                       // threads 0 .. 8*KC-1 stage the strip columns
    if (strip_tile && tid < 8 * KC && tid / KC < g.nstrip) strip_ptr = g.b_ptrs[g.nb_main + tid / KC] + koff;
    auto issue = [&](uint64_t kb, int stage) {
        cplx* dst = sm + (size_t)stage * STAGE_ELEMS;
        const bool in = kb + (uint64_t)koff < k1;
#pragma unroll
        for (int j = 0; j < PER; ++j) {
            const bool ok = in && colptr[j] != nullptr;
            cp_async16(dst + soff[j], ok ? (const void*)(colptr[j] + kb) : (const void*)g.a_ptrs, ok ? 16 : 0);
        }
        if (STRIP && strip_tile && tid < 8 * KC) {
            const bool ok = in && strip_ptr != nullptr;
            cp_async16(dst + (MT + NT + tid / KC) * S + koff, ok ? (const void*)(strip_ptr + kb) : (const void*)g.a_ptrs, ok ? 16 : 0);
        }
    };

    double cre[BM][BN][2], cim[BM][BN][2];
#pragma unroll
    for (int i = 0; i < BM; ++i)
#pragma unroll
        for (int j = 0; j < BN; ++j) { cre[i][j][0] = cre[i][j][1] = cim[i][j][0] = cim[i][j][1] = 0.0; }
    // blocks entirely in the padding, or (symmetric case) entirely below the diagonal, do no tensor work
    bool blk[BM][BN];
#pragma unroll
    for (int i = 0; i < BM; ++i)
#pragma unroll
        for (int j = 0; j < BN; ++j) {
            const int r0 = mt * MT + (wm * BM + i) * 8, c0 = nt * NT + (wn * BN + j) * 8;
            blk[i][j] = strip_warp ? r0 < g.na : (r0 < g.na && c0 < g.nb_main && !(g.symmetric && c0 + 8 <= r0));
        }
    bool all_valid = true;      // warp-uniform: the common case runs without a predicate on every mma.sync
#pragma unroll
    for (int i = 0; i < BM; ++i)
#pragma unroll
        for (int j = 0; j < BN; ++j) all_valid = all_valid && blk[i][j];

    const uint64_t nchunks = k1 > k0 ? (k1 - k0 + KC - 1) / KC : 0;
#pragma unroll
    for (int s = 0; s < STAGES - 1; ++s) {
        if ((uint64_t)s < nchunks) issue(k0 + (uint64_t)s * KC, s);
        cp_async_commit();
    }
    const int fr = lane >> 2, fk = lane & 3;
    for (uint64_t ch = 0; ch < nchunks; ++ch) {
        cp_async_wait<STAGES - 2>();
        __syncthreads();            // chunk `ch` has landed for every thread; stage (ch-1)%STAGES is free again
        {
            const uint64_t nx = ch + STAGES - 1;
            if (nx < nchunks) issue(k0 + nx * KC, (int)(nx % STAGES));
            cp_async_commit();
        }
        const cplx* sA = sm + (size_t)(ch % STAGES) * STAGE_ELEMS;
        const cplx* sB = sA + MT * S;
        if (all_valid) gram_chunk<WM, WN, BM, BN, KC, true>(sA, sB, wm, wn, fr, fk, blk, cre, cim);
        else gram_chunk<WM, WN, BM, BN, KC, false>(sA, sB, wm, wn, fr, fk, blk, cre, cim);
    }
    cp_async_wait<0>();
    const int Npad = g.npad;
    const int Mpad = g.mtiles * MT;
    cplx* out = g.partial + (size_t)ks * Mpad * Npad;
#pragma unroll
    for (int i = 0; i < BM; ++i)
#pragma unroll
        for (int j = 0; j < BN; ++j) {
            const int row = mt * MT + (wm * BM + i) * 8 + (lane >> 2);
            const int col = (strip_warp ? g.ntiles * NT : nt * NT + (wn * BN + j) * 8) + (lane & 3) * 2;
            cplx z0, z1;
            z0.x = cre[i][j][0]; z0.y = cim[i][j][0];
            z1.x = cre[i][j][1]; z1.y = cim[i][j][1];
            out[(size_t)row * Npad + col] = z0;
            out[(size_t)row * Npad + col + 1] = z1;
        }
}

// Thin Gram: a handful of column pairs (the blocked schedule at 30 qubits contracts 5 resident with 2 streaming
// columns at a time).  The tensor-pipe kernel would stage mostly padding and keep only a few KB per SM in
// flight; here every thread streams amplitudes with coalesced 128-bit loads of all NA + NB columns and keeps the
// NA x NB complex accumulators in registers, so the kernel runs at HBM speed.  Same partial layout and second
// stage as the tensor-pipe kernel (one "k-split" per CTA).
template <int NA, int NB>
__global__ void __launch_bounds__(256) qgt_gram_thin_kernel(GramLaunch g) {
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const uint64_t per = (g.D + (uint64_t)g.ksplit - 1) / (uint64_t)g.ksplit;
    const uint64_t k0 = (uint64_t)blockIdx.x * per;
    const uint64_t k1 = (k0 + per < g.D) ? k0 + per : g.D;
    const cplx* ap[NA];
    const cplx* bp[NB];
#pragma unroll
    for (int j = 0; j < NA; ++j) ap[j] = j < g.na ? g.a_ptrs[j] : nullptr;
#pragma unroll
    for (int l = 0; l < NB; ++l) bp[l] = l < g.nb ? g.b_ptrs[l] : nullptr;
    double cr[NA][NB], ci[NA][NB];
#pragma unroll
    for (int j = 0; j < NA; ++j)
#pragma unroll
        for (int l = 0; l < NB; ++l) { cr[j][l] = 0.0; ci[j][l] = 0.0; }
    for (uint64_t i = k0 + (uint64_t)tid; i < k1; i += 256) {
        cplx a[NA], b[NB];
#pragma unroll
        for (int j = 0; j < NA; ++j) { if (ap[j]) a[j] = ap[j][i]; else { a[j].x = 0.0; a[j].y = 0.0; } }
#pragma unroll
        for (int l = 0; l < NB; ++l) { if (bp[l]) b[l] = bp[l][i]; else { b[l].x = 0.0; b[l].y = 0.0; } }
#pragma unroll
        for (int j = 0; j < NA; ++j)
#pragma unroll
            for (int l = 0; l < NB; ++l) {      // conj(a) * b
                cr[j][l] = fma(a[j].x, b[l].x, cr[j][l]); cr[j][l] = fma(a[j].y, b[l].y, cr[j][l]);
                ci[j][l] = fma(a[j].x, b[l].y, ci[j][l]); ci[j][l] = fma(-a[j].y, b[l].x, ci[j][l]);
            }
    }
    __shared__ double red[8][NA * NB * 2];
#pragma unroll
    for (int j = 0; j < NA; ++j)
#pragma unroll
        for (int l = 0; l < NB; ++l) {
            double xr = cr[j][l], xi = ci[j][l];
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) { xr += __shfl_down_sync(0xffffffffu, xr, o); xi += __shfl_down_sync(0xffffffffu, xi, o); }
            if (lane == 0) { red[warp][(j * NB + l) * 2] = xr; red[warp][(j * NB + l) * 2 + 1] = xi; }
        }
    __syncthreads();
    if (tid < NA * NB) {
        double xr = 0.0, xi = 0.0;
        for (int w = 0; w < 8; ++w) { xr += red[w][tid * 2]; xi += red[w][tid * 2 + 1]; }      // fixed order: deterministic
        cplx z; z.x = xr; z.y = xi;
        g.partial[((size_t)blockIdx.x * NA + tid / NB) * NB + tid % NB] = z;
    }
}

__global__ void qgt_gram_reduce_kernel(const cplx* partial, int ksplit, int Mpad, int Npad, int na, int nb, int nb_main, int strip_col0,
                                       const int* a_ids, const int* b_ids, cplx* C, int ldc, int symmetric) {
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= na * nb) return;
    const int i = idx / nb, j = idx % nb;
    if (symmetric && j < i) return;
    const int pj = j < nb_main ? j : strip_col0 + (j - nb_main);      // strip columns sit behind the tile grid
    double sr = 0.0, si = 0.0;
    for (int ks = 0; ks < ksplit; ++ks) {           // fixed order: deterministic
        const cplx z = partial[((size_t)ks * Mpad + i) * Npad + pj];
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

// tile shape with the least padded work: 64x64 for big blocks, 32x32 when that wastes less, 32x16 for the
// few-column Grams of the blocked schedule
static int g_gram_tile_override = 0;
void set_gram_tile_override(int t) { g_gram_tile_override = t; }

GramShape gram_shape(int na, int nb) {
    GramShape s;
    s.thin = 0;
    if (g_gram_tile_override == 0) {
        const int thin_shapes[5][2] = {{8, 2}, {4, 4}, {2, 8}, {16, 1}, {1, 16}};
        for (const auto& t : thin_shapes)
            if (na <= t[0] && nb <= t[1]) { s.MT = t[0]; s.NT = t[1]; s.thin = 1; return s; }
    }
    if (g_gram_tile_override == 64) { s.MT = 64; s.NT = 64; return s; }
    if (g_gram_tile_override == 32) { s.MT = 32; s.NT = 32; return s; }
    if (na <= 32 && nb <= 16) { s.MT = 32; s.NT = 16; return s; }
    // 32x32 tiles measured faster than 64x64 on B200 at every size tried (less padding, more CTAs per SM)
    s.MT = 32; s.NT = 32;
    return s;
}

size_t gram_configure(GramLaunch& g, GramShape shp) {
    g.nstrip = 0;
    g.nb_main = g.nb;
    if (shp.thin) { g.mtiles = g.ntiles = 1; g.npad = shp.NT; return (size_t)shp.MT * shp.NT; }
    const int extra = g.nb - g.na;
    if (g.symmetric && shp.MT == 32 && shp.NT == 32 && extra >= 1 && extra <= 8) { g.nstrip = extra; g.nb_main = g.na; }
    g.mtiles = (g.na + shp.MT - 1) / shp.MT;
    g.ntiles = (g.nb_main + shp.NT - 1) / shp.NT;
    g.npad = g.ntiles * shp.NT + (g.nstrip ? 8 : 0);
    return (size_t)g.mtiles * shp.MT * g.npad;
}

template <int WM, int WN, int BM, int BN, int KC, int STAGES>
static cudaError_t launch_gram_t(const GramLaunch& g, cudaStream_t st) {
    constexpr int MT = WM * BM * 8, NT = WN * BN * 8;
    constexpr bool STRIP = (WM == 2 && WN == 4 && BM == 2 && BN == 1);
    constexpr size_t smem = (size_t)STAGES * (MT + NT + (STRIP ? 8 : 0)) * (KC + 4) * sizeof(cplx);
    auto kern = qgt_gram_kernel<WM, WN, BM, BN, KC, STAGES>;
    static bool attr_set = false;
    if (!attr_set) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        attr_set = true;
    }
    const unsigned grid = (unsigned)(g.mtiles * g.ntiles * g.ksplit);
    kern<<<grid, WM * WN * 32, smem, st>>>(g);
    return cudaGetLastError();
}

template <int WM, int WN, int BM, int BN, int KC, int STAGES>
static int gram_occupancy_t() {
    constexpr int MT = WM * BM * 8, NT = WN * BN * 8;
    constexpr bool STRIP = (WM == 2 && WN == 4 && BM == 2 && BN == 1);
    constexpr size_t smem = (size_t)STAGES * (MT + NT + (STRIP ? 8 : 0)) * (KC + 4) * sizeof(cplx);
    static int occ = 0;
    if (!occ) {
        auto kern = qgt_gram_kernel<WM, WN, BM, BN, KC, STAGES>;
        cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        int o = 0;
        if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&o, kern, WM * WN * 32, smem) != cudaSuccess || o < 1) o = 1;
        occ = o;
    }
    return occ;
}

// Split-K factor.  The CTAs that do work (tiles on or above the diagonal of a symmetric launch) times the split
// should fill a whole number of waves of resident CTAs: ~18 CTAs per SM in total when there is enough work
// (fine-grained work evens out diagonal / padded tiles), at most 512 splits, at least 256 amplitudes per CTA.  A
// one-tile streaming Gram with 512 splits on 444 resident CTAs ran 1.15 waves in the time of 2.
int gram_choose_ksplit(const GramLaunch& g, GramShape shp, int num_sms) {
    int occ;
    if (shp.thin) {
        static int thin_occ[5] = {0, 0, 0, 0, 0};
        const int which = shp.MT == 8 ? 0 : shp.MT == 4 ? 1 : shp.MT == 2 ? 2 : shp.MT == 16 ? 3 : 4;
        if (!thin_occ[which]) {
            int o = 0;
            cudaError_t e = which == 0 ? cudaOccupancyMaxActiveBlocksPerMultiprocessor(&o, qgt_gram_thin_kernel<8, 2>, 256, 0)
                          : which == 1 ? cudaOccupancyMaxActiveBlocksPerMultiprocessor(&o, qgt_gram_thin_kernel<4, 4>, 256, 0)
                          : which == 2 ? cudaOccupancyMaxActiveBlocksPerMultiprocessor(&o, qgt_gram_thin_kernel<2, 8>, 256, 0)
                          : which == 3 ? cudaOccupancyMaxActiveBlocksPerMultiprocessor(&o, qgt_gram_thin_kernel<16, 1>, 256, 0)
                                       : cudaOccupancyMaxActiveBlocksPerMultiprocessor(&o, qgt_gram_thin_kernel<1, 16>, 256, 0);
            thin_occ[which] = (e == cudaSuccess && o > 0) ? o : 1;
        }
        occ = thin_occ[which];
    }
    else if (shp.MT == 64) occ = gram_occupancy_t<2, 4, 4, 2, 8, 4>();
    else if (shp.NT == 32) occ = gram_occupancy_t<2, 4, 2, 1, 16, 3>();
    else occ = gram_occupancy_t<4, 2, 1, 1, 16, 4>();
    int active = 0;
    for (int mt = 0; mt < g.mtiles; ++mt)
        for (int nt = 0; nt < g.ntiles; ++nt)
            if (!(g.symmetric && (nt + 1) * shp.NT <= mt * shp.MT)) active++;
    if (active < 1) active = 1;
    const uint64_t slots = (uint64_t)num_sms * occ;
    const uint64_t by_len = g.D / 256 > 1 ? g.D / 256 : 1;
    const uint64_t cap = shp.thin ? 1024 : 512;
    int waves = shp.thin ? 1 : (18 + occ - 1) / occ;      // thin: pure streaming, one balanced wave
    uint64_t ks = 1;
    for (; waves >= 1; --waves) {
        ks = (uint64_t)waves * slots / (uint64_t)active;
        if (ks <= cap) break;
    }
    if (ks > cap) ks = cap;
    if (ks > by_len) ks = by_len;
    return (int)(ks < 1 ? 1 : ks);
}

cudaError_t launch_gram(const GramLaunch& g, GramShape shp, cudaStream_t st) {
    if (g.mtiles * g.ntiles * g.ksplit == 0) return cudaSuccess;
    if (shp.thin) {
        if (shp.MT == 8) qgt_gram_thin_kernel<8, 2><<<g.ksplit, 256, 0, st>>>(g);
        else if (shp.MT == 4) qgt_gram_thin_kernel<4, 4><<<g.ksplit, 256, 0, st>>>(g);
        else if (shp.MT == 2) qgt_gram_thin_kernel<2, 8><<<g.ksplit, 256, 0, st>>>(g);
        else if (shp.MT == 16) qgt_gram_thin_kernel<16, 1><<<g.ksplit, 256, 0, st>>>(g);
        else qgt_gram_thin_kernel<1, 16><<<g.ksplit, 256, 0, st>>>(g);
        return cudaGetLastError();
    }
    if (shp.MT == 64) return launch_gram_t<2, 4, 4, 2, 8, 4>(g, st);
    if (shp.NT == 32) return launch_gram_t<2, 4, 2, 1, 16, 3>(g, st);
    // deeper rings (6 stages) and 32-amplitude chunks were measured on the HBM-bound streaming Grams: no gain
    return launch_gram_t<4, 2, 1, 1, 16, 4>(g, st);
}

cudaError_t launch_gram_reduce(const GramLaunch& g, GramShape shp, const int* a_ids, const int* b_ids,
                               cplx* C, int ldc, cudaStream_t st) {
    const int total = g.na * g.nb;
    if (total == 0) return cudaSuccess;
    qgt_gram_reduce_kernel<<<(total + 255) / 256, 256, 0, st>>>(g.partial, g.ksplit, g.mtiles * shp.MT, g.npad, g.na, g.nb, g.nb_main,
                                                             g.ntiles * shp.NT, a_ids, b_ids, C, ldc, g.symmetric);
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

// dst_i = E_z(goff + i) * src_i: the diagonal cost observable applied to a state (dst may equal src)
__global__ void __launch_bounds__(256) qgt_cost_apply_kernel(cplx* dst, const cplx* src, uint64_t D, QgtCostTable ct, uint64_t goff) {
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < D; i += stride) {
        const double e = qgt_cost_energy(ct, goff + i);
        cplx z = src[i];
        z.x *= e; z.y *= e;
        dst[i] = z;
    }
}

cudaError_t launch_cost_apply(cplx* dst, const cplx* src, uint64_t D, QgtCostTable ct, uint64_t goff, cudaStream_t st) {
    const uint64_t want = (D + 255) / 256;
    const unsigned grid = (unsigned)(want < 148 * 16 ? want : 148 * 16);
    qgt_cost_apply_kernel<<<grid, 256, 0, st>>>(dst, src, D, ct, goff);
    return cudaGetLastError();
}

cudaError_t launch_cost_dot(const cplx* a, const cplx* b, uint64_t D, QgtCostTable ct, uint64_t goff, double* out2, cudaStream_t st) {
    cudaError_t e = cudaMemsetAsync(out2, 0, 2 * sizeof(double), st);
    if (e != cudaSuccess) return e;
    const uint64_t want = (D + 255) / 256;
    const unsigned grid = (unsigned)(want < 148 * 8 ? want : 148 * 8);
    qgt_cost_dot_kernel<<<grid, 256, 0, st>>>(a, b, D, ct, goff, out2);
    return cudaGetLastError();
}

// ------------------------------------------------------------------------------------------------
// roofline denominators measured in place (MEASURED_PEAKS.json carries no FP64 tensor figure)
// ------------------------------------------------------------------------------------------------
// four independent DMMA accumulator chains per warp, each with its own operand registers: the shape that reaches the
// pipe's peak on B200 (tools/peaks.cu, tools/dmma_ilp.cu)
__global__ void __launch_bounds__(1024) qgt_peak_dmma_kernel(double* out, int iters) {
    double a[4], b[4], c[4][2];
#pragma unroll
    for (int i = 0; i < 4; i++) { a[i] = 1.0 + (threadIdx.x + i) * 1e-9; b[i] = 1.0 - (threadIdx.x + 3 * i) * 1e-9; c[i][0] = 0; c[i][1] = 0; }
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int r = 0; r < 4; r++)
#pragma unroll
            for (int i = 0; i < 4; i++)
                asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                             : "+d"(c[i][0]), "+d"(c[i][1]) : "d"(a[i]), "d"(b[i]));
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < 4; i++) s += c[i][0] + c[i][1];
    out[(size_t)blockIdx.x * blockDim.x + threadIdx.x] = s;
}

cudaError_t measure_dmma_peak(int num_sms, double* scratch /* num_sms * 1024 doubles */, cudaStream_t st, double* tflops) {
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    const int iters = 8192;
    float best = 1e30f;
    cudaError_t e = cudaSuccess;
    for (int r = 0; r < 6 && e == cudaSuccess; r++) {
        cudaEventRecord(e0, st);
        qgt_peak_dmma_kernel<<<num_sms, 1024, 0, st>>>(scratch, iters);
        cudaEventRecord(e1, st);
        e = cudaEventSynchronize(e1);
        float ms = 0.f;
        cudaEventElapsedTime(&ms, e0, e1);
        if (r > 0 && ms < best) best = ms;
    }
    cudaEventDestroy(e0); cudaEventDestroy(e1);
    if (e == cudaSuccess) e = cudaGetLastError();
    *tflops = (double)num_sms * 32 * iters * 16 * 512.0 / best * 1e-9;      // m8n8k4 = 2*8*8*4 flop
    return e;
}

}  // namespace qgt

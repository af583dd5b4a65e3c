// fused.cu — fused gate sweep + transition-matrix kernel of the QGT hot path (sm_100a).
//
//   qgt_fused_kernel     advances one derivative column lambda AND the marching state phi through a run in
//                        lockstep (two 2^K-amplitude tiles on chip, every dense stage on the FP64 tensor pipe as in
//                        the plain sweep kernel) and, after every stage in which a parameter occurs, accumulates the
//                        8x8 transition matrix over the stage's three matrix qubits
//                            rho[c][a] = sum_rest phi'[c, rest] * conj(lambda'[a, rest])
//                        with DMMAs whose operands are the C fragments the stage just produced (no extra
//                        shared-memory traffic).  <lambda| G |phi> for every generator G of that stage is then a
//                        64-term contraction (qgt_rho_contract_kernel), so the overlaps <d_mu psi|d_nu psi> are
//                        obtained without ever writing the later column or re-reading the earlier ones in a Gram pass.
//   qgt_rho_reduce_kernel   deterministic sum of the per-CTA partial transition matrices.
//   qgt_rho_contract_kernel A[mu][nu] += sum Gt_nu[a][c] rho_mu[c][a].
//
// Replaces the three 2^n-long dot loops per tensor element of compute_quantum_geometric_tensor (reference
// src/quantum_geometric/core/quantum_geometric_tensor_network.c:1127-1175) and the P^2 D triple loop of
// diffgeo_compute_fubini_study / _berry_curvature (distributed/differential_geometry.c:2819-2906) whenever the
// derivative columns do not all fit in HBM (and, by option, when they do).
#include <type_traits>

#include "kernels.cuh"
#include "mma_common.cuh"

namespace qgt {

// mma.sync as a VOLATILE asm: without it the compiler treats the instruction as a pure function and if-converts
// "diagonal-real form ? 4 DMMAs : 8 DMMAs" into 8 unconditional ones (measured: twice the tensor work on every
// diagonal-real stage).
__device__ __forceinline__ void dmma884v(double& c0, double& c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                 : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}
#define dmma884 dmma884v

// A fragments of one 8x8 stage matrix for this lane (row lane>>2, columns lane&3 and 4 + lane&3)
struct StageFrag {
    double m0x, m0y, m1x, m1y;   // dense: complex elements; diagonal-real form: real parts only ...
    double dx, dy;               // ... and the row's phase
    bool diag_real;
};

__device__ __forceinline__ StageFrag qgt_load_frag(const cplx* M, bool diag_real, int lane) {
    StageFrag f;
    f.diag_real = diag_real;
    if (diag_real) {
        f.m0x = reinterpret_cast<const double*>(M)[lane];
        f.m1x = reinterpret_cast<const double*>(M)[32 + lane];
        f.m0y = 0.0; f.m1y = 0.0;
        const cplx d = M[64 + (lane >> 2)];
        f.dx = d.x; f.dy = d.y;
    } else {
        const cplx m0 = M[lane], m1 = M[32 + lane];
        f.m0x = m0.x; f.m0y = m0.y; f.m1x = m1.x; f.m1y = m1.y;
        f.dx = 1.0; f.dy = 0.0;
    }
    return f;
}

// 8 vectors through one stage: B operand (v0 = amplitudes k, v1 = amplitudes 4+k of vector lane>>2) -> C fragment
// (o0, o1 = row lane>>2 of vectors 2(lane&3), 2(lane&3)+1)
__device__ __forceinline__ void qgt_apply8(const StageFrag& f, const cplx& v0, const cplx& v1, cplx& o0, cplx& o1) {
    double cr0 = 0.0, cr1 = 0.0, ci0 = 0.0, ci1 = 0.0;
    dmma884(cr0, cr1, f.m0x, v0.x);
    dmma884(ci0, ci1, f.m0x, v0.y);
    dmma884(cr0, cr1, f.m1x, v1.x);
    dmma884(ci0, ci1, f.m1x, v1.y);
    if (f.diag_real) {
        o0.x = f.dx * cr0 - f.dy * ci0; o0.y = f.dx * ci0 + f.dy * cr0;
        o1.x = f.dx * cr1 - f.dy * ci1; o1.y = f.dx * ci1 + f.dy * cr1;
    } else {
        dmma884(cr0, cr1, -f.m0y, v0.y);
        dmma884(ci0, ci1, f.m0y, v0.x);
        dmma884(cr0, cr1, -f.m1y, v1.y);
        dmma884(ci0, ci1, f.m1y, v1.x);
        o0.x = cr0; o0.y = ci0; o1.x = cr1; o1.y = ci1;
    }
}

// rho[c][a] += sum over the 4 vectors of this k-step of p[c] * conj(l[a]).  Both operands are C fragments of the
// stage (lane = (row, vector)), which is exactly the A / B operand layout of the next DMMA: p is the A operand
// (row c = lane>>2), l the B operand (column a = lane>>2), the vector index lane&3 is summed over.  3M complex
// product: T1 = pr lr, T2 = pi li, T3 = (pr + pi)(lr - li);  Re = T1 + T2, Im = T3 - T1 + T2  (three independent
// accumulator chains, 6 DMMAs + 4 additions per 8 vectors instead of 8 DMMAs).
__device__ __forceinline__ void qgt_rho3(double (&t)[6], const cplx& p, const cplx& l) {
    dmma884(t[0], t[1], p.x, l.x);
    dmma884(t[2], t[3], p.y, l.y);
    dmma884(t[4], t[5], p.x + p.y, l.x - l.y);
}

// the same with the matrix form fixed at compile time: callers branch ONCE per stage on the (warp-uniform) form, so the
// four extra DMMAs of the dense form sit behind a real branch instead of being predicated off one by one (predicated-off
// DMMAs still travel through the tensor pipe: measured 2x pipe-busy time on diagonal-real stages)
template <bool DR>
__device__ __forceinline__ void qgt_apply8t(const StageFrag& f, const cplx& v0, const cplx& v1, cplx& o0, cplx& o1) {
    double cr0 = 0.0, cr1 = 0.0, ci0 = 0.0, ci1 = 0.0;
    dmma884(cr0, cr1, f.m0x, v0.x);
    dmma884(ci0, ci1, f.m0x, v0.y);
    dmma884(cr0, cr1, f.m1x, v1.x);
    dmma884(ci0, ci1, f.m1x, v1.y);
    if (DR) {
        o0.x = f.dx * cr0 - f.dy * ci0; o0.y = f.dx * ci0 + f.dy * cr0;
        o1.x = f.dx * cr1 - f.dy * ci1; o1.y = f.dx * ci1 + f.dy * cr1;
    } else {
        dmma884(cr0, cr1, -f.m0y, v0.y);
        dmma884(ci0, ci1, f.m0y, v0.x);
        dmma884(cr0, cr1, -f.m1y, v1.y);
        dmma884(ci0, ci1, f.m1y, v1.x);
        o0.x = cr0; o0.y = ci0; o1.x = cr1; o1.y = ci1;
    }
}

__device__ __forceinline__ cplx qgt_cmul(const cplx& a, const cplx& b) {
    cplx o; o.x = a.x * b.x - a.y * b.y; o.y = a.x * b.y + a.y * b.x; return o;
}

// ---- mbarrier + bulk-async copy (the trajectory tiles of phi arrive as one 2^K * 16-byte cp.async.bulk) ----------
__device__ __forceinline__ unsigned qgt_smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void qgt_mbar_init(uint64_t* bar, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(qgt_smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void qgt_mbar_expect_tx(uint64_t* bar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(qgt_smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void qgt_bulk_g2s(void* dst, const void* src, unsigned bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 :: "r"(qgt_smem_u32(dst)), "l"(src), "r"(bytes), "r"(qgt_smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void qgt_mbar_wait(uint64_t* bar, unsigned parity) {
    unsigned ok = 0, spins = 0;
    do {
        asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}"
                     : "=r"(ok) : "r"(qgt_smem_u32(bar)), "r"(parity) : "memory");
        if (!ok && ++spins > (1u << 24)) __trap();      // a lost copy must not hang the device
    } while (!ok);
}

// pending phases of the sub-pass's thread diagonals for the two result slots of group g (C-fragment layout)
__device__ __forceinline__ void qgt_pend_pair(const QgtDevSubPass& sp, const QgtDevThrDiag* tdiags, uint64_t gwarp, int g, int k,
                                              cplx& p0, cplx& p1) {
    p0.x = 1.0; p0.y = 0.0; p1 = p0;
    uint64_t gc0 = gwarp;
    const int t0 = g * 8 + 2 * k;
#pragma unroll
    for (int i = 0; i < 5; ++i) if ((t0 >> i) & 1) gc0 |= sp.g_thr[i];
    const uint64_t gc1 = gc0 | sp.g_thr[0];
    for (int t = sp.tdiag_begin; t < sp.tdiag_end; ++t) {
        qgt_thread_diag(p0, tdiags[t], gc0);
        qgt_thread_diag(p1, tdiags[t], gc1);
    }
}

// number of dense stages of a run = stage_end of its last sub-pass (stage indices are run-relative)
__device__ __forceinline__ int subs_stage_count(const QgtDevSubPass* subs, int nsub) { return subs[nsub - 1].stage_end; }

#define QGT_FDBG_NO_RHO 1      // timing experiments (option fused_debug; results are wrong by construction)
#define QGT_FDBG_NO_B 2
#define QGT_FDBG_NO_REDUCE 4
#define QGT_FDBG_NO_GLOBAL 8
#define QGT_FDBG_NO_BCOPY 16   // trajectory mode: no bulk copies / waits (rho is taken against stale shared memory)

// Shared memory: [tile A: the column][tile B: phi][matrix pool of the run][override matrices of the CTA's item]
//                [sub-pass descriptors][lookup tables][rho accumulators: rho_blocks x 128 doubles]
//                [scratch: 2 x warps x 128 doubles][block index of each warp's partial: 2 x 8 ints]
// A CTA owns ONE item and a contiguous chunk of tiles (blockIdx = chunk * nitems + item, so CTAs that run together
// touch the same phi tiles and find them in L2); its transition matrices stay in shared memory until the end.
//
// TRAJ = false: phi's tile is loaded next to the column's and advanced through the same stages (recomputed per item).
// TRAJ = true:  a separate launch of the self item (phi alone) has written phi's tile after every transition-matrix
//               stage as a contiguous image of the swizzled tile ("trajectory"); the other items fetch those images
//               with one cp.async.bulk per stage (mbarrier completion) and spend no tensor work on phi.
template <bool TRAJ>
__global__ void __launch_bounds__(256, 2) qgt_fused_kernel(FusedLaunch a) {
    constexpr int N = 8;
    constexpr int OVR_ELEMS = QGT_VARIANT_STRIDE(N) << QGT_MAX_VARIANT_BITS;
    extern __shared__ __align__(128) unsigned char qgt_fsmem_raw[];
    __shared__ QgtDevRun run;
    __shared__ __align__(8) uint64_t bar_b;
    const int tid = threadIdx.x, T = blockDim.x, lane = tid & 31, warp = tid >> 5, nwarps = T >> 5;
    for (int i = tid; i < (int)(sizeof(QgtDevRun) / 4); i += T)
        reinterpret_cast<uint32_t*>(&run)[i] = reinterpret_cast<const uint32_t*>(&a.runs[a.run_idx])[i];
    if (TRAJ && tid == 0) qgt_mbar_init(&bar_b, 1);
    __syncthreads();
    const int item_idx = (int)(blockIdx.x % (unsigned)a.nitems);
    const uint64_t chunk = blockIdx.x / (unsigned)a.nitems;
    const QgtSweepItem& it = a.items[item_idx];
    const bool self = it.self != 0;
    const int rho_from = it.rho_from;
    const int dbg = a.debug;
    // phi is only needed while transition matrices remain to be taken
    const bool pair = !TRAJ && !self && it.phi_dst != nullptr;       // phi is advanced through the whole run and written back
    const bool use_b = !self && (pair || rho_from <= run.last_rho_stage) && !(dbg & QGT_FDBG_NO_B);
    const size_t tile_elems = (size_t)1 << run.K;

    cplx* tileA = reinterpret_cast<cplx*>(qgt_fsmem_raw);
    cplx* tileB = tileA + tile_elems;
    cplx* spool = tileB + tile_elems;
    cplx* sovr = spool + run.mat_count;
    QgtDevSubPass* subs = reinterpret_cast<QgtDevSubPass*>(sovr + OVR_ELEMS);
    QgtFastSub* fast = reinterpret_cast<QgtFastSub*>(subs + run.nsub);
    QgtFastWarp* fwarp = reinterpret_cast<QgtFastWarp*>(fast + run.nsub);
    uint32_t* flane = reinterpret_cast<uint32_t*>(fwarp + 8 * run.nsub);
    double* rho_acc = reinterpret_cast<double*>((reinterpret_cast<uintptr_t>(flane + 32 * run.nsub) + 15) & ~(uintptr_t)15);
    double* scratch = rho_acc + (size_t)run.rho_blocks * 128;
    int* wblk = reinterpret_cast<int*>(scratch + (size_t)2 * nwarps * 128);
    QgtDevStage* stages = reinterpret_cast<QgtDevStage*>(wblk + 16);          // the run's stage descriptors, staged once
    const int nstages_run = run.nsub > 0 ? subs_stage_count(a.subs + run.sub_off, run.nsub) : 0;
    const QgtDevThrDiag* tdiags = a.tdiags + run.tdiag_off;
    {
        const uint32_t* gst = reinterpret_cast<const uint32_t*>(a.stages + run.stage_off);
        for (int i = tid; i < nstages_run * (int)(sizeof(QgtDevStage) / 4); i += T) reinterpret_cast<uint32_t*>(stages)[i] = gst[i];
    }
    __syncthreads();
    {
        const cplx* gpool = a.pool + run.mat_off;
        for (int i = tid; i < run.mat_count; i += T) spool[i] = gpool[i];
        const uint32_t* gs = reinterpret_cast<const uint32_t*>(a.subs + run.sub_off);
        uint32_t* ss = reinterpret_cast<uint32_t*>(subs);
        for (int i = tid; i < run.nsub * (int)(sizeof(QgtDevSubPass) / 4); i += T) ss[i] = gs[i];
        for (int i = tid; i < run.rho_blocks * 128; i += T) rho_acc[i] = 0.0;
        if (it.ovr_kind == 1) {
            const cplx* g = reinterpret_cast<const cplx*>(it.ovr_mat);
            const int cnt = QGT_VARIANT_STRIDE(N) << stages[it.ovr_index].nvar;
            for (int i = tid; i < cnt && i < OVR_ELEMS; i += T) sovr[i] = g[i];
        }
    }
    __syncthreads();
    qgt_fast_build(run, subs, stages, fast, fwarp, flane, tid, T);
    __syncthreads();

    const QgtIoMap<3> io = qgt_make_iomap<3>(run, tid);
    const int k = lane & 3;
    const uint64_t tau0 = chunk * (uint64_t)a.tiles_per_cta;
    const uint64_t tau1 = tau0 + (uint64_t)a.tiles_per_cta < a.ntiles ? tau0 + (uint64_t)a.tiles_per_cta : a.ntiles;
    // TRAJ: the transition-matrix stages of this item in order; B copy j (0, 1, ...) = stage j % nrs of tile j / nrs
    int nrs = 0;
    int rs_first = -1;
    if (TRAJ && use_b)
        for (int sg = rho_from > 0 ? rho_from : 0; sg <= run.last_rho_stage; ++sg)
            if (stages[sg].rho_off >= 0) { if (rs_first < 0) rs_first = sg; nrs++; }
    unsigned bphase = 0;
    auto issue_b = [&](uint64_t tau, int sg) {        // one thread: fetch phi's image after stage sg of tile tau
        qgt_mbar_expect_tx(&bar_b, (unsigned)(tile_elems * sizeof(cplx)));
        qgt_bulk_g2s(tileB, a.traj[stages[sg].traj_ord] + tau * tile_elems, (unsigned)(tile_elems * sizeof(cplx)), &bar_b);
    };
    if (TRAJ && nrs > 0 && tid == 0 && tau0 < tau1) issue_b(tau0, rs_first);
    int par = 0;                                     // scratch buffer the next transition matrix goes to
    for (uint64_t tau = tau0; tau < tau1; ++tau) {
        const uint64_t tilebase = qgt_tile_base(run, a.tile_off + tau);
        const uint64_t tileg = tilebase | a.gprefix;
        if (!(dbg & QGT_FDBG_NO_GLOBAL)) {
            const cplx* srcA = reinterpret_cast<const cplx*>(it.src);
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const uint32_t idx = (uint32_t)tid + (uint32_t)i * (uint32_t)T;
                const uint64_t g = tilebase | qgt_io_offset<3>(io, i);
                cp_async16(tileA + qgt_swz(idx), srcA + g, 16);
                if (!TRAJ && use_b) cp_async16(tileB + qgt_swz(idx), a.phi + g, 16);
            }
            cp_async_commit();
            cp_async_wait<0>();
        }
        __syncthreads();
        for (int s = 0; s < run.nsub; ++s) {
            const QgtDevSubPass& sp = subs[s];
            const QgtFastWarp fw = fwarp[s * 8 + warp];
            const uint32_t lt = flane[s * 32 + lane];
            const uint32_t baseB = fw.s ^ (lt & 0xffffu), baseC = fw.s ^ (lt >> 16);
            const uint64_t gwarp = tileg | fw.g;
            const uint32_t gx1 = sp.s_thr[3], gx2 = sp.s_thr[4], sr2 = sp.s_reg[2], st0 = sp.s_thr[0];
            const bool has_tdiag = sp.tdiag_end > sp.tdiag_begin;
            const int nstage = sp.stage_end - sp.stage_begin;
            if (nstage == 0) {                       // thread diagonals only (parameter-free by construction)
                if (has_tdiag) {
#pragma unroll
                    for (int g = 0; g < 4; ++g) {
                        const uint32_t gx = ((g & 1) ? gx1 : 0u) ^ ((g & 2) ? gx2 : 0u);
                        cplx p0, p1;
                        qgt_pend_pair(sp, tdiags, gwarp, g, k, p0, p1);
                        tileA[baseC ^ gx] = qgt_cmul(p0, tileA[baseC ^ gx]);
                        tileA[baseC ^ gx ^ st0] = qgt_cmul(p1, tileA[baseC ^ gx ^ st0]);
                        if (!TRAJ && use_b) {
                            tileB[baseC ^ gx] = qgt_cmul(p0, tileB[baseC ^ gx]);
                            tileB[baseC ^ gx ^ st0] = qgt_cmul(p1, tileB[baseC ^ gx ^ st0]);
                        }
                    }
                }
                __syncthreads();
                continue;
            }
            for (int sg = sp.stage_begin; sg < sp.stage_end; ++sg) {
                const QgtDevStage& st = stages[sg];
                const int var = qgt_variant_index(st, gwarp);
                const bool ovr = (it.ovr_kind == 1 && sg == it.ovr_index);
                const cplx* Mb = spool + st.mat_off + var * QGT_VARIANT_STRIDE(N);
                const StageFrag fb = qgt_load_frag(Mb, st.form == QGT_FORM_DIAG_REAL, lane);
                const StageFrag fa = ovr ? qgt_load_frag(sovr + var * QGT_VARIANT_STRIDE(N), it.ovr_form == QGT_FORM_DIAG_REAL, lane) : fb;
                const bool last = (sg == sp.stage_end - 1);
                const bool rho_stage = st.rho_off >= 0;
                const bool do_rho = rho_stage && sg >= rho_from && (self || use_b) && !(dbg & QGT_FDBG_NO_RHO);
                const bool apply_b = !TRAJ && use_b && (pair || sg <= run.last_rho_stage);
                const bool fetch_b = TRAJ && use_b && rho_stage && sg >= rho_from;      // this stage consumes a trajectory image
                double t6[6] = {0.0, 0.0, 0.0, 0.0, 0.0, 0.0};
                if (TRAJ) {
                    // all four groups through the stage first (results stay in registers), then - as late as possible - wait
                    // for phi's image of this stage and take the transition matrix
                    cplx ra0[4], ra1[4];
#pragma unroll
                    for (int h = 0; h < 2; ++h) {
                        cplx va0[2], va1[2];
#pragma unroll
                        for (int g2 = 0; g2 < 2; ++g2) {
                            const uint32_t gx = (g2 ? gx1 : 0u) ^ (h ? gx2 : 0u);
                            va0[g2] = tileA[baseB ^ gx]; va1[g2] = tileA[baseB ^ gx ^ sr2];
                        }
                        __syncwarp();             // every lane has read the groups' slots before any is overwritten
#pragma unroll
                        for (int g2 = 0; g2 < 2; ++g2) {
                            const uint32_t gx = (g2 ? gx1 : 0u) ^ (h ? gx2 : 0u);
                            cplx a0, a1;
                            qgt_apply8(fa, va0[g2], va1[g2], a0, a1);
                            if (last && has_tdiag) {
                                cplx p0, p1;
                                qgt_pend_pair(sp, tdiags, gwarp, h * 2 + g2, k, p0, p1);
                                a0 = qgt_cmul(p0, a0); a1 = qgt_cmul(p1, a1);
                            }
                            tileA[baseC ^ gx] = a0;
                            tileA[baseC ^ gx ^ st0] = a1;
                            ra0[h * 2 + g2] = a0; ra1[h * 2 + g2] = a1;
                        }
                    }
                    if (fetch_b) { qgt_mbar_wait(&bar_b, bphase); bphase ^= 1; }
                    if (do_rho) {
#pragma unroll
                        for (int g = 0; g < 4; ++g) {
                            if (self) { qgt_rho3(t6, ra0[g], ra0[g]); qgt_rho3(t6, ra1[g], ra1[g]); }
                            else {
                                const uint32_t gx = ((g & 1) ? gx1 : 0u) ^ ((g & 2) ? gx2 : 0u);
                                const cplx b0 = tileB[baseC ^ gx], b1 = tileB[baseC ^ gx ^ st0];      // the image is in C-fragment order already
                                qgt_rho3(t6, b0, ra0[g]); qgt_rho3(t6, b1, ra1[g]);
                            }
                        }
                    }
                } else {
                    // two halves of two 8-vector groups each: operand loads first, then the tensor work.  (Instantiating this
                    // body per matrix form, as the direct kernel does, was measured here: the 128-register kernel spills and
                    // runs 28 % slower, so the dense form's extra DMMAs stay predicated.)
#pragma unroll
                    for (int h = 0; h < 2; ++h) {
                        cplx va0[2], va1[2], vb0[2], vb1[2];
#pragma unroll
                        for (int g2 = 0; g2 < 2; ++g2) {
                            const uint32_t gx = (g2 ? gx1 : 0u) ^ (h ? gx2 : 0u);
                            va0[g2] = tileA[baseB ^ gx]; va1[g2] = tileA[baseB ^ gx ^ sr2];
                            if (apply_b) { vb0[g2] = tileB[baseB ^ gx]; vb1[g2] = tileB[baseB ^ gx ^ sr2]; }
                        }
                        __syncwarp();             // every lane has read the groups' slots before any is overwritten
#pragma unroll
                        for (int g2 = 0; g2 < 2; ++g2) {
                            const uint32_t gx = (g2 ? gx1 : 0u) ^ (h ? gx2 : 0u);
                            cplx a0, a1, b0, b1;
                            qgt_apply8(fa, va0[g2], va1[g2], a0, a1);
                            cplx p0, p1;
                            if (last && has_tdiag) {
                                qgt_pend_pair(sp, tdiags, gwarp, h * 2 + g2, k, p0, p1);
                                a0 = qgt_cmul(p0, a0); a1 = qgt_cmul(p1, a1);
                            }
                            tileA[baseC ^ gx] = a0;
                            tileA[baseC ^ gx ^ st0] = a1;
                            if (apply_b) {
                                qgt_apply8(fb, vb0[g2], vb1[g2], b0, b1);
                                if (last && has_tdiag) { b0 = qgt_cmul(p0, b0); b1 = qgt_cmul(p1, b1); }
                                tileB[baseC ^ gx] = b0;
                                tileB[baseC ^ gx ^ st0] = b1;
                            }
                            if (do_rho) {
                                if (self) { qgt_rho3(t6, a0, a0); qgt_rho3(t6, a1, a1); }
                                else { qgt_rho3(t6, b0, a0); qgt_rho3(t6, b1, a1); }
                            }
                        }
                    }
                }
                if (do_rho) {
                    double* sc = scratch + ((size_t)par * nwarps + warp) * 128 + lane * 4;
                    *reinterpret_cast<double2*>(sc) = make_double2(t6[0] + t6[2], t6[1] + t6[3]);
                    *reinterpret_cast<double2*>(sc + 2) = make_double2(t6[4] - t6[0] + t6[2], t6[5] - t6[1] + t6[3]);
                    if (lane == 0) wblk[par * 8 + warp] = st.rho_off + var;
                }
                __syncthreads();                  // tile coherent for the next stage, trajectory image consumed, scratch complete
                if (fetch_b && tid == 0) {
                    // next image this CTA needs: the item's next transition-matrix stage, else the first one of the next tile
                    int nx = -1;
                    for (int q2 = sg + 1; q2 <= run.last_rho_stage; ++q2) if (stages[q2].rho_off >= 0) { nx = q2; break; }
                    if (nx >= 0) issue_b(tau, nx);
                    else if (tau + 1 < tau1) issue_b(tau + 1, rs_first);
                }
                if (TRAJ && self && rho_stage && !(dbg & QGT_FDBG_NO_GLOBAL)) {
                    // phi pass: publish the tile after this stage as a contiguous image of the swizzled tile
                    cplx* img = a.traj[st.traj_ord] + tau * tile_elems;
#pragma unroll
                    for (int i = 0; i < 8; ++i) {
                        const uint32_t idx = (uint32_t)tid + (uint32_t)i * (uint32_t)T;
                        img[idx] = tileA[idx];
                    }
                    __syncthreads();              // the image is read out before the next stage overwrites the tile
                }
                if (do_rho) {
                    if (!(dbg & QGT_FDBG_NO_REDUCE)) {
                        // fixed-order sum over the warps (deterministic); a warp's block was selected by its variant
                        for (int e = tid; e < 128; e += T)
                            for (int w = 0; w < nwarps; ++w)
                                rho_acc[(size_t)wblk[par * 8 + w] * 128 + e] += scratch[((size_t)par * nwarps + w) * 128 + e];
                    }
                    par ^= 1;
                }
            }
        }
        if (!(dbg & QGT_FDBG_NO_GLOBAL)) {
            qgt_phase_store<3>(io, tileA, reinterpret_cast<cplx*>(it.dst), tilebase, tid, T, it.accumulate != 0);
            if (pair && use_b) qgt_phase_store<3>(io, tileB, reinterpret_cast<cplx*>(it.phi_dst), tilebase, tid, T, false);
        }
    }
    __syncthreads();
    double* out = a.rho_partial + (size_t)blockIdx.x * run.rho_blocks * 128;
    for (int i = tid; i < run.rho_blocks * 128; i += T) out[i] = rho_acc[i];
}

// ------------------------------------------------------------------------------------------------
// Pipelined form of the trajectory-mode kernel for full-size tiles (K = 11): ONE persistent CTA of 16 warps per
// SM, two column-tile buffers (the next tile streams in with cp.async while this one is processed) and a ring of two
// trajectory-tile buffers filled by cp.async.bulk with mbarrier completion, issued two stages ahead by one thread.
// With 2 CTAs x 8 warps and single buffers (kernel above) a CTA spends most of a tile waiting for its own loads;
// here the tensor pipe only idles at the per-stage barrier.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(512, 1) qgt_fused_pipe_kernel(FusedLaunch a) {
    constexpr int N = 8, T = 512, NW = 16;
    constexpr int OVR_ELEMS = QGT_VARIANT_STRIDE(N) << QGT_MAX_VARIANT_BITS;
    constexpr size_t TILE = (size_t)1 << 11;
    constexpr unsigned TILE_BYTES = (unsigned)(TILE * sizeof(cplx));
    extern __shared__ __align__(128) unsigned char qgt_fsmem_raw[];
    __shared__ QgtDevRun run;
    __shared__ __align__(8) uint64_t bar_b[2];
    __shared__ int rs_list[QGT_MAX_TRAJ];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    for (int i = tid; i < (int)(sizeof(QgtDevRun) / 4); i += T)
        reinterpret_cast<uint32_t*>(&run)[i] = reinterpret_cast<const uint32_t*>(&a.runs[a.run_idx])[i];
    if (tid == 0) { qgt_mbar_init(&bar_b[0], 1); qgt_mbar_init(&bar_b[1], 1); }
    __syncthreads();
    const int item_idx = (int)(blockIdx.x % (unsigned)a.nitems);
    const uint64_t chunk = blockIdx.x / (unsigned)a.nitems;
    const QgtSweepItem& it = a.items[item_idx];
    const bool self = it.self != 0;
    const int rho_from = it.rho_from;
    const int dbg = a.debug;
    const bool use_b = !self && rho_from <= run.last_rho_stage && !(dbg & QGT_FDBG_NO_B);

    cplx* tileA = reinterpret_cast<cplx*>(qgt_fsmem_raw);           // two buffers
    cplx* tileB = tileA + 2 * TILE;                                  // two buffers
    cplx* spool = tileB + 2 * TILE;
    cplx* sovr = spool + run.mat_count;
    QgtDevSubPass* subs = reinterpret_cast<QgtDevSubPass*>(sovr + OVR_ELEMS);
    QgtFastWarp* fwarp = reinterpret_cast<QgtFastWarp*>(subs + run.nsub);            // [nsub][16]
    uint32_t* flane = reinterpret_cast<uint32_t*>(fwarp + NW * run.nsub);             // [nsub][32]
    double* rho_acc = reinterpret_cast<double*>((reinterpret_cast<uintptr_t>(flane + 32 * run.nsub) + 15) & ~(uintptr_t)15);
    double* scratch = rho_acc + (size_t)run.rho_blocks * 128;                         // [2][16][128]
    int* wblk = reinterpret_cast<int*>(scratch + (size_t)2 * NW * 128);               // [2][16]
    QgtDevStage* stages = reinterpret_cast<QgtDevStage*>(wblk + 2 * NW);
    const int nstages_run = run.nsub > 0 ? subs_stage_count(a.subs + run.sub_off, run.nsub) : 0;
    const QgtDevThrDiag* tdiags = a.tdiags + run.tdiag_off;
    {
        const uint32_t* gst = reinterpret_cast<const uint32_t*>(a.stages + run.stage_off);
        for (int i = tid; i < nstages_run * (int)(sizeof(QgtDevStage) / 4); i += T) reinterpret_cast<uint32_t*>(stages)[i] = gst[i];
        const cplx* gpool = a.pool + run.mat_off;
        for (int i = tid; i < run.mat_count; i += T) spool[i] = gpool[i];
        const uint32_t* gs = reinterpret_cast<const uint32_t*>(a.subs + run.sub_off);
        uint32_t* ss = reinterpret_cast<uint32_t*>(subs);
        for (int i = tid; i < run.nsub * (int)(sizeof(QgtDevSubPass) / 4); i += T) ss[i] = gs[i];
        for (int i = tid; i < run.rho_blocks * 128; i += T) rho_acc[i] = 0.0;
    }
    __syncthreads();
    if (it.ovr_kind == 1) {
        const cplx* g = reinterpret_cast<const cplx*>(it.ovr_mat);
        const int cnt = QGT_VARIANT_STRIDE(N) << stages[it.ovr_index].nvar;
        for (int i = tid; i < cnt && i < OVR_ELEMS; i += T) sovr[i] = g[i];
    }
    // lookup tables: lane part of the operand / result slots, warp part (thread bits 4..7 with 16 warps on 256 vectors)
    for (int w = tid; w < run.nsub * 32; w += T) {
        const int s = w >> 5, l = w & 31, q = l >> 2, kk = l & 3;
        const QgtDevSubPass& sp = subs[s];
        const uint32_t b = ((q & 1) ? sp.s_thr[0] : 0u) ^ ((q & 2) ? sp.s_thr[1] : 0u) ^ ((q & 4) ? sp.s_thr[2] : 0u) ^
                           ((kk & 1) ? sp.s_reg[0] : 0u) ^ ((kk & 2) ? sp.s_reg[1] : 0u);
        const uint32_t c = ((kk & 1) ? sp.s_thr[1] : 0u) ^ ((kk & 2) ? sp.s_thr[2] : 0u) ^ ((q & 1) ? sp.s_reg[0] : 0u) ^
                           ((q & 2) ? sp.s_reg[1] : 0u) ^ ((q & 4) ? sp.s_reg[2] : 0u);
        flane[w] = b | (c << 16);
        if (l < NW) {
            QgtFastWarp fw; fw.g = 0; fw.s = 0; fw.pad = 0;
            for (int i = 4; i < 8; ++i)
                if ((l >> (i - 4)) & 1) { fw.s ^= sp.s_thr[i]; fw.g |= sp.g_thr[i]; }
            fwarp[s * NW + l] = fw;
        }
    }
    int nrs = 0;
    if (tid == 0 && use_b)
        for (int sg = rho_from > 0 ? rho_from : 0; sg <= run.last_rho_stage; ++sg)
            if (stages[sg].rho_off >= 0 && nrs < QGT_MAX_TRAJ) rs_list[nrs++] = sg;
    if (tid == 0) { for (int i = nrs; i < QGT_MAX_TRAJ; ++i) rs_list[i] = -1; }
    __syncthreads();
    if (use_b) { nrs = 0; for (int i = 0; i < QGT_MAX_TRAJ; ++i) if (rs_list[i] >= 0) nrs++; }

    const QgtIoMap<2> io = qgt_make_iomap<2>(run, tid);
    const int k = lane & 3;
    const uint64_t tau0 = chunk * (uint64_t)a.tiles_per_cta;
    const uint64_t tau1 = tau0 + (uint64_t)a.tiles_per_cta < a.ntiles ? tau0 + (uint64_t)a.tiles_per_cta : a.ntiles;
    const int ntl = (int)(tau1 - tau0);
    const int total_images = ntl * nrs;
    auto issue_b = [&](int j) {                       // one thread: image j of this CTA into ring slot j & 1
        const uint64_t tau = tau0 + (uint64_t)(j / nrs);
        const int sg = rs_list[j % nrs];
        qgt_mbar_expect_tx(&bar_b[j & 1], TILE_BYTES);
        qgt_bulk_g2s(tileB + (size_t)(j & 1) * TILE, a.traj[stages[sg].traj_ord] + tau * TILE, TILE_BYTES, &bar_b[j & 1]);
    };
    if (tid == 0 && !(dbg & QGT_FDBG_NO_BCOPY)) { if (total_images > 0) issue_b(0); if (total_images > 1) issue_b(1); }
    int jc = 0;                                       // next image to consume (uniform over the CTA)
    auto load_tile = [&](uint64_t tau, cplx* buf) {
        const uint64_t tb = qgt_tile_base(run, a.tile_off + tau);
        const cplx* srcA = reinterpret_cast<const cplx*>(it.src);
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const uint32_t idx = (uint32_t)tid + (uint32_t)i * (uint32_t)T;
            cp_async16(buf + qgt_swz(idx), srcA + (tb | qgt_io_offset<2>(io, i)), 16);
        }
    };
    if (ntl > 0 && !(dbg & QGT_FDBG_NO_GLOBAL)) load_tile(tau0, tileA);
    cp_async_commit();
    int par = 0;
    for (int ti = 0; ti < ntl; ++ti) {
        const uint64_t tau = tau0 + (uint64_t)ti;
        cplx* cur = tileA + (size_t)(ti & 1) * TILE;
        if (ti + 1 < ntl && !(dbg & QGT_FDBG_NO_GLOBAL)) load_tile(tau + 1, tileA + (size_t)((ti + 1) & 1) * TILE);
        cp_async_commit();
        cp_async_wait<1>();
        __syncthreads();
        const uint64_t tilebase = qgt_tile_base(run, a.tile_off + tau);
        const uint64_t tileg = tilebase | a.gprefix;
        for (int s = 0; s < run.nsub; ++s) {
            const QgtDevSubPass& sp = subs[s];
            const QgtFastWarp fw = fwarp[s * NW + warp];
            const uint32_t lt = flane[s * 32 + lane];
            const uint32_t baseB = fw.s ^ (lt & 0xffffu), baseC = fw.s ^ (lt >> 16);
            const uint64_t gwarp = tileg | fw.g;
            const uint32_t gx1 = sp.s_thr[3], sr2 = sp.s_reg[2], st0 = sp.s_thr[0];
            const bool has_tdiag = sp.tdiag_end > sp.tdiag_begin;
            const int nstage = sp.stage_end - sp.stage_begin;
            // pending thread-diagonal phases of this warp's two groups: vector index bits 0..3 <- (2k | j, g), bits 4.. <- warp
            auto pend_pair = [&](int g, cplx& p0, cplx& p1) {
                p0.x = 1.0; p0.y = 0.0; p1 = p0;
                uint64_t gc0 = gwarp;
                const int t0 = g * 8 + 2 * k;
#pragma unroll
                for (int i = 0; i < 4; ++i) if ((t0 >> i) & 1) gc0 |= sp.g_thr[i];
                const uint64_t gc1 = gc0 | sp.g_thr[0];
                for (int t = sp.tdiag_begin; t < sp.tdiag_end; ++t) {
                    qgt_thread_diag(p0, tdiags[t], gc0);
                    qgt_thread_diag(p1, tdiags[t], gc1);
                }
            };
            if (nstage == 0) {
                if (has_tdiag) {
#pragma unroll
                    for (int g = 0; g < 2; ++g) {
                        const uint32_t gx = g ? gx1 : 0u;
                        cplx p0, p1;
                        pend_pair(g, p0, p1);
                        cur[baseC ^ gx] = qgt_cmul(p0, cur[baseC ^ gx]);
                        cur[baseC ^ gx ^ st0] = qgt_cmul(p1, cur[baseC ^ gx ^ st0]);
                    }
                }
                __syncthreads();
                continue;
            }
            for (int sg = sp.stage_begin; sg < sp.stage_end; ++sg) {
                const QgtDevStage& st = stages[sg];
                const int var = qgt_variant_index(st, gwarp);
                const bool ovr = (it.ovr_kind == 1 && sg == it.ovr_index);
                const StageFrag fa = ovr ? qgt_load_frag(sovr + var * QGT_VARIANT_STRIDE(N), it.ovr_form == QGT_FORM_DIAG_REAL, lane)
                                         : qgt_load_frag(spool + st.mat_off + var * QGT_VARIANT_STRIDE(N), st.form == QGT_FORM_DIAG_REAL, lane);
                const bool last = (sg == sp.stage_end - 1);
                const bool rho_stage = st.rho_off >= 0;
                const bool fetch_b = use_b && rho_stage && sg >= rho_from;
                const bool do_rho = rho_stage && sg >= rho_from && (self || use_b) && !(dbg & QGT_FDBG_NO_RHO);
                cplx ra0[2], ra1[2];
                {
                    cplx va0[2], va1[2];
#pragma unroll
                    for (int g = 0; g < 2; ++g) {
                        const uint32_t gx = g ? gx1 : 0u;
                        va0[g] = cur[baseB ^ gx]; va1[g] = cur[baseB ^ gx ^ sr2];
                    }
                    __syncwarp();                 // every lane has read the groups' slots before any is overwritten
#pragma unroll
                    for (int g = 0; g < 2; ++g) {
                        const uint32_t gx = g ? gx1 : 0u;
                        qgt_apply8(fa, va0[g], va1[g], ra0[g], ra1[g]);
                        if (last && has_tdiag) {
                            cplx p0, p1;
                            pend_pair(g, p0, p1);
                            ra0[g] = qgt_cmul(p0, ra0[g]); ra1[g] = qgt_cmul(p1, ra1[g]);
                        }
                        cur[baseC ^ gx] = ra0[g];
                        cur[baseC ^ gx ^ st0] = ra1[g];
                    }
                }
                const cplx* B = tileB + (size_t)(jc & 1) * TILE;
                if (fetch_b && !(dbg & QGT_FDBG_NO_BCOPY)) qgt_mbar_wait(&bar_b[jc & 1], (unsigned)(jc >> 1) & 1u);
                if (do_rho) {
                    double t6[6] = {0.0, 0.0, 0.0, 0.0, 0.0, 0.0};
#pragma unroll
                    for (int g = 0; g < 2; ++g) {
                        if (self) { qgt_rho3(t6, ra0[g], ra0[g]); qgt_rho3(t6, ra1[g], ra1[g]); }
                        else {
                            const uint32_t gx = g ? gx1 : 0u;
                            const cplx b0 = B[baseC ^ gx], b1 = B[baseC ^ gx ^ st0];       // the image is in C-fragment order already
                            qgt_rho3(t6, b0, ra0[g]); qgt_rho3(t6, b1, ra1[g]);
                        }
                    }
                    double* sc = scratch + ((size_t)par * NW + warp) * 128 + lane * 4;
                    *reinterpret_cast<double2*>(sc) = make_double2(t6[0] + t6[2], t6[1] + t6[3]);
                    *reinterpret_cast<double2*>(sc + 2) = make_double2(t6[4] - t6[0] + t6[2], t6[5] - t6[1] + t6[3]);
                    if (lane == 0) wblk[par * NW + warp] = st.rho_off + var;
                }
                __syncthreads();                  // tile coherent for the next stage, image consumed, scratch complete
                if (fetch_b) {
                    if (tid == 0 && jc + 2 < total_images && !(dbg & QGT_FDBG_NO_BCOPY)) issue_b(jc + 2);
                    jc++;
                }
                if (self && rho_stage && !(dbg & QGT_FDBG_NO_GLOBAL)) {
                    // phi pass: publish the tile after this stage as a contiguous image of the swizzled tile
                    cplx* img = a.traj[st.traj_ord] + tau * TILE;
#pragma unroll
                    for (int i = 0; i < 4; ++i) img[tid + i * T] = cur[tid + i * T];
                    __syncthreads();              // read out before the next stage overwrites the tile
                }
                if (do_rho) {
                    if (tid < 128 && !(dbg & QGT_FDBG_NO_REDUCE)) {
                        // fixed-order sum over the warps (deterministic); a warp's block was selected by its variant
#pragma unroll 4
                        for (int w = 0; w < NW; ++w)
                            rho_acc[(size_t)wblk[par * NW + w] * 128 + tid] += scratch[((size_t)par * NW + w) * 128 + tid];
                    }
                    par ^= 1;
                }
            }
        }
        if (!(dbg & QGT_FDBG_NO_GLOBAL))
            qgt_phase_store<2>(io, cur, reinterpret_cast<cplx*>(it.dst), tilebase, tid, T, it.accumulate != 0);
    }
    cp_async_wait<0>();
    __syncthreads();
    double* out = a.rho_partial + (size_t)blockIdx.x * run.rho_blocks * 128;
    for (int i = tid; i < run.rho_blocks * 128; i += T) out[i] = rho_acc[i];
}

// ------------------------------------------------------------------------------------------------
// Lean form of the trajectory-mode kernel for the common case: 11-qubit tiles, every sub-pass = one dense stage
// without thread diagonals (a hardware-efficient ansatz).  16 warps per CTA, 64 registers, TWO CTAs per SM (32
// warps hide each other's barriers and shared-memory latencies, as the 4 x 8-warp plain sweep kernel does); all
// per-stage constants come from small shared-memory tables built once per CTA.
// ------------------------------------------------------------------------------------------------
constexpr int QGT_DIRECT_MAX_SUBS = 48;       // sub-passes (= stages) of a run the direct kernel keeps variant masks for
struct QgtLeanSub {          // 32 bytes per sub-pass (= per stage)
    uint32_t gx1, sr2, st0;  // slot XOR terms: thread bit 3 (second group), matrix bit 2, thread bit 0
    uint32_t mat_off;        // variant 0 of the stage matrix in the run's pool (complex elements)
    uint32_t vm0, vm1;       // variant-selecting bits, as positions in a packed 32-bit word (see qgt_lean_pack)
    int32_t rho_off;         // first transition-matrix block, -1 = none
    int32_t info;            // bit 0 diagonal-real form, bits 8.. trajectory ordinal
};

__global__ void __launch_bounds__(512, 2) qgt_fused_lean_kernel(FusedLaunch a) {
    constexpr int N = 8, T = 512, NW = 16;
    constexpr int OVR_ELEMS = QGT_VARIANT_STRIDE(N) << QGT_MAX_VARIANT_BITS;
    constexpr int TILE = 1 << 11;
    constexpr unsigned TILE_BYTES = (unsigned)(TILE * sizeof(cplx));
    extern __shared__ __align__(128) cplx qgt_lsm[];
    __shared__ QgtDevRun run;
    __shared__ __align__(8) uint64_t bar_b;
    __shared__ int rs_list[QGT_MAX_TRAJ];
    __shared__ int wblk[NW];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    for (int i = tid; i < (int)(sizeof(QgtDevRun) / 4); i += T)
        reinterpret_cast<uint32_t*>(&run)[i] = reinterpret_cast<const uint32_t*>(&a.runs[a.run_idx])[i];
    if (tid == 0) qgt_mbar_init(&bar_b, 1);
    __syncthreads();
    const int item_idx = (int)(blockIdx.x % (unsigned)a.nitems);
    const int chunk = (int)(blockIdx.x / (unsigned)a.nitems);
    const QgtSweepItem& it = a.items[item_idx];
    const bool self = it.self != 0;
    const int rho_from = it.rho_from;
    const bool use_b = !self && rho_from <= run.last_rho_stage;
    const int nsub = run.nsub;
    const int dbg = a.debug;

    // shared-memory carve-up (every region a multiple of 16 bytes)
    cplx* tileA = qgt_lsm;
    cplx* tileB = tileA + TILE;
    cplx* spool = tileB + TILE;
    cplx* sovr = spool + run.mat_count;
    double* rho_acc = reinterpret_cast<double*>(sovr + OVR_ELEMS);
    double* scratch = rho_acc + run.rho_blocks * 128;                  // [16][128]
    QgtLeanSub* lsub = reinterpret_cast<QgtLeanSub*>(scratch + NW * 128);
    uint32_t* flane = reinterpret_cast<uint32_t*>(lsub + nsub);        // [nsub][32]
    uint2* fwarp = reinterpret_cast<uint2*>(flane + 32 * nsub);        // [nsub][16]: (slot part, variant-bit word)
    {
        const cplx* gpool = a.pool + run.mat_off;
        for (int i = tid; i < run.mat_count; i += T) spool[i] = gpool[i];
        for (int i = tid; i < run.rho_blocks * 128; i += T) rho_acc[i] = 0.0;
        if (it.ovr_kind == 1) {
            const cplx* g = reinterpret_cast<const cplx*>(it.ovr_mat);
            const int cnt = QGT_VARIANT_STRIDE(N) << a.stages[run.stage_off + it.ovr_index].nvar;
            for (int i = tid; i < cnt && i < OVR_ELEMS; i += T) sovr[i] = g[i];
        }
        // tables.  The variant of a warp depends on two global-index bits; pack "is bit vm set in (warp part | tile part)"
        // as: word bit 0/1 = contribution of the warp bits to variant bit 0/1 (tile part added per tile from vm masks)
        const QgtDevSubPass* gsubs = a.subs + run.sub_off;
        const QgtDevStage* gstages = a.stages + run.stage_off;
        for (int w = tid; w < nsub * 32; w += T) {
            const int sI = w >> 5, l = w & 31, q = l >> 2, kk = l & 3;
            const QgtDevSubPass& sp = gsubs[sI];
            const QgtDevStage& st = gstages[sp.stage_begin];
            const uint32_t b = ((q & 1) ? sp.s_thr[0] : 0u) ^ ((q & 2) ? sp.s_thr[1] : 0u) ^ ((q & 4) ? sp.s_thr[2] : 0u) ^
                               ((kk & 1) ? sp.s_reg[0] : 0u) ^ ((kk & 2) ? sp.s_reg[1] : 0u);
            const uint32_t c = ((kk & 1) ? sp.s_thr[1] : 0u) ^ ((kk & 2) ? sp.s_thr[2] : 0u) ^ ((q & 1) ? sp.s_reg[0] : 0u) ^
                               ((q & 2) ? sp.s_reg[1] : 0u) ^ ((q & 4) ? sp.s_reg[2] : 0u);
            flane[w] = b | (c << 16);
            if (l < NW) {
                uint32_t sl = 0; uint64_t g = 0;
                for (int i = 4; i < 8; ++i)
                    if ((l >> (i - 4)) & 1) { sl ^= sp.s_thr[i]; g |= sp.g_thr[i]; }
                uint32_t vb = 0;
                if (st.nvar > 0 && (g & st.vmask[0])) vb |= 1u;
                if (st.nvar > 1 && (g & st.vmask[1])) vb |= 2u;
                fwarp[sI * NW + l] = make_uint2(sl, vb);
            }
            if (l == 31) {
                QgtLeanSub ls;
                ls.gx1 = sp.s_thr[3]; ls.sr2 = sp.s_reg[2]; ls.st0 = sp.s_thr[0];
                ls.mat_off = (uint32_t)st.mat_off;
                ls.vm0 = 0; ls.vm1 = 0;
                ls.rho_off = st.rho_off;
                ls.info = (st.form == QGT_FORM_DIAG_REAL ? 1 : 0) | ((st.traj_ord < 0 ? 0xff : st.traj_ord) << 8) | (st.nvar << 4);
                lsub[sI] = ls;
            }
        }
        if (tid == 0) {
            int n = 0;
            if (use_b)
                for (int sg = rho_from > 0 ? rho_from : 0; sg <= run.last_rho_stage; ++sg)
                    if (gstages[sg].rho_off >= 0 && n < QGT_MAX_TRAJ) rs_list[n++] = sg;
            for (int i = n; i < QGT_MAX_TRAJ; ++i) rs_list[i] = -1;
        }
    }
    __syncthreads();
    int nrs = 0;
#pragma unroll
    for (int i = 0; i < QGT_MAX_TRAJ; ++i) if (rs_list[i] >= 0) nrs++;
    const QgtDevStage* gstages = a.stages + run.stage_off;

    const QgtIoMap<2> io = qgt_make_iomap<2>(run, tid);
    const int tau0 = chunk * a.tiles_per_cta;
    const int ntl = (int)((uint64_t)tau0 + a.tiles_per_cta < a.ntiles ? (uint64_t)a.tiles_per_cta : a.ntiles - (uint64_t)tau0);
    const int total_images = ntl * nrs;
    auto issue_b = [&](int j) {                       // one thread: image j of this CTA
        const int sg = rs_list[j % nrs];
        qgt_mbar_expect_tx(&bar_b, TILE_BYTES);
        qgt_bulk_g2s(tileB, a.traj[gstages[sg].traj_ord] + (size_t)(tau0 + j / nrs) * TILE, TILE_BYTES, &bar_b);
    };
    if (tid == 0 && total_images > 0 && !(dbg & QGT_FDBG_NO_BCOPY)) issue_b(0);
    int jc = 0;
    const int ovr_stage = it.ovr_kind == 1 ? it.ovr_index : -1;
    const bool ovr_dr = it.ovr_form == QGT_FORM_DIAG_REAL;
    const cplx* srcA = reinterpret_cast<const cplx*>(it.src);
    cplx* dstA = reinterpret_cast<cplx*>(it.dst);
    for (int ti = 0; ti < ntl; ++ti) {
        const uint64_t tilebase = qgt_tile_base(run, a.tile_off + (uint64_t)(tau0 + ti));
        const uint64_t tileg = tilebase | a.gprefix;
        if (!(dbg & QGT_FDBG_NO_GLOBAL)) {
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const uint32_t idx = (uint32_t)tid + (uint32_t)i * (uint32_t)T;
                cp_async16(tileA + qgt_swz(idx), srcA + (tilebase | qgt_io_offset<2>(io, i)), 16);
            }
        }
        cp_async_commit();
        cp_async_wait<0>();
        __syncthreads();
        for (int s = 0; s < nsub; ++s) {
            const QgtLeanSub ls = lsub[s];
            const uint2 fw = fwarp[s * NW + warp];
            const uint32_t lt = flane[s * 32 + lane];
            const uint32_t baseB = fw.x ^ (lt & 0xffffu), baseC = fw.x ^ (lt >> 16);
            int var = (int)fw.y;
            const int nvar = (ls.info >> 4) & 3;
            if (nvar) {                               // tile part of the variant bits
                const QgtDevStage& st = gstages[s];
                if (tileg & st.vmask[0]) var |= 1;
                if (nvar > 1 && (tileg & st.vmask[1])) var |= 2;
            }
            const bool ovr = (s == ovr_stage);
            const bool dr = ovr ? ovr_dr : (ls.info & 1);
            const cplx* M = ovr ? sovr + var * QGT_VARIANT_STRIDE(N) : spool + ls.mat_off + var * QGT_VARIANT_STRIDE(N);
            const StageFrag fa = qgt_load_frag(M, dr, lane);
            const bool rho_stage = ls.rho_off >= 0;
            const bool fetch_b = use_b && rho_stage && s >= rho_from;
            const bool do_rho = rho_stage && s >= rho_from && (self || use_b) && !(dbg & QGT_FDBG_NO_RHO);
            cplx ra0[2], ra1[2];
            {
                const cplx va00 = tileA[baseB], va01 = tileA[baseB ^ ls.sr2];
                const cplx va10 = tileA[baseB ^ ls.gx1], va11 = tileA[baseB ^ ls.gx1 ^ ls.sr2];
                __syncwarp();                         // every lane has read the groups' slots before any is overwritten
                qgt_apply8(fa, va00, va01, ra0[0], ra1[0]);
                qgt_apply8(fa, va10, va11, ra0[1], ra1[1]);
                tileA[baseC] = ra0[0]; tileA[baseC ^ ls.st0] = ra1[0];
                tileA[baseC ^ ls.gx1] = ra0[1]; tileA[baseC ^ ls.gx1 ^ ls.st0] = ra1[1];
            }
            if (fetch_b && !(dbg & QGT_FDBG_NO_BCOPY)) qgt_mbar_wait(&bar_b, (unsigned)jc & 1u);
            if (do_rho) {
                double t6[6] = {0.0, 0.0, 0.0, 0.0, 0.0, 0.0};
                if (self) {
                    qgt_rho3(t6, ra0[0], ra0[0]); qgt_rho3(t6, ra1[0], ra1[0]);
                    qgt_rho3(t6, ra0[1], ra0[1]); qgt_rho3(t6, ra1[1], ra1[1]);
                } else {
                    const cplx b00 = tileB[baseC], b01 = tileB[baseC ^ ls.st0];
                    const cplx b10 = tileB[baseC ^ ls.gx1], b11 = tileB[baseC ^ ls.gx1 ^ ls.st0];
                    qgt_rho3(t6, b00, ra0[0]); qgt_rho3(t6, b01, ra1[0]);
                    qgt_rho3(t6, b10, ra0[1]); qgt_rho3(t6, b11, ra1[1]);
                }
                double* sc = scratch + warp * 128 + lane * 4;
                *reinterpret_cast<double2*>(sc) = make_double2(t6[0] + t6[2], t6[1] + t6[3]);
                *reinterpret_cast<double2*>(sc + 2) = make_double2(t6[4] - t6[0] + t6[2], t6[5] - t6[1] + t6[3]);
                if (lane == 0) wblk[warp] = ls.rho_off + var;
            }
            __syncthreads();                          // tile coherent for the next stage, image consumed, scratch complete
            if (fetch_b) {
                if (tid == 0 && jc + 1 < total_images && !(dbg & QGT_FDBG_NO_BCOPY)) issue_b(jc + 1);
                jc++;
            }
            if (self && rho_stage && !(dbg & QGT_FDBG_NO_GLOBAL)) {
                cplx* img = a.traj[(ls.info >> 8) & 0xff] + (size_t)(tau0 + ti) * TILE;
#pragma unroll
                for (int i = 0; i < 4; ++i) img[tid + i * T] = tileA[tid + i * T];
            }
            if (do_rho) {
                // warps 0-3: one double of the block per lane, summed over the 16 warps in fixed order (deterministic); a
                // warp's partial belongs to the block its variant selected
                if (tid < 128 && !(dbg & QGT_FDBG_NO_REDUCE)) {
                    const int nb = 1 << nvar;
                    for (int v = 0; v < nb; ++v) {
                        double acc = 0.0;
#pragma unroll
                        for (int w = 0; w < NW; ++w) {
                            const double x = scratch[w * 128 + tid];
                            if (nb == 1 || wblk[w] == ls.rho_off + v) acc += x;
                        }
                        rho_acc[(ls.rho_off + v) * 128 + tid] += acc;
                    }
                }
            }
            if (do_rho || (self && rho_stage)) __syncthreads();    // scratch / tile read out before they are rewritten
        }
        if (!(dbg & QGT_FDBG_NO_GLOBAL)) qgt_phase_store<2>(io, tileA, dstA, tilebase, tid, T, it.accumulate != 0);
    }
    __syncthreads();
    double* out = a.rho_partial + (size_t)blockIdx.x * run.rho_blocks * 128;
    for (int i = tid; i < run.rho_blocks * 128; i += T) out[i] = rho_acc[i];
}

size_t fused_lean_smem_bytes(int mat_count, int nsub, int rho_blocks) {
    return (sizeof(cplx) << 11) * 2 + sizeof(cplx) * (size_t)mat_count + sizeof(cplx) * (size_t)(QGT_VARIANT_STRIDE(8) << QGT_MAX_VARIANT_BITS) +
           (size_t)rho_blocks * 128 * sizeof(double) + (size_t)16 * 128 * sizeof(double) +
           (size_t)nsub * (sizeof(QgtLeanSub) + 32 * sizeof(uint32_t) + 16 * sizeof(uint2));
}

// ------------------------------------------------------------------------------------------------
// Direct form of the trajectory mode (default for 11-qubit tiles whose sub-passes are single stages): phi's tile
// after a stage is published in FRAGMENT ORDER - element ((warp * 4 + group) * 32 + lane) * 2 + j is the j-th result
// of that lane - so a consumer lane fetches exactly its own two C-fragment values per group with two coalesced
// 16-byte read-only loads straight into registers.  No second tile in shared memory, no staging copy: a CTA needs
// its own 32 KB tile + tables, three CTAs of 8 warps fit on an SM and their barriers / load phases interleave as in
// the plain sweep kernel.
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ cplx qgt_ldg_nc(const cplx* p) {
    cplx v;
    asm volatile("ld.global.nc.v2.f64 {%0, %1}, [%2];" : "=d"(v.x), "=d"(v.y) : "l"(p));
    return v;
}

__global__ void __launch_bounds__(256, 3) qgt_fused_direct_kernel(FusedLaunch a) {
    constexpr int N = 8, T = 256, NW = 8;
    constexpr int TILE = 1 << 11;
    extern __shared__ __align__(128) cplx qgt_dsm[];
    __shared__ QgtDevRun run;
    __shared__ int wblk[2 * NW];
    __shared__ uint64_t svm[QGT_DIRECT_MAX_SUBS][2];     // the stages' variant masks (not read from global memory per stage)
    __shared__ int tvar[QGT_DIRECT_MAX_SUBS];            // per tile: the tile's part of every stage's variant index
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    for (int i = tid; i < (int)(sizeof(QgtDevRun) / 4); i += T)
        reinterpret_cast<uint32_t*>(&run)[i] = reinterpret_cast<const uint32_t*>(&a.runs[a.run_idx])[i];
    __syncthreads();
    const int item_idx = (int)(blockIdx.x % (unsigned)a.nitems);
    const int chunk = (int)(blockIdx.x / (unsigned)a.nitems);
    const QgtSweepItem& it = a.items[item_idx];
    const bool self = it.self != 0;
    const int rho_from = it.rho_from;
    const int dbg = a.debug;
    const bool use_b = !self && rho_from <= run.last_rho_stage;
    const int nsub = run.nsub;

    cplx* tileA = qgt_dsm;
    cplx* spool = tileA + TILE;
    double* rho_acc = reinterpret_cast<double*>(spool + run.mat_count);
    double* scratch = rho_acc + run.rho_blocks * 128;                  // [2][8][128]
    QgtLeanSub* lsub = reinterpret_cast<QgtLeanSub*>(scratch + 2 * NW * 128);
    uint32_t* flane = reinterpret_cast<uint32_t*>(lsub + nsub);        // [nsub][32]
    uint2* fwarp = reinterpret_cast<uint2*>(flane + 32 * nsub);        // [nsub][8]: (slot part, variant bits of the warp)
    const QgtDevStage* gstages = a.stages + run.stage_off;
    {
        const cplx* gpool = a.pool + run.mat_off;
        for (int i = tid; i < run.mat_count; i += T) spool[i] = gpool[i];
        for (int i = tid; i < run.rho_blocks * 128; i += T) rho_acc[i] = 0.0;
        const QgtDevSubPass* gsubs = a.subs + run.sub_off;
        for (int i = tid; i < nsub * 2; i += T) {
            const QgtDevStage& st = gstages[gsubs[i >> 1].stage_begin];
            svm[i >> 1][i & 1] = (i & 1) < st.nvar ? st.vmask[i & 1] : 0ull;
        }
        for (int w = tid; w < nsub * 32; w += T) {
            const int sI = w >> 5, l = w & 31, q = l >> 2, kk = l & 3;
            const QgtDevSubPass& sp = gsubs[sI];
            const QgtDevStage& st = gstages[sp.stage_begin];
            const uint32_t b = ((q & 1) ? sp.s_thr[0] : 0u) ^ ((q & 2) ? sp.s_thr[1] : 0u) ^ ((q & 4) ? sp.s_thr[2] : 0u) ^
                               ((kk & 1) ? sp.s_reg[0] : 0u) ^ ((kk & 2) ? sp.s_reg[1] : 0u);
            const uint32_t c = ((kk & 1) ? sp.s_thr[1] : 0u) ^ ((kk & 2) ? sp.s_thr[2] : 0u) ^ ((q & 1) ? sp.s_reg[0] : 0u) ^
                               ((q & 2) ? sp.s_reg[1] : 0u) ^ ((q & 4) ? sp.s_reg[2] : 0u);
            flane[w] = b | (c << 16);
            if (l < NW) {
                uint32_t sl = 0; uint64_t g = 0;
                for (int i = 5; i < 8; ++i)
                    if ((l >> (i - 5)) & 1) { sl ^= sp.s_thr[i]; g |= sp.g_thr[i]; }
                uint32_t vb = 0;
                if (st.nvar > 0 && (g & st.vmask[0])) vb |= 1u;
                if (st.nvar > 1 && (g & st.vmask[1])) vb |= 2u;
                fwarp[sI * NW + l] = make_uint2(sl, vb);
            }
            if (l == 31) {
                QgtLeanSub ls;
                ls.gx1 = sp.s_thr[3]; ls.sr2 = sp.s_reg[2]; ls.st0 = sp.s_thr[0];
                ls.mat_off = (uint32_t)st.mat_off;
                ls.vm0 = sp.s_thr[4]; ls.vm1 = 0;                      // vm0 doubles as the second group bit (gx2) here
                ls.rho_off = st.rho_off;
                ls.info = (st.form == QGT_FORM_DIAG_REAL ? 1 : 0) | ((st.traj_ord < 0 ? 0xff : st.traj_ord) << 8) | (st.nvar << 4);
                lsub[sI] = ls;
            }
        }
    }
    __syncthreads();

    const QgtIoMap<3> io = qgt_make_iomap<3>(run, tid);
    const int tau0 = chunk * a.tiles_per_cta;
    const int ntl = (int)((uint64_t)tau0 + a.tiles_per_cta < a.ntiles ? (uint64_t)a.tiles_per_cta : a.ntiles - (uint64_t)tau0);
    const int ovr_stage = it.ovr_kind == 1 ? it.ovr_index : -1;
    const bool ovr_dr = it.ovr_form == QGT_FORM_DIAG_REAL;
    const cplx* ovr_mat = reinterpret_cast<const cplx*>(it.ovr_mat);
    const cplx* srcA = reinterpret_cast<const cplx*>(it.src);
    cplx* dstA = reinterpret_cast<cplx*>(it.dst);
    const int frag_off = (warp * 4 * 32 + lane) * 2;                   // + g * 64 + j
    int par = 0;
    for (int ti = 0; ti < ntl; ++ti) {
        const uint64_t tilebase = qgt_tile_base(run, a.tile_off + (uint64_t)(tau0 + ti));
        const uint64_t tileg = tilebase | a.gprefix;
        if (!(dbg & QGT_FDBG_NO_GLOBAL)) {
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const uint32_t idx = (uint32_t)tid + (uint32_t)i * (uint32_t)T;
                cp_async16(tileA + qgt_swz(idx), srcA + (tilebase | qgt_io_offset<3>(io, i)), 16);
            }
        }
        cp_async_commit();
        // the tile's variant bits of all stages, once per tile (after the last stage's barrier nobody reads the previous
        // tile's entries any more); the stage loop then needs neither the tile's global index nor the 64-bit masks
        if (tid < nsub) tvar[tid] = ((tileg & svm[tid][0]) ? 1 : 0) | ((tileg & svm[tid][1]) ? 2 : 0);
        cp_async_wait<0>();
        __syncthreads();
        for (int s = 0; s < nsub; ++s) {
            const QgtLeanSub ls = lsub[s];
            const uint2 fw = fwarp[s * NW + warp];
            const uint32_t lt = flane[s * 32 + lane];
            const uint32_t baseB = fw.x ^ (lt & 0xffffu), baseC = fw.x ^ (lt >> 16);
            const uint32_t gx1 = ls.gx1, gx2 = ls.vm0;
            const int var = (int)fw.y | tvar[s];
            const int nvar = (ls.info >> 4) & 3;
            const bool ovr = (s == ovr_stage);
            const bool dr = ovr ? ovr_dr : (ls.info & 1);
            const StageFrag fa = ovr ? qgt_load_frag(ovr_mat + var * QGT_VARIANT_STRIDE(N), dr, lane)
                                     : qgt_load_frag(spool + ls.mat_off + var * QGT_VARIANT_STRIDE(N), dr, lane);
            const bool rho_stage = ls.rho_off >= 0;
            const bool fetch_b = use_b && rho_stage && s >= rho_from && !(dbg & QGT_FDBG_NO_BCOPY);
            const bool do_rho = rho_stage && s >= rho_from && (self || use_b) && !(dbg & QGT_FDBG_NO_RHO);
            cplx* img = rho_stage ? a.traj[(ls.info >> 8) & 0xff] + (size_t)(tau0 + ti) * TILE + frag_off : nullptr;
            // phi's fragments of this stage, one half (two groups) at a time: the first half is requested now, the second
            // while the first half's groups go through the stage, and each half is contracted right after its groups, so
            // only two groups of results and two halves of phi are live (the four-group form spilled phi's fragments to
            // local memory straight after the load, which made every warp wait for L2 where the prefetch was meant to hide it)
            cplx vb[2][2][2];
            if (fetch_b) {
#pragma unroll
                for (int g2 = 0; g2 < 2; ++g2) { vb[0][g2][0] = qgt_ldg_nc(img + g2 * 64); vb[0][g2][1] = qgt_ldg_nc(img + g2 * 64 + 1); }
            }
            double t6[6] = {0.0, 0.0, 0.0, 0.0, 0.0, 0.0};
            auto groups = [&](auto dr_tag) {
                constexpr bool DR = decltype(dr_tag)::value;
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    cplx va0[2], va1[2], ra0[2], ra1[2];
#pragma unroll
                    for (int g2 = 0; g2 < 2; ++g2) {
                        const uint32_t gx = (g2 ? gx1 : 0u) ^ (h ? gx2 : 0u);
                        va0[g2] = tileA[baseB ^ gx]; va1[g2] = tileA[baseB ^ gx ^ ls.sr2];
                    }
                    __syncwarp();                     // every lane has read the groups' slots before any is overwritten
                    if (h == 0 && fetch_b) {          // behind the barrier: the compiler must not hoist these next to the first half
#pragma unroll
                        for (int g2 = 0; g2 < 2; ++g2) { vb[1][g2][0] = qgt_ldg_nc(img + (2 + g2) * 64); vb[1][g2][1] = qgt_ldg_nc(img + (2 + g2) * 64 + 1); }
                    }
#pragma unroll
                    for (int g2 = 0; g2 < 2; ++g2) {
                        const uint32_t gx = (g2 ? gx1 : 0u) ^ (h ? gx2 : 0u);
                        qgt_apply8t<DR>(fa, va0[g2], va1[g2], ra0[g2], ra1[g2]);
                        tileA[baseC ^ gx] = ra0[g2];
                        tileA[baseC ^ gx ^ ls.st0] = ra1[g2];
                    }
                    if (self && rho_stage && !(dbg & QGT_FDBG_NO_GLOBAL)) {
#pragma unroll
                        for (int g2 = 0; g2 < 2; ++g2) { img[(h * 2 + g2) * 64] = ra0[g2]; img[(h * 2 + g2) * 64 + 1] = ra1[g2]; }
                    }
                    if (do_rho) {
#pragma unroll
                        for (int g2 = 0; g2 < 2; ++g2) {
                            if (self) { qgt_rho3(t6, ra0[g2], ra0[g2]); qgt_rho3(t6, ra1[g2], ra1[g2]); }
                            else { qgt_rho3(t6, vb[h][g2][0], ra0[g2]); qgt_rho3(t6, vb[h][g2][1], ra1[g2]); }
                        }
                    }
                }
            };
            if (dr) groups(std::true_type{}); else groups(std::false_type{});
            if (do_rho) {
                double* sc = scratch + (par * NW + warp) * 128 + lane * 4;
                *reinterpret_cast<double2*>(sc) = make_double2(t6[0] + t6[2], t6[1] + t6[3]);
                *reinterpret_cast<double2*>(sc + 2) = make_double2(t6[4] - t6[0] + t6[2], t6[5] - t6[1] + t6[3]);
                if (lane == 0) wblk[par * NW + warp] = ls.rho_off + var;
            }
            __syncthreads();                          // tile coherent for the next stage, scratch complete
            if (do_rho) {
                // every warp sums 16 of the 128 doubles over the 8 warps as a fixed tree (deterministic); a warp's partial
                // belongs to the block its variant selected: the lower half-warp collects variant 0 (2), the upper one
                // variant 1 (3), so one pass serves both blocks of a one-bit stage and a stage without variants needs no
                // selection at all.  (The serial 8-term chain per block, run by 16 lanes, cost 9 % of the kernel: its
                // additions queue behind the other warps' DMMAs on the FP64 pipe.)  The scratch is double-buffered: no
                // second barrier.
                if (!(dbg & QGT_FDBG_NO_REDUCE)) {
                    const int e = warp * 16 + (lane & 15);
                    const double* sp = scratch + par * NW * 128 + e;
                    if (nvar == 0) {
                        if (lane < 16) {
                            double x[NW];
#pragma unroll
                            for (int w = 0; w < NW; ++w) x[w] = sp[w * 128];
                            rho_acc[ls.rho_off * 128 + e] += ((x[0] + x[1]) + (x[2] + x[3])) + ((x[4] + x[5]) + (x[6] + x[7]));
                        }
                    } else {
                        double x[NW];
                        int bl[NW];
#pragma unroll
                        for (int w = 0; w < NW; ++w) { x[w] = sp[w * 128]; bl[w] = wblk[par * NW + w] - ls.rho_off; }
                        for (int v = lane >> 4; v < (1 << nvar); v += 2) {
                            double y[NW];
#pragma unroll
                            for (int w = 0; w < NW; ++w) y[w] = (bl[w] == v) ? x[w] : 0.0;
                            rho_acc[(ls.rho_off + v) * 128 + e] += ((y[0] + y[1]) + (y[2] + y[3])) + ((y[4] + y[5]) + (y[6] + y[7]));
                        }
                    }
                }
                par ^= 1;
            }
        }
        if (!(dbg & QGT_FDBG_NO_GLOBAL)) qgt_phase_store<3>(io, tileA, dstA, tilebase, tid, T, it.accumulate != 0);
    }
    __syncthreads();
    double* out = a.rho_partial + (size_t)blockIdx.x * run.rho_blocks * 128;
    for (int i = tid; i < run.rho_blocks * 128; i += T) out[i] = rho_acc[i];
}

size_t fused_direct_smem_bytes(int mat_count, int nsub, int rho_blocks) {
    return (sizeof(cplx) << 11) + sizeof(cplx) * (size_t)mat_count + (size_t)rho_blocks * 128 * sizeof(double) +
           (size_t)2 * 8 * 128 * sizeof(double) + (size_t)nsub * (sizeof(QgtLeanSub) + 32 * sizeof(uint32_t) + 8 * sizeof(uint2));
}

size_t fused_pipe_smem_bytes(int mat_count, int nsub, int rho_blocks, int nstages) {
    return (sizeof(cplx) << 11) * 4 + sizeof(cplx) * (size_t)mat_count + sizeof(cplx) * (size_t)(QGT_VARIANT_STRIDE(8) << QGT_MAX_VARIANT_BITS) +
           (size_t)nsub * (sizeof(QgtDevSubPass) + 16 * sizeof(QgtFastWarp) + 32 * sizeof(uint32_t)) + 16 +
           (size_t)rho_blocks * 128 * sizeof(double) + (size_t)2 * 16 * 128 * sizeof(double) + 32 * sizeof(int) + (size_t)nstages * sizeof(QgtDevStage);
}

// tiles per CTA and number of tile chunks: ~8 waves of resident CTAs (2 per SM) so that the tail is short and the
// per-CTA set-up (matrix pool, lookup tables, rho flush) is amortised over at least 8 tiles
bool fused_uses_pipe(int K, int use_traj, int pipeline, int mat_count, int nsub, int rho_blocks, int nstages) {
    return use_traj && pipeline == 1 && K == 11 && fused_pipe_smem_bytes(mat_count, nsub, rho_blocks, nstages) <= 220 * 1024;
}

bool fused_uses_direct(int K, int use_traj, int pipeline, int all_simple, int mat_count, int nsub, int rho_blocks) {
    return use_traj && pipeline == 3 && K == 11 && all_simple && nsub <= QGT_DIRECT_MAX_SUBS &&
           fused_direct_smem_bytes(mat_count, nsub, rho_blocks) <= 110 * 1024;
}

// kind: 0 = generic 8-warp kernel (2 CTAs per SM), 1 = persistent pipelined kernel (1 CTA per SM), 2 = direct kernel (3 CTAs per SM)
void fused_geometry(uint64_t ntiles, int nitems, int num_sms, int kind, bool use_traj, int* tiles_per_cta, int* tile_groups) {
    const uint64_t items = (uint64_t)(nitems > 0 ? nitems : 1);
    const uint64_t work = ntiles * items;
    uint64_t tpc;
    if (kind == 2 && work >= (uint64_t)num_sms * 3 * 5 * 8) {
        // all CTAs of a launch do the same amount of work, so they finish wave by wave: a whole number of waves of resident
        // CTAs (3 per SM) leaves no SM idle in the last one.  Many short waves beat few long ones (28 qubits, 29 columns:
        // 5 waves 9.93 s, 16 waves 9.43 s, 64-128 waves 9.23 s, 512 waves 9.41 s per evaluation): concurrently running CTAs then
        // work on a narrow band of tiles and share phi's trajectory images in L2.  At least 16 tiles per CTA pay for its
        // prologue (stage matrices and lookup tables into shared memory).
        static int waves = 0;
        if (!waves) { const char* e = std::getenv("QGT_B200_DIRECT_WAVES"); waves = e ? std::atoi(e) : 64; if (waves < 1) waves = 64; }
        const uint64_t slots = (uint64_t)num_sms * 3 * (uint64_t)waves;
        uint64_t tg = slots / items;
        if (tg < 1) tg = 1;
        tpc = (ntiles + tg - 1) / tg;
        if (tpc < 16) tpc = 16;
    } else {
        // the persistent kernel keeps one CTA per SM: fewer, longer CTAs (its prologue is paid per CTA); the adjoint gradient's
        // single pair item (no trajectory): every CTA leaves a partial transition-matrix buffer behind that the reduce kernel
        // has to walk, so no more CTAs than two waves of resident ones
        // generic kernel: 64 waves as well (30 qubits, 6 columns per launch: 8 waves 74.1 s, 32 waves 71.0 s, 128 waves 70.2 s)
        static int gwaves = 0;
        if (!gwaves) { const char* e = std::getenv("QGT_B200_GENERIC_WAVES"); gwaves = e ? std::atoi(e) : 64; if (gwaves < 1) gwaves = 64; }
        const uint64_t target = kind == 1 ? (uint64_t)num_sms * 6 : (nitems <= 2 && !use_traj) ? (uint64_t)num_sms * 4 : (uint64_t)num_sms * 2 * (uint64_t)gwaves;
        tpc = (work + target - 1) / target;
        if (tpc < 8) {
            // at least 8 tiles per CTA amortise its prologue - unless that leaves SMs idle (few items on a small state)
            const uint64_t fill = (work + (uint64_t)num_sms * 2 - 1) / ((uint64_t)num_sms * 2);
            tpc = fill < 8 ? (fill < 1 ? 1 : fill) : 8;
        }
    }
    if (tpc < 1) tpc = 1;
    if (tpc > ntiles) tpc = ntiles;
    *tiles_per_cta = (int)tpc;
    *tile_groups = (int)((ntiles + tpc - 1) / tpc);
}

size_t fused_smem_bytes(int K, int mat_count, int nsub, int rho_blocks, int nstages) {
    const int T = 1 << (K - 3);
    const int nwarps = T / 32 > 0 ? T / 32 : 1;
    return (sizeof(cplx) << K) * 2 + sizeof(cplx) * (size_t)mat_count + sizeof(cplx) * (size_t)(QGT_VARIANT_STRIDE(8) << QGT_MAX_VARIANT_BITS) +
           (size_t)nsub * (sizeof(QgtDevSubPass) + QGT_FAST_BYTES_PER_SUB) + 16 + (size_t)rho_blocks * 128 * sizeof(double) +
           (size_t)2 * nwarps * 128 * sizeof(double) + 16 * sizeof(int) + (size_t)nstages * sizeof(QgtDevStage);
}

cudaError_t launch_fused(const FusedLaunch& a, int K, int mat_count, int nsub, int rho_blocks, int nstages, cudaStream_t st) {
    if (K < 8 || K > 11) return cudaErrorInvalidValue;
    if (a.nitems <= 0 || a.ntiles == 0) return cudaSuccess;
    const size_t smem = fused_smem_bytes(K, mat_count, nsub, rho_blocks, nstages);
    if (smem > 200 * 1024) return cudaErrorInvalidValue;
    const unsigned grid = (unsigned)a.nitems * (unsigned)a.tile_groups;
    static bool attr_set = false;
    if (!attr_set) {
        cudaError_t e = cudaFuncSetAttribute(qgt_fused_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
        if (e == cudaSuccess) e = cudaFuncSetAttribute(qgt_fused_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
        if (e != cudaSuccess) return e;
        attr_set = true;
    }
    if (fused_uses_direct(K, a.use_traj, a.pipeline, a.all_simple, mat_count, nsub, rho_blocks)) {
        const size_t dsmem = fused_direct_smem_bytes(mat_count, nsub, rho_blocks);
        static bool dattr = false;
        if (!dattr) {
            cudaError_t e = cudaFuncSetAttribute(qgt_fused_direct_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 110 * 1024);
            if (e != cudaSuccess) return e;
            dattr = true;
        }
        qgt_fused_direct_kernel<<<(unsigned)a.nitems * (unsigned)a.tile_groups, 256, dsmem, st>>>(a);
        return cudaGetLastError();
    }
    if (a.use_traj && a.pipeline == 2 && K == 11 && a.all_simple && fused_lean_smem_bytes(mat_count, nsub, rho_blocks) <= 112 * 1024) {
        const size_t lsmem = fused_lean_smem_bytes(mat_count, nsub, rho_blocks);
        static bool lattr = false;
        if (!lattr) {
            cudaError_t e = cudaFuncSetAttribute(qgt_fused_lean_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 112 * 1024);
            if (e != cudaSuccess) return e;
            lattr = true;
        }
        qgt_fused_lean_kernel<<<grid, 512, lsmem, st>>>(a);
        return cudaGetLastError();
    }
    if (fused_uses_pipe(K, a.use_traj, a.pipeline, mat_count, nsub, rho_blocks, nstages)) {
        const size_t psmem = fused_pipe_smem_bytes(mat_count, nsub, rho_blocks, nstages);
        {
            static bool pattr = false;
            if (!pattr) {
                cudaError_t e = cudaFuncSetAttribute(qgt_fused_pipe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024);
                if (e != cudaSuccess) return e;
                pattr = true;
            }
            qgt_fused_pipe_kernel<<<grid, 512, psmem, st>>>(a);
            return cudaGetLastError();
        }
    }
    if (a.use_traj) qgt_fused_kernel<true><<<grid, 1 << (K - 3), smem, st>>>(a);
    else qgt_fused_kernel<false><<<grid, 1 << (K - 3), smem, st>>>(a);
    return cudaGetLastError();
}

// rho[item][i] = sum over tile chunks of partial[chunk][item][i], fixed order
__global__ void qgt_rho_reduce_kernel(const double* partial, int groups, int nitems, int per_item, double* rho) {
    const size_t total = (size_t)nitems * per_item;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
        double s = 0.0;
        for (int g = 0; g < groups; ++g) s += partial[(size_t)g * total + i];
        rho[i] = s;
    }
}

// few outputs, many partials (the adjoint gradient's single pair item): one warp per output element, the lanes walk the
// CTA partials with stride 32 and are combined in a fixed order, so the result is still reproducible
__global__ void __launch_bounds__(256) qgt_rho_reduce_warp_kernel(const double* partial, int groups, size_t total, double* rho) {
    const size_t i = (size_t)blockIdx.x * 8 + (threadIdx.x >> 5);
    if (i >= total) return;
    const int lane = threadIdx.x & 31;
    double s = 0.0;
    for (int g = lane; g < groups; g += 32) s += partial[(size_t)g * total + i];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_down_sync(0xffffffffu, s, o);
    if (lane == 0) rho[i] = s;
}

__global__ void qgt_add_doubles_kernel(double* dst, const double* src, size_t n) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) dst[i] += src[i];
}

cudaError_t launch_add_doubles(double* dst, const double* src, size_t n, cudaStream_t st) {
    if (n == 0) return cudaSuccess;
    qgt_add_doubles_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(dst, src, n);
    return cudaGetLastError();
}

cudaError_t launch_rho_reduce(const double* partial, int groups, int nitems, int per_item, double* rho, cudaStream_t st) {
    const size_t total = (size_t)nitems * per_item;
    if (total == 0) return cudaSuccess;
    if (total <= 16384 && groups >= 64) {
        qgt_rho_reduce_warp_kernel<<<(unsigned)((total + 7) / 8), 256, 0, st>>>(partial, groups, total, rho);
        return cudaGetLastError();
    }
    const unsigned grid = (unsigned)((total + 255) / 256 < 148 * 8 ? (total + 255) / 256 : 148 * 8);
    qgt_rho_reduce_kernel<<<grid, 256, 0, st>>>(partial, groups, nitems, per_item, rho);
    return cudaGetLastError();
}

// One 64-thread block per output element A[out]: sum over the group's entries (item, stage) and the stage's variant
// blocks of  sum_e X[e] * rho[e]  (complex, both in the C-fragment order of a transition-matrix block).
__global__ void __launch_bounds__(64) qgt_rho_contract_kernel(const double* rho, int per_item, const double* xpool,
                                                              const QgtContractGroup* groups, const QgtContractEntry* entries, cplx* A) {
    const QgtContractGroup grp = groups[blockIdx.x];
    const int e = threadIdx.x;                       // complex element: lane = e >> 1, column parity j = e & 1
    const int off = (e >> 1) * 4 + (e & 1);
    double sr = 0.0, si = 0.0;
    for (int i = grp.begin; i < grp.end; ++i) {
        const QgtContractEntry en = entries[i];
        const double* rb = rho + (size_t)en.item * per_item + (size_t)en.rho_off * 128;
        const double* xb = xpool + (size_t)en.x_off * 128;
        for (int v = 0; v < en.nblocks; ++v) {
            const double rr = rb[v * 128 + off], ri = rb[v * 128 + off + 2];
            const double xr = xb[v * 128 + off], xi = xb[v * 128 + off + 2];
            sr += xr * rr - xi * ri;
            si += xr * ri + xi * rr;
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) { sr += __shfl_xor_sync(0xffffffffu, sr, o); si += __shfl_xor_sync(0xffffffffu, si, o); }
    __shared__ double w[4];
    if ((e & 31) == 0) { w[(e >> 5) * 2] = sr; w[(e >> 5) * 2 + 1] = si; }
    __syncthreads();
    if (e == 0) {
        cplx z = A[grp.out];
        z.x += w[0] + w[2];
        z.y += w[1] + w[3];
        A[grp.out] = z;
    }
}

cudaError_t launch_rho_contract(const double* rho, int per_item, const double* xpool, const QgtContractGroup* groups, int ngroups,
                                const QgtContractEntry* entries, cplx* A, cudaStream_t st) {
    if (ngroups <= 0) return cudaSuccess;
    qgt_rho_contract_kernel<<<ngroups, 64, 0, st>>>(rho, per_item, xpool, groups, entries, A);
    return cudaGetLastError();
}

}  // namespace qgt

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
#include "kernels.cuh"
#include "mma_common.cuh"

namespace qgt {

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
// (row c = lane>>2), l the B operand (column a = lane>>2), the vector index lane&3 is summed over.
// r[0], r[1] = Re rho[c][2(lane&3)], [.. + 1];  r[2], r[3] = Im.
__device__ __forceinline__ void qgt_rho4(double (&r)[4], const cplx& p, const cplx& l) {
    dmma884(r[0], r[1], p.x, l.x);
    dmma884(r[2], r[3], p.y, l.x);
    dmma884(r[0], r[1], p.y, l.y);
    dmma884(r[2], r[3], -p.x, l.y);
}

__device__ __forceinline__ cplx qgt_cmul(const cplx& a, const cplx& b) {
    cplx o; o.x = a.x * b.x - a.y * b.y; o.y = a.x * b.y + a.y * b.x; return o;
}

// Shared memory: [tile A: the column][tile B: phi][matrix pool of the run][override matrices of the CTA's item]
//                [sub-pass descriptors][lookup tables][rho accumulators: rho_blocks x 128 doubles]
//                [scratch: 2 x (warps) x 128 doubles]
// A CTA owns ONE item and a contiguous chunk of tiles (blockIdx = chunk * nitems + item, so CTAs that run together
// read the same phi tiles and find them in L2); its transition matrices stay in shared memory until the end.
__global__ void __launch_bounds__(256, 2) qgt_fused_kernel(FusedLaunch a) {
    constexpr int N = 8;
    constexpr int OVR_ELEMS = QGT_VARIANT_STRIDE(N) << QGT_MAX_VARIANT_BITS;
    extern __shared__ __align__(16) unsigned char qgt_fsmem_raw[];
    __shared__ QgtDevRun run;
    const int tid = threadIdx.x, T = blockDim.x, lane = tid & 31, warp = tid >> 5, nwarps = T >> 5;
    for (int i = tid; i < (int)(sizeof(QgtDevRun) / 4); i += T)
        reinterpret_cast<uint32_t*>(&run)[i] = reinterpret_cast<const uint32_t*>(&a.runs[a.run_idx])[i];
    __syncthreads();
    const int item_idx = (int)(blockIdx.x % (unsigned)a.nitems);
    const uint64_t chunk = blockIdx.x / (unsigned)a.nitems;
    const QgtSweepItem& it = a.items[item_idx];
    const bool self = it.self != 0;
    const int rho_from = it.rho_from;
    const bool use_b = !self && rho_from <= run.last_rho_stage;      // phi is only needed while transition matrices remain to be taken

    cplx* tileA = reinterpret_cast<cplx*>(qgt_fsmem_raw);
    cplx* tileB = tileA + ((size_t)1 << run.K);
    cplx* spool = tileB + ((size_t)1 << run.K);
    cplx* sovr = spool + run.mat_count;
    QgtDevSubPass* subs = reinterpret_cast<QgtDevSubPass*>(sovr + OVR_ELEMS);
    QgtFastSub* fast = reinterpret_cast<QgtFastSub*>(subs + run.nsub);
    QgtFastWarp* fwarp = reinterpret_cast<QgtFastWarp*>(fast + run.nsub);
    uint32_t* flane = reinterpret_cast<uint32_t*>(fwarp + 8 * run.nsub);
    double* rho_acc = reinterpret_cast<double*>((reinterpret_cast<uintptr_t>(flane + 32 * run.nsub) + 15) & ~(uintptr_t)15);
    double* scratch = rho_acc + (size_t)run.rho_blocks * 128;
    {
        const cplx* gpool = a.pool + run.mat_off;
        for (int i = tid; i < run.mat_count; i += T) spool[i] = gpool[i];
        const uint32_t* gs = reinterpret_cast<const uint32_t*>(a.subs + run.sub_off);
        uint32_t* ss = reinterpret_cast<uint32_t*>(subs);
        for (int i = tid; i < run.nsub * (int)(sizeof(QgtDevSubPass) / 4); i += T) ss[i] = gs[i];
        for (int i = tid; i < run.rho_blocks * 128; i += T) rho_acc[i] = 0.0;
        if (it.ovr_kind == 1) {
            const cplx* g = reinterpret_cast<const cplx*>(it.ovr_mat);
            const int cnt = QGT_VARIANT_STRIDE(N) << a.stages[run.stage_off + it.ovr_index].nvar;
            for (int i = tid; i < cnt && i < OVR_ELEMS; i += T) sovr[i] = g[i];
        }
    }
    __syncthreads();
    const QgtDevStage* stages = a.stages + run.stage_off;
    const QgtDevThrDiag* tdiags = a.tdiags + run.tdiag_off;
    qgt_fast_build(run, subs, stages, fast, fwarp, flane, tid, T);
    __syncthreads();

    const QgtIoMap<3> io = qgt_make_iomap<3>(run, tid);
    const int q = lane >> 2, k = lane & 3;
    const uint64_t tau0 = chunk * (uint64_t)a.tiles_per_cta;
    const uint64_t tau1 = tau0 + (uint64_t)a.tiles_per_cta < a.ntiles ? tau0 + (uint64_t)a.tiles_per_cta : a.ntiles;
    int par = 0;                                     // scratch buffer the next transition matrix goes to
    for (uint64_t tau = tau0; tau < tau1; ++tau) {
        const uint64_t tilebase = qgt_tile_base(run, tau);
        const uint64_t tileg = tilebase | a.gprefix;
        {
            const cplx* srcA = reinterpret_cast<const cplx*>(it.src);
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const uint32_t idx = (uint32_t)tid + (uint32_t)i * (uint32_t)T;
                const uint64_t g = tilebase | qgt_io_offset<3>(io, i);
                cp_async16(tileA + qgt_swz(idx), srcA + g, 16);
                if (use_b) cp_async16(tileB + qgt_swz(idx), a.phi + g, 16);
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
            // thread diagonals of the sub-pass (parameter-free by construction): one pending phase per result slot,
            // applied with the last stage to BOTH tiles (it cancels inside rho)
            const bool has_tdiag = sp.tdiag_end > sp.tdiag_begin;
            cplx pend0[4], pend1[4];
            if (has_tdiag) {
#pragma unroll
                for (int g = 0; g < 4; ++g) {
                    pend0[g].x = 1.0; pend0[g].y = 0.0; pend1[g] = pend0[g];
                    uint64_t gc0 = gwarp;
                    const int t0 = g * 8 + 2 * k;
#pragma unroll
                    for (int i = 0; i < 5; ++i) if ((t0 >> i) & 1) gc0 |= sp.g_thr[i];
                    const uint64_t gc1 = gc0 | sp.g_thr[0];
                    for (int t = sp.tdiag_begin; t < sp.tdiag_end; ++t) {
                        qgt_thread_diag(pend0[g], tdiags[t], gc0);
                        qgt_thread_diag(pend1[g], tdiags[t], gc1);
                    }
                }
            }
            const int nstage = sp.stage_end - sp.stage_begin;
            if (nstage == 0) {
                if (has_tdiag) {
#pragma unroll
                    for (int g = 0; g < 4; ++g) {
                        const uint32_t gx = ((g & 1) ? gx1 : 0u) ^ ((g & 2) ? gx2 : 0u);
                        tileA[baseC ^ gx] = qgt_cmul(pend0[g], tileA[baseC ^ gx]);
                        tileA[baseC ^ gx ^ st0] = qgt_cmul(pend1[g], tileA[baseC ^ gx ^ st0]);
                        if (use_b) {
                            tileB[baseC ^ gx] = qgt_cmul(pend0[g], tileB[baseC ^ gx]);
                            tileB[baseC ^ gx ^ st0] = qgt_cmul(pend1[g], tileB[baseC ^ gx ^ st0]);
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
                const cplx* Ma = ovr ? sovr + var * QGT_VARIANT_STRIDE(N) : Mb;
                const StageFrag fb = qgt_load_frag(Mb, st.form == QGT_FORM_DIAG_REAL, lane);
                const StageFrag fa = ovr ? qgt_load_frag(Ma, it.ovr_form == QGT_FORM_DIAG_REAL, lane) : fb;
                const bool last = (sg == sp.stage_end - 1);
                const bool do_rho = st.rho_off >= 0 && sg >= rho_from;
                const bool need_b = use_b && sg <= run.last_rho_stage;
                double r[4] = {0.0, 0.0, 0.0, 0.0};
#pragma unroll
                for (int g = 0; g < 4; ++g) {
                    const uint32_t gx = ((g & 1) ? gx1 : 0u) ^ ((g & 2) ? gx2 : 0u);
                    const cplx va0 = tileA[baseB ^ gx], va1 = tileA[baseB ^ gx ^ sr2];
                    cplx vb0, vb1;
                    if (need_b) { vb0 = tileB[baseB ^ gx]; vb1 = tileB[baseB ^ gx ^ sr2]; }
                    __syncwarp();                 // every lane has read the group's slots before any is overwritten
                    cplx a0, a1, b0, b1;
                    qgt_apply8(fa, va0, va1, a0, a1);
                    if (last && has_tdiag) { a0 = qgt_cmul(pend0[g], a0); a1 = qgt_cmul(pend1[g], a1); }
                    tileA[baseC ^ gx] = a0;
                    tileA[baseC ^ gx ^ st0] = a1;
                    if (need_b) {
                        qgt_apply8(fb, vb0, vb1, b0, b1);
                        if (last && has_tdiag) { b0 = qgt_cmul(pend0[g], b0); b1 = qgt_cmul(pend1[g], b1); }
                        tileB[baseC ^ gx] = b0;
                        tileB[baseC ^ gx ^ st0] = b1;
                    }
                    if (do_rho) {
                        if (self) { qgt_rho4(r, a0, a0); qgt_rho4(r, a1, a1); }
                        else { qgt_rho4(r, b0, a0); qgt_rho4(r, b1, a1); }
                    }
                }
                if (do_rho) {
                    double* sc = scratch + ((size_t)par * nwarps + warp) * 128 + lane * 4;
                    *reinterpret_cast<double2*>(sc) = make_double2(r[0], r[1]);
                    *reinterpret_cast<double2*>(sc + 2) = make_double2(r[2], r[3]);
                }
                __syncthreads();                  // next stage / sub-pass reads slots other warps' lanes wrote; scratch is complete
                if (do_rho) {
                    // fixed-order sum over the warps (deterministic); a warp's block is selected by its variant
                    for (int e = tid; e < 128; e += T) {
                        for (int w = 0; w < nwarps; ++w) {
                            const int vw = qgt_variant_index(st, tileg | fwarp[s * 8 + w].g);
                            rho_acc[(size_t)(st.rho_off + vw) * 128 + e] += scratch[((size_t)par * nwarps + w) * 128 + e];
                        }
                    }
                    par ^= 1;
                }
            }
        }
        qgt_phase_store<3>(io, tileA, reinterpret_cast<cplx*>(it.dst), tilebase, tid, T, it.accumulate != 0);
    }
    __syncthreads();
    double* out = a.rho_partial + (size_t)blockIdx.x * run.rho_blocks * 128;
    for (int i = tid; i < run.rho_blocks * 128; i += T) out[i] = rho_acc[i];
}

// tiles per CTA and number of tile chunks: ~8 waves of resident CTAs (2 per SM) so that the tail is short and the
// per-CTA set-up (matrix pool, lookup tables, rho flush) is amortised over at least 8 tiles
void fused_geometry(uint64_t ntiles, int nitems, int num_sms, int* tiles_per_cta, int* tile_groups) {
    const uint64_t target = (uint64_t)num_sms * 2 * 8;
    uint64_t tpc = (ntiles * (uint64_t)(nitems > 0 ? nitems : 1) + target - 1) / target;
    if (tpc < 8) tpc = 8;
    if (tpc > ntiles) tpc = ntiles;
    *tiles_per_cta = (int)tpc;
    *tile_groups = (int)((ntiles + tpc - 1) / tpc);
}

size_t fused_smem_bytes(int K, int mat_count, int nsub, int rho_blocks) {
    const int T = 1 << (K - 3);
    const int nwarps = T / 32 > 0 ? T / 32 : 1;
    return (sizeof(cplx) << K) * 2 + sizeof(cplx) * (size_t)mat_count + sizeof(cplx) * (size_t)(QGT_VARIANT_STRIDE(8) << QGT_MAX_VARIANT_BITS) +
           (size_t)nsub * (sizeof(QgtDevSubPass) + QGT_FAST_BYTES_PER_SUB) + 16 + (size_t)rho_blocks * 128 * sizeof(double) +
           (size_t)2 * nwarps * 128 * sizeof(double);
}

cudaError_t launch_fused(const FusedLaunch& a, int K, int mat_count, int nsub, int rho_blocks, cudaStream_t st) {
    if (K < 8 || K > 11) return cudaErrorInvalidValue;
    if (a.nitems <= 0 || a.ntiles == 0) return cudaSuccess;
    const size_t smem = fused_smem_bytes(K, mat_count, nsub, rho_blocks);
    if (smem > 200 * 1024) return cudaErrorInvalidValue;
    static bool attr_set = false;
    if (!attr_set) {
        cudaError_t e = cudaFuncSetAttribute(qgt_fused_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
        if (e != cudaSuccess) return e;
        attr_set = true;
    }
    const unsigned grid = (unsigned)a.nitems * (unsigned)a.tile_groups;
    qgt_fused_kernel<<<grid, 1 << (K - 3), smem, st>>>(a);
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

cudaError_t launch_rho_reduce(const double* partial, int groups, int nitems, int per_item, double* rho, cudaStream_t st) {
    const size_t total = (size_t)nitems * per_item;
    if (total == 0) return cudaSuccess;
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

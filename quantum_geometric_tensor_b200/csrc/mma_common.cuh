// mma_common.cuh - device helpers shared by the sweep kernels (kernels.cu) and the fused sweep + transition-matrix
// kernel (fused.cu): the FP64 tensor-pipe instruction, cp.async, and the per-CTA lookup tables of the tensor-only paths.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "dev_structs.h"
#include "sweep_core.cuh"

namespace qgt {

__device__ __forceinline__ void dmma884(double& c0, double& c1, double a, double b) {
    asm("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
        : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gmem_src, int src_bytes) {
    const unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" :: "r"(d), "l"(gmem_src), "r"(src_bytes) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" :: "n"(N) : "memory"); }

// Per-CTA lookup tables of the tensor-only kernel, built once per launch in shared memory: everything a
// sub-pass needs that depends only on (sub-pass, warp, lane), so the per-tile code is loads + DMMAs + stores.
struct QgtFastSub {
    uint64_t vm0, vm1;           // variant-selecting global index bits of the (single) stage, 0 when unused
    uint32_t mat_off, stage;     // variant 0 in the run's pool; stage index (override test)
    uint32_t sr2, st0, gx1, gx2; // slot XOR terms: matrix bit 2, thread bit 0, thread bits 3 and 4
    uint32_t simple, form;       // exactly one dense stage and no thread diagonal; QGT_FORM_* of that stage
    int32_t rho_off, pad;        // fused kernel: the stage's first transition-matrix block, -1 = none
};
struct QgtFastWarp { uint64_t g; uint32_t s; uint32_t pad; };   // warp-index bits: global index / swizzled slot
#define QGT_FAST_BYTES_PER_SUB (sizeof(QgtFastSub) + 8 * sizeof(QgtFastWarp) + 32 * sizeof(uint32_t))

__device__ __forceinline__ void qgt_fast_build(const QgtDevRun& run, const QgtDevSubPass* subs, const QgtDevStage* stages,
                                               QgtFastSub* fast, QgtFastWarp* fwarp, uint32_t* flane, int tid, int T) {
    const int nthr_bits = run.K - 3;
    for (int w = tid; w < run.nsub * 32; w += T) {
        const int s = w >> 5, lane = w & 31, q = lane >> 2, k = lane & 3;
        const QgtDevSubPass& sp = subs[s];
        const uint32_t b = ((q & 1) ? sp.s_thr[0] : 0u) ^ ((q & 2) ? sp.s_thr[1] : 0u) ^ ((q & 4) ? sp.s_thr[2] : 0u) ^
                           ((k & 1) ? sp.s_reg[0] : 0u) ^ ((k & 2) ? sp.s_reg[1] : 0u);
        const uint32_t c = ((k & 1) ? sp.s_thr[1] : 0u) ^ ((k & 2) ? sp.s_thr[2] : 0u) ^ ((q & 1) ? sp.s_reg[0] : 0u) ^
                           ((q & 2) ? sp.s_reg[1] : 0u) ^ ((q & 4) ? sp.s_reg[2] : 0u);
        flane[w] = b | (c << 16);
        if (lane < 8) {
            QgtFastWarp fw; fw.g = 0; fw.s = 0; fw.pad = 0;
            for (int i = 5; i < nthr_bits; ++i)
                if ((lane >> (i - 5)) & 1) { fw.s ^= sp.s_thr[i]; fw.g |= sp.g_thr[i]; }
            fwarp[s * 8 + lane] = fw;
        }
        if (lane == 8) {
            QgtFastSub f;
            f.simple = (sp.stage_end - sp.stage_begin == 1 && sp.tdiag_end == sp.tdiag_begin) ? 1u : 0u;
            f.stage = (uint32_t)sp.stage_begin; f.form = QGT_FORM_DENSE;
            f.vm0 = f.vm1 = 0; f.mat_off = 0; f.rho_off = -1; f.pad = 0;
            if (f.simple) {
                const QgtDevStage& st = stages[sp.stage_begin];
                f.mat_off = (uint32_t)st.mat_off;
                f.form = (uint32_t)st.form;
                f.rho_off = st.rho_off;
                if (st.nvar > 0) f.vm0 = st.vmask[0];
                if (st.nvar > 1) f.vm1 = st.vmask[1];
            }
            f.sr2 = sp.s_reg[2]; f.st0 = sp.s_thr[0]; f.gx1 = sp.s_thr[3]; f.gx2 = sp.s_thr[4];
            fast[s] = f;
        }
    }
}

}  // namespace qgt

// kernels.cuh — launch wrappers of the sm_100a kernels (definitions in kernels.cu).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "dev_structs.h"
#include "sweep_core.cuh"

namespace qgt {

struct SweepLaunch {
    const QgtDevRun* runs;
    const QgtDevSubPass* subs;
    const QgtDevStage* stages;
    const QgtDevThrDiag* tdiags;
    const QgtDevCost* costs;
    const cplx* pool;
    int run_idx;
    const QgtSweepItem* items;
    int nitems;
    uint64_t ntiles;
    int tiles_per_item;          // consecutive tiles a CTA processes per item set-up; power of two dividing ntiles
    uint64_t gprefix;            // rank bits of a sharded state, OR-ed into every global index used by masks
    int mma_only;                // every sub-pass of the run qualifies for the tensor-pipe path (host-checked)
    int double_buffer;           // tensor-only kernel: two tile buffers (3 CTAs/SM) instead of one (4 CTAs/SM)
    int debug_skip;              // timing experiments only: 1 no tile load, 2 no store, 4 no sub-passes
    int use_mma;                 // dense stages on the FP64 tensor pipe (DMMA) where the sub-pass allows it
    const double* cost_ein;      // ... and ein[2^K] behind them (derivative items need the energy itself)
    const cplx* cost_phase;      // runs with a cost pass: [cost op of the run][2^K] phases exp(-i angle ein[idx]) built before the
                                 // launch (launch_cost_phase_tables), or null = one sincos per amplitude
    QgtCostTable ct;
};

// fused sweep + transition-matrix launch (fused.cu): grid = nitems x tile_groups, CTA (chunk, item) handles tiles
// [chunk * tiles_per_cta, ...) of its item
struct FusedLaunch {
    const QgtDevRun* runs;
    const QgtDevSubPass* subs;
    const QgtDevStage* stages;
    const QgtDevThrDiag* tdiags;
    const cplx* pool;
    int run_idx;
    const QgtSweepItem* items;
    int nitems;
    const cplx* phi;             // the marching state at the start of the run (second tile of every non-self item)
    uint64_t ntiles;             // tiles of this launch: tile_off .. tile_off + ntiles - 1 of the state
    uint64_t tile_off;           // (a launch may cover one range of the tiles only: ranged trajectory mode, see capi.cu)
    int tiles_per_cta, tile_groups;
    uint64_t gprefix;
    double* rho_partial;         // [tile_groups][nitems][rho_blocks * 128]
    int use_traj;                // phi's tiles after every transition-matrix stage come from the trajectory columns (written
                                 // by the launch of the self item alone) instead of being recomputed per item
    int debug;                   // timing experiments only (QGT_FDBG_* bits; results are wrong by construction)
    int pipeline;                // trajectory mode, K = 11: 1 = the persistent 16-warp kernel with double-buffered tiles,
                                 // 2 = the lean 2 x 16-warp kernel (runs whose sub-passes are all one stage, no thread diagonal)
    int all_simple;              // every sub-pass of the run is exactly one dense stage without thread diagonals
    cplx* traj[QGT_MAX_TRAJ];    // trajectory images: [ntiles][2^K] images of the swizzled tile, indexed from the launch's first tile
};

struct GramLaunch {
    const cplx* const* a_ptrs;   // device array of na column pointers
    const cplx* const* b_ptrs;   // device array of nb column pointers
    int na, nb;
    uint64_t D;                  // amplitudes per column (local shard length)
    int ksplit;
    int mtiles, ntiles;
    int symmetric;               // b list starts with the a list: tiles strictly below the diagonal are skipped
    int nb_main;                 // columns of B covered by the tile grid (nb - nstrip)
    int nstrip;                  // symmetric 32x32 launches: the last <= 8 columns of B (the projection column psi) are
                                 // contracted by the two warps that idle in diagonal tiles instead of a padded tile column
    int npad;                    // row stride of `partial`: ntiles*NT (+ 8 strip columns)
    cplx* partial;               // [ksplit][mtiles*MT][npad]
};

struct GramShape { int MT, NT; int thin; };   // thin: few pairs, register accumulators, HBM-bound (MT x NT = operand slots)

// K, R: tile / register qubits of the run; grid is chosen inside
cudaError_t launch_sweep(const SweepLaunch& a, int K, int R, int B, int mat_count, int nsub, int has_cost, int num_sms, cudaStream_t st);
cudaError_t launch_cost_phase_tables(const SweepLaunch& a, int K, int ncost, cplx* out, cudaStream_t st);

bool fused_uses_pipe(int K, int use_traj, int pipeline, int mat_count, int nsub, int rho_blocks, int nstages);
bool fused_uses_direct(int K, int use_traj, int pipeline, int all_simple, int mat_count, int nsub, int rho_blocks);
// kind: 0 generic kernel, 1 persistent pipelined kernel, 2 direct kernel
void fused_geometry(uint64_t ntiles, int nitems, int num_sms, int kind, bool use_traj, int* tiles_per_cta, int* tile_groups);
size_t fused_smem_bytes(int K, int mat_count, int nsub, int rho_blocks, int nstages);
cudaError_t launch_fused(const FusedLaunch& a, int K, int mat_count, int nsub, int rho_blocks, int nstages, cudaStream_t st);
cudaError_t launch_rho_reduce(const double* partial, int groups, int nitems, int per_item, double* rho, cudaStream_t st);
cudaError_t launch_add_doubles(double* dst, const double* src, size_t n, cudaStream_t st);      // dst[i] += src[i]
cudaError_t launch_rho_contract(const double* rho, int per_item, const double* xpool, const QgtContractGroup* groups, int ngroups,
                                const QgtContractEntry* entries, cplx* A, cudaStream_t st);

GramShape gram_shape(int na, int nb);
// fills mtiles / ntiles / nb_main / nstrip / npad from na, nb, symmetric; returns the partial elements per k-split
size_t gram_configure(GramLaunch& g, GramShape shp);
// split-K factor of a configured launch (needs na, nb, D, symmetric and the fields gram_configure fills)
int gram_choose_ksplit(const GramLaunch& g, GramShape shp, int num_sms);
void set_gram_tile_override(int t);   // tuning: 0 = automatic, 32 or 64 = force that square tile
cudaError_t launch_gram(const GramLaunch& g, GramShape shp, cudaStream_t st);
// C[(a_ids[i]), (b_ids[j])] = sum_ks partial ; mirrored conj ; ldc = leading dimension of C
cudaError_t launch_gram_reduce(const GramLaunch& g, GramShape shp, const int* a_ids, const int* b_ids,
                               cplx* C, int ldc, cudaStream_t st);
// Q = C[0:P,0:P] - v v^H with v = C[:,P];  any output may be null
cudaError_t launch_finalize(const cplx* C, int P, double* metric, double* berry, cplx* q_full, cudaStream_t st);

// FP64 tensor-pipe (DMMA.8x8x4) peak of this device, best of 5 timed launches
cudaError_t measure_dmma_peak(int num_sms, double* scratch, cudaStream_t st, double* tflops);

cudaError_t launch_init_state(cplx* dst, uint64_t D, int initial_state, double plus_amp, uint64_t global_offset, cudaStream_t st);
cudaError_t launch_norm2(const cplx* src, uint64_t D, double* out /*device, zeroed inside*/, cudaStream_t st);
cudaError_t launch_axpy(cplx* dst, const cplx* src, uint64_t D, double ar, double ai, cudaStream_t st);   // dst += a*src
// expectation of the diagonal cost:  out[0] = sum |psi|^2 E ;  out2 = sum conj(a) b E  (re, im)
cudaError_t launch_cost_apply(cplx* dst, const cplx* src, uint64_t D, QgtCostTable ct, uint64_t global_offset, cudaStream_t st);   // dst = E_z * src
cudaError_t launch_cost_dot(const cplx* a, const cplx* b, uint64_t D, QgtCostTable ct, uint64_t global_offset,
                            double* out2 /*device, 2 doubles, zeroed inside*/, cudaStream_t st);

}  // namespace qgt

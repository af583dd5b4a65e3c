// ctx.hpp — internal definitions behind the opaque handles of include/qgt_b200.h
#pragma once
#include <cuda_runtime.h>

#include <string>
#include <vector>

#include "../../include/qgt_b200.h"
#include "kernels.cuh"
#include "plan.hpp"

namespace qgt {

extern thread_local std::string g_last_error;
int fail(int status, const std::string& msg);
int cuda_fail(cudaError_t e, const char* what);

// grow-only device buffer cached in the context
struct DevBuf {
    void* ptr = nullptr;
    size_t bytes = 0;
    int reserve(size_t n);
    void release();
};

// CUDA-event stopwatch on the library's stream; categories 0 = sweep, 1 = gram / contraction, 2 = other, 3 = exchange
struct Timer {
    bool enabled = true;
    std::vector<cudaEvent_t> events;
    std::vector<int> cats;
    std::vector<std::string> labels;
    bool trace = false;          // QGT_B200_TRACE=1: print every launch with its device time to stderr
    size_t used = 0;
    void begin(cudaStream_t st, int cat, const char* label = nullptr);
    void end(cudaStream_t st);
    void collect(double ms[4]);
    ~Timer();
};

struct DistState;   // dist.cu

// host side of a fused evaluation: what the final assembly needs besides the device results
struct FusedHost {
    struct StageGen {
        int run = 0, rho_off = 0, nvar = 0;
        std::vector<int> params;
        std::vector<double> gens;           // stage_generators(): [param][variant][a][c] complex
    };
    std::vector<StageGen> stages;           // every stage with parameters of the runs whose self transition matrices were taken
    std::vector<size_t> self_off;           // per run: offset (doubles) of its self transition matrices in rho_self, npos = none
    size_t self_doubles = 0;
};

}  // namespace qgt

struct qgt_b200_ctx {
    int device = 0;
    int num_sms = 148;
    cudaStream_t stream = nullptr;
    cudaStream_t side_stream = nullptr;      // ranged trajectory mode: phi's launch of the next range runs here, next to the columns' launch
    cudaEvent_t ev_side[2] = {nullptr, nullptr}, ev_main[2] = {nullptr, nullptr}, ev_group = nullptr;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    size_t ws_limit = 0;
    size_t max_slots = 0;
    int use_mma = 1;
    int double_buffer = 0;
    int tiles_per_item = 0;      // 0 = automatic
    int debug_skip = 0;          // timing experiments only (results are wrong): 1 no tile load, 2 no store, 4 no sub-passes
    double ms_prog_pack = 0.0;
    qgt::PlanOptions opt;
    qgt::DevBuf arena, img_runs, img_subs, img_stages, img_tdiags, img_costs, img_pool, ovr_pool, items, aux, partial, cmat, outbuf, edges, vweights, scratch;
    qgt::DevBuf cost_phase;                               // per-launch phase tables of a run's cost ops
    int cost_tables = 1;                                  // 0 = one sincos per amplitude in the cost pass (round-1 form)
    qgt::DevBuf partial_side, rho_side;                   // the side stream's own partial sums and reduced transition matrices
    qgt::DevBuf fx_pool, fx_tab, rho, rho_self, amat;     // fused schedule: evolved generators, contraction tables, transition matrices, A
    qgt::FusedHost fused_host;
    int psi_phys_slot = 0;       // arena column holding the program's psi slot after run_program (exchanges rotate columns through a spare)
    std::vector<int> img_stage_form, img_run_stage_off;   // QGT_FORM_* of every stage of the uploaded plan (flop accounting)
    int fused_overlap = 1;       // ranged trajectory mode: phi's launch of the next range on the side stream (0 = everything in order)
    int fused_traj = -1;         // trajectory mode of the fused schedule: -1 automatic, 0 never, 1 whenever it fits
    int fused_debug = 0;         // timing experiments only
    int fused_pipeline = 3;      // trajectory mode at K = 11: 3 = direct kernel (fragment-order trajectory, 3 CTAs x 8 warps) where the
                                 // run qualifies, 2 = lean 2 x 16-warp kernel, 1 = persistent double-buffered 16-warp kernel,
                                 // 0 = the generic 8-warp kernel
    int fused_mode = -1;         // -1 automatic (fused when the columns do not all fit), 0 never, 1 whenever the plan qualifies
    void* pinned = nullptr;
    QgtCostTable cost = {nullptr, 0, nullptr, 0};
    std::vector<QgtCostTable> seg_cost;      // sharded states: one cost table per mapped segment (remapped qubits)
    qgt::Timer timer;
    qgt_b200_stats stats = {};
    // multi-GPU
    int rank = 0, world = 1;
    qgt::DistState* dist = nullptr;
};

struct qgt_b200_state {
    qgt_b200_ctx* ctx = nullptr;
    int n = 0;          // global qubits
    int nloc = 0;       // qubits held locally (n - log2 world)
    uint64_t D = 0;     // local amplitudes
    cplx* d = nullptr;
    bool owns = false;
};

namespace qgt {

// executor pieces shared with dist.cu
int upload_plan(qgt_b200_ctx* c, const qgt_b200_circuit& circ, const CircuitPlan& plan, PlanImage& img);
int upload_cost_table(qgt_b200_ctx* c, const qgt_b200_circuit& circ);
int apply_plan_inplace(qgt_b200_ctx* c, const CircuitPlan& plan, cplx* d, uint64_t D);
int run_program(qgt_b200_ctx* c, const qgt_b200_circuit& circ, const CircuitPlan& plan, const Program& prog,
                cplx* arena, uint64_t D, cplx* cmat);
size_t workspace_slots(qgt_b200_ctx* c, uint64_t D, size_t reserve_bytes);
// fused schedule or the Gram schedule for this plan and slot count (option "fused")
bool choose_fused(const qgt_b200_ctx* c, const CircuitPlan& plan, size_t slots);
// after run_program of a fused program: (allreduce over ranks,) download A and the self transition matrices and
// assemble metric / Berry curvature / full Q on the host; any output may be null
int fused_finish(qgt_b200_ctx* c, const CircuitPlan& plan, double* metric, double* berry, double* q_full);
void stats_begin(qgt_b200_ctx* c);
int stats_end(qgt_b200_ctx* c);

// dist.cu
void dist_shutdown(qgt_b200_ctx* c);
int dist_allreduce_host(qgt_b200_ctx* c, double* v, int n);   // sum over ranks, no-op for world == 1
int dist_apply_circuit(qgt_b200_state* s, const qgt_b200_circuit* circ, const double* theta);
int dist_exchange(qgt_b200_ctx* c, cplx* col, uint64_t D, unsigned mask);   // grouped qubit exchange of one column, in place (via scratch)
int dist_exchange_multi(qgt_b200_ctx* c, const cplx* src, cplx* dst, uint64_t D, unsigned mask);   // the same, out of place, no extra copy
int dist_allreduce_device(qgt_b200_ctx* c, double* d_buf, size_t count);   // in place, stream ordered
int upload_segment_costs(qgt_b200_ctx* c, const qgt_b200_circuit& circ, const std::vector<MappedSegment>& segs);   // per-segment cost tables (remapped qubits)
int dist_qgt(qgt_b200_ctx* c, const qgt_b200_circuit* circ, const double* theta,
             double* metric, double* berry, double* q_full, qgt_b200_state* psi_out);

}  // namespace qgt

// emul.cpp — host emulation of the CUDA sweep kernel.   TEST INFRASTRUCTURE ONLY.
//
// Compiles quantum_geometric_tensor_b200/csrc/sweep_core.cuh (the per-thread phases the kernel is
// made of) and plan.cpp for the host and runs one sweep launch item with threads as a loop and
// __syncthreads() as phase boundaries.  This checks, without a GPU, the parts of the kernel that are
// pure index arithmetic: tile/base deposit, swizzle, thread-bit permutation, register-qubit binding
// of controls/parities, the derivative override.  It is never part of the product.
#include <cstring>
#include <string>
#include <vector>

#include "../../quantum_geometric_tensor_b200/csrc/plan.hpp"
#include "../../quantum_geometric_tensor_b200/csrc/sweep_core.cuh"

using namespace qgt;

template <int R>
static void emul_item(const QgtDevRun& run, const QgtDevSubPass* subs, const QgtDevOp* ops, const QgtSweepItem& it,
                      const QgtCostTable& ct, uint64_t D) {
    const int T = 1 << (run.K - R);
    const uint64_t ntiles = D >> run.K;
    std::vector<cplx> tile((size_t)1 << run.K);
    for (uint64_t tau = 0; tau < ntiles; tau++) {
        const uint64_t tilebase = qgt_tile_base(run, tau);
        for (int tid = 0; tid < T; tid++) {
            const QgtIoMap<R> io = qgt_make_iomap<R>(run, tid);
            qgt_phase_load<R>(io, tile.data(), (const cplx*)it.src, tilebase, tid, T);
        }
        for (int s = 0; s < run.nsub; s++)
            for (int tid = 0; tid < T; tid++)
                qgt_phase_subpass<R>(run, subs[s], ops, it.ovr_op, it.ovr, tile.data(), tilebase, tid, ct);
        for (int tid = 0; tid < T; tid++) {
            const QgtIoMap<R> io = qgt_make_iomap<R>(run, tid);
            qgt_phase_store<R>(io, tile.data(), (cplx*)it.dst, tilebase, tid, T, it.accumulate != 0);
        }
    }
}

extern "C" int emul_num_runs(const qgt_b200_circuit* circ, const double* theta, int K, int R) {
    PlanOptions opt; opt.tile_qubits = K; opt.reg_qubits = R;
    CircuitPlan plan; std::string err;
    if (build_plan(*circ, theta, opt, plan, err)) return -1;
    return (int)plan.runs.size();
}

// one sweep item: dst (+)= run[run_idx](src) with op `ovr_op` replaced by its derivative (or -1)
extern "C" int emul_sweep(const qgt_b200_circuit* circ, const double* theta, int K, int R, int run_idx,
                          const double* src, double* dst, int ovr_op, int accumulate) {
    PlanOptions opt; opt.tile_qubits = K; opt.reg_qubits = R;
    CircuitPlan plan; std::string err;
    if (build_plan(*circ, theta, opt, plan, err)) return -1;
    if (run_idx < 0 || run_idx >= (int)plan.runs.size()) return -2;
    PlanImage img;
    build_image(plan, img);
    const Run& run = plan.runs[run_idx];
    const QgtDevRun& dr = img.runs[run_idx];
    QgtSweepItem it;
    std::memset(&it, 0, sizeof it);
    it.src = src; it.dst = dst; it.ovr_op = ovr_op; it.accumulate = accumulate ? 1u : 0u;
    if (ovr_op >= 0) {
        const int sp = find_subpass(run, ovr_op);
        if (sp < 0) return -3;
        it.ovr = bind_op(run, run.subs[sp], run.ops[ovr_op], true);
    }
    std::vector<QgtDevEdge> ed(circ->num_edges);
    for (size_t k = 0; k < circ->num_edges; k++) { ed[k].i = circ->edges[k].i; ed[k].j = circ->edges[k].j; ed[k].w = circ->edges[k].weight; }
    QgtCostTable ct{ed.data(), (int)ed.size(), circ->vertex_weights, circ->num_qubits};
    const uint64_t D = (uint64_t)1 << circ->num_qubits;
    const int Reff = run.subs.empty() ? R : (int)run.subs[0].reg_local.size();
    const QgtDevSubPass* subs = img.subs.data() + dr.sub_off;
    const QgtDevOp* ops = img.ops.data() + dr.ops_off;
    switch (Reff) {
    case 1: emul_item<1>(dr, subs, ops, it, ct, D); break;
    case 2: emul_item<2>(dr, subs, ops, it, ct, D); break;
    case 3: emul_item<3>(dr, subs, ops, it, ct, D); break;
    default: return -4;
    }
    return 0;
}

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

static int g_emul_batch = 0;
extern "C" void emul_set_batch(int b) { g_emul_batch = b; }
static int g_emul_cost_tables = 1;      // the cost pass through the phase tables (the device default) or one sincos per amplitude
extern "C" void emul_set_cost_tables(int v) { g_emul_cost_tables = v; }

template <int R, int B>
static void emul_item(const QgtDevRun& run, const PlanImage& img, const QgtSweepItem& it, const QgtCostTable& ct, uint64_t D) {
    const int T = 1 << (run.K - R - B);
    const uint64_t ntiles = D >> run.K;
    std::vector<cplx> tile((size_t)1 << run.K);
    const QgtDevSubPass* subs = img.subs.data() + run.sub_off;
    QgtSubCtx cx;
    cx.stages = img.stages.data() + run.stage_off;
    cx.tdiags = img.tdiags.data() + run.tdiag_off;
    // pool of the run followed by the item's override matrices, as the kernel lays them out in shared memory
    std::vector<cplx> pool(run.mat_count + (QGT_VARIANT_STRIDE(1 << R) << QGT_MAX_VARIANT_BITS));
    std::memcpy(pool.data(), reinterpret_cast<const cplx*>(img.pool.data()) + run.mat_off, run.mat_count * sizeof(cplx));
    if (it.ovr_kind == 1)
        std::memcpy(pool.data() + run.mat_count, it.ovr_mat, sizeof(cplx) * (QGT_VARIANT_STRIDE(1 << R) << cx.stages[it.ovr_index].nvar));
    cx.pool = pool.data();
    cx.ovr_mat_off = run.mat_count;
    cx.ovr_kind = it.ovr_kind; cx.ovr_index = it.ovr_index; cx.ovr_form = it.ovr_form;
    cx.ovr_tdiag = &it.ovr_tdiag;
    std::vector<cplx> ptab;
    std::vector<double> etab;
    if (run.has_cost && g_emul_cost_tables) {
        etab.resize((size_t)1 << run.K);
        for (uint32_t idx = 0; idx < (1u << run.K); idx++) etab[idx] = qgt_cost_ein(run, ct, idx);
    }
    std::vector<double> cost_buf(qgt_cost_smem_doubles(run.K, ct.num_edges, !g_emul_cost_tables));
    QgtCostSmem cost_sm = qgt_cost_smem_carve(cost_buf.data(), run.K, ct.num_edges, !g_emul_cost_tables);
    if (run.has_cost) for (int tid = 0; tid < T; tid++) qgt_cost_build_ein(run, ct, cost_sm, tid, T);
    for (uint64_t tau = 0; tau < ntiles; tau++) {
        const uint64_t tilebase = qgt_tile_base(run, tau);
        for (int tid = 0; tid < T; tid++) {
            const QgtIoMap<R + B> io = qgt_make_iomap<R + B>(run, tid);
            qgt_phase_load<R + B>(io, tile.data(), (const cplx*)it.src, tilebase, tid, T);
        }
        for (int s = 0; s < run.nsub; s++)
            for (int tid = 0; tid < T; tid++) {
                if (subs[s].nreg == 0) {
                    const QgtDevCost& co = (it.ovr_kind == 3 && subs[s].cost == it.ovr_index) ? it.ovr_cost : img.costs[run.cost_off + subs[s].cost];
                    if (tid == 0) {          // the kernel's three barrier-separated steps, all threads of each step in turn
                        for (int t1 = 0; t1 < T; t1++)
                            for (int t2 = t1; t2 < QGT_COST_LIN_THREADS; t2 += T) qgt_cost_tile_lin(run, ct, cost_sm, tilebase, t2);
                        for (int t2 = 0; t2 < T; t2++) qgt_cost_tile_tables(run, cost_sm, t2, T, co.angle);
                        // the per-launch phase table of this cost op (a small kernel on the device)
                        ptab.resize((size_t)1 << run.K);
                        for (uint32_t idx = 0; idx < (1u << run.K); idx++)
                            ptab[idx] = qgt_cost_phase_entry(run, ct, img.costs[run.cost_off + subs[s].cost].angle, idx);
                    }
                    qgt_phase_cost(run, co, tile.data(), cost_sm, g_emul_cost_tables ? ptab.data() : nullptr, etab.data(), tid, T);
                } else {
                    qgt_phase_subpass<R, B>(run, subs[s], cx, tile.data(), tilebase, tid);
                }
            }
        for (int tid = 0; tid < T; tid++) {
            const QgtIoMap<R + B> io = qgt_make_iomap<R + B>(run, tid);
            qgt_phase_store<R + B>(io, tile.data(), (cplx*)it.dst, tilebase, tid, T, it.accumulate != 0);
        }
    }
}

extern "C" int emul_num_runs(const qgt_b200_circuit* circ, const double* theta, int K, int R) {
    PlanOptions opt; opt.tile_qubits = K; opt.reg_qubits = R; opt.batch_qubits = g_emul_batch;
    CircuitPlan plan; std::string err;
    if (build_plan(*circ, theta, opt, plan, err)) return -1;
    return (int)plan.runs.size();
}

// one sweep item: dst (+)= run[run_idx](src) with op `ovr_op` replaced by its derivative (or -1)
extern "C" int emul_sweep(const qgt_b200_circuit* circ, const double* theta, int K, int R, int run_idx,
                          const double* src, double* dst, int ovr_op, int accumulate, const int* extra, int nextra) {
    PlanOptions opt; opt.tile_qubits = K; opt.reg_qubits = R; opt.batch_qubits = g_emul_batch;
    CircuitPlan plan; std::string err;
    if (build_plan(*circ, theta, opt, plan, err)) return -1;
    if (run_idx < 0 || run_idx >= (int)plan.runs.size()) return -2;
    PlanImage img;
    build_image(plan, img);
    const Run& run = plan.runs[run_idx];
    const QgtDevRun& dr = img.runs[run_idx];
    QgtSweepItem it;
    std::memset(&it, 0, sizeof it);
    it.src = src; it.dst = dst; it.ovr_kind = 0; it.ovr_index = -1; it.accumulate = accumulate ? 1u : 0u;
    std::vector<double> mats;
    if (ovr_op >= 0) {
        const OpLocation loc = locate_op(run, ovr_op);
        if (loc.kind == 0) return -3;
        it.ovr_kind = loc.kind; it.ovr_index = loc.index;
        const SubPass& sp = run.subs[loc.sub];
        if (loc.kind == 1) {
            int first = 0;
            for (int s2 = 0; s2 < loc.sub; s2++) first += (int)run.subs[s2].stages.size();
            std::vector<int> dops(1, ovr_op);
            for (int e = 0; e < nextra; e++) dops.push_back(extra[e]);
            it.ovr_form = stage_matrices_sum(run, sp, sp.stages[loc.index - first], dops, mats);
            it.ovr_mat = mats.data();
        } else if (loc.kind == 2) it.ovr_tdiag = make_tdiag(run.ops[ovr_op], true);
        else it.ovr_cost = make_cost(run.ops[ovr_op], true);
    }
    std::vector<QgtDevEdge> ed(circ->num_edges);
    for (size_t k = 0; k < circ->num_edges; k++) { ed[k].i = circ->edges[k].i; ed[k].j = circ->edges[k].j; ed[k].w = circ->edges[k].weight; }
    QgtCostTable ct{ed.data(), (int)ed.size(), circ->vertex_weights, circ->num_qubits};
    const uint64_t D = (uint64_t)1 << circ->num_qubits;
    switch (plan.R * 2 + plan.B) {
    case 2: emul_item<1, 0>(dr, img, it, ct, D); break;
    case 3: emul_item<1, 1>(dr, img, it, ct, D); break;
    case 4: emul_item<2, 0>(dr, img, it, ct, D); break;
    case 5: emul_item<2, 1>(dr, img, it, ct, D); break;
    case 6: emul_item<3, 0>(dr, img, it, ct, D); break;
    case 7: emul_item<3, 1>(dr, img, it, ct, D); break;
    default: return -4;
    }
    return 0;
}

// host-side planning cost (plan + image + program + derivative matrices), milliseconds per call
#include <chrono>
extern "C" double emul_host_plan_ms(const qgt_b200_circuit* circ, const double* theta, int K, int R, size_t slots, int reps, double* parts) {
    PlanOptions opt; opt.tile_qubits = K; opt.reg_qubits = R; opt.batch_qubits = g_emul_batch;
    using clk = std::chrono::steady_clock;
    double t_plan = 0, t_img = 0, t_prog = 0, t_ovr = 0;
    for (int rep = 0; rep < reps; rep++) {
        CircuitPlan plan; std::string err;
        auto t0 = clk::now();
        if (build_plan(*circ, theta, opt, plan, err)) return -1;
        auto t1 = clk::now();
        PlanImage img; build_image(plan, img);
        auto t2 = clk::now();
        Program prog; if (build_qgt_program(plan, slots, false, prog, err)) return -2;
        auto t3 = clk::now();
        std::vector<double> mats; size_t total = 0;
        for (const Instr& in : prog.instrs) {
            if (in.kind != INSTR_SWEEP) continue;
            const Run& run = plan.runs[in.run];
            for (const SweepCol& sc : in.cols) {
                if (sc.ovr_op < 0) continue;
                const OpLocation loc = locate_op(run, sc.ovr_op);
                if (loc.kind == 1) {
                    int first = 0;
                    for (int s2 = 0; s2 < loc.sub; s2++) first += (int)run.subs[s2].stages.size();
                    std::vector<int> dops(1, sc.ovr_op);
                    dops.insert(dops.end(), sc.ovr_extra.begin(), sc.ovr_extra.end());
                    stage_matrices_sum(run, run.subs[loc.sub], run.subs[loc.sub].stages[loc.index - first], dops, mats);
                    total += mats.size();
                }
            }
        }
        auto t4 = clk::now();
        t_plan += std::chrono::duration<double, std::milli>(t1 - t0).count();
        t_img += std::chrono::duration<double, std::milli>(t2 - t1).count();
        t_prog += std::chrono::duration<double, std::milli>(t3 - t2).count();
        t_ovr += std::chrono::duration<double, std::milli>(t4 - t3).count();
    }
    if (parts) { parts[0] = t_plan / reps; parts[1] = t_img / reps; parts[2] = t_prog / reps; parts[3] = t_ovr / reps; }
    return (t_plan + t_img + t_prog + t_ovr) / reps;
}

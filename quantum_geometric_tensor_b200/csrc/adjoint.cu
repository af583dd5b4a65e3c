// adjoint.cu — energy gradient of a parameterised circuit by the adjoint method:
// what feeds the natural-gradient step next to the metric.  Replaces the reference's per-parameter re-simulation
// (algorithms/qaoa.c:489-558: two circuit executions per parameter; core/quantum_geometric_gradient.c:2887 consumes the
// result) by ONE forward circuit and ONE backward pass over two states, whatever the number of parameters.
//
//   E = <psi|H|psi>,  H = E_z of the circuit's edge list / vertex weights (sum of w (1 - Z_i Z_j)/2 and v_q Z_q terms)
//   backward pass on the inverse circuit: chi_0 = psi, Lambda_0 = H psi,
//   dE/dtheta_mu = -2 Re sum_j <Lambda_j| (d_mu S_j) S_j^+ |chi_j>                       (plan.hpp)
//
// Fusable plans (the ansatz circuits): the fused kernel advances chi and Lambda in lockstep and accumulates the 8x8
// transition matrices of every stage with a parameter; P dot products collapse into 64-term contractions.  Other plans
// (QAOA cost layers, registers below 8 qubits): derivative columns are spawned per run and contracted by a Gram.
#include <cuda_runtime.h>

#include <algorithm>
#include <chrono>
#include <cstring>
#include <vector>

#include "ctx.hpp"

using namespace qgt;

extern "C" int qgt_b200_expectation_gradient(qgt_b200_ctx* c, const qgt_b200_circuit* circ, const double* theta,
                                              double* energy, double* grad) {
    if (!c) return fail(QGT_B200_ERR_INVALID_ARG, "ctx is NULL");
    if (!circ) return fail(QGT_B200_ERR_INVALID_ARG, "circuit is NULL");
    if (circ->num_params < 0) return fail(QGT_B200_ERR_INVALID_ARG, "num_params < 0");
    if (circ->num_params > 0 && !theta) return fail(QGT_B200_ERR_INVALID_ARG, "theta is NULL");
    cudaSetDevice(c->device);
    const auto t_wall0 = std::chrono::steady_clock::now();
    const int n = circ->num_qubits, P = circ->num_params;
    // a state sharded over the ranks of the communicator: every rank holds 2^nloc amplitudes, runs the same programs and
    // contributes partial transition matrices / dot products, summed with one allreduce each
    const bool sharded = c->world > 1;
    int gbits = 0;
    while ((1 << gbits) < c->world) gbits++;
    const int nloc = n - gbits;
    if (sharded && nloc < 4) return fail(QGT_B200_ERR_INVALID_ARG, "state too small for this many ranks");
    const uint64_t D = (uint64_t)1 << nloc;
    const uint64_t goff = (uint64_t)c->rank << nloc;
    int rc;
    std::string err;

    // plans of U and of U^+ (same options: same tiles, same fusion rules)
    CircuitPlan fwd, inv;
    std::vector<MappedSegment> fsegs, isegs;
    std::vector<qgt_b200_gate> inv_gates;
    invert_circuit(*circ, inv_gates);
    qgt_b200_circuit icirc = *circ;
    icirc.gates = inv_gates.data(); icirc.num_gates = inv_gates.size();
    if (sharded) {
        // the forward plan restores the identity qubit layout (H and the inverse plan assume it); the inverse plan need not
        if ((rc = build_plan_sharded(*circ, theta, c->opt, nloc, true, fwd, fsegs, err))) return fail(rc, err);
        if ((rc = build_plan_sharded(icirc, theta, c->opt, nloc, false, inv, isegs, err))) return fail(rc, err);
    } else {
        if ((rc = build_plan(*circ, theta, c->opt, fwd, err))) return fail(rc, err);
        if ((rc = build_plan(icirc, theta, c->opt, inv, err))) return fail(rc, err);
    }
    const bool want_grad = grad != nullptr && P > 0;
    const bool fused = want_grad && c->fused_mode != 0 && plan_supports_fused(inv);

    size_t slots = workspace_slots(c, D, (size_t)256 << 20);
    if (slots < 5) return fail(QGT_B200_ERR_NO_MEMORY, "workspace too small: the adjoint gradient needs 5 statevector-sized columns");
    int scratch = 0;
    if (want_grad && !fused) {
        int most = 1;
        for (const Run& run : inv.runs) {
            std::vector<char> seen(P, 0);
            int cnt = 0;
            for (const ParamOcc& oc : run.occ) if (!seen[oc.param]) { seen[oc.param] = 1; cnt++; }
            most = std::max(most, cnt);
        }
        // sharded: every rank must build the same programs, whatever its free memory
        scratch = sharded ? std::min(most, 1) : (int)std::min<size_t>((size_t)most, slots - 4);
    }
    const int num_slots = 3 + scratch;
    if ((rc = c->arena.reserve((size_t)(num_slots + 1) * D * sizeof(cplx)))) return rc;      // + the spare column of the exchanges
    const size_t cm = (size_t)(P + 1) * (P + 1);
    if ((rc = c->cmat.reserve(std::max<size_t>(16, cm * sizeof(cplx))))) return rc;
    if ((rc = c->scratch.reserve(256))) return rc;
    cplx* arena = (cplx*)c->arena.ptr;
    cplx* chi = arena;                  // slot 0
    cplx* lam = arena + 2 * D;          // slot 2

    stats_begin(c);
    PlanImage img;
    // forward: psi = U |init>
    if ((rc = upload_plan(c, *circ, fwd, img))) return rc;
    if (sharded && (rc = upload_segment_costs(c, *circ, fsegs))) return rc;
    cudaError_t e = launch_init_state(chi, D, circ->initial_state, std::pow(2.0, -0.5 * n), goff, c->stream);
    if (e != cudaSuccess) return cuda_fail(e, "init launch");
    if ((rc = apply_plan_inplace(c, fwd, chi, D))) return rc;
    // Lambda = H psi, E = <psi|Lambda>   (identity layout: global index = rank bits | local index)
    c->timer.begin(c->stream, 2);
    e = launch_cost_apply(lam, chi, D, c->cost, goff, c->stream);
    double h[2] = {0.0, 0.0};
    if (e == cudaSuccess) e = launch_cost_dot(chi, chi, D, c->cost, goff, (double*)c->scratch.ptr, c->stream);
    c->timer.end(c->stream);
    if (e == cudaSuccess) e = cudaMemcpyAsync(h, c->scratch.ptr, sizeof h, cudaMemcpyDeviceToHost, c->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(c->stream);
    if (e != cudaSuccess) return cuda_fail(e, "cost observable");
    if ((rc = dist_allreduce_host(c, h, 1))) return rc;
    c->stats.other_launches += 2;
    if (energy) *energy = h[0];

    if (want_grad) {
        std::fill(grad, grad + P, 0.0);
        PlanImage iimg;
        if ((rc = upload_plan(c, icirc, inv, iimg))) return rc;
        if (sharded && (rc = upload_segment_costs(c, icirc, isegs))) return rc;
        std::vector<double> row((size_t)P * 2);
        if (fused) {
            Program prog;
            if ((rc = build_gradient_fused_program(inv, prog))) return fail(rc, "gradient program");
            if ((rc = run_program(c, icirc, inv, prog, arena, D, (cplx*)c->cmat.ptr))) return rc;
            if ((rc = dist_allreduce_device(c, (double*)c->amat.ptr, row.size()))) return rc;       // row 0 of A
            e = cudaMemcpyAsync(row.data(), c->amat.ptr, row.size() * sizeof(double), cudaMemcpyDeviceToHost, c->stream);
            if (e == cudaSuccess) e = cudaStreamSynchronize(c->stream);
            if (e != cudaSuccess) return cuda_fail(e, "gradient download");
            for (int mu = 0; mu < P; mu++) grad[mu] = -2.0 * row[2 * (size_t)mu];
        } else {
            std::vector<Program> progs;
            cplx* d_row = (cplx*)c->cmat.ptr + (size_t)P * (P + 1);       // C[P][nu], nu = 0..P-1
            for (int r = 0; r < (int)inv.runs.size(); r++) {
                if (inv.runs[r].exchange_gbit >= 0) {
                    // sharded state: both states go through the exchange in place (the per-run programs below assume every
                    // slot stays where it is, so the exchange is not left to run_program's column rotation)
                    if ((rc = dist_exchange(c, chi, D, inv.runs[r].exchange_mask))) return rc;
                    if ((rc = dist_exchange(c, lam, D, inv.runs[r].exchange_mask))) return rc;
                    continue;
                }
                if ((rc = build_gradient_run_programs(inv, r, scratch, progs))) return fail(rc, "gradient program");
                for (const Program& g : progs) {
                    bool has_gram = false;
                    for (const Instr& in : g.instrs) if (in.kind == INSTR_GRAM) has_gram = true;
                    if (has_gram) {
                        e = cudaMemsetAsync(c->cmat.ptr, 0, cm * sizeof(cplx), c->stream);
                        if (e != cudaSuccess) return cuda_fail(e, "memset");
                    }
                    if ((rc = run_program(c, icirc, inv, g, arena, D, (cplx*)c->cmat.ptr))) return rc;
                    if (!has_gram) continue;
                    if ((rc = dist_allreduce_device(c, (double*)d_row, row.size()))) return rc;
                    e = cudaMemcpyAsync(row.data(), d_row, row.size() * sizeof(double), cudaMemcpyDeviceToHost, c->stream);
                    if (e == cudaSuccess) e = cudaStreamSynchronize(c->stream);
                    if (e != cudaSuccess) return cuda_fail(e, "gradient download");
                    for (int mu = 0; mu < P; mu++) grad[mu] += -2.0 * row[2 * (size_t)mu];
                }
            }
        }
    }
    if ((rc = stats_end(c))) return rc;
    c->stats.num_runs = (int)(fwd.runs.size() + inv.runs.size());
    c->stats.resident_columns = 1;
    c->stats.blocks = 1;
    c->stats.tile_qubits = fwd.runs.empty() ? 0 : fwd.runs[0].K;
    c->stats.fused = fused ? 1 : 0;
    c->stats.ms_wall = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t_wall0).count();
    return QGT_B200_OK;
}

// ---- natural-gradient optimiser: metric + adjoint gradient + regularised solve per step --------
// Replaces the host loop around compute_quantum_geometric_tensor / compute_regularized_natural_gradient
// (core/quantum_geometric_gradient.c:2887, consumed by hybrid/classical_optimization_engine.c): both device evaluations
// per step are single calls, the P x P solve stays on the host as in the reference.
extern "C" int qgt_b200_natural_gradient_step(qgt_b200_ctx* c, const qgt_b200_circuit* circ, const double* theta, double learning_rate,
                                               const qgt_b200_natgrad_config* cfg, double* theta_out, double* energy, double* lambda_used) {
    if (!c || !circ || !theta_out) return fail(QGT_B200_ERR_INVALID_ARG, "ctx/circuit/theta_out is NULL");
    const int P = circ->num_params;
    if (P <= 0) return fail(QGT_B200_ERR_INVALID_ARG, "the circuit has no parameters");
    if (!theta) return fail(QGT_B200_ERR_INVALID_ARG, "theta is NULL");
    std::vector<double> metric((size_t)P * P), grad(P), dx(P);
    int rc = qgt_b200_qgt(c, circ, theta, metric.data(), nullptr, nullptr, nullptr);
    if (rc) return rc;
    double e = 0.0;
    if ((rc = qgt_b200_expectation_gradient(c, circ, theta, &e, grad.data()))) return rc;
    if ((rc = qgt_b200_natural_gradient(c, metric.data(), grad.data(), (size_t)P, cfg, dx.data(), lambda_used))) return rc;
    for (int i = 0; i < P; i++) theta_out[i] = theta[i] - learning_rate * dx[i];
    if (energy) *energy = e;
    return QGT_B200_OK;
}

extern "C" int qgt_b200_natural_gradient_descent(qgt_b200_ctx* c, const qgt_b200_circuit* circ, double* theta, int iterations,
                                                  double learning_rate, const qgt_b200_natgrad_config* cfg, double* history) {
    if (!c || !circ || !theta) return fail(QGT_B200_ERR_INVALID_ARG, "ctx/circuit/theta is NULL");
    if (iterations < 0) return fail(QGT_B200_ERR_INVALID_ARG, "iterations < 0");
    const int P = circ->num_params;
    std::vector<double> next((size_t)std::max(1, P));
    int rc;
    for (int it = 0; it < iterations; it++) {
        double e = 0.0;
        if ((rc = qgt_b200_natural_gradient_step(c, circ, theta, learning_rate, cfg, next.data(), &e, nullptr))) return rc;
        if (history) history[it] = e;
        std::copy(next.begin(), next.begin() + P, theta);
    }
    if (history) {
        double e = 0.0;
        if ((rc = qgt_b200_expectation_gradient(c, circ, theta, &e, nullptr))) return rc;
        history[iterations] = e;
    }
    return QGT_B200_OK;
}

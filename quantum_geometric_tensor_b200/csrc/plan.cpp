// plan.cpp — circuit lowering, commutation-aware fusion into runs / sub-passes, QGT column schedule.
//
// Replaces, on the host side, the per-gate loop of sim_execute_circuit
// (reference src/quantum_geometric/hardware/quantum_simulator.c:499-533), the gate tables of
// apply_gate_by_type (:188-283) / generate_gate_matrix (hardware/quantum_simulator_cpu.c:659-746),
// and the per-element derivative/overlap orchestration of compute_quantum_geometric_tensor
// (core/quantum_geometric_tensor_network.c:1058-1181) with one static schedule per QGT evaluation.
#include "plan.hpp"

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <set>
#include <sstream>

namespace qgt {

static const double kPi = 3.14159265358979323846;
static const double kSqrt2 = 1.41421356237309504880;  // same literal as quantum_simulator.c:36

static inline void setc(double* m, int k, double re, double im) { m[2 * k] = re; m[2 * k + 1] = im; }
static inline uint64_t bit(int q) { return (uint64_t)1 << q; }

static double gate_angle(const qgt_b200_gate& g, const double* theta) {
    return g.param >= 0 ? g.scale * theta[g.param] + g.angle : g.angle;
}

static bool is_two_qubit(int kind) {
    switch (kind) {
    case QGT_B200_GATE_CNOT: case QGT_B200_GATE_CY: case QGT_B200_GATE_CZ: case QGT_B200_GATE_SWAP:
    case QGT_B200_GATE_CRX: case QGT_B200_GATE_CRY: case QGT_B200_GATE_CRZ: case QGT_B200_GATE_CH:
    case QGT_B200_GATE_ZZ: return true;
    default: return false;
    }
}

static bool is_parametric(int kind) {
    switch (kind) {
    case QGT_B200_GATE_RX: case QGT_B200_GATE_RY: case QGT_B200_GATE_RZ: case QGT_B200_GATE_U1:
    case QGT_B200_GATE_PHASE: case QGT_B200_GATE_CRX: case QGT_B200_GATE_CRY: case QGT_B200_GATE_CRZ:
    case QGT_B200_GATE_ZZ: case QGT_B200_GATE_COST: return true;
    default: return false;
    }
}

int lower_gate(const qgt_b200_circuit& c, const double* theta, int gi, std::vector<LoweredOp>& out, std::string& err) {
    const qgt_b200_gate& g = c.gates[gi];
    const int n = c.num_qubits;
    char buf[160];
    if (g.kind != QGT_B200_GATE_COST && (g.target < 0 || g.target >= n)) {
        snprintf(buf, sizeof buf, "gate %d: target %d out of range for %d qubits", gi, g.target, n);
        err = buf; return QGT_B200_ERR_CIRCUIT;
    }
    if (is_two_qubit(g.kind) && (g.control < 0 || g.control >= n || g.control == g.target)) {
        snprintf(buf, sizeof buf, "gate %d: control %d invalid (target %d, %d qubits)", gi, g.control, g.target, n);
        err = buf; return QGT_B200_ERR_CIRCUIT;
    }
    if (g.param >= c.num_params) {
        snprintf(buf, sizeof buf, "gate %d: parameter %d out of range (%d parameters)", gi, g.param, c.num_params);
        err = buf; return QGT_B200_ERR_CIRCUIT;
    }
    if (g.param >= 0 && !is_parametric(g.kind)) {
        snprintf(buf, sizeof buf, "gate %d: kind %d takes no parameter", gi, g.kind);
        err = buf; return QGT_B200_ERR_CIRCUIT;
    }
    const double a = (theta || g.param < 0) ? gate_angle(g, theta) : g.angle;
    const double s = g.scale;
    const double cs = std::cos(a / 2.0), sn = std::sin(a / 2.0);
    LoweredOp op;
    op.gate = gi;
    op.param = g.param;
    op.target = g.target;
    const uint64_t cb = is_two_qubit(g.kind) ? bit(g.control) : 0;

    auto diag = [&](double d0r, double d0i, double d1r, double d1i) {
        op.type = QGT_OP_DIAG; op.target = -1; op.pmask = bit(g.target);
        setc(op.m, 0, d0r, d0i); setc(op.m, 1, d1r, d1i);
    };
    auto rz_like = [&]() {   // d0 = e^{-ia/2}, d1 = e^{+ia/2}; derivative -i/2 d0, +i/2 d1 (times scale)
        setc(op.m, 0, cs, -sn); setc(op.m, 1, cs, sn);
        setc(op.dm, 0, -0.5 * s * sn, -0.5 * s * cs); setc(op.dm, 1, -0.5 * s * sn, 0.5 * s * cs);
    };

    switch (g.kind) {
    case QGT_B200_GATE_I: return QGT_B200_OK;
    case QGT_B200_GATE_X: op.type = QGT_OP_PERM; break;
    case QGT_B200_GATE_CNOT: op.type = QGT_OP_PERM; op.cmask = cb; break;
    case QGT_B200_GATE_Y:
    case QGT_B200_GATE_CY:
        op.type = QGT_OP_U; op.cmask = cb;
        setc(op.m, 0, 0, 0); setc(op.m, 1, 0, -1); setc(op.m, 2, 0, 1); setc(op.m, 3, 0, 0); break;
    case QGT_B200_GATE_Z: diag(1, 0, -1, 0); break;
    case QGT_B200_GATE_CZ: diag(1, 0, -1, 0); op.cmask = cb; break;
    case QGT_B200_GATE_H:
    case QGT_B200_GATE_CH: {
        const double h = 1.0 / kSqrt2;
        op.type = QGT_OP_UREAL; op.cmask = cb;
        setc(op.m, 0, h, 0); setc(op.m, 1, h, 0); setc(op.m, 2, h, 0); setc(op.m, 3, -1.0 / kSqrt2, 0); break; }
    case QGT_B200_GATE_S: diag(1, 0, 0, 1); break;
    case QGT_B200_GATE_SDG: diag(1, 0, 0, -1); break;
    case QGT_B200_GATE_T: diag(1, 0, std::cos(kPi / 4.0), std::sin(kPi / 4.0)); break;
    case QGT_B200_GATE_TDG: diag(1, 0, std::cos(kPi / 4.0), -std::sin(kPi / 4.0)); break;
    case QGT_B200_GATE_SX:
        op.type = QGT_OP_U;
        setc(op.m, 0, 0.5, 0.5); setc(op.m, 1, 0.5, -0.5); setc(op.m, 2, 0.5, -0.5); setc(op.m, 3, 0.5, 0.5); break;
    case QGT_B200_GATE_RX:
    case QGT_B200_GATE_CRX:
        op.type = QGT_OP_URX; op.cmask = cb;
        setc(op.m, 0, cs, 0); setc(op.m, 1, 0, -sn); setc(op.m, 2, 0, -sn); setc(op.m, 3, cs, 0);
        setc(op.dm, 0, -0.5 * s * sn, 0); setc(op.dm, 1, 0, -0.5 * s * cs);
        setc(op.dm, 2, 0, -0.5 * s * cs); setc(op.dm, 3, -0.5 * s * sn, 0);
        break;
    case QGT_B200_GATE_RY:
    case QGT_B200_GATE_CRY:
        op.type = QGT_OP_UREAL; op.cmask = cb;
        setc(op.m, 0, cs, 0); setc(op.m, 1, -sn, 0); setc(op.m, 2, sn, 0); setc(op.m, 3, cs, 0);
        setc(op.dm, 0, -0.5 * s * sn, 0); setc(op.dm, 1, -0.5 * s * cs, 0);
        setc(op.dm, 2, 0.5 * s * cs, 0); setc(op.dm, 3, -0.5 * s * sn, 0);
        break;
    case QGT_B200_GATE_RZ:
    case QGT_B200_GATE_CRZ:
        op.type = QGT_OP_DIAG; op.target = -1; op.pmask = bit(g.target); op.cmask = cb;
        rz_like(); break;
    case QGT_B200_GATE_ZZ:
        op.type = QGT_OP_DIAG; op.target = -1; op.pmask = bit(g.target) | bit(g.control);
        rz_like(); break;
    case QGT_B200_GATE_U1:
    case QGT_B200_GATE_PHASE:
        diag(1, 0, std::cos(a), std::sin(a));
        setc(op.dm, 0, 0, 0); setc(op.dm, 1, -s * std::sin(a), s * std::cos(a)); break;
    case QGT_B200_GATE_SWAP: {   // three CNOTs, quantum_simulator.c:269-275
        LoweredOp p1 = op; p1.type = QGT_OP_PERM; p1.target = g.target; p1.cmask = bit(g.control);
        LoweredOp p2 = op; p2.type = QGT_OP_PERM; p2.target = g.control; p2.cmask = bit(g.target);
        out.push_back(p1); out.push_back(p2); out.push_back(p1);
        return QGT_B200_OK; }
    case QGT_B200_GATE_COST:
        op.type = QGT_OP_COST; op.target = -1;
        op.m[0] = a; op.m[1] = 0; op.dm[0] = a; op.dm[1] = s; op.dflags = QGT_FLAG_COST_DERIV; break;
    default:
        snprintf(buf, sizeof buf, "gate %d: unsupported kind %d", gi, g.kind);
        err = buf; return QGT_B200_ERR_UNSUPPORTED;
    }
    if (g.kind == QGT_B200_GATE_CRX || g.kind == QGT_B200_GATE_CRY || g.kind == QGT_B200_GATE_CRZ)
        op.dflags |= QGT_FLAG_ZERO_CTRL_FAIL;
    out.push_back(op);
    return QGT_B200_OK;
}

// ---- fusion into runs ------------------------------------------------------------------------
namespace {

struct Dag {
    std::vector<std::vector<int>> succ;
    std::vector<int> indeg;
};

uint64_t diag_qubits(const LoweredOp& op, int n) {
    if (op.type == QGT_OP_COST) return n >= 64 ? ~0ull : (bit(n) - 1);
    return op.cmask | op.pmask;
}

Dag build_dag(const std::vector<LoweredOp>& ops, int n) {
    Dag d;
    const int N = (int)ops.size();
    d.succ.assign(N, {});
    d.indeg.assign(N, 0);
    std::vector<int> last_nd(n, -1);
    std::vector<std::vector<int>> diag_since(n);
    std::vector<std::set<int>> preds(N);
    for (int i = 0; i < N; i++) {
        const uint64_t dq = diag_qubits(ops[i], n);
        for (int q = 0; q < n; q++) {
            if (!(dq >> q & 1)) continue;
            if (last_nd[q] >= 0) preds[i].insert(last_nd[q]);
            diag_since[q].push_back(i);
        }
        if (ops[i].target >= 0) {
            const int q = ops[i].target;
            if (last_nd[q] >= 0) preds[i].insert(last_nd[q]);
            for (int j : diag_since[q]) if (j != i) preds[i].insert(j);
            last_nd[q] = i;
            diag_since[q].clear();
        }
    }
    for (int i = 0; i < N; i++)
        for (int p : preds[i]) { d.succ[p].push_back(i); d.indeg[i]++; }
    return d;
}

// Thread-bit -> tile-position map of a sub-pass.
//  * positions that select a matrix variant go last (they become warp-index bits, so a warp shares one
//    matrix: required by the tensor-pipe path); returns false when they do not all fit above bit 4;
//  * the shared-memory swizzle XORs bit p of the index into bank-group bit (p mod 3).  The tensor-pipe
//    path loads with lanes spanning {thread bit 0, matrix bits 0,1} and stores with lanes spanning
//    {matrix bit 0, thread bits 1,2}; the register path spans thread bits 0..2.  Pick thread bits 0..2
//    so those triples have distinct residues where possible (conflict-free quarter-warps).
//  * the order of the matrix qubits is free (the host builds the matrices in whatever order is chosen), so all
//    orders are tried together with the thread-bit choice.
bool make_tperm(int K, std::vector<int>& reg_local, const std::vector<int>& batch_local,
                const std::vector<int>& variant_local, std::vector<int>& tperm) {
    std::vector<char> taken(K, 0), is_var(K, 0);
    for (int r : reg_local) taken[r] = 1;
    for (int r : batch_local) taken[r] = 1;
    for (int v : variant_local) if (!taken[v]) is_var[v] = 1;
    std::vector<int> freep, varp;
    for (int p = 0; p < K; p++) if (!taken[p]) (is_var[p] ? varp : freep).push_back(p);
    int best[3] = {-1, -1, -1}, best_score = -1;
    const int nf = (int)freep.size();
    std::vector<int> order = reg_local, best_order = reg_local;
    std::sort(order.begin(), order.end());
    // the score depends on the positions' residues only: try residue triples and take the first free positions of
    // each class (27 triples x 6 orders instead of nf^3 x 6 position triples)
    std::vector<int> by_res[3];
    for (int i = 0; i < nf; i++) by_res[freep[i] % 3].push_back(i);
    do {
        const int r0 = order.size() > 0 ? order[0] % 3 : -1, r1 = order.size() > 1 ? order[1] % 3 : -1;
        for (int pa = 0; pa < 3 && best_score < 10; pa++) for (int pb = 0; pb < 3 && best_score < 10; pb++) for (int pc = 0; pc < 3; pc++) {
            int used[3] = {0, 0, 0};
            const int a = used[pa] < (int)by_res[pa].size() ? by_res[pa][used[pa]++] : -1;
            const int b2 = used[pb] < (int)by_res[pb].size() ? by_res[pb][used[pb]++] : -1;
            const int c = used[pc] < (int)by_res[pc].size() ? by_res[pc][used[pc]++] : -1;
            if (a < 0 || b2 < 0 || c < 0) continue;
            int score = 0;
            if (pa != r0 && pa != r1 && r0 != r1) score += 4;          // tensor-path loads
            if (pb != r0 && pc != r0 && pb != pc) score += 4;          // tensor-path stores
            if (pa != pb && pa != pc && pb != pc) score += 2;          // register path
            if (score > best_score) { best_score = score; best[0] = a; best[1] = b2; best[2] = c; best_order = order; }
        }
    } while (best_score < 10 && order.size() == 3 && std::next_permutation(order.begin(), order.end()));
    if (best_score >= 0) reg_local = best_order;
    tperm.clear();
    if (best_score >= 0) {
        for (int k = 0; k < 3; k++) tperm.push_back(freep[best[k]]);
        for (int i = 0; i < nf; i++) if (i != best[0] && i != best[1] && i != best[2]) tperm.push_back(freep[i]);
    } else {
        tperm = freep;
    }
    const bool warp_uniform = (int)tperm.size() >= 5;       // variant bits start at thread bit >= 5
    for (int p : varp) tperm.push_back(p);
    return warp_uniform || varp.empty();
}

}  // namespace

// qubit that decides where the scheduler places an op: the target of a 2x2 / permutation op, the lowest local qubit
// of a PARAMETERISED diagonal op, -1 for free diagonal ops and the cost layer
static int sched_target(const LoweredOp& o, int nl) {
    if (o.target >= 0) return o.target;
    if (o.type == QGT_OP_DIAG && o.param >= 0)
        for (int q = 0; q < nl; q++) if ((o.pmask | o.cmask) >> q & 1) return q;
    return -1;
}

OpLocation locate_op(const Run& run, int op_index) {
    OpLocation loc;
    int stage_base = 0, tdiag_base = 0, cost_base = 0;
    for (size_t s = 0; s < run.subs.size(); s++) {
        const SubPass& sp = run.subs[s];
        if (op_index >= sp.op_begin && op_index < sp.op_end) {
            loc.sub = (int)s;
            if (sp.is_cost) { loc.kind = 3; loc.index = cost_base; return loc; }
            for (size_t k = 0; k < sp.stages.size(); k++)
                for (int o : sp.stages[k].ops) if (o == op_index) { loc.kind = 1; loc.index = stage_base + (int)k; return loc; }
            for (size_t k = 0; k < sp.tdiags.size(); k++)
                if (sp.tdiags[k] == op_index) { loc.kind = 2; loc.index = tdiag_base + (int)k; return loc; }
            return loc;
        }
        stage_base += (int)sp.stages.size();
        tdiag_base += (int)sp.tdiags.size();
        cost_base += sp.is_cost ? 1 : 0;
    }
    return loc;
}

QgtDevThrDiag make_tdiag(const LoweredOp& op, bool derivative) {
    QgtDevThrDiag t;
    std::memset(&t, 0, sizeof t);
    t.cmask = op.cmask; t.pmask = op.pmask;
    std::memcpy(t.d, derivative ? op.dm : op.m, 4 * sizeof(double));
    t.flags = derivative ? op.dflags : op.flags;
    return t;
}

QgtDevCost make_cost(const LoweredOp& op, bool derivative) {
    QgtDevCost c;
    std::memset(&c, 0, sizeof c);
    c.angle = derivative ? op.dm[0] : op.m[0];
    c.dscale = derivative ? op.dm[1] : op.m[1];
    c.flags = derivative ? op.dflags : op.flags;
    return c;
}

// all variants of a stage's matrix, row-major N x N complex each (dense[(p*N*N + i*N + j)*2])
static void stage_dense(const Run& run, const SubPass& sp, const Stage& st, int deriv_op, std::vector<double>& dense) {
    const int R = (int)sp.reg_local.size();
    const int N = 1 << R;
    const int nvar = (int)st.vqubits.size();
    dense.assign((size_t)(1 << nvar) * N * N * 2, 0.0);
    // masks of each op over the register combo c and the variant pattern p
    struct Bound { uint32_t creg, cvar, preg, pvar; int j; };
    std::vector<Bound> bound(st.ops.size());
    for (size_t k = 0; k < st.ops.size(); k++) {
        const LoweredOp& op = run.ops[st.ops[k]];
        Bound bd = {0, 0, 0, 0, -1};
        for (int r = 0; r < R; r++) {
            const int q = run.tile_qubits[sp.reg_local[r]];
            if (op.cmask >> q & 1) bd.creg |= 1u << r;
            if (op.pmask >> q & 1) bd.preg |= 1u << r;
            if (op.target == q) bd.j = r;
        }
        for (int v = 0; v < nvar; v++) {
            if (op.cmask >> st.vqubits[v] & 1) bd.cvar |= 1u << v;
            if (op.pmask >> st.vqubits[v] & 1) bd.pvar |= 1u << v;
        }
        bound[k] = bd;
    }
    std::vector<double> M((size_t)N * N * 2), T((size_t)N * N * 2);
    for (int p = 0; p < (1 << nvar); p++) {
        std::fill(M.begin(), M.end(), 0.0);
        for (int i = 0; i < N; i++) M[2 * (i * N + i)] = 1.0;
        for (size_t k = 0; k < st.ops.size(); k++) {
            const int oi = st.ops[k];
            const LoweredOp& op = run.ops[oi];
            const Bound& bd = bound[k];
            const bool der = (oi == deriv_op);
            const double* m = der ? op.dm : op.m;
            const bool zero_fail = ((der ? op.dflags : op.flags) & QGT_FLAG_ZERO_CTRL_FAIL) != 0;
            const bool var_ok = ((uint32_t)p & bd.cvar) == bd.cvar;
            double u[8];
            if (op.type == QGT_OP_PERM) { const double x[8] = {0, 0, 1, 0, 1, 0, 0, 0}; std::memcpy(u, x, sizeof u); }
            else if (op.type != QGT_OP_DIAG) std::memcpy(u, m, sizeof u);
            const int pvar_par = __builtin_popcount((uint32_t)p & bd.pvar) & 1;
            // T = A * M with A having at most two non-zeros per row
            for (int c = 0; c < N; c++) {
                double* trow = &T[2 * (size_t)c * N];
                const double* mrow = &M[2 * (size_t)c * N];
                const bool ok = var_ok && (((uint32_t)c & bd.creg) == bd.creg);
                if (!ok) {
                    if (zero_fail) std::fill(trow, trow + 2 * N, 0.0);
                    else std::memcpy(trow, mrow, 2 * N * sizeof(double));
                    continue;
                }
                if (op.type == QGT_OP_DIAG) {
                    const int par = pvar_par ^ (__builtin_popcount((uint32_t)c & bd.preg) & 1);
                    const double pr = m[2 * par], pi = m[2 * par + 1];
                    for (int l = 0; l < N; l++) {
                        trow[2 * l] = pr * mrow[2 * l] - pi * mrow[2 * l + 1];
                        trow[2 * l + 1] = pr * mrow[2 * l + 1] + pi * mrow[2 * l];
                    }
                } else {
                    const int bsel = (c >> bd.j) & 1;
                    const double* r0 = &M[2 * (size_t)(c & ~(1 << bd.j)) * N];
                    const double* r1 = &M[2 * (size_t)(c | (1 << bd.j)) * N];
                    const double ar = u[4 * bsel], ai = u[4 * bsel + 1], br = u[4 * bsel + 2], bi = u[4 * bsel + 3];
                    for (int l = 0; l < N; l++) {
                        trow[2 * l] = ar * r0[2 * l] - ai * r0[2 * l + 1] + br * r1[2 * l] - bi * r1[2 * l + 1];
                        trow[2 * l + 1] = ar * r0[2 * l + 1] + ai * r0[2 * l] + br * r1[2 * l + 1] + bi * r1[2 * l];
                    }
                }
            }
            M.swap(T);
        }
        std::memcpy(&dense[(size_t)p * N * N * 2], M.data(), M.size() * sizeof(double));
    }
}

// Device layout of a stage's variants (see dev_structs.h).  8x8 matrices whose rows all have a constant
// phase, M[i][j] = d_i * r_ij with r real, are stored in QGT_FORM_DIAG_REAL (decided for all variants of
// the stage together); everything else dense.
static int pack_stage(int N, int nvariants, const std::vector<double>& dense, std::vector<double>& out, bool allow_parity) {
    const size_t vstride = (size_t)QGT_VARIANT_STRIDE(N) * 2;      // doubles per variant
    out.assign((size_t)nvariants * vstride, 0.0);
    bool diag_real = (N == 8);
    std::vector<double> d((size_t)nvariants * N * 2, 0.0), rm((size_t)nvariants * N * N, 0.0);
    for (int p = 0; p < nvariants && diag_real; p++) {
        const double* M = &dense[(size_t)p * N * N * 2];
        for (int i = 0; i < N && diag_real; i++) {
            int jm = 0; double best = -1.0;
            for (int j = 0; j < N; j++) {
                const double a = M[2 * (i * N + j)] * M[2 * (i * N + j)] + M[2 * (i * N + j) + 1] * M[2 * (i * N + j) + 1];
                if (a > best) { best = a; jm = j; }
            }
            const double mag = std::sqrt(best);
            double pr = 1.0, pi = 0.0;
            if (mag > 0.0) { pr = M[2 * (i * N + jm)] / mag; pi = M[2 * (i * N + jm) + 1] / mag; }
            d[((size_t)p * N + i) * 2] = pr; d[((size_t)p * N + i) * 2 + 1] = pi;
            for (int j = 0; j < N; j++) {
                const double xr = M[2 * (i * N + j)], xi = M[2 * (i * N + j) + 1];
                const double re = xr * pr + xi * pi, im = xi * pr - xr * pi;       // x * conj(phase)
                if (std::fabs(im) > 1e-14 * mag) { diag_real = false; break; }
                rm[((size_t)p * N + i) * N + j] = re;
            }
        }
    }
    // the X-rotation structure (dev_structs.h, QGT_FORM_PARITY): real part on even, imaginary part on odd index distance
    bool parity = !diag_real && allow_parity && N == 8;
    for (int p = 0; p < nvariants && parity; p++) {
        const double* M = &dense[(size_t)p * N * N * 2];
        double big = 0.0;
        for (int e = 0; e < N * N * 2; e++) big = std::max(big, std::fabs(M[e]));
        for (int i = 0; i < N && parity; i++)
            for (int j = 0; j < N; j++) {
                const int odd = __builtin_popcount((unsigned)(i ^ j)) & 1;
                if (std::fabs(M[2 * (i * N + j) + (odd ? 0 : 1)]) > 1e-14 * big) { parity = false; break; }
            }
    }
    if (parity) {
        for (int p = 0; p < nvariants; p++) {
            double* dst = &out[(size_t)p * vstride];
            const double* M = &dense[(size_t)p * N * N * 2];
            auto comp = [](int idx) { return ((((idx >> 2) ^ (idx >> 1) ^ idx) & 1) << 2) | (idx & 3); };     // (p, b1, b0) -> component
            for (int q = 0; q < 8; q++)
                for (int k = 0; k < 4; k++) {
                    const int r = comp(q), ce = comp(k), co = comp(4 + k);
                    const double re_e = M[2 * (r * N + ce)], im_e = M[2 * (r * N + ce) + 1];
                    const double re_o = M[2 * (r * N + co)], im_o = M[2 * (r * N + co) + 1];
                    // even rows (q < 4): [A_ee | -B_eo] for X, [A_ee | B_eo] for Y; odd rows: [B_oe | A_oo] for X, [-B_oe | A_oo] for Y
                    dst[2 * QGT_MIDX(N, q, k)] = q < 4 ? re_e : im_e;
                    dst[2 * QGT_MIDX(N, q, k) + 1] = q < 4 ? re_e : -im_e;
                    dst[2 * QGT_MIDX(N, q, 4 + k)] = q < 4 ? -im_o : re_o;
                    dst[2 * QGT_MIDX(N, q, 4 + k) + 1] = q < 4 ? im_o : re_o;
                }
        }
        return QGT_FORM_PARITY;
    }
    for (int p = 0; p < nvariants; p++) {
        double* dst = &out[(size_t)p * vstride];
        const double* M = &dense[(size_t)p * N * N * 2];
        for (int i = 0; i < N; i++) {
            for (int j = 0; j < N; j++) {
                if (diag_real) dst[QGT_MIDX(N, i, j)] = rm[((size_t)p * N + i) * N + j];
                else { dst[2 * QGT_MIDX(N, i, j)] = M[2 * (i * N + j)]; dst[2 * QGT_MIDX(N, i, j) + 1] = M[2 * (i * N + j) + 1]; }
            }
            if (diag_real) { dst[2 * (N * N + i)] = d[((size_t)p * N + i) * 2]; dst[2 * (N * N + i) + 1] = d[((size_t)p * N + i) * 2 + 1]; }
        }
    }
    return diag_real ? QGT_FORM_DIAG_REAL : QGT_FORM_DENSE;
}

int stage_matrices(const Run& run, const SubPass& sp, const Stage& st, int deriv_op, std::vector<double>& out) {
    std::vector<double> dense;
    stage_dense(run, sp, st, deriv_op, dense);
    return pack_stage(1 << (int)sp.reg_local.size(), 1 << (int)st.vqubits.size(), dense, out, run.parity_form);
}

int stage_matrices_sum(const Run& run, const SubPass& sp, const Stage& st, const std::vector<int>& deriv_ops, std::vector<double>& out) {
    std::vector<double> one, sum;
    for (int op : deriv_ops) {
        stage_dense(run, sp, st, op, one);
        if (sum.empty()) sum = one;
        else for (size_t i = 0; i < sum.size(); i++) sum[i] += one[i];
    }
    return pack_stage(1 << (int)sp.reg_local.size(), 1 << (int)st.vqubits.size(), sum, out, run.parity_form);
}

// A derivative column spawned at sub-pass s of a run recomputes the s earlier sub-passes phi has been through (the
// kernel applies the whole run to every item).  Where many parameters are born inside one long run - the first run
// of an ansatz, which fuses the light cone of its tile through every layer - the run is executed in pieces: same tile,
// consecutive sub-pass ranges, so the rest of the plan is unchanged.  Cost model in stage applications: a piece of s
// sub-passes costs (columns alive at its end) * (s + 1) (one pass over memory is worth about one stage); the cut points
// minimising the sum are found by dynamic programming; every piece also pays two kernel launches (~10 us, i.e.
// ~4 stage applications of a 2^20 state), which keeps small states in one piece.  `born` / `alive` carry the
// parameters seen so far.
static void split_run_by_births(Run& run, bool enable, int state_qubits, std::vector<char>& born, int& alive, std::vector<Run>& out) {
    const int S = (int)run.subs.size();
    std::vector<int> births(S, 0);
    int total = 0;
    for (int s = 0; s < S; s++)
        for (int i = run.subs[s].op_begin; i < run.subs[s].op_end; i++) {
            const int p = run.ops[i].param;
            if (p >= 0 && p < (int)born.size() && !born[p]) { born[p] = 1; births[s]++; total++; }
        }
    std::vector<int> cuts;                                   // piece boundaries (sub-pass indices), ascending
    if (enable && S >= 4 && total >= 4) {
        std::vector<long> cum(S + 1, 0);
        for (int s = 0; s < S; s++) cum[s + 1] = cum[s] + births[s];
        const double launch_units = state_qubits < 0 ? 0.0 : 4.2 * std::ldexp(1.0, 20 - std::min(state_qubits, 40));
        std::vector<double> best(S + 1, 0.0);
        std::vector<int> from(S + 1, 0);
        for (int j = 1; j <= S; j++) {
            best[j] = -1.0;
            for (int i = 0; i < j; i++) {
                const double cost = best[i] + (double)(alive + cum[j]) * (j - i + 1) + launch_units;
                if (best[j] < 0.0 || cost < best[j]) { best[j] = cost; from[j] = i; }
            }
        }
        for (int j = S; j > 0; j = from[j]) cuts.push_back(j);
        std::reverse(cuts.begin(), cuts.end());
    } else {
        cuts.push_back(S);
    }
    alive += total;
    if (cuts.size() <= 1) { out.push_back(std::move(run)); return; }
    int lo = 0;
    for (int hi : cuts) {
        Run piece;
        piece.K = run.K; piece.tile_qubits = run.tile_qubits; piece.other_qubits = run.other_qubits;
        piece.exchange_gbit = run.exchange_gbit; piece.exchange_mask = run.exchange_mask; piece.segment = run.segment;
        const int base = run.subs[lo].op_begin, end = run.subs[hi - 1].op_end;
        piece.ops.assign(run.ops.begin() + base, run.ops.begin() + end);
        for (int s = lo; s < hi; s++) {
            SubPass sp = run.subs[s];
            sp.op_begin -= base; sp.op_end -= base;
            for (Stage& st : sp.stages) for (int& o : st.ops) o -= base;
            for (int& o : sp.tdiags) o -= base;
            piece.subs.push_back(std::move(sp));
        }
        out.push_back(std::move(piece));
        lo = hi;
    }
}

int build_plan(const qgt_b200_circuit& c, const double* theta, const PlanOptions& opt_in, CircuitPlan& plan, std::string& err) {
    const int n = c.num_qubits;
    if (n < 1 || n > QGT_MAX_QUBITS) { err = "num_qubits out of range"; return QGT_B200_ERR_INVALID_ARG; }
    if (c.num_gates && !c.gates) { err = "gates is NULL"; return QGT_B200_ERR_INVALID_ARG; }
    PlanOptions opt = opt_in;
    if (const char* e = std::getenv("QGT_B200_BIRTH_CUT")) opt.birth_cut = std::atoi(e);   // test hook: 2 forces splits on small states
    opt.reg_qubits = std::max(1, std::min(opt.reg_qubits, 3));
    opt.batch_qubits = std::max(0, std::min(opt.batch_qubits, (int)QGT_MAX_REG_QUBITS - opt.reg_qubits));
    opt.tile_qubits = std::max(opt.reg_qubits + opt.batch_qubits + 2,
                               std::min(opt.tile_qubits, std::min((int)QGT_MAX_TILE_QUBITS, opt.reg_qubits + opt.batch_qubits + 8)));
    const int nl = (opt.local_qubits > 0 && opt.local_qubits < n) ? opt.local_qubits : n;   // qubits inside one shard
    const int K = std::min(nl, opt.tile_qubits);
    const int R = std::min(opt.reg_qubits, K);
    const int B = std::min(opt.batch_qubits, K - R);
    // the forced low qubits must leave room for at least R freely chosen tile qubits
    const int L = (K == nl) ? K : std::max(0, std::min(opt.low_qubits, K - R));
    plan = CircuitPlan();
    plan.n = n; plan.nloc = nl; plan.P = c.num_params; plan.opt = opt; plan.K = K; plan.R = R; plan.B = B;
    plan.first_run.assign(std::max(0, c.num_params), -1);
    plan.last_run.assign(std::max(0, c.num_params), -1);

    std::vector<LoweredOp> ops;
    std::vector<int> gate_order;
    {
        int wf = opt.wavefront;
        if (const char* e = std::getenv("QGT_B200_WAVEFRONT")) wf = std::atoi(e);
        if (wf > 0 && nl == n && n - wf >= 2 && c.num_gates > 0) {
            std::vector<MappedSegment> dummy;
            std::string werr;
            if (map_circuit_sharded(c, n - wf, false, dummy, werr, &gate_order) != QGT_B200_OK || gate_order.size() != c.num_gates) gate_order.clear();
        }
    }
    for (size_t g = 0; g < c.num_gates; g++) {
        int rc = lower_gate(c, theta, gate_order.empty() ? (int)g : gate_order[g], ops, err);
        if (rc) return rc;
    }
    const int N = (int)ops.size();
    for (const LoweredOp& o : ops)
        if (o.target >= nl) {
            err = "a non-diagonal gate targets a rank (global) qubit: the circuit must be mapped first";
            return QGT_B200_ERR_CIRCUIT;
        }
    Dag dag = build_dag(ops, n);
    std::vector<char> done(N, 0);
    std::vector<int> ready;
    for (int i = 0; i < N; i++) if (dag.indeg[i] == 0) ready.push_back(i);
    int scheduled = 0;
    // Parameters already born before the current run (see split_run_by_births).
    std::vector<char> born(std::max(1, c.num_params), 0);
    int alive = 1;

    while (scheduled < N) {
        Run run;
        run.K = K;

        std::vector<char> inS(n, 0);
        int sizeS = 0;
        for (int q = 0; q < L; q++) { inS[q] = 1; sizeS++; }
        if (K == nl) { for (int q = 0; q < nl; q++) inS[q] = 1; sizeS = nl; }
        struct RawSub { std::vector<int> regq; int b, e; bool is_cost = false; };
        std::vector<RawSub> raw;
        bool run_full = false;
        while (!run_full) {
            RawSub sp; sp.b = (int)run.ops.size(); sp.is_cost = false;
            std::vector<char> inR(n, 0);
            int sizeR = 0;
            for (;;) {
                if ((int)run.ops.size() >= opt.max_ops_per_run) { run_full = true; break; }
                int best = -1, best_pri = 99;
                for (int i : ready) {
                    const LoweredOp& o = ops[i];
                    int pri;
                    // 2x2 ops first (on a held register qubit, then a new register qubit of the tile, then a
                    // new tile qubit); diagonal ops last so that those on register qubits end up adjacent
                    // and merge into one table; the cost layer gets a tile-level pass of its own.
                    // A diagonal gate that carries a parameter is placed like a 2x2 gate on its (lowest local) qubit:
                    // its generator has to sit in a dense stage, and waiting for a sub-pass that holds the qubit anyway
                    // costs nothing, whereas taking it "for free" now would add an identity x phase stage of its own.
                    const int tq = sched_target(o, nl);
                    if (o.type == QGT_OP_COST) pri = (sp.b == (int)run.ops.size()) ? 4 : 99;
                    else if (tq < 0) pri = 3;
                    else if (inR[tq]) pri = 0;
                    else if (sizeR < R && inS[tq]) pri = 1;
                    else if (sizeR < R && sizeS < K) pri = 2;
                    else continue;
                    if (pri == 99) continue;
                    if (pri < best_pri || (pri == best_pri && i < best)) { best = i; best_pri = pri; }
                }
                if (best < 0) break;
                const LoweredOp& o = ops[best];
                if (o.type == QGT_OP_COST) sp.is_cost = true;
                const int tqb = sched_target(o, nl);
                if (tqb >= 0) {
                    if (!inS[tqb]) { inS[tqb] = 1; sizeS++; }
                    if (!inR[tqb]) { inR[tqb] = 1; sizeR++; sp.regq.push_back(tqb); }
                }
                run.ops.push_back(o);
                done[best] = 1; scheduled++;
                ready.erase(std::find(ready.begin(), ready.end(), best));
                for (int sidx : dag.succ[best]) if (--dag.indeg[sidx] == 0) ready.push_back(sidx);
                if (sp.is_cost) break;            // a cost pass holds exactly one op
            }
            sp.e = (int)run.ops.size();
            if (sp.e == sp.b) break;          // nothing fits a fresh sub-pass: the run is complete
            raw.push_back(sp);
        }
        if (run.ops.empty()) { err = "planner made no progress"; return QGT_B200_ERR_INTERNAL; }
        // pad the tile to K qubits with the lowest unused ones
        for (int q = 0; q < nl && sizeS < K; q++) if (!inS[q]) { inS[q] = 1; sizeS++; }
        for (int q = 0; q < nl; q++) (inS[q] ? run.tile_qubits : run.other_qubits).push_back(q);
        std::vector<int> local_of(n, -1);
        for (int j = 0; j < K; j++) local_of[run.tile_qubits[j]] = j;
        for (const RawSub& rs : raw) {
            SubPass sp;
            sp.op_begin = rs.b; sp.op_end = rs.e;
            sp.is_cost = rs.is_cost;
            if (rs.is_cost) { run.subs.push_back(sp); continue; }
            sp.nreg_used = (int)rs.regq.size();
            std::vector<char> used(K, 0);
            for (int q : rs.regq) { sp.reg_local.push_back(local_of[q]); used[local_of[q]] = 1; }
            for (int p = K - 1; p >= 0 && (int)sp.reg_local.size() < R; p--)
                if (!used[p]) { sp.reg_local.push_back(p); used[p] = 1; }
            std::sort(sp.reg_local.begin(), sp.reg_local.end());
            uint64_t regmask = 0;
            for (int p : sp.reg_local) regmask |= bit(run.tile_qubits[p]);
            for (int i = rs.b; i < rs.e; i++) {
                const LoweredOp& o = run.ops[i];
                const uint64_t used_bits = o.cmask | o.pmask | (o.target >= 0 ? bit(o.target) : 0);
                // a diagonal gate off the matrix qubits is one phase per thread - unless it carries a parameter: its
                // generator must sit in a dense stage (as an identity x phase matrix selected by variant bits) so that
                // the fused schedule finds <column| G |phi> in the stage's transition matrix
                if (o.type == QGT_OP_DIAG && (used_bits & regmask) == 0 && o.param < 0) { sp.tdiags.push_back(i); continue; }
                std::vector<int> need;               // non-register qubits this op depends on
                for (int q = 0; q < n; q++) if ((used_bits & ~regmask) >> q & 1) need.push_back(q);
                bool fits = !sp.stages.empty();
                if (fits) {
                    std::vector<int> merged = sp.stages.back().vqubits;
                    for (int q : need) if (std::find(merged.begin(), merged.end(), q) == merged.end()) merged.push_back(q);
                    if ((int)merged.size() > QGT_MAX_VARIANT_BITS) fits = false;
                    else sp.stages.back().vqubits = merged;
                }
                if (!fits) { Stage st; st.vqubits = need; sp.stages.push_back(st); }
                sp.stages.back().ops.push_back(i);
            }
            // batch positions: free tile positions, preferably not selecting a matrix variant of this
            // sub-pass (then both halves of a thread always share their matrix), highest first
            {
                std::vector<char> is_var(K, 0);
                for (const Stage& stg : sp.stages)
                    for (int q : stg.vqubits) if (local_of[q] >= 0) is_var[local_of[q]] = 1;
                for (int pass = 0; pass < 2 && (int)sp.batch_local.size() < B; pass++)
                    for (int p = K - 1; p >= 0 && (int)sp.batch_local.size() < B; p--)
                        if (!used[p] && (pass == 1 || !is_var[p])) { sp.batch_local.push_back(p); used[p] = 1; }
            }
            {
                std::vector<int> variant_local;
                for (const Stage& stg : sp.stages)
                    for (int q : stg.vqubits) if (local_of[q] >= 0) variant_local.push_back(local_of[q]);
                sp.mma_ok = make_tperm(K, sp.reg_local, sp.batch_local, variant_local, sp.tperm);
            }
            run.subs.push_back(sp);
        }
        std::vector<Run> pieces;
        split_run_by_births(run, opt.birth_cut != 0, opt.birth_cut == 2 ? -1 : nl, born, alive, pieces);
        for (Run& piece : pieces) {
            const int ridx = (int)plan.runs.size();
            {   // transition-matrix layout of the fused schedule
                int sidx = 0;
                for (SubPass& sp : piece.subs)
                    for (Stage& st : sp.stages) {
                        st.params.clear();
                        for (int o : st.ops) {
                            const int p = piece.ops[o].param;
                            if (p >= 0 && std::find(st.params.begin(), st.params.end(), p) == st.params.end()) st.params.push_back(p);
                        }
                        st.rho_off = -1; st.traj_ord = -1;
                        if (!st.params.empty()) {
                            st.rho_off = piece.rho_blocks;
                            piece.rho_blocks += 1 << (int)st.vqubits.size();
                            piece.last_rho_stage = sidx;
                            st.traj_ord = piece.rho_stages++;
                        }
                        sidx++;
                    }
            }
            for (int i = 0; i < (int)piece.ops.size(); i++) {
                const int p = piece.ops[i].param;
                if (p < 0) continue;
                piece.occ.push_back({p, i});
                if (plan.first_run[p] < 0) plan.first_run[p] = ridx;
                plan.last_run[p] = ridx;
            }
            plan.runs.push_back(std::move(piece));
        }
    }
    // circuits with a cost pass stay in the sweep kernel (plan_supports_fused), which knows the X-rotation matrix form
    bool any_cost = false;
    for (const Run& run : plan.runs) for (const SubPass& sp : run.subs) any_cost = any_cost || sp.is_cost;
    for (Run& run : plan.runs) run.parity_form = any_cost && opt.parity_form != 0;
    return QGT_B200_OK;
}

void build_image(const CircuitPlan& plan, PlanImage& img) {
    img = PlanImage();
    std::vector<double> mats;
    for (const Run& run : plan.runs) {
        QgtDevRun dr;
        std::memset(&dr, 0, sizeof dr);
        dr.K = run.K; dr.n = plan.nloc;
        dr.nsub = (int)run.subs.size();
        dr.sub_off = (int)img.subs.size();
        dr.stage_off = (int)img.stages.size();
        dr.tdiag_off = (int)img.tdiags.size();
        dr.cost_off = (int)img.costs.size();
        dr.mat_off = (int)(img.pool.size() / 2);
        for (size_t j = 0; j < run.tile_qubits.size(); j++) dr.tq[j] = (int8_t)run.tile_qubits[j];   // empty for an exchange pseudo-run
        for (size_t j = 0; j < run.other_qubits.size(); j++) dr.ntq[j] = (int8_t)run.other_qubits[j];
        for (const SubPass& sp : run.subs) {
            QgtDevSubPass ds;
            std::memset(&ds, 0, sizeof ds);
            ds.nreg = sp.is_cost ? 0 : (int)sp.reg_local.size();
            ds.stage_begin = (int)img.stages.size() - dr.stage_off;
            ds.tdiag_begin = (int)img.tdiags.size() - dr.tdiag_off;
            ds.cost = (int)img.costs.size() - dr.cost_off;
            ds.mma_ok = sp.mma_ok ? 1 : 0;
            auto swz = [](uint32_t idx) { return idx ^ ((idx >> 3) & 7u) ^ ((idx >> 6) & 7u) ^ ((idx >> 9) & 7u); };
            for (size_t r = 0; r < sp.reg_local.size() + sp.batch_local.size(); r++) {
                const int p = r < sp.reg_local.size() ? sp.reg_local[r] : sp.batch_local[r - sp.reg_local.size()];
                ds.s_reg[r] = swz(1u << p);
                ds.g_reg[r] = bit(run.tile_qubits[p]);
            }
            for (size_t t = 0; t < sp.tperm.size(); t++) {
                ds.s_thr[t] = swz(1u << sp.tperm[t]);
                ds.g_thr[t] = bit(run.tile_qubits[sp.tperm[t]]);
            }
            if (sp.is_cost) img.costs.push_back(make_cost(run.ops[sp.op_begin], false));
            for (int o : sp.tdiags) img.tdiags.push_back(make_tdiag(run.ops[o], false));
            for (const Stage& st : sp.stages) {
                QgtDevStage d;
                std::memset(&d, 0, sizeof d);
                d.nvar = (int16_t)st.vqubits.size();
                for (int k = 0; k < d.nvar; k++) d.vmask[k] = bit(st.vqubits[k]);
                d.mat_off = (int)(img.pool.size() / 2) - dr.mat_off;
                d.rho_off = st.rho_off;
                d.traj_ord = st.traj_ord;
                d.form = (int16_t)stage_matrices(run, sp, st, -1, mats);
                img.pool.insert(img.pool.end(), mats.begin(), mats.end());
                img.stages.push_back(d);
            }
            ds.stage_end = (int)img.stages.size() - dr.stage_off;
            ds.tdiag_end = (int)img.tdiags.size() - dr.tdiag_off;
            img.subs.push_back(ds);
        }
        dr.mat_count = (int)(img.pool.size() / 2) - dr.mat_off;
        for (const SubPass& sp : run.subs) if (sp.is_cost) dr.has_cost = 1;
        dr.rho_blocks = run.rho_blocks;
        dr.last_rho_stage = run.last_rho_stage;
        img.runs.push_back(dr);
    }
}

// ---- sharded states --------------------------------------------------------------------------------
namespace {
// qubits a gate acts on non-diagonally (they must be local when the gate runs)
void nondiag_qubits(const qgt_b200_gate& g, int out[2], int& cnt) {
    cnt = 0;
    switch (g.kind) {
    case QGT_B200_GATE_X: case QGT_B200_GATE_Y: case QGT_B200_GATE_H: case QGT_B200_GATE_SX:
    case QGT_B200_GATE_RX: case QGT_B200_GATE_RY: case QGT_B200_GATE_CNOT: case QGT_B200_GATE_CY:
    case QGT_B200_GATE_CH: case QGT_B200_GATE_CRX: case QGT_B200_GATE_CRY:
        out[cnt++] = g.target; break;
    case QGT_B200_GATE_SWAP: out[cnt++] = g.target; out[cnt++] = g.control; break;
    default: break;
    }
}
}  // namespace

int map_circuit_sharded(const qgt_b200_circuit& c, int nloc, bool restore_identity,
                        std::vector<MappedSegment>& segs, std::string& err, std::vector<int>* order) {
    const int n = c.num_qubits;
    bool batch_exchanges = true;
    if (const char* e = std::getenv("QGT_B200_BATCH_EXCHANGES")) batch_exchanges = std::atoi(e) != 0;   // 0: one rank bit per exchange
    segs.clear();
    if (nloc < 2 || nloc > n) { err = "invalid shard size"; return QGT_B200_ERR_INVALID_ARG; }
    std::vector<int> phys(n), logical(n);
    for (int q = 0; q < n; q++) { phys[q] = q; logical[q] = q; }
    MappedSegment cur;
    cur.phys_of_logical = phys;
    auto close_segment = [&](unsigned mask) {
        cur.exchange_mask = mask;
        cur.exchange_gbit = mask ? __builtin_ctz(mask) : -1;
        segs.push_back(cur);
        cur = MappedSegment();
    };
    auto local_swap = [&](int pa, int pb) {          // physical positions, both local
        if (pa == pb) return;
        qgt_b200_gate sw = {QGT_B200_GATE_SWAP, pa, pb, -1, 0.0, 1.0};
        cur.gates.push_back(sw);
        const int la = logical[pa], lb = logical[pb];
        phys[la] = pb; logical[pb] = la;
        phys[lb] = pa; logical[pa] = lb;
    };
    // Bring the logical qubits `in` (all on rank bits) into the top local positions in ONE exchange, evicting `out` (local
    // logical qubits, same count).  The rank bits involved, ascending, pair with local positions nloc-k .. nloc-1.
    auto bring_in = [&](std::vector<int> in, const std::vector<int>& out) {
        const int k = (int)in.size();
        std::sort(in.begin(), in.end(), [&](int a, int b) { return phys[a] < phys[b]; });
        for (int i = 0; i < k; i++) {                // victim i must sit at local position nloc-k+i
            const int want = nloc - k + i;
            // a victim already parked at a wanted position keeps it if possible: place victims greedily
            local_swap(phys[out[i]], want);
        }
        unsigned mask = 0;
        for (int q : in) mask |= 1u << (phys[q] - nloc);
        close_segment(mask);
        for (int i = 0; i < k; i++) {
            const int q = in[i], v = logical[nloc - k + i];
            const int pg = phys[q], pl = nloc - k + i;
            phys[q] = pl; logical[pl] = q;
            phys[v] = pg; logical[pg] = v;
        }
        cur.phys_of_logical = phys;
    };
    // ---- gate DAG: two uses of a qubit commute when both are diagonal on it (controls count as diagonal) ----
    const size_t NG = c.num_gates;
    std::vector<std::vector<int>> g_nd(NG), g_dg(NG), succ(NG);
    std::vector<int> indeg(NG, 0);
    {
        std::vector<int> last_nd(n, -1);
        std::vector<std::vector<int>> diag_since(n);
        for (size_t gi = 0; gi < NG; gi++) {
            const qgt_b200_gate& g = c.gates[gi];
            int nd[2], cnt;
            nondiag_qubits(g, nd, cnt);
            for (int k = 0; k < cnt; k++) {
                if (nd[k] < 0 || nd[k] >= n) { err = "gate qubit out of range"; return QGT_B200_ERR_CIRCUIT; }
                g_nd[gi].push_back(nd[k]);
            }
            auto is_nd = [&](int q) { return std::find(g_nd[gi].begin(), g_nd[gi].end(), q) != g_nd[gi].end(); };
            if (g.kind == QGT_B200_GATE_COST) { for (int q = 0; q < n; q++) g_dg[gi].push_back(q); }
            else if (g.kind != QGT_B200_GATE_I) {
                if (g.target >= 0 && g.target < n && !is_nd(g.target)) g_dg[gi].push_back(g.target);
                if (is_two_qubit(g.kind) && g.control >= 0 && g.control < n && !is_nd(g.control)) g_dg[gi].push_back(g.control);
            }
            std::set<int> preds;
            for (int q : g_dg[gi]) { if (last_nd[q] >= 0) preds.insert(last_nd[q]); diag_since[q].push_back((int)gi); }
            for (int q : g_nd[gi]) {
                if (last_nd[q] >= 0) preds.insert(last_nd[q]);
                for (int j : diag_since[q]) preds.insert(j);
                last_nd[q] = (int)gi; diag_since[q].clear();
            }
            for (int pgi : preds) { succ[pgi].push_back((int)gi); indeg[gi]++; }
        }
    }
    std::vector<char> emitted(NG, 0);
    std::set<int> ready;
    for (size_t gi = 0; gi < NG; gi++) if (indeg[gi] == 0) ready.insert((int)gi);
    // next not-yet-emitted gate (by list position) that acts on logical qubit v non-diagonally
    auto next_use = [&](int v) {
        for (size_t gj = 0; gj < NG; gj++)
            if (!emitted[gj] && std::find(g_nd[gj].begin(), g_nd[gj].end(), v) != g_nd[gj].end()) return gj;
        return NG + 1;
    };
    size_t done = 0;
    while (done < NG) {
        // everything that can run with the current placement, in list order (gates float across exchanges they do not
        // depend on, so a segment is as long as the data dependencies allow and fuses into few runs)
        bool progress = true;
        while (progress) {
            progress = false;
            for (auto it = ready.begin(); it != ready.end();) {
                const int gi = *it;
                bool local = true;
                for (int q : g_nd[gi]) if (phys[q] >= nloc) local = false;
                if (!local) { ++it; continue; }
                qgt_b200_gate pg = c.gates[gi];
                if (pg.kind != QGT_B200_GATE_COST) {
                    pg.target = phys[pg.target];
                    if (pg.control >= 0 && pg.control < n) pg.control = phys[pg.control];
                }
                cur.gates.push_back(pg);
                if (order) order->push_back(gi);
                emitted[gi] = 1; done++;
                it = ready.erase(it);
                for (int sidx : succ[gi]) if (--indeg[sidx] == 0) { ready.insert(sidx); progress = true; }
            }
        }
        if (done == NG) break;
        // every ready gate waits for a qubit on a rank bit.  Bring in: the rank-bit qubits in order of need (those of ready
        // gates first, by list position); evict: local qubits no ready gate needs, farthest next use first (Belady).  A pair
        // joins the batch while the incoming qubit is needed before the outgoing one.
        std::vector<int> need_now(n, 0);
        for (int gi : ready) for (int q : g_nd[gi]) need_now[q] = 1;
        const int first_ready = *ready.begin();
        std::vector<std::pair<size_t, int>> globals, locals;
        for (int v = 0; v < n; v++) {
            if (phys[v] >= nloc) {
                size_t when = next_use(v);
                if (need_now[v]) {
                    bool in_first = std::find(g_nd[first_ready].begin(), g_nd[first_ready].end(), v) != g_nd[first_ready].end();
                    when = in_first ? 0 : std::min(when, (size_t)first_ready + 1);
                }
                globals.push_back({when, v});
            } else if (!need_now[v]) {
                locals.push_back({next_use(v), v});
            }
        }
        std::sort(globals.begin(), globals.end());
        std::sort(locals.begin(), locals.end(), [](const std::pair<size_t, int>& x, const std::pair<size_t, int>& y) {
            return x.first != y.first ? x.first > y.first : x.second < y.second; });
        std::vector<int> in, out;
        size_t must_count = 0;
        for (const auto& gl : globals) if (gl.first == 0) must_count++;
        for (size_t i = 0; i < globals.size() && i < locals.size(); i++) {
            const bool must = globals[i].first == 0;
            if (!must && !(batch_exchanges && globals[i].first < locals[i].first)) break;
            in.push_back(globals[i].second); out.push_back(locals[i].second);
        }
        if (in.size() < must_count || in.empty()) { err = "no local qubit available to exchange"; return QGT_B200_ERR_INTERNAL; }
        bring_in(in, out);
    }
    if (restore_identity) {
        // undo the permutation: fix the rank bits first (each needs its own logical qubit local at the top)
        for (int pq = n - 1; pq >= nloc; pq--) {
            if (logical[pq] == pq) continue;
            const int want = pq;                       // logical qubit that belongs on this rank bit
            if (phys[want] >= nloc) {                  // it sits on another rank bit: bring it in first
                int victim = -1;
                for (int v = 0; v < n; v++) if (phys[v] < nloc && v < nloc) { victim = v; break; }
                if (victim < 0) for (int v = 0; v < n; v++) if (phys[v] < nloc) { victim = v; break; }
                bring_in(std::vector<int>(1, want), std::vector<int>(1, victim));
            }
            // now `want` is local: move it to the top and exchange it with rank bit pq
            const int top = nloc - 1;
            if (phys[want] != top) {
                const int w = logical[top], pv = phys[want];
                qgt_b200_gate sw = {QGT_B200_GATE_SWAP, pv, top, -1, 0.0, 1.0};
                cur.gates.push_back(sw);
                phys[w] = pv; logical[pv] = w;
                phys[want] = top; logical[top] = want;
            }
            const int other = logical[pq];
            close_segment(1u << (pq - nloc));
            phys[want] = pq; logical[pq] = want;
            phys[other] = top; logical[top] = other;
            cur.phys_of_logical = phys;
        }
        // then sort the local qubits with local SWAPs
        for (int pq = 0; pq < nloc; pq++) {
            if (logical[pq] == pq) continue;
            const int pv = phys[pq], w = logical[pq];
            qgt_b200_gate sw = {QGT_B200_GATE_SWAP, pv, pq, -1, 0.0, 1.0};
            cur.gates.push_back(sw);
            phys[w] = pv; logical[pv] = w;
            phys[pq] = pq; logical[pq] = pq;
        }
    }
    close_segment(0);
    return QGT_B200_OK;
}

int build_plan_sharded(const qgt_b200_circuit& c, const double* theta, const PlanOptions& opt_in, int nloc, bool restore_identity,
                       CircuitPlan& plan, std::vector<MappedSegment>& segs, std::string& err) {
    int rc = map_circuit_sharded(c, nloc, restore_identity, segs, err);
    if (rc) return rc;
    PlanOptions opt = opt_in;
    opt.local_qubits = nloc;
    plan = CircuitPlan();
    for (size_t si = 0; si < segs.size(); si++) {
        qgt_b200_circuit sub = c;
        sub.gates = segs[si].gates.data();
        sub.num_gates = segs[si].gates.size();
        CircuitPlan sp;
        opt.birth_cut = opt_in.birth_cut && si == 0;      // later segments start with columns alive the planner call cannot see
        if ((rc = build_plan(sub, theta, opt, sp, err))) return rc;
        if (si == 0) { plan = sp; plan.runs.clear(); }
        for (Run& r : sp.runs) { r.segment = (int)si; plan.runs.push_back(std::move(r)); }
        if (segs[si].exchange_gbit >= 0) {
            Run ex;
            ex.K = plan.K;
            ex.exchange_gbit = segs[si].exchange_gbit;
            ex.exchange_mask = segs[si].exchange_mask;
            ex.segment = (int)si;
            plan.runs.push_back(ex);
        }
    }
    plan.first_run.assign(std::max(0, c.num_params), -1);
    plan.last_run.assign(std::max(0, c.num_params), -1);
    for (size_t r = 0; r < plan.runs.size(); r++)
        for (const ParamOcc& oc : plan.runs[r].occ) {
            if (plan.first_run[oc.param] < 0) plan.first_run[oc.param] = (int)r;
            plan.last_run[oc.param] = (int)r;
        }
    return QGT_B200_OK;
}

// ---- QGT column schedule ---------------------------------------------------------------------
namespace {

struct Sched {
    const CircuitPlan& plan;
    Program& prog;
    int psi;                         // slot of the marching state phi
    int psi_alt = -1;                // spare slot: phi advances out of place inside the columns' launch (resident mode)
    int ckpt;                        // slot of the rolling checkpoint (blocked mode) or -1
    int ckpt_time = 0;               // the checkpoint holds the state BEFORE run ckpt_time
    std::vector<int> res_slots, str_slots;

    Sched(const CircuitPlan& p, Program& g) : plan(p), prog(g), psi(0), ckpt(-1) {}

    void sweep(int run, const std::vector<SweepCol>& cols) {
        if (cols.empty()) return;
        Instr in; in.kind = INSTR_SWEEP; in.run = run; in.cols = cols;
        prog.instrs.push_back(std::move(in));
    }
    void gram(const std::vector<int>& as, const std::vector<int>& aid, const std::vector<int>& bs, const std::vector<int>& bid) {
        if (as.empty() || bs.empty()) return;
        Instr in; in.kind = INSTR_GRAM; in.a_slots = as; in.a_ids = aid; in.b_slots = bs; in.b_ids = bid;
        prog.instrs.push_back(std::move(in));
    }
    void copy(int src, int dst) { Instr in; in.kind = INSTR_COPY; in.src = src; in.dst = dst; prog.instrs.push_back(in); }
    void init(int dst) { Instr in; in.kind = INSTR_INIT; in.dst = dst; prog.instrs.push_back(in); }

    // accumulate-spawn launches: occurrences beyond the first of a parameter in one run (and every
    // occurrence of an already alive column) write dst += ..., so no two items of a launch may share dst
    void emit_accumulates(int run, std::vector<SweepCol>& acc) {
        while (!acc.empty()) {
            std::vector<SweepCol> now, later;
            std::set<int> dsts;
            for (const SweepCol& c : acc) (dsts.insert(c.dst).second ? now : later).push_back(c);
            sweep(run, now);
            acc.swap(later);
        }
    }

    // occurrences of parameter p in run r, grouped by dense stage (product rule inside a stage = one item)
    std::vector<SweepCol> occurrence_items(int r, int p, int src, int dst) {
        const Run& run = plan.runs[r];
        std::vector<SweepCol> items;
        std::vector<int> stage_of_item;
        for (const ParamOcc& oc : run.occ) {
            if (oc.param != p) continue;
            const OpLocation loc = locate_op(run, oc.op);
            int found = -1;
            if (loc.kind == 1)
                for (size_t k = 0; k < items.size(); k++) if (stage_of_item[k] == loc.index) found = (int)k;
            if (found >= 0) { items[found].ovr_extra.push_back(oc.op); continue; }
            SweepCol sc; sc.src = src; sc.dst = dst; sc.ovr_op = oc.op; sc.accumulate = true;
            items.push_back(sc);
            stage_of_item.push_back(loc.kind == 1 ? loc.index : -1 - (int)items.size());
        }
        return items;
    }

    // March phi from run r0 with `resident` columns (parameters) and, per run, the streaming
    // parameters in `streaming` (all born at or after r0).  diag: also emit the resident x resident
    // (+psi) Gram at T_res.
    void march(int r0, const std::vector<int>& resident, const std::vector<int>& streaming, bool diag, bool to_end) {
        const int R = (int)plan.runs.size();
        const int P = plan.P;
        std::vector<int> slot_of(P, -1);
        std::vector<char> is_res(P, 0), is_str(P, 0), alive(P, 0);
        int T_res = r0;
        for (size_t i = 0; i < resident.size(); i++) {
            slot_of[resident[i]] = res_slots[i]; is_res[resident[i]] = 1;
            T_res = std::max(T_res, plan.last_run[resident[i]]);
        }
        for (int p : streaming) is_str[p] = 1;
        std::vector<int> free_str(str_slots.rbegin(), str_slots.rend());
        size_t str_left = streaming.size();
        bool diag_done = !diag;
        std::vector<int> res_ids(resident.begin(), resident.end());
        std::vector<int> res_sl;
        for (int p : resident) res_sl.push_back(slot_of[p]);

        for (int r = r0; r < R; r++) {
            const Run& run = plan.runs[r];
            // 1. advance every alive column; first occurrences of resident parameters born here
            std::vector<SweepCol> A, acc;
            std::vector<int> alive_str_now;
            for (int p = 0; p < P; p++) if (alive[p]) { A.push_back({slot_of[p], slot_of[p], -1, false}); if (is_str[p]) alive_str_now.push_back(p); }
            std::vector<int> born_str;
            std::vector<char> spawned_here(P, 0);
            std::vector<char> seen_here(P, 0);
            for (const ParamOcc& oc : run.occ) {
                const int p = oc.param;
                if (seen_here[p]) continue;
                seen_here[p] = 1;
                if (is_res[p] || (is_str[p] && alive[p])) {
                    std::vector<SweepCol> items = occurrence_items(r, p, psi, slot_of[p]);
                    for (size_t k = 0; k < items.size(); k++) {
                        if (k == 0 && !alive[p]) { items[k].accumulate = false; A.push_back(items[k]); spawned_here[p] = 1; }
                        else acc.push_back(items[k]);
                    }
                } else if (is_str[p]) {
                    born_str.push_back(p); spawned_here[p] = 1;
                }
            }
            // phi rides in the same launch, written to the spare slot: the spawns of this run (here and below) keep
            // reading the old phi, and one launch per run disappears
            const bool phi_oop = psi_alt >= 0 && run.exchange_gbit < 0;
            if (phi_oop) A.push_back({psi, psi_alt, -1, false});
            sweep(r, A);
            emit_accumulates(r, acc);
            for (int p = 0; p < P; p++) if (spawned_here[p] && is_res[p]) alive[p] = 1;

            const bool res_complete = (r >= T_res);
            // 2. streaming columns that were alive and are ready now
            if (res_complete) {
                std::vector<int> bs, bid;
                for (int p : alive_str_now)
                    if (plan.last_run[p] <= r) { bs.push_back(slot_of[p]); bid.push_back(p); }
                if (!bs.empty()) {
                    gram(res_sl, res_ids, bs, bid);
                    for (int p : bid) { alive[p] = 0; free_str.push_back(slot_of[p]); slot_of[p] = -1; str_left--; }
                }
            }
            // 3. streaming parameters born in this run, in rounds of the free slots
            size_t bi = 0;
            // persistent ones first (they keep their slot beyond this run)
            std::stable_sort(born_str.begin(), born_str.end(), [&](int x, int y) {
                const bool px = !(res_complete && plan.last_run[x] <= r), py = !(res_complete && plan.last_run[y] <= r);
                return px > py;
            });
            while (bi < born_str.size()) {
                std::vector<SweepCol> S, sacc;
                std::vector<int> round;
                while (bi < born_str.size() && !free_str.empty()) {
                    const int p = born_str[bi++];
                    slot_of[p] = free_str.back(); free_str.pop_back();
                    round.push_back(p);
                    std::vector<SweepCol> items = occurrence_items(r, p, psi, slot_of[p]);
                    for (size_t k = 0; k < items.size(); k++) {
                        if (k == 0) { items[k].accumulate = false; S.push_back(items[k]); }
                        else sacc.push_back(items[k]);
                    }
                }
                sweep(r, S);
                emit_accumulates(r, sacc);
                std::vector<int> bs, bid;
                for (int p : round) {
                    if (res_complete && plan.last_run[p] <= r) { bs.push_back(slot_of[p]); bid.push_back(p); }
                    else alive[p] = 1;
                }
                if (!bs.empty()) {
                    gram(res_sl, res_ids, bs, bid);
                    for (int p : bid) { free_str.push_back(slot_of[p]); slot_of[p] = -1; str_left--; }
                }
                if (round.empty()) break;   // no free slot at all: select_fit() prevents this
            }
            // 4. phi itself
            if (phi_oop) std::swap(psi, psi_alt);
            else sweep(r, {{psi, psi, -1, false}});
            // 5. resident x (resident, psi) at the first time every resident column is complete
            if (!diag_done && res_complete) {
                std::vector<int> bs = res_sl, bid = res_ids;
                bs.push_back(psi); bid.push_back(P);
                gram(res_sl, res_ids, bs, bid);
                diag_done = true;
            }
            if (diag_done && str_left == 0 && !to_end) return;
        }
    }
};

// choose the streaming parameters one march can serve with c slots; the rest waits for another pass
void select_fit(const CircuitPlan& plan, const std::vector<int>& pending, int c, int T_res,
                std::vector<int>& take, std::vector<int>& rest) {
    const int R = (int)plan.runs.size();
    std::vector<int> occ(R + 1, 0);
    take.clear(); rest.clear();
    // one slot stays free for the transient parameters (born and ready in the same run) - when there are any
    bool any_transient = false;
    for (int p : pending) any_transient = any_transient || std::max(plan.last_run[p], T_res) == plan.first_run[p];
    const int cap = any_transient ? c - 1 : c;
    for (int p : pending) {
        const int f = plan.first_run[p];
        const int ready = std::max(plan.last_run[p], T_res);
        if (ready == f) { take.push_back(p); continue; }     // transient: uses the reserved slot
        bool ok = true;
        for (int r = f; r <= ready; r++) if (occ[r] >= cap) { ok = false; break; }
        if (ok) { for (int r = f; r <= ready; r++) occ[r]++; take.push_back(p); }
        else rest.push_back(p);
    }
    if (take.empty() && !rest.empty()) { take.push_back(rest.front()); rest.erase(rest.begin()); }
}

}  // namespace

// HBM traffic of a Gram-schedule program in column reads / writes of 16 bytes per amplitude (a sweep item reads and writes
// its column, a Gram reads each operand once, a copy reads and writes)
static double program_traffic(const Program& prog) {
    double t = 0.0;
    for (const Instr& in : prog.instrs) {
        if (in.kind == INSTR_SWEEP || in.kind == INSTR_FUSED) t += 2.0 * (double)in.cols.size();
        else if (in.kind == INSTR_GRAM) t += (double)in.a_slots.size() + (double)in.b_slots.size();
        else if (in.kind == INSTR_COPY) t += 2.0;
    }
    return t;
}

static int build_qgt_program_split(const CircuitPlan& plan, size_t total_slots, bool want_psi, Program& prog, std::string& err, int c_forced);

int build_qgt_program(const CircuitPlan& plan, size_t total_slots, bool want_psi, Program& prog, std::string& err) {
    int Pa = 0;
    for (int p = 0; p < plan.P; p++) if (plan.first_run[p] >= 0) Pa++;
    const bool blocked = !((size_t)Pa + 1 <= total_slots);
    if (!blocked || total_slots < 5 || Pa > 64) return build_qgt_program_split(plan, total_slots, want_psi, prog, err, 0);
    // few parameters on a state that leaves room for a handful of columns (30-qubit QAOA: 16 parameters, 9 columns): how the
    // columns are split between the resident block and the streaming ones decides how often the block is marched again
    // (every march serves c - 1 long-lived streaming columns).  The programs are small: build one per split, keep the one
    // that moves the fewest bytes.
    const int avail = (int)total_slots - 2;
    int rc = QGT_B200_ERR_NO_MEMORY;
    double best = -1.0;
    for (int c = 2; c <= avail - 1; c++) {
        Program trial; std::string terr;
        const int trc = build_qgt_program_split(plan, total_slots, want_psi, trial, terr, c);
        if (trc != QGT_B200_OK) { if (best < 0.0) { rc = trc; err = terr; } continue; }
        const double t = program_traffic(trial);
        if (best < 0.0 || t < best) { best = t; prog = std::move(trial); rc = QGT_B200_OK; }
    }
    return rc;
}

static int build_qgt_program_split(const CircuitPlan& plan, size_t total_slots, bool want_psi, Program& prog, std::string& err, int c_forced) {
    prog = Program();
    const int P = plan.P;
    const int R = (int)plan.runs.size();
    std::vector<int> ord;
    for (int p = 0; p < P; p++) if (plan.first_run[p] >= 0) ord.push_back(p);
    std::stable_sort(ord.begin(), ord.end(), [&](int a, int b) { return plan.first_run[a] < plan.first_run[b]; });
    const int Pa = (int)ord.size();

    Sched s(plan, prog);
    int b, c;
    if ((size_t)Pa + 1 <= total_slots) { b = Pa; c = 0; }
    else {
        if (total_slots < 5) { err = "workspace too small: need at least 5 statevector-sized columns"; return QGT_B200_ERR_NO_MEMORY; }
        const int avail = (int)total_slots - 2;             // phi + rolling checkpoint
        c = c_forced > 0 ? std::min(c_forced, avail - 1) : std::max(2, std::min(avail / 3, 16));
        b = avail - c;
    }
    const bool blocked = (c > 0);
    int next = 0;
    s.psi = next++;
    if (blocked) s.ckpt = next++;
    for (int i = 0; i < b; i++) s.res_slots.push_back(next++);
    for (int i = 0; i < c; i++) s.str_slots.push_back(next++);
    if (!blocked && Pa > 0 && (size_t)next + 1 <= total_slots) s.psi_alt = next++;
    prog.num_slots = next; prog.psi_slot = s.psi; prog.resident = b; prog.streaming = c;

    if (!blocked) {
        s.init(s.psi);
        prog.blocks = Pa ? 1 : 0;
        s.march(0, ord, {}, true, want_psi);
        prog.psi_slot = s.psi;               // phi may have ended in the spare slot
        prog.psi_final = want_psi;
        return QGT_B200_OK;
    }
    s.init(s.ckpt);
    s.ckpt_time = 0;
    for (int lo = 0; lo < Pa; lo += b) {
        const int hi = std::min(Pa, lo + b);
        std::vector<int> resident(ord.begin() + lo, ord.begin() + hi);
        std::vector<int> pending(ord.begin() + hi, ord.end());
        const int r0 = plan.first_run[resident.front()];
        int T_res = r0;
        for (int p : resident) T_res = std::max(T_res, plan.last_run[p]);
        while (s.ckpt_time < r0) { s.sweep(s.ckpt_time, {{s.ckpt, s.ckpt, -1, false}}); s.ckpt_time++; }
        bool first_pass = true;
        do {
            std::vector<int> take, rest;
            select_fit(plan, pending, c, T_res, take, rest);
            s.copy(s.ckpt, s.psi);
            s.march(r0, resident, take, first_pass, false);
            pending.swap(rest);
            first_pass = false;
        } while (!pending.empty());
        prog.blocks++;
    }
    if (want_psi) {
        while (s.ckpt_time < R) { s.sweep(s.ckpt_time, {{s.ckpt, s.ckpt, -1, false}}); s.ckpt_time++; }
        s.copy(s.ckpt, s.psi);
        prog.psi_final = true;
    }
    return QGT_B200_OK;
}

void stage_generators(const Run& run, const SubPass& sp, const Stage& st, std::vector<double>& out) {
    const int R = (int)sp.reg_local.size();
    const int N = 1 << R;
    const int nv = 1 << (int)st.vqubits.size();
    std::vector<double> M, one, Nsum;
    stage_dense(run, sp, st, -1, M);
    out.assign(st.params.size() * (size_t)nv * N * N * 2, 0.0);
    for (size_t k = 0; k < st.params.size(); k++) {
        Nsum.assign((size_t)nv * N * N * 2, 0.0);
        for (int o : st.ops) {
            if (run.ops[o].param != st.params[k]) continue;
            stage_dense(run, sp, st, o, one);
            for (size_t i = 0; i < Nsum.size(); i++) Nsum[i] += one[i];
        }
        for (int v = 0; v < nv; v++) {
            const double* Mv = &M[(size_t)v * N * N * 2];
            const double* Nv = &Nsum[(size_t)v * N * N * 2];
            double* G = &out[(k * (size_t)nv + v) * N * N * 2];
            for (int a = 0; a < N; a++)
                for (int c = 0; c < N; c++) {           // G[a][c] = sum_j N[a][j] conj(M[c][j])
                    double gr = 0.0, gi = 0.0;
                    for (int j = 0; j < N; j++) {
                        const double nr = Nv[2 * (a * N + j)], ni = Nv[2 * (a * N + j) + 1];
                        const double mr = Mv[2 * (c * N + j)], mi = Mv[2 * (c * N + j) + 1];
                        gr += nr * mr + ni * mi;
                        gi += ni * mr - nr * mi;
                    }
                    G[2 * (a * N + c)] = gr; G[2 * (a * N + c) + 1] = gi;
                }
        }
    }
}

// ---- fused schedule --------------------------------------------------------------------------------------
bool plan_supports_fused(const CircuitPlan& plan) {
    if (plan.R != 3 || plan.B != 0 || plan.K < 8) return false;
    for (const Run& run : plan.runs) {
        if (run.exchange_gbit >= 0) continue;
        for (const SubPass& sp : run.subs) {
            if (sp.is_cost) return false;                       // TODO(cost): the QAOA cost pass has no fused form yet
            if (!sp.mma_ok) return false;
        }
        for (const ParamOcc& oc : run.occ)
            if (locate_op(run, oc.op).kind != 1) return false;
    }
    return true;
}

int build_fused_program(const CircuitPlan& plan, size_t total_slots, bool want_psi, Program& prog, std::string& err, int traj_mode) {
    prog = Program();
    const int P = plan.P;
    const int R = (int)plan.runs.size();
    std::vector<int> ord;
    for (int p = 0; p < P; p++) if (plan.first_run[p] >= 0) ord.push_back(p);
    std::stable_sort(ord.begin(), ord.end(), [&](int a, int b) { return plan.first_run[a] < plan.first_run[b]; });
    const int Pa = (int)ord.size();
    if (Pa == 0) return build_qgt_program(plan, total_slots, want_psi, prog, err);
    int R_last = 0;
    for (int p : ord) R_last = std::max(R_last, plan.last_run[p]);
    // trajectory mode: Tm extra columns hold phi after every transition-matrix stage of the current run; the items then
    // fetch those tiles instead of recomputing phi (a third fewer tensor operations per column pass).  Worth it as long as
    // at least two resident columns remain: per useful column pass and stage ~372 (1 + 1/b) DMMAs against
    // ~552 + 372/b' when phi is recomputed (b' = b + Tm).
    int Tm = 0;
    for (const Run& run : plan.runs) Tm = std::max(Tm, run.rho_stages);
    bool traj = traj_mode != 0 && Tm >= 1 && Tm <= QGT_MAX_TRAJ;
    int ranges = 1;
    if (traj) {
        const bool fits_all = (size_t)Pa + 2 + Tm <= total_slots;
        if (!fits_all) {
            // blocked: every column spent on images is a resident column less (30 qubits: 9 columns fit, 5 transition-matrix
            // stages per run).  The executor then walks a run's launches range by range over the tiles - phi's launch of one
            // eighth of the tiles, then the columns' launches of the same eighth - and the images of all stages of a range
            // share one column.
            int min_tiles_log2 = 64;
            for (const Run& run : plan.runs) if (run.exchange_gbit < 0) min_tiles_log2 = std::min(min_tiles_log2, plan.nloc - run.K);
            // ... when that buys at least a quarter more resident columns: at 28 qubits (39 columns, 4-5 stages) it would be 35
            // instead of 31, the same 8-9 blocks, and eight times as many, shorter launches cost more than they save (8.90 s
            // against 8.79 s per evaluation); at 30 qubits 5 columns instead of none (70.9 s -> 44.0 s)
            const long b_full = (long)total_slots - 3 - Tm, b_ranged = (long)total_slots - 4;
            // (16 ranges when 8 cannot hold two ranges' images at once: the executor overlaps phi's launch of the next range with
            // the columns' launches of the current one)
            if (min_tiles_log2 >= 3 && Tm <= 8 && (b_full < 2 || b_ranged * 4 >= b_full * 5)) ranges = (2 * Tm > 8 && min_tiles_log2 >= 4) ? 16 : 8;
            const long b_traj = ranges > 1 ? b_ranged : b_full;
            if (b_traj < (traj_mode == 1 ? 1 : 2)) traj = false;
        }
    }
    if (!traj) ranges = 1;
    const int extra = traj ? (ranges > 1 ? 1 : Tm) : 0;
    const bool blocked = (size_t)Pa + 2 + extra > total_slots;
    if (blocked && total_slots < (size_t)4 + extra) { err = "workspace too small: need at least 4 statevector-sized columns"; return QGT_B200_ERR_NO_MEMORY; }
    const int b = blocked ? (int)total_slots - 3 - extra : Pa;
    Sched s(plan, prog);
    int next = 0;
    int phi = next++, phi_alt = next++;
    const int ckpt = blocked ? next++ : -1;
    for (int i = 0; i < extra; i++) prog.traj_slots.push_back(next++);
    for (int i = 0; i < b; i++) s.res_slots.push_back(next++);
    prog.num_slots = next; prog.resident = b; prog.streaming = 0; prog.fused = true; prog.traj_ranges = ranges;
    const int nblocks = (Pa + b - 1) / b;
    prog.blocks = nblocks;
    // block order: 1, 2, ... (the rolling checkpoint only moves forward), then block 0 from a fresh initial state - it
    // marches through every run, so its phi doubles as the final state and the self transition matrices are its
    std::vector<int> order;
    for (int i = 1; i < nblocks; i++) order.push_back(i);
    order.push_back(0);
    int ckpt_time = 0;
    if (blocked && nblocks > 1) s.init(ckpt);
    auto fused = [&](int run, int phi_slot, std::vector<SweepCol>& cols) {
        if (cols.empty()) return;
        Instr in; in.kind = INSTR_FUSED; in.run = run; in.phi = phi_slot; in.traj = traj;
        if (traj && cols.size() > 1 && cols[0].self) {
            // trajectory mode: phi alone goes first and publishes its tiles, the columns follow in a second launch
            Instr first = in;
            first.cols.assign(1, cols[0]);
            prog.instrs.push_back(std::move(first));
            in.cols.assign(cols.begin() + 1, cols.end());
        } else {
            in.cols = cols;
        }
        prog.instrs.push_back(std::move(in));
    };
    for (int blk : order) {
        const int lo = blk * b, hi = std::min(Pa, lo + b);
        std::vector<int> slot_of(P, -1);
        std::vector<char> is_res(P, 0), alive(P, 0);
        for (int i = lo; i < hi; i++) { slot_of[ord[i]] = s.res_slots[i - lo]; is_res[ord[i]] = 1; }
        const int r0 = plan.first_run[ord[lo]];
        if (blk == 0) s.init(phi);
        else {
            while (ckpt_time < r0) { s.sweep(ckpt_time, {{ckpt, ckpt, -1, false}}); ckpt_time++; }
            s.copy(ckpt, phi);
        }
        const int r_end = (blk == 0 && want_psi) ? R - 1 : R_last;
        for (int r = r0; r <= r_end; r++) {
            const Run& run = plan.runs[r];
            if (run.exchange_gbit >= 0 || r > R_last) {            // exchange pseudo-run, or phi alone after the last parameter
                std::vector<SweepCol> A;
                A.push_back({phi, phi, -1, false});
                for (int p = 0; p < P; p++) if (alive[p]) A.push_back({slot_of[p], slot_of[p], -1, false});
                s.sweep(r, A);
                continue;
            }
            std::vector<SweepCol> A, acc;
            {
                SweepCol c(phi, phi_alt, -1, false);
                c.id = P; c.self = true;
                c.rho_from = (blk == 0) ? 0 : 1 << 20;         // the self transition matrices are taken once, in block 0
                A.push_back(c);
            }
            for (int p = 0; p < P; p++)
                if (alive[p]) { SweepCol c(slot_of[p], slot_of[p], -1, false); c.id = p; c.rho_from = 0; A.push_back(c); }
            std::vector<char> seen_here(P, 0);
            std::vector<int> born;
            for (const ParamOcc& oc : run.occ) {
                const int p = oc.param;
                if (seen_here[p] || !is_res[p]) continue;
                seen_here[p] = 1;
                std::vector<SweepCol> items = s.occurrence_items(r, p, phi, slot_of[p]);
                for (size_t k = 0; k < items.size(); k++) {
                    items[k].id = p;
                    items[k].rho_from = locate_op(run, items[k].ovr_op).index + 1;
                    if (k == 0 && !alive[p]) { items[k].accumulate = false; A.push_back(items[k]); born.push_back(p); }
                    else acc.push_back(items[k]);
                }
            }
            fused(r, phi, A);
            while (!acc.empty()) {                                 // no two items of a launch may share a destination
                std::vector<SweepCol> now, later;
                std::set<int> dsts;
                for (const SweepCol& c : acc) (dsts.insert(c.dst).second ? now : later).push_back(c);
                fused(r, phi, now);
                acc.swap(later);
            }
            for (int p : born) alive[p] = 1;
            std::swap(phi, phi_alt);
        }
    }
    prog.psi_slot = phi;
    prog.psi_final = want_psi;
    return QGT_B200_OK;
}

// ---- adjoint-gradient programs (run on the plan of the INVERSE circuit, see adjoint.cu) ---------------------------
// Slots: 0 = chi (starts as psi = U|init>), 1 = chi's out-of-place twin, 2 = Lambda (starts as H psi), 3.. = scratch.
void invert_circuit(const qgt_b200_circuit& c, std::vector<qgt_b200_gate>& out) {
    out.clear();
    for (size_t k = c.num_gates; k-- > 0;) {
        qgt_b200_gate g = c.gates[k];
        switch (g.kind) {
        case QGT_B200_GATE_S: g.kind = QGT_B200_GATE_SDG; break;
        case QGT_B200_GATE_SDG: g.kind = QGT_B200_GATE_S; break;
        case QGT_B200_GATE_T: g.kind = QGT_B200_GATE_TDG; break;
        case QGT_B200_GATE_TDG: g.kind = QGT_B200_GATE_T; break;
        case QGT_B200_GATE_SX: out.push_back(g); out.push_back(g); break;          // SX^4 = 1: the inverse is SX^3
        default:
            if (is_parametric(g.kind)) { g.angle = -g.angle; g.scale = -g.scale; }  // angle_eff = scale * theta + angle -> its negative
            break;
        }
        out.push_back(g);
    }
}

int build_gradient_fused_program(const CircuitPlan& plan, Program& prog) {
    prog = Program();
    prog.num_slots = 3; prog.fused = true; prog.resident = 1; prog.blocks = 1;
    const int chi = 0, lam = 2;
    for (int r = 0; r < (int)plan.runs.size(); r++) {
        const Run& run = plan.runs[r];
        Instr in; in.run = r;
        if (run.exchange_gbit >= 0 || run.rho_stages == 0) {   // an exchange of a sharded state, or no parameter in this run: both states just move
            in.kind = INSTR_SWEEP;
            in.cols.push_back(SweepCol(chi, chi, -1, false));
            in.cols.push_back(SweepCol(lam, lam, -1, false));
            prog.instrs.push_back(std::move(in));
            continue;
        }
        // one item per tile: Lambda and chi are staged together, both go through every stage and both are written back
        // (pair mode), the transition matrices of the stages with parameters are taken on the way
        in.kind = INSTR_FUSED; in.phi = chi; in.traj = false;
        SweepCol col(lam, lam, -1, false);
        col.id = 0; col.rho_from = 0;              // row 0 of A collects <Lambda| G_nu |chi> for every nu
        col.phi_dst = chi;
        in.cols.push_back(col);
        prog.instrs.push_back(std::move(in));
    }
    prog.psi_slot = chi;
    return QGT_B200_OK;
}

int build_gradient_run_programs(const CircuitPlan& plan, int r, int scratch_slots, std::vector<Program>& progs) {
    progs.clear();
    const Run& run = plan.runs[r];
    if (scratch_slots < 1) return QGT_B200_ERR_UNSUPPORTED;
    const int P = plan.P, chi = 0, lam = 2;
    if (run.exchange_gbit >= 0) {                  // sharded state: both states go through the exchange
        Program g; g.num_slots = 3 + scratch_slots;
        Sched s(plan, g);
        s.sweep(r, {SweepCol(chi, chi, -1, false), SweepCol(lam, lam, -1, false)});
        progs.push_back(std::move(g));
        return QGT_B200_OK;
    }
    std::vector<int> params;
    {
        std::vector<char> seen(P, 0);
        for (const ParamOcc& oc : run.occ) if (!seen[oc.param]) { seen[oc.param] = 1; params.push_back(oc.param); }
    }
    auto base = [&]() { Program g; g.num_slots = 3 + scratch_slots; return g; };
    if (params.empty()) {
        Program g = base();
        Sched s(plan, g);
        s.sweep(r, {SweepCol(chi, chi, -1, false), SweepCol(lam, lam, -1, false)});
        progs.push_back(std::move(g));
        return QGT_B200_OK;
    }
    for (size_t lo = 0; lo < params.size(); lo += (size_t)scratch_slots) {
        const size_t hi = std::min(params.size(), lo + (size_t)scratch_slots);
        Program g = base();
        Sched s(plan, g);
        std::vector<SweepCol> A, acc;
        if (lo == 0) A.push_back(SweepCol(lam, lam, -1, false));       // Lambda moves to the end of the run first
        std::vector<int> slots, ids;
        for (size_t k = lo; k < hi; k++) {
            const int slot = 3 + (int)(k - lo);
            std::vector<SweepCol> items = s.occurrence_items(r, params[k], chi, slot);
            for (size_t j = 0; j < items.size(); j++) {
                if (j == 0) { items[j].accumulate = false; A.push_back(items[j]); }
                else acc.push_back(items[j]);
            }
            slots.push_back(slot); ids.push_back(params[k]);
        }
        s.sweep(r, A);
        s.emit_accumulates(r, acc);
        s.gram({lam}, {P}, slots, ids);                                  // C[P][nu] = <Lambda_after| (d_nu T) chi_before>
        progs.push_back(std::move(g));
    }
    {
        Program g = base();
        Sched s(plan, g);
        s.sweep(r, {SweepCol(chi, chi, -1, false)});
        progs.push_back(std::move(g));
    }
    return QGT_B200_OK;
}

// ---- JSON dump (tests interpret this on the CPU) ------------------------------------------------
static void jarr(std::ostringstream& o, const std::vector<int>& v) {
    o << "[";
    for (size_t i = 0; i < v.size(); i++) o << (i ? "," : "") << v[i];
    o << "]";
}

static void jop(std::ostringstream& o, const LoweredOp& op, bool deriv) {
    char b[64];
    o << "{\"type\":" << op.type << ",\"target\":" << op.target << ",\"cmask\":" << op.cmask
      << ",\"pmask\":" << op.pmask << ",\"flags\":" << (deriv ? op.dflags : op.flags) << ",\"gate\":" << op.gate
      << ",\"param\":" << op.param << ",\"m\":[";
    const double* m = deriv ? op.dm : op.m;
    for (int i = 0; i < 8; i++) { snprintf(b, sizeof b, "%.17g", m[i]); o << (i ? "," : "") << b; }
    o << "]}";
}

std::string dump_json(const qgt_b200_circuit& c, const CircuitPlan& plan, const Program* prog,
                      const std::vector<MappedSegment>* segs) {
    std::ostringstream o;
    o << "{\"n\":" << plan.n << ",\"nloc\":" << plan.nloc << ",\"P\":" << plan.P << ",\"K\":" << (plan.runs.empty() ? 0 : plan.runs[0].K)
      << ",\"initial_state\":" << c.initial_state << ",\"runs\":[";
    for (size_t r = 0; r < plan.runs.size(); r++) {
        const Run& run = plan.runs[r];
        o << (r ? "," : "") << "{\"exchange\":" << run.exchange_gbit << ",\"exchange_mask\":" << run.exchange_mask << ",\"segment\":" << run.segment << ",\"tile\":"; jarr(o, run.tile_qubits);
        o << ",\"subs\":[";
        for (size_t s = 0; s < run.subs.size(); s++) {
            const SubPass& sp = run.subs[s];
            o << (s ? "," : "") << "{\"reg\":"; jarr(o, sp.reg_local);
            o << ",\"batch\":"; jarr(o, sp.batch_local);
            o << ",\"tperm\":"; jarr(o, sp.tperm);
            o << ",\"forms\":[";             // QGT_FORM_* of each dense stage (1 = complex diagonal x real matrix)
            { std::vector<double> tmp; for (size_t k = 0; k < sp.stages.size(); k++) o << (k ? "," : "") << stage_matrices(run, sp, sp.stages[k], -1, tmp); }
            o << "],\"stages\":[";
            for (size_t k = 0; k < sp.stages.size(); k++) {
                o << (k ? "," : "") << "{\"ops\":"; jarr(o, sp.stages[k].ops);
                o << ",\"params\":"; jarr(o, sp.stages[k].params);
                o << ",\"rho_off\":" << sp.stages[k].rho_off << ",\"nvar\":" << sp.stages[k].vqubits.size() << "}";
            }
            o << "],\"tdiags\":"; jarr(o, sp.tdiags);
            o << ",\"ops\":[" << sp.op_begin << "," << sp.op_end << "]}";
        }
        o << "],\"ops\":[";
        for (size_t i = 0; i < run.ops.size(); i++) { o << (i ? "," : ""); jop(o, run.ops[i], false); }
        o << "],\"dops\":{";
        bool firstd = true;
        for (const ParamOcc& oc : run.occ) { o << (firstd ? "" : ",") << "\"" << oc.op << "\":"; jop(o, run.ops[oc.op], true); firstd = false; }
        o << "}}";
    }
    o << "]";
    if (segs) {
        o << ",\"segments\":[";
        for (size_t i = 0; i < segs->size(); i++) {
            o << (i ? "," : "") << "{\"phys\":"; jarr(o, (*segs)[i].phys_of_logical);
            o << ",\"exchange\":" << (*segs)[i].exchange_gbit << ",\"exchange_mask\":" << (*segs)[i].exchange_mask << ",\"gates\":" << (*segs)[i].gates.size() << "}";
        }
        o << "]";
    }
    if (prog) {
        o << ",\"program\":{\"slots\":" << prog->num_slots << ",\"psi\":" << prog->psi_slot << ",\"resident\":" << prog->resident
          << ",\"streaming\":" << prog->streaming << ",\"blocks\":" << prog->blocks << ",\"psi_final\":" << (prog->psi_final ? 1 : 0)
          << ",\"fused\":" << (prog->fused ? 1 : 0) << ",\"traj_ranges\":" << prog->traj_ranges << ",\"traj\":";
        jarr(o, prog->traj_slots);
        o << ",\"instrs\":[";
        for (size_t i = 0; i < prog->instrs.size(); i++) {
            const Instr& in = prog->instrs[i];
            o << (i ? "," : "");
            if (in.kind == INSTR_SWEEP || in.kind == INSTR_FUSED) {
                o << "{\"k\":\"" << (in.kind == INSTR_FUSED ? "fused" : "sweep") << "\",\"run\":" << in.run << ",\"phi\":" << in.phi << ",\"cols\":[";
                for (size_t j = 0; j < in.cols.size(); j++)
                {
                    o << (j ? "," : "") << "[" << in.cols[j].src << "," << in.cols[j].dst << "," << in.cols[j].ovr_op << "," << (in.cols[j].accumulate ? 1 : 0) << ",";
                    jarr(o, in.cols[j].ovr_extra);
                    if (in.kind == INSTR_FUSED) o << "," << in.cols[j].id << "," << in.cols[j].rho_from << "," << (in.cols[j].self ? 1 : 0) << "," << in.cols[j].phi_dst;
                    o << "]";
                }
                o << "]}";
            } else if (in.kind == INSTR_GRAM) {
                o << "{\"k\":\"gram\",\"a\":"; jarr(o, in.a_slots); o << ",\"aid\":"; jarr(o, in.a_ids);
                o << ",\"b\":"; jarr(o, in.b_slots); o << ",\"bid\":"; jarr(o, in.b_ids); o << "}";
            } else if (in.kind == INSTR_COPY) {
                o << "{\"k\":\"copy\",\"src\":" << in.src << ",\"dst\":" << in.dst << "}";
            } else {
                o << "{\"k\":\"init\",\"dst\":" << in.dst << "}";
            }
        }
        o << "]}";
    }
    o << "}";
    return o.str();
}

}  // namespace qgt

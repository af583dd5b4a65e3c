// plan.hpp — host-side planner: circuit -> fused runs -> column schedule ("program").
// Pure C++17, no CUDA: compiled into libqgt_b200.so and (for CPU tests) exercised through
// qgt_b200_plan_dump().
#pragma once
#include <cstdint>
#include <string>
#include <vector>

#include "../../include/qgt_b200.h"
#include "dev_structs.h"

namespace qgt {

struct PlanOptions {
    int tile_qubits = 11;   // K
    int reg_qubits = 3;     // R: qubits of a dense stage matrix
    int batch_qubits = 0;   // B: extra amplitudes per thread (2^B halves share every matrix element)
    int low_qubits = 4;     // L: lowest qubits always in the tile (contiguous 16*2^L bytes)
    int max_ops_per_run = 160;
    int birth_cut = 1;      // execute a run in which many parameters are born in pieces (split_run_by_births); 2 = ignore the launch cost
    int wavefront = 2;      // unsharded plans: lower the gates in the order the sharded mapper would emit them with this many
                            // virtual rank qubits (gates that wait for them are deferred, the rest runs ahead): a dependency-
                            // respecting reordering that lets the run builder work deep before it works wide
    int parity_form = 1;    // plans with a cost pass: pack X-rotation stage matrices in QGT_FORM_PARITY (4 DMMAs per 8 vectors instead of 8)
    int local_qubits = 0;   // sharded states: qubits >= local_qubits are rank bits (diagonal use / controls only); 0 = all local
};

// an op in global-qubit terms, before it is bound to a sub-pass
struct LoweredOp {
    int type = QGT_OP_U;
    int target = -1;            // non-diagonal target qubit (U-types), -1 for DIAG/COST
    uint64_t cmask = 0;         // control qubits (global mask)
    uint64_t pmask = 0;         // DIAG parity qubits (global mask)
    uint32_t flags = 0;
    double m[8] = {0};
    int gate = -1;              // index of the source gate
    int param = -1;             // parameter the source gate carries, or -1
    // derivative op (valid when param >= 0): same type/target/masks, different payload
    double dm[8] = {0};
    uint32_t dflags = 0;
};

// a dense stage of a sub-pass: the lowered ops (in order) whose product forms the 2^R x 2^R matrix, and the
// non-register qubits that select among its variants
struct Stage {
    std::vector<int> ops;
    std::vector<int> vqubits;       // global qubit numbers, variant bit k <-> vqubits[k]
    std::vector<int> params;        // distinct parameters with an occurrence among `ops`, in order of appearance
    int rho_off = -1;               // fused schedule: first of the stage's 2^nvar transition-matrix blocks inside an
                                    // item's rho buffer, -1 when no parameter occurs in the stage
    int traj_ord = -1;              // ordinal among the run's stages with parameters (which trajectory column holds phi after it)
};

struct SubPass {
    std::vector<int> reg_local;     // local bit positions in registers (ascending, padded to R); empty for a cost pass
    int nreg_used = 0;
    std::vector<int> batch_local;   // local bit positions of the batch qubits (B entries)
    std::vector<int> tperm;         // thread bit -> local position
    int op_begin = 0, op_end = 0;   // lowered ops
    bool is_cost = false;           // tile-level pass holding exactly one COST op
    bool mma_ok = false;            // variant-selecting tile qubits all sit on warp-index thread bits
    std::vector<Stage> stages;
    std::vector<int> tdiags;        // lowered diagonal ops that touch no register qubit
};

struct ParamOcc { int param; int op; };   // op = index into Run::ops

struct Run {
    int K = 0;
    std::vector<int> tile_qubits;   // ascending global positions, size K
    std::vector<int> other_qubits;  // ascending, size n-K
    std::vector<LoweredOp> ops;     // in execution order
    std::vector<SubPass> subs;
    std::vector<ParamOcc> occ;      // parameterised ops in this run
    int exchange_gbit = -1;         // pseudo-run of a sharded state: >= 0 marks an exchange (lowest rank bit of exchange_mask)
    unsigned exchange_mask = 0;     // rank bits b_0 < b_1 < ... swapped, in ONE grouped all-to-all, with the top local qubits
                                    // nloc-k, ..., nloc-1 (b_i <-> nloc-k+i); k = popcount
    int segment = 0;                // index of the mapped segment the run belongs to (selects the cost table)
    bool parity_form = false;       // 8x8 stage matrices with the X-rotation structure are packed in QGT_FORM_PARITY (plans with a cost
                                    // pass: they run in the sweep kernel, the only one that knows the form)
    int rho_blocks = 0;             // fused schedule: 64-element transition-matrix blocks per item (sum of 2^nvar over stages with parameters)
    int last_rho_stage = -1;        // run-relative index of the last stage with a parameter occurrence
    int rho_stages = 0;             // number of stages with a parameter occurrence
};

struct CircuitPlan {
    int n = 0, P = 0;
    int nloc = 0;                   // qubits addressed inside one shard (== n on a single GPU)
    int K = 0, R = 0, B = 0;        // tile / matrix / batch qubits actually used (clamped to n)
    PlanOptions opt;
    std::vector<Run> runs;
    std::vector<int> first_run, last_run;   // per parameter, -1 when the parameter has no gate
};

// device-format image of a plan (what gets uploaded)
struct PlanImage {
    std::vector<QgtDevRun> runs;
    std::vector<QgtDevSubPass> subs;
    std::vector<QgtDevStage> stages;
    std::vector<QgtDevThrDiag> tdiags;
    std::vector<QgtDevCost> costs;
    std::vector<double> pool;            // matrices, interleaved (re, im)
};

// where a lowered op of a run ended up on the device
struct OpLocation {
    int kind = 0;       // 1 dense stage, 2 thread diagonal, 3 cost   (QgtSweepItem::ovr_kind)
    int sub = -1;
    int index = -1;     // stage / tdiag / cost index relative to the run's first one
};

int lower_gate(const qgt_b200_circuit& c, const double* theta, int gate_index, std::vector<LoweredOp>& out, std::string& err);
int build_plan(const qgt_b200_circuit& c, const double* theta, const PlanOptions& opt, CircuitPlan& plan, std::string& err);
void build_image(const CircuitPlan& plan, PlanImage& img);
OpLocation locate_op(const Run& run, int op_index);
// all variants of a stage's matrix, (re, im) interleaved, row-major 2^R x 2^R each; deriv_op >= 0 replaces
// that lowered op by its derivative
// in the device layout of dev_structs.h; returns the QGT_FORM_* chosen
int stage_matrices(const Run& run, const SubPass& sp, const Stage& st, int deriv_op, std::vector<double>& out);
// product rule over several ops of one stage: sum of the single-derivative matrices
int stage_matrices_sum(const Run& run, const SubPass& sp, const Stage& st, const std::vector<int>& deriv_ops, std::vector<double>& out);
QgtDevThrDiag make_tdiag(const LoweredOp& op, bool derivative);
QgtDevCost make_cost(const LoweredOp& op, bool derivative);

// ---- column schedule ---------------------------------------------------------------------------
enum InstrKind { INSTR_SWEEP = 0, INSTR_GRAM = 1, INSTR_COPY = 2, INSTR_INIT = 3, INSTR_FUSED = 4 };

struct SweepCol {
    SweepCol() {}
    SweepCol(int s, int d, int o, bool a) : src(s), dst(d), ovr_op(o), accumulate(a) {}
    int src = 0, dst = 0;       // slots
    int ovr_op = -1;            // index into the run's ops, -1 = none
    std::vector<int> ovr_extra; // further ops of the same parameter inside the same dense stage: the item applies
                                // the product-rule sum  sum_j (stage with op j replaced by its derivative)
    bool accumulate = false;
    // fused schedule (INSTR_FUSED)
    int id = -1;                // parameter whose column this is; P = the marching state phi itself
    int rho_from = 0;           // first stage (run-relative) whose transition matrix <column| . |phi> is contracted:
                                // 0 for a column that merely advances, (stage of the spawning occurrence) + 1 for a spawn
    bool self = false;          // the item IS phi: one tile, rho = <phi| . |phi>
    int phi_dst = -1;           // pair mode: slot the advanced phi is written to by this item (no separate phi item), -1 = none
};

struct Instr {
    int kind = INSTR_SWEEP;
    int run = -1;
    int phi = -1;                           // FUSED: slot of phi at the start of the run (second tile of every item)
    bool traj = false;                      // FUSED: phi's tiles come from / go to the program's trajectory columns
    std::vector<SweepCol> cols;             // SWEEP, FUSED
    std::vector<int> a_slots, a_ids;        // GRAM: <a|b> for every pair; id = parameter index, P = psi
    std::vector<int> b_slots, b_ids;
    int src = -1, dst = -1;                 // COPY: dst <- src ; INIT: dst <- initial state
};

struct Program {
    int num_slots = 0;
    int psi_slot = 0;           // slot holding phi while marching
    int resident = 0;           // b
    int streaming = 0;          // c
    int blocks = 0;
    bool psi_final = false;     // psi_slot holds U(theta)|init> at the end
    bool fused = false;         // built by build_fused_program: no Gram instructions, Q assembled from transition matrices
    std::vector<int> traj_slots;// fused schedule, trajectory mode: columns holding phi after every transition-matrix stage of
                                // the current run (tile images); empty = phi is recomputed next to every column
    int traj_ranges = 1;        // > 1: the executor walks every run's trajectory launches range by range over the tiles
                                // (phi's launch of a range, then the columns' launches of the same range), so the images of
                                // all transition-matrix stages of a run share ONE column: traj_slots has one entry and image t
                                // of a range lives at column + t * D / traj_ranges
    std::vector<Instr> instrs;
};

// total_slots = number of 2^n-amplitude columns that fit in the workspace
int build_qgt_program(const CircuitPlan& plan, size_t total_slots, bool want_psi, Program& prog, std::string& err);

// Fused schedule ("apply-gate-and-contract", SURVEY.md section 7 hard-part 1c).  For mu born before nu,
//   <d_mu psi|d_nu psi> = <lambda_mu(t)| G_nu |phi(t)>   at the time t of nu's gate,
// lambda_mu = the derivative column propagated to t, G_nu = (dU_nu) U_nu^+.  The kernel advances a column and phi
// through a run in lockstep (two tiles on chip) and, after every dense stage in which a parameter occurs, accumulates
// the 8x8 transition matrix rho[c][a] = sum_rest phi'[c, rest] conj(lambda'[a, rest]) over the stage's 3 matrix
// qubits; <lambda|G|phi> = sum_{a,c} Gt[a][c] rho[c][a] with Gt = (derivative stage matrix) x (stage matrix)^+.
// No streaming column is ever written and no Gram pass re-reads the resident block: every column costs one pass per
// run it is alive in, whatever the number of column slots.  Pairs inside one stage, the diagonal and the projections
// <d_mu psi|psi> come from rho of phi with itself.  b = total_slots - 3 resident columns per block (phi, its
// out-of-place twin and a rolling checkpoint take three slots).
bool plan_supports_fused(const CircuitPlan& plan);
// traj_mode: 0 never, 1 whenever it fits, -1 automatic (when at least two resident columns remain)
int build_fused_program(const CircuitPlan& plan, size_t total_slots, bool want_psi, Program& prog, std::string& err, int traj_mode = -1);
// index of element (c, a, re/im) inside a 128-double transition-matrix block: the order in which the DMMA C fragments
// of a warp hold it (lane = c*4 + a/2 owns columns a, a^1)
static inline int rho_index(int c, int a, int part) { return (c * 4 + (a >> 1)) * 4 + part * 2 + (a & 1); }
// Evolved generators of one stage: for every parameter of st.params (same order) and every variant, the matrix
// Gt = N M^+ (N = sum of the stage matrices with one occurrence replaced by its derivative), row-major 8x8 complex
// (re, im); out[(k * nvariants + v) * 128 + (a * 8 + c) * 2]
void stage_generators(const Run& run, const SubPass& sp, const Stage& st, std::vector<double>& out);

// ---- adjoint energy gradient ---------------------------------------------------------------------------
// E(theta) = <psi|H|psi>, H diagonal.  With the inverse circuit C' = U^+ run on chi_0 = psi and Lambda_0 = H psi,
//   dE/dtheta_mu = -2 Re <Lambda_j| (d_mu S_j) S_j^+ |chi_j>   summed over the stages S_j of C' that carry theta_mu
// (chi_j, Lambda_j = the two states after stage j): the same evolved-generator x transition-matrix contraction as the
// fused QGT schedule, with Lambda as the only column.  One forward circuit + one backward pass of two states.
void invert_circuit(const qgt_b200_circuit& c, std::vector<qgt_b200_gate>& out);        // gates of U^+, same parameter indices
// slots 0 = chi, 1 = chi's twin, 2 = Lambda; row 0 of the A matrix receives <Lambda|G_nu|chi>
int build_gradient_fused_program(const CircuitPlan& inverse_plan, Program& prog);
// plans the fused path cannot take (cost layers, tiny tiles): per run, derivative columns (d_nu T) chi are spawned into
// `scratch_slots` columns (slots 3..) and contracted with the advanced Lambda by a Gram, C[P][nu]; the programs of one
// run must be executed in order and C read after each
int build_gradient_run_programs(const CircuitPlan& inverse_plan, int run, int scratch_slots, std::vector<Program>& progs);

// ---- sharded states: logical -> physical qubit mapping ------------------------------------------------
// A state of n qubits sharded over 2^g ranks keeps physical qubits nloc.. (nloc = n - g) in the rank index.
// Gates that act non-diagonally on such a qubit need it local: the mapper walks the gate list, and when a
// gate needs a currently-global logical qubit it evicts the local qubit whose next non-diagonal use is
// farthest away (moved to physical position nloc-1 by a local SWAP first), emitting an EXCHANGE of that rank
// bit with physical qubit nloc-1.  Gates are re-expressed on physical qubits; a segment ends at each exchange.
struct MappedSegment {
    std::vector<qgt_b200_gate> gates;    // physical qubit numbers; targets of non-diagonal gates are < nloc
    std::vector<int> phys_of_logical;    // mapping in force for these gates (cost tables are remapped with it)
    int exchange_gbit = -1;              // after the gates: an exchange follows (lowest bit of exchange_mask); -1 = none
    unsigned exchange_mask = 0;          // rank bits swapped with the top popcount(mask) local qubits (see Run::exchange_mask)
};
int map_circuit_sharded(const qgt_b200_circuit& c, int nloc, bool restore_identity,
                        std::vector<MappedSegment>& segs, std::string& err, std::vector<int>* order = nullptr);
// plan of a mapped circuit: runs of every segment in order, with exchange pseudo-runs in between
int build_plan_sharded(const qgt_b200_circuit& c, const double* theta, const PlanOptions& opt, int nloc, bool restore_identity,
                       CircuitPlan& plan, std::vector<MappedSegment>& segs, std::string& err);

std::string dump_json(const qgt_b200_circuit& c, const CircuitPlan& plan, const Program* prog,
                      const std::vector<MappedSegment>* segs = nullptr);

}  // namespace qgt

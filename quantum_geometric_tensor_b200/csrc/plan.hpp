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
    int reg_qubits = 3;     // R
    int low_qubits = 4;     // L: lowest qubits always in the tile (contiguous 16*2^L bytes)
    int max_ops_per_run = 160;
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

struct SubPass {
    std::vector<int> reg_local;     // local bit positions in registers (ascending, padded to R)
    int nreg_used = 0;
    std::vector<int> tperm;         // thread bit -> local position
    int op_begin = 0, op_end = 0;
};

struct ParamOcc { int param; int op; };   // op = index into Run::ops

struct Run {
    int K = 0;
    std::vector<int> tile_qubits;   // ascending global positions, size K
    std::vector<int> other_qubits;  // ascending, size n-K
    std::vector<LoweredOp> ops;     // in execution order
    std::vector<SubPass> subs;
    std::vector<ParamOcc> occ;      // parameterised ops in this run
};

struct CircuitPlan {
    int n = 0, P = 0;
    PlanOptions opt;
    std::vector<Run> runs;
    std::vector<int> first_run, last_run;   // per parameter, -1 when the parameter has no gate
};

// device-format image of a plan (what gets uploaded)
struct PlanImage {
    std::vector<QgtDevRun> runs;
    std::vector<QgtDevOp> ops;
    std::vector<QgtDevSubPass> subs;
};

int lower_gate(const qgt_b200_circuit& c, const double* theta, int gate_index, std::vector<LoweredOp>& out, std::string& err);
int build_plan(const qgt_b200_circuit& c, const double* theta, const PlanOptions& opt, CircuitPlan& plan, std::string& err);
void build_image(const CircuitPlan& plan, PlanImage& img);
QgtDevOp bind_op(const Run& run, const SubPass& sp, const LoweredOp& op, bool derivative);
int find_subpass(const Run& run, int op_index);

// ---- column schedule ---------------------------------------------------------------------------
enum InstrKind { INSTR_SWEEP = 0, INSTR_GRAM = 1, INSTR_COPY = 2, INSTR_INIT = 3 };

struct SweepCol {
    int src = 0, dst = 0;       // slots
    int ovr_op = -1;            // index into the run's ops, -1 = none
    bool accumulate = false;
};

struct Instr {
    int kind = INSTR_SWEEP;
    int run = -1;
    std::vector<SweepCol> cols;             // SWEEP
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
    std::vector<Instr> instrs;
};

// total_slots = number of 2^n-amplitude columns that fit in the workspace
int build_qgt_program(const CircuitPlan& plan, size_t total_slots, bool want_psi, Program& prog, std::string& err);

std::string dump_json(const qgt_b200_circuit& c, const CircuitPlan& plan, const Program* prog);

}  // namespace qgt

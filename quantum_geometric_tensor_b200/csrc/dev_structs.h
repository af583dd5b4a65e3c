// dev_structs.h — plain structs shared by the host planner (plan.cpp) and the CUDA kernels.
//
// A circuit is cut into RUNS.  One run = one pass over the statevector ("sweep"): a CTA stages a
// TILE of 2^K amplitudes (the K "tile qubits", always including the lowest L qubits so global
// accesses stay contiguous) in shared memory and applies every op of the run to it.  Inside a run
// ops are grouped into SUB-PASSES: during a sub-pass each thread holds 2^R amplitudes in registers
// (the R "register qubits" of that sub-pass) and applies all ops whose non-diagonal target is one
// of them without touching shared memory.
#ifndef QGT_DEV_STRUCTS_H
#define QGT_DEV_STRUCTS_H

#include <stdint.h>

#define QGT_MAX_TILE_QUBITS 12
#define QGT_MAX_QUBITS 40
#define QGT_MAX_REG_QUBITS 4
#define QGT_MAX_TRAJ 8           // transition-matrix stages per run the trajectory mode of the fused kernel supports

// lowered-op types (host side; the device only sees dense stages, thread diagonals and the cost pass)
enum QgtOpType {
    QGT_OP_U = 0,        // dense complex 2x2
    QGT_OP_UREAL = 1,    // 2x2 with real entries (RY, H and their derivatives)
    QGT_OP_URX = 2,      // real diagonal, imaginary off-diagonal (RX and its derivative)
    QGT_OP_PERM = 3,     // exchange the pair (X, CNOT)
    QGT_OP_DIAG = 4,     // amp *= parity(g & pmask) ? d1 : d0   when (g & cmask) == cmask
    QGT_OP_COST = 5      // amp *= exp(-i a E(g)),  E from the edge list (QAOA cost layer); a tile-level pass
};

enum QgtOpFlags {
    QGT_FLAG_ZERO_CTRL_FAIL = 1,  // derivative of a controlled rotation: amplitudes with control=0 become 0
    QGT_FLAG_COST_DERIV = 2       // multiply additionally by (-i * m[1] * E(g))
};

#define QGT_MAX_VARIANT_BITS 2
// matrix variants are laid out QGT_VARIANT_STRIDE(N) complex elements apart: the odd padding keeps two
// variants read by lanes of one quarter-warp in different shared-memory bank groups
#define QGT_VARIANT_STRIDE(N) ((N) * (N) + (N) + 1)
// Matrix forms.  QGT_FORM_DENSE: N*N complex elements.  QGT_FORM_DIAG_REAL: the matrix factors as
// (complex diagonal D) x (real matrix Rm) - true for any stage whose qubits see rotations about Y / CNOT /
// H before their diagonal gates (RZ, CZ, phases), i.e. for every layer of a hardware-efficient ansatz.
// Rm is stored as N*N packed doubles (element (i, j) at double index QGT_MIDX(N, i, j)) at the start of the
// variant, D in the N complex elements at index N*N; the tensor-pipe path then needs 4 DMMAs per 8 vectors
// instead of 8 and one complex multiply per result.
#define QGT_FORM_DENSE 0
#define QGT_FORM_DIAG_REAL 1
// QGT_FORM_PARITY (8x8, plans with a cost pass = sweep kernel only): the real part lives on the elements with an even
// number of differing index bits, the imaginary part on the odd ones - any product of X rotations on the stage's qubits
// (the QAOA mixer).  In the basis ordered (parity, bit 1, bit 0) the matrix is [[A_ee, i B_eo], [i B_oe, A_oo]] with
// real 4x4 blocks, and with the inputs split into even / odd components v_e, v_o
//     X = [A_ee; B_oe] v_e.re + [-B_eo; A_oo] v_o.im = [y_e.re; y_o.im],   Y = [A_ee; -B_oe] v_e.im + [B_eo; A_oo] v_o.re = [y_e.im; y_o.re]
// takes 4 DMMAs per 8 vectors and no other arithmetic.  Slot QGT_MIDX(8, q, k) of the variant holds (x, y) = the two
// stacked matrices' element (row q, even column k); slot QGT_MIDX(8, q, 4 + k) those of the odd column k (q, k in the
// parity basis: index (p, b1, b0) is component ((p ^ b1 ^ b0) << 2) | (b1 << 1) | b0).
#define QGT_FORM_PARITY 2
// element (i, j) of a stage matrix inside its variant.  8x8 matrices are stored in DMMA A-fragment order
// (lane (r, k) reads M[r][k] and M[r][4+k]: 32 consecutive elements per load, conflict-free); smaller
// ones row-major.
#define QGT_MIDX(N, i, j) ((N) == 8 ? ((((j) >> 2) << 5) + ((i) << 2) + ((j) & 3)) : ((i) * (N) + (j)))

// One dense stage of a sub-pass: all gates that touch the sub-pass's R register qubits, multiplied together
// on the host into a 2^R x 2^R complex matrix.  Gates controlled by (or diagonal on) up to
// QGT_MAX_VARIANT_BITS non-register qubits give 2^nvar matrix variants; a thread picks its variant from
// the bits of its global index.
typedef struct QgtDevStage {
    int16_t  nvar;                           // number of selecting bits
    int16_t  form;                           // QGT_FORM_* of every variant of this stage
    int32_t  mat_off;                        // offset of variant 0 in the run's matrix pool, in complex elements
    uint64_t vmask[QGT_MAX_VARIANT_BITS];    // global index bit selecting variant bit k
    int32_t  rho_off;                        // fused kernel: first transition-matrix block of this stage in an item's rho buffer (one
                                             // block of 64 complex per variant), -1 when no parameter occurs in the stage
    int32_t  traj_ord;                       // ordinal of the stage among the run's transition-matrix stages (trajectory column), -1 = none
} QgtDevStage;                               // 32 bytes

// a diagonal gate (or the derivative of one) that touches no register qubit: one phase per thread
typedef struct QgtDevThrDiag {
    uint64_t cmask;                          // controls (global index bits), all must be 1
    uint64_t pmask;                          // parity bits (global index bits)
    double   d[4];                           // d0, d1 (re, im)
    uint32_t flags;
    uint32_t pad;
} QgtDevThrDiag;                             // 56 bytes

// the cost layer (tile-level pass)
typedef struct QgtDevCost {
    double   angle;
    double   dscale;
    uint32_t flags;
    uint32_t pad;
} QgtDevCost;

// A sub-pass, with everything index-related precomputed on the host.  The shared-memory swizzle is linear
// over XOR, so the swizzled slot of an amplitude is the XOR of per-bit contributions: s_thr[i] for thread
// bit i, s_reg[r] for register bit r (matrix qubits first, then batch qubits); g_* are the matching global
// amplitude-index bits.
typedef struct QgtDevSubPass {
    int32_t  nreg;                           // matrix qubits R; 0 marks a tile-level pass holding one COST op
    int32_t  stage_begin, stage_end;         // dense stages (indices into the run's stage array)
    int32_t  tdiag_begin, tdiag_end;         // thread diagonals (indices into the run's thread-diagonal array)
    int32_t  cost;                           // index into the run's cost array when nreg == 0
    int32_t  mma_ok;                         // every stage's variant is uniform over a warp: tensor-pipe path applies
    int32_t  pad;
    uint32_t s_thr[QGT_MAX_TILE_QUBITS];
    uint32_t s_reg[QGT_MAX_REG_QUBITS];
    uint64_t g_thr[QGT_MAX_TILE_QUBITS];
    uint64_t g_reg[QGT_MAX_REG_QUBITS];
} QgtDevSubPass;                             // 224 bytes

typedef struct QgtDevRun {
    int32_t K;                           // tile qubits (== num_qubits when the state is smaller than a tile)
    int32_t n;                           // total qubits
    int32_t nsub;
    int32_t sub_off;                     // offsets into the plan-wide arrays
    int32_t stage_off;
    int32_t tdiag_off;
    int32_t cost_off;
    int32_t mat_off;                     // first matrix element of this run in the matrix pool
    int32_t mat_count;                   // complex elements of this run in the pool
    int8_t  tq[QGT_MAX_TILE_QUBITS];     // tile qubits: local bit j <-> global bit tq[j], ascending
    int8_t  ntq[QGT_MAX_QUBITS];         // the n-K other qubits, ascending (tile id bits are deposited here)
    int32_t has_cost;                    // the run contains a cost-layer pass (energy tables are staged in shared memory)
    int32_t rho_blocks;                  // fused kernel: transition-matrix blocks per item
    int32_t last_rho_stage;              // fused kernel: last stage (run-relative) with a parameter occurrence, -1 = none
    int32_t pad2;
} QgtDevRun;

// one column of a batched sweep launch
typedef struct QgtSweepItem {
    const void* src;                     // complex double [2^n]
    void*       dst;                     // may equal src (in place)
    uint32_t    accumulate;              // dst += result instead of dst = result
    int32_t     ovr_kind;                // 0 none, 1 dense stage, 2 thread diagonal, 3 cost
    int32_t     ovr_index;               // which stage / thread diagonal / cost entry of the run is replaced
    int32_t     ovr_form;                // kind 1: QGT_FORM_* of the replacement matrices
    const void* ovr_mat;                 // kind 1: the replacement matrices (all variants) in global memory
    QgtDevThrDiag ovr_tdiag;             // kind 2
    QgtDevCost    ovr_cost;              // kind 3
    // fused kernel only
    int32_t     self;                    // the item is phi itself: one tile, rho = <phi| . |phi>
    int32_t     rho_from;                // first stage (run-relative) whose transition matrix is accumulated
    void*       phi_dst;                 // pair mode (adjoint gradient): phi's tile goes through EVERY stage and is stored here
                                         // (may equal the launch's phi: in place); null = phi is only carried for the matrices
} QgtSweepItem;

typedef struct QgtDevEdge { int32_t i, j; double w; } QgtDevEdge;

// fused schedule: A[out] += sum over entries [begin, end) of  sum_{variant blocks} sum_e X[e] * rho[e]
typedef struct QgtContractEntry {
    int32_t item;        // item of the launch whose transition matrices are contracted
    int32_t rho_off;     // first block of the stage inside the item's rho buffer
    int32_t nblocks;     // variants of the stage
    int32_t x_off;       // first block of the evolved generator in the X pool (same element order as rho)
} QgtContractEntry;
typedef struct QgtContractGroup { int32_t begin, end, out, pad; } QgtContractGroup;

#endif

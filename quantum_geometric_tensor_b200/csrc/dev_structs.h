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

enum QgtOpType {
    QGT_OP_U = 0,      // dense complex 2x2 on a register qubit
    QGT_OP_UREAL = 1,  // 2x2 with real entries (RY, H and their derivatives)
    QGT_OP_URX = 2,    // real diagonal, imaginary off-diagonal (RX and its derivative)
    QGT_OP_PERM = 3,   // exchange the pair (X, CNOT)
    QGT_OP_DIAG = 4,   // amp *= parity(g & pmask) ? d1 : d0   when (g & cmask) == cmask
    QGT_OP_COST = 5    // amp *= exp(-i a E(g)),  E from the edge list (QAOA cost layer)
};

enum QgtOpFlags {
    QGT_FLAG_ZERO_CTRL_FAIL = 1,  // derivative of a controlled rotation: amplitudes with control=0 become 0
    QGT_FLAG_COST_DERIV = 2       // multiply additionally by (-i * m[1] * E(g))
};

typedef struct QgtDevOp {
    int32_t  type;
    int32_t  tbit;     // U-types: index (0..R-1) of the target inside the sub-pass's register qubits
    uint32_t creg;     // controls that are register qubits of the sub-pass (mask over register index)
    uint32_t preg;     // DIAG: parity bits that are register qubits (mask over register index)
    uint64_t cmask;    // remaining controls, as a mask over GLOBAL amplitude-index bits
    uint64_t pmask;    // DIAG: remaining parity bits, global mask
    uint32_t flags;
    uint32_t pad;
    double   m[8];     // U: m00,m01,m10,m11 (re,im);  DIAG: d0,d1 (re,im);  COST: m[0]=angle, m[1]=derivative scale
} QgtDevOp;            // 112 bytes

typedef struct QgtDevSubPass {
    int32_t nreg;                        // register qubits used (<= R); unused ones are padded with free tile positions
    int32_t op_begin, op_end;            // range in the run's op array
    int32_t pad;
    int8_t  regq[QGT_MAX_REG_QUBITS];    // LOCAL bit positions (0..K-1) held in registers, ascending
    int8_t  tperm[QGT_MAX_TILE_QUBITS];  // thread bit i -> LOCAL bit position
} QgtDevSubPass;                         // 32 bytes

typedef struct QgtDevRun {
    int32_t K;                           // tile qubits (== num_qubits when the state is smaller than a tile)
    int32_t n;                           // total qubits
    int32_t nsub;
    int32_t nops;
    int32_t ops_off;                     // offsets into the global op / sub-pass arrays
    int32_t sub_off;
    int8_t  tq[QGT_MAX_TILE_QUBITS];     // tile qubits: local bit j <-> global bit tq[j], ascending
    int8_t  ntq[QGT_MAX_QUBITS];         // the n-K other qubits, ascending (tile id bits are deposited here)
    int32_t pad;
} QgtDevRun;

// one column of a batched sweep launch
typedef struct QgtSweepItem {
    const void* src;                     // complex double [2^n]
    void*       dst;                     // may equal src (in place)
    int32_t     ovr_op;                  // index into the run's ops that is replaced by `ovr`, or -1
    uint32_t    accumulate;              // dst += result instead of dst = result
    QgtDevOp    ovr;                     // the derivative op (generator folded into the gate)
} QgtSweepItem;

typedef struct QgtDevEdge { int32_t i, j; double w; } QgtDevEdge;

#endif

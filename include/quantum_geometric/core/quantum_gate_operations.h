/* Forwarding header: the reference's include path for the gate objects (create_quantum_gate, ...).
 * Everything is declared in include/qgt_compat.h (see there for the reference lines each block follows). */
#ifndef QGT_B200_FWD_CORE_QUANTUM_GATE_OPERATIONS_H
#define QGT_B200_FWD_CORE_QUANTUM_GATE_OPERATIONS_H
#include "../../qgt_compat.h"
#endif

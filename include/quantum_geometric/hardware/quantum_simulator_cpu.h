/* Forwarding header: the reference's include path for this part of the QGT hot path.
 * Everything is declared in include/qgt_compat.h (see there for the reference lines each block follows). */
#ifndef QGT_B200_FWD_HARDWARE_QUANTUM_SIMULATOR_CPU_H
#define QGT_B200_FWD_HARDWARE_QUANTUM_SIMULATOR_CPU_H
#include "../../qgt_compat.h"
#endif

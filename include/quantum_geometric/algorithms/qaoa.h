/* Forwarding header: the reference's include path for the QAOA driver (qaoa_init, qaoa_apply_circuit, ...).
 * Everything is declared in include/qgt_compat.h (see there for the reference lines each block follows). */
#ifndef QGT_B200_FWD_ALGORITHMS_QAOA_H
#define QGT_B200_FWD_ALGORITHMS_QAOA_H
#include "../../qgt_compat.h"
#endif

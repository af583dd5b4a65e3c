/* Forwarding header: the reference's include path for the generic tensor objects (geometric_tensor_create, ...).
 * Everything is declared in include/qgt_compat.h (see there for the reference lines each block follows). */
#ifndef QGT_B200_FWD_CORE_QUANTUM_GEOMETRIC_TENSOR_H
#define QGT_B200_FWD_CORE_QUANTUM_GEOMETRIC_TENSOR_H
#include "../../qgt_compat.h"
#endif

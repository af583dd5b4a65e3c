/* Forwarding header: the reference's include path for the device-memory seam (qg_gpu_*, gpu_malloc, ...).
 * Everything is declared in include/qgt_compat.h (see there for the reference lines each block follows). */
#ifndef QGT_B200_FWD_CORE_QUANTUM_GEOMETRIC_GPU_H
#define QGT_B200_FWD_CORE_QUANTUM_GEOMETRIC_GPU_H
#include "../../qgt_compat.h"
#endif

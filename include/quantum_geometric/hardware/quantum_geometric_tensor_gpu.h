/* Forwarding header: the reference's include path for the QGT-on-GPU seam (GPUContext, compute_quantum_*_gpu).
 * Everything is declared in include/qgt_compat.h (see there for the reference lines each block follows). */
#ifndef QGT_B200_FWD_HARDWARE_QUANTUM_GEOMETRIC_TENSOR_GPU_H
#define QGT_B200_FWD_HARDWARE_QUANTUM_GEOMETRIC_TENSOR_GPU_H
#include "../../qgt_compat.h"
#endif

/* Forwarding header: the reference's include path for the pluggable compute-backend seam (ComputeBackendOps,
 * compute_register_backend, ...).  Everything is declared in include/qgt_compute_backend.h. */
#ifndef QGT_B200_FWD_SUPERCOMPUTER_COMPUTE_TYPES_H
#define QGT_B200_FWD_SUPERCOMPUTER_COMPUTE_TYPES_H
#include "../../qgt_compute_backend.h"
#endif

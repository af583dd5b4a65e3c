/* Forwarding header: the reference's include path for the numerical-backend set-up calls.
 * Everything is declared in include/qgt_compat.h (see there for the reference lines each block follows). */
#ifndef QGT_B200_FWD_CORE_NUMERICAL_BACKEND_H
#define QGT_B200_FWD_CORE_NUMERICAL_BACKEND_H
#include "../../qgt_compat.h"
#endif

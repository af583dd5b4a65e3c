/* compat_common.h — shared plumbing of the reference-API wrappers: one lazily created context per process. */
#ifndef QGT_COMPAT_COMMON_H
#define QGT_COMPAT_COMMON_H
#include "qgt_b200.h"
#include "qgt_compat.h"

qgt_b200_ctx* qgt_compat_ctx(void);                 /* NULL (and an error recorded) when no device is usable */
void qgt_compat_set_error(const char* where, int status);
/* reference gate kind -> qgt_b200 gate; returns 0 when the kind has no statevector meaning here */
int qgt_compat_convert_gate(gate_type_t type, uint32_t target, uint32_t control, double angle, int param, qgt_b200_gate* out);
#endif

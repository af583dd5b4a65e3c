/* Empty stand-in for <cblas.h>: distributed/differential_geometry.c includes it but the two
 * routines the oracle uses (diffgeo_compute_fubini_study / _berry_curvature) call no BLAS. */
#ifndef QGT_ORACLE_SHIM_CBLAS_H
#define QGT_ORACLE_SHIM_CBLAS_H
#endif

/* Pre-included (gcc -include) when compiling the reference's supercomputer/compute_simd.c for oracle/_ref: on x86-64
 * that file expects a compute_simd_avx2.c the reference tree does not contain, so its scalar fallback (the code the CPU
 * backend's arithmetic is defined by) is selected by hiding the architecture macro from the reference's own
 * compute_types.h.  System headers are pulled in first, with the macro intact. */
#include <math.h>
#include <stdbool.h>
#include <stddef.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#undef __x86_64__
#undef _M_X64

/*
 * tensor_compat.c — the generic rank-N tensor objects of the reference's core API
 * (include/quantum_geometric/core/quantum_geometric_tensor.h:13-100, src/.../core/quantum_geometric_tensor.c).
 *
 * Small host algebra on ComplexFloat arrays (no 2^n-sized operand ever passes through it): SURVEY.md section 8b lists it
 * among the "core API kept as host C".  The reference's version cannot be linked here (it needs core/simd_operations.c,
 * which does not compile, and the qrng module), so the entry points its tests/test_quantum_geometric_tensor.c and
 * tests/test_quantum_geometric_tensor_init.c drive are provided natively: create / destroy / clone, initialise (zero,
 * identity, data, random), add / subtract / scale / conjugate / adjoint, multiply (matrix product on the last / first
 * index), contract (any index pairs), transpose (any permutation), outer and inner product, norm, trace, the Hermitian
 * and unitary predicates and the validators.  Same argument checks and error codes as the reference
 * (QGT_CHECK_NULL / QGT_CHECK_ARGUMENT -> QGT_ERROR_INVALID_PARAMETER), with one deliberate difference: a tensor whose
 * component array is gone validates as QGT_ERROR_INVALID_STATE — what the reference's own test asserts
 * (tests/test_quantum_geometric_tensor.c:122-126) and its implementation does not return.
 */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "compat_common.h"

#define HERMITIAN_TOL 1e-6f

static size_t count_elements(const size_t* dims, size_t rank, int* ok) {
    size_t total = 1;
    *ok = dims != NULL && rank > 0;
    for (size_t i = 0; *ok && i < rank; i++) {
        if (dims[i] == 0 || total > SIZE_MAX / dims[i]) *ok = 0;
        else total *= dims[i];
    }
    return total;
}

qgt_error_t geometric_tensor_create(quantum_geometric_tensor_t** tensor, geometric_tensor_type_t type, const size_t* dimensions, size_t rank) {
    if (!tensor) return QGT_ERROR_INVALID_PARAMETER;
    int ok;
    const size_t total = count_elements(dimensions, rank, &ok);
    if (!ok) return QGT_ERROR_INVALID_PARAMETER;
    const size_t aligned = (total + 7) & ~(size_t)7;
    quantum_geometric_tensor_t* t = (quantum_geometric_tensor_t*)calloc(1, sizeof *t);
    size_t* dims = (size_t*)malloc(rank * sizeof(size_t));
    ComplexFloat* comps = (ComplexFloat*)aligned_alloc(64, ((aligned * sizeof(ComplexFloat) + 63) / 64) * 64);
    if (!t || !dims || !comps) { free(t); free(dims); free(comps); return QGT_ERROR_MEMORY_ALLOCATION; }
    memset(comps, 0, aligned * sizeof(ComplexFloat));
    memcpy(dims, dimensions, rank * sizeof(size_t));
    t->type = type; t->dimensions = dims; t->rank = rank; t->dimension = dims[0]; t->components = comps;
    t->total_elements = total; t->aligned_elements = aligned;
    t->is_hermitian = type == GEOMETRIC_TENSOR_HERMITIAN; t->is_unitary = type == GEOMETRIC_TENSOR_UNITARY;
    t->is_symmetric = type == GEOMETRIC_TENSOR_SYMMETRIC || type == GEOMETRIC_TENSOR_HERMITIAN;
    t->hardware = HARDWARE_TYPE_CPU; t->mem_type = QGT_MEM_STANDARD;
    *tensor = t;
    return QGT_SUCCESS;
}

void geometric_tensor_destroy(quantum_geometric_tensor_t* t) {
    if (!t) return;
    free(t->dimensions); free(t->components); free(t->auxiliary_data); free(t);
}

qgt_error_t geometric_tensor_clone(quantum_geometric_tensor_t** dest, const quantum_geometric_tensor_t* src) {
    if (!dest || !src || !src->components) return QGT_ERROR_INVALID_PARAMETER;
    qgt_error_t err = geometric_tensor_create(dest, src->type, src->dimensions, src->rank);
    if (err) return err;
    memcpy((*dest)->components, src->components, src->total_elements * sizeof(ComplexFloat));
    (*dest)->is_hermitian = src->is_hermitian; (*dest)->is_unitary = src->is_unitary; (*dest)->is_symmetric = src->is_symmetric;
    return QGT_SUCCESS;
}

qgt_error_t geometric_tensor_initialize(quantum_geometric_tensor_t* t, const ComplexFloat* data) {
    if (!t || !t->components || !data) return QGT_ERROR_INVALID_PARAMETER;
    memcpy(t->components, data, t->total_elements * sizeof(ComplexFloat));
    return QGT_SUCCESS;
}

qgt_error_t geometric_tensor_initialize_zero(quantum_geometric_tensor_t* t) {
    if (!t || !t->components) return QGT_ERROR_INVALID_PARAMETER;
    memset(t->components, 0, t->aligned_elements * sizeof(ComplexFloat));
    return QGT_SUCCESS;
}

qgt_error_t geometric_tensor_initialize_identity(quantum_geometric_tensor_t* t) {
    if (!t || !t->components || !t->dimensions) return QGT_ERROR_INVALID_PARAMETER;
    for (size_t i = 1; i < t->rank; i++) if (t->dimensions[i] != t->dimensions[0]) return QGT_ERROR_INVALID_PARAMETER;
    memset(t->components, 0, t->aligned_elements * sizeof(ComplexFloat));
    size_t stride = 0, s = 1;
    for (size_t i = 0; i < t->rank; i++) { stride += s; s *= t->dimensions[0]; }      /* (i, i, ..., i) */
    for (size_t i = 0; i < t->dimensions[0]; i++) t->components[i * stride].real = 1.0f;
    return QGT_SUCCESS;
}

/* unit-modulus random entries (Hermitian-typed square matrices come out Hermitian), as quantum_geometric_tensor.c:257-343;
   the generator is a fixed-seed 64-bit LCG instead of the reference's qrng module */
static unsigned long long g_tensor_rng = 0x2545F4914F6CDD1Dull;
static float tensor_uniform(void) {
    g_tensor_rng = g_tensor_rng * 6364136223846793005ull + 1442695040888963407ull;
    return (float)((g_tensor_rng >> 40) / (double)(1ull << 24));
}

qgt_error_t geometric_tensor_initialize_random(quantum_geometric_tensor_t* t, float min_val, float max_val) {
    if (!t || !t->components) return QGT_ERROR_INVALID_PARAMETER;
    if (min_val >= max_val) return QGT_ERROR_INVALID_ARGUMENT;
    const float range = max_val - min_val;
    ComplexFloat* d = t->components;
    const bool herm = t->type == GEOMETRIC_TENSOR_HERMITIAN && t->rank == 2 && t->dimensions[0] == t->dimensions[1];
    const size_t n = herm ? t->dimensions[0] : 0;
    for (size_t k = 0; k < t->total_elements; k++) {
        const size_t i = herm ? k / n : 0, j = herm ? k % n : 0;
        if (herm && j < i) continue;
        float re = min_val + tensor_uniform() * range, im = min_val + tensor_uniform() * range;
        if (herm && i == j) { d[k].real = re; d[k].imag = 0.0f; continue; }
        const float mag = sqrtf(re * re + im * im);
        if (mag > 0) { re /= mag; im /= mag; }
        d[k].real = re; d[k].imag = im;
        if (herm) { d[j * n + i].real = re; d[j * n + i].imag = -im; }
    }
    return QGT_SUCCESS;
}

static bool same_shape(const quantum_geometric_tensor_t* a, const quantum_geometric_tensor_t* b) {
    if (a->rank != b->rank) return false;
    for (size_t i = 0; i < a->rank; i++) if (a->dimensions[i] != b->dimensions[i]) return false;
    return true;
}

/* give `result` the shape dims[rank] (its buffers are re-allocated when the element count differs) */
static qgt_error_t reshape(quantum_geometric_tensor_t* result, const size_t* dims, size_t rank) {
    int ok;
    const size_t total = count_elements(dims, rank, &ok);
    if (!ok) return QGT_ERROR_INVALID_PARAMETER;
    const size_t aligned = (total + 7) & ~(size_t)7;
    if (!result->components || aligned > result->aligned_elements) {
        ComplexFloat* c = (ComplexFloat*)aligned_alloc(64, ((aligned * sizeof(ComplexFloat) + 63) / 64) * 64);
        if (!c) return QGT_ERROR_MEMORY_ALLOCATION;
        free(result->components);
        result->components = c;
    }
    if (rank != result->rank || !result->dimensions) {
        size_t* d = (size_t*)malloc(rank * sizeof(size_t));
        if (!d) return QGT_ERROR_MEMORY_ALLOCATION;
        free(result->dimensions);
        result->dimensions = d;
    }
    memcpy(result->dimensions, dims, rank * sizeof(size_t));
    result->rank = rank; result->dimension = dims[0]; result->total_elements = total; result->aligned_elements = aligned;
    memset(result->components, 0, aligned * sizeof(ComplexFloat));
    return QGT_SUCCESS;
}

static qgt_error_t elementwise(quantum_geometric_tensor_t* r, const quantum_geometric_tensor_t* a, const quantum_geometric_tensor_t* b, float sign) {
    if (!r || !a || !b || !a->components || !b->components) return QGT_ERROR_INVALID_PARAMETER;
    if (!same_shape(a, b)) return QGT_ERROR_INVALID_PARAMETER;
    if (r != a && r != b) { qgt_error_t e = reshape(r, a->dimensions, a->rank); if (e) return e; }
    for (size_t i = 0; i < a->total_elements; i++) {
        const ComplexFloat x = a->components[i], y = b->components[i];
        r->components[i].real = x.real + sign * y.real; r->components[i].imag = x.imag + sign * y.imag;
    }
    return QGT_SUCCESS;
}

qgt_error_t geometric_tensor_add(quantum_geometric_tensor_t* r, const quantum_geometric_tensor_t* a, const quantum_geometric_tensor_t* b) { return elementwise(r, a, b, 1.0f); }
qgt_error_t geometric_tensor_subtract(quantum_geometric_tensor_t* r, const quantum_geometric_tensor_t* a, const quantum_geometric_tensor_t* b) { return elementwise(r, a, b, -1.0f); }

qgt_error_t geometric_tensor_scale(quantum_geometric_tensor_t* r, const quantum_geometric_tensor_t* t, ComplexFloat s) {
    if (!r || !t || !t->components) return QGT_ERROR_INVALID_PARAMETER;
    if (r != t) { qgt_error_t e = reshape(r, t->dimensions, t->rank); if (e) return e; }
    for (size_t i = 0; i < t->total_elements; i++) {
        const ComplexFloat x = t->components[i];
        r->components[i].real = x.real * s.real - x.imag * s.imag; r->components[i].imag = x.real * s.imag + x.imag * s.real;
    }
    return QGT_SUCCESS;
}

qgt_error_t geometric_tensor_conjugate(quantum_geometric_tensor_t* r, const quantum_geometric_tensor_t* t) {
    if (!r || !t || !t->components) return QGT_ERROR_INVALID_PARAMETER;
    if (r != t) { qgt_error_t e = reshape(r, t->dimensions, t->rank); if (e) return e; }
    for (size_t i = 0; i < t->total_elements; i++) { r->components[i].real = t->components[i].real; r->components[i].imag = -t->components[i].imag; }
    return QGT_SUCCESS;
}

/* general contraction: the listed index pairs are summed; the result carries a's free indices, then b's */
qgt_error_t geometric_tensor_contract(quantum_geometric_tensor_t* r, const quantum_geometric_tensor_t* a, const quantum_geometric_tensor_t* b,
                                      const size_t* ia, const size_t* ib, size_t n) {
    if (!r || !a || !b || !a->components || !b->components || (n && (!ia || !ib))) return QGT_ERROR_INVALID_PARAMETER;
    if (r == a || r == b || n > a->rank || n > b->rank || a->rank > 16 || b->rank > 16) return QGT_ERROR_INVALID_PARAMETER;
    bool ca[16] = {0}, cb[16] = {0};
    for (size_t k = 0; k < n; k++) {
        if (ia[k] >= a->rank || ib[k] >= b->rank || ca[ia[k]] || cb[ib[k]]) return QGT_ERROR_INVALID_PARAMETER;
        if (a->dimensions[ia[k]] != b->dimensions[ib[k]]) return QGT_ERROR_DIMENSION_MISMATCH;
        ca[ia[k]] = cb[ib[k]] = true;
    }
    size_t sa[16], sb[16], odims[32], orank = 0, fa[16], fb[16], nfa = 0, nfb = 0;
    sa[a->rank - 1] = 1; for (size_t i = a->rank - 1; i-- > 0;) sa[i] = sa[i + 1] * a->dimensions[i + 1];
    sb[b->rank - 1] = 1; for (size_t i = b->rank - 1; i-- > 0;) sb[i] = sb[i + 1] * b->dimensions[i + 1];
    for (size_t i = 0; i < a->rank; i++) if (!ca[i]) { fa[nfa++] = i; odims[orank++] = a->dimensions[i]; }
    for (size_t i = 0; i < b->rank; i++) if (!cb[i]) { fb[nfb++] = i; odims[orank++] = b->dimensions[i]; }
    const size_t one = 1;
    qgt_error_t e = reshape(r, orank ? odims : &one, orank ? orank : 1);
    if (e) return e;
    size_t csize = 1;
    for (size_t k = 0; k < n; k++) csize *= a->dimensions[ia[k]];
    for (size_t o = 0; o < r->total_elements; o++) {
        size_t rem = o, offa = 0, offb = 0;
        for (size_t k = orank; k-- > 0;) {
            const size_t idx = rem % odims[k]; rem /= odims[k];
            if (k >= nfa) offb += idx * sb[fb[k - nfa]]; else offa += idx * sa[fa[k]];
        }
        double sr = 0.0, si = 0.0;
        for (size_t c = 0; c < csize; c++) {
            size_t rc = c, pa = offa, pb = offb;
            for (size_t k = n; k-- > 0;) { const size_t idx = rc % a->dimensions[ia[k]]; rc /= a->dimensions[ia[k]]; pa += idx * sa[ia[k]]; pb += idx * sb[ib[k]]; }
            const ComplexFloat x = a->components[pa], y = b->components[pb];
            sr += (double)x.real * y.real - (double)x.imag * y.imag; si += (double)x.real * y.imag + (double)x.imag * y.real;
        }
        r->components[o].real = (float)sr; r->components[o].imag = (float)si;
    }
    return QGT_SUCCESS;
}

/* last index of a with the first index of b: the matrix product for rank 2 (quantum_geometric_tensor.c:599-830) */
qgt_error_t geometric_tensor_multiply(quantum_geometric_tensor_t* r, const quantum_geometric_tensor_t* a, const quantum_geometric_tensor_t* b) {
    if (!r || !a || !b) return QGT_ERROR_INVALID_PARAMETER;
    if (a->rank != b->rank) return QGT_ERROR_INVALID_PARAMETER;
    if (a->dimensions[a->rank - 1] != b->dimensions[0]) return QGT_ERROR_INVALID_DIMENSION;
    const size_t la = a->rank - 1, fb = 0;
    return geometric_tensor_contract(r, a, b, &la, &fb, 1);
}

qgt_error_t geometric_tensor_outer_product(quantum_geometric_tensor_t* r, const quantum_geometric_tensor_t* a, const quantum_geometric_tensor_t* b) {
    return geometric_tensor_contract(r, a, b, NULL, NULL, 0);
}

qgt_error_t geometric_tensor_inner_product(ComplexFloat* out, const quantum_geometric_tensor_t* a, const quantum_geometric_tensor_t* b) {
    if (!out || !a || !b || !a->components || !b->components || !same_shape(a, b)) return QGT_ERROR_INVALID_PARAMETER;
    double sr = 0.0, si = 0.0;
    for (size_t i = 0; i < a->total_elements; i++) {          /* <a|b> = sum conj(a) b */
        const ComplexFloat x = a->components[i], y = b->components[i];
        sr += (double)x.real * y.real + (double)x.imag * y.imag; si += (double)x.real * y.imag - (double)x.imag * y.real;
    }
    out->real = (float)sr; out->imag = (float)si;
    return QGT_SUCCESS;
}

qgt_error_t geometric_tensor_transpose(quantum_geometric_tensor_t* r, const quantum_geometric_tensor_t* t, const size_t* perm) {
    if (!r || !t || !perm || !t->components || r == t || t->rank > 16) return QGT_ERROR_INVALID_PARAMETER;
    bool used[16] = {0};
    size_t nd[16], st[16];
    for (size_t i = 0; i < t->rank; i++) {
        if (perm[i] >= t->rank || used[perm[i]]) return QGT_ERROR_INVALID_ARGUMENT;
        used[perm[i]] = true;
        nd[i] = t->dimensions[perm[i]];
    }
    st[t->rank - 1] = 1; for (size_t i = t->rank - 1; i-- > 0;) st[i] = st[i + 1] * t->dimensions[i + 1];
    qgt_error_t e = reshape(r, nd, t->rank);
    if (e) return e;
    for (size_t o = 0; o < r->total_elements; o++) {           /* result[i_0..] = tensor[index perm[k] = i_k] */
        size_t rem = o, src = 0;
        for (size_t k = t->rank; k-- > 0;) { src += (rem % nd[k]) * st[perm[k]]; rem /= nd[k]; }
        r->components[o] = t->components[src];
    }
    return QGT_SUCCESS;
}

qgt_error_t geometric_tensor_adjoint(quantum_geometric_tensor_t* r, const quantum_geometric_tensor_t* t) {
    if (!r || !t || t->rank != 2) return QGT_ERROR_INVALID_PARAMETER;
    const size_t perm[2] = {1, 0};
    qgt_error_t e = geometric_tensor_transpose(r, t, perm);
    return e ? e : geometric_tensor_conjugate(r, r);
}

qgt_error_t geometric_tensor_norm(float* norm, const quantum_geometric_tensor_t* t) {
    if (!norm || !t || !t->components) return QGT_ERROR_INVALID_PARAMETER;
    double s = 0.0;
    for (size_t i = 0; i < t->total_elements; i++) s += (double)t->components[i].real * t->components[i].real + (double)t->components[i].imag * t->components[i].imag;
    *norm = (float)sqrt(s);
    return QGT_SUCCESS;
}

qgt_error_t geometric_tensor_trace(ComplexFloat* trace, const quantum_geometric_tensor_t* t) {
    if (!trace || !t || !t->components || t->rank != 2 || t->dimensions[0] != t->dimensions[1]) return QGT_ERROR_INVALID_PARAMETER;
    double sr = 0.0, si = 0.0;
    for (size_t i = 0; i < t->dimensions[0]; i++) { sr += t->components[i * t->dimensions[0] + i].real; si += t->components[i * t->dimensions[0] + i].imag; }
    trace->real = (float)sr; trace->imag = (float)si;
    return QGT_SUCCESS;
}

bool geometric_tensor_is_hermitian(const quantum_geometric_tensor_t* t) {
    if (!t || !t->components || t->rank != 2 || t->dimensions[0] != t->dimensions[1]) return false;
    const size_t n = t->dimensions[0];
    for (size_t i = 0; i < n; i++) {
        if (fabsf(t->components[i * n + i].imag) > HERMITIAN_TOL) return false;
        for (size_t j = i + 1; j < n; j++) {
            const ComplexFloat a = t->components[i * n + j], b = t->components[j * n + i];
            if (fabsf(a.real - b.real) > HERMITIAN_TOL || fabsf(a.imag + b.imag) > HERMITIAN_TOL) return false;
        }
    }
    return true;
}

bool geometric_tensor_is_unitary(const quantum_geometric_tensor_t* t) {
    if (!t || !t->components || t->rank != 2 || t->dimensions[0] != t->dimensions[1]) return false;
    const size_t n = t->dimensions[0];
    for (size_t i = 0; i < n; i++)
        for (size_t j = 0; j < n; j++) {              /* (U U^+)_ij */
            double sr = 0.0, si = 0.0;
            for (size_t k = 0; k < n; k++) {
                const ComplexFloat a = t->components[i * n + k], b = t->components[j * n + k];
                sr += (double)a.real * b.real + (double)a.imag * b.imag; si += (double)a.imag * b.real - (double)a.real * b.imag;
            }
            if (fabs(sr - (i == j ? 1.0 : 0.0)) > 1e-5 || fabs(si) > 1e-5) return false;
        }
    return true;
}

qgt_error_t geometric_tensor_validate_dimensions(const quantum_geometric_tensor_t* t) {
    if (!t || !t->dimensions) return QGT_ERROR_INVALID_PARAMETER;
    int ok;
    const size_t total = count_elements(t->dimensions, t->rank, &ok);
    if (!ok || total != t->total_elements || ((total + 7) & ~(size_t)7) != t->aligned_elements) return QGT_ERROR_INVALID_PARAMETER;
    return QGT_SUCCESS;
}

qgt_error_t geometric_tensor_validate(const quantum_geometric_tensor_t* t) {
    if (!t || !t->dimensions) return QGT_ERROR_INVALID_PARAMETER;
    if (!t->components) return QGT_ERROR_INVALID_STATE;
    qgt_error_t e = geometric_tensor_validate_dimensions(t);
    if (e) return e;
    for (size_t i = 0; i < t->total_elements; i++)
        if (!isfinite(t->components[i].real) || !isfinite(t->components[i].imag)) return -24;      /* QGT_ERROR_NUMERICAL_INSTABILITY */
    return QGT_SUCCESS;
}

qgt_error_t geometric_tensor_validate_compatibility(const quantum_geometric_tensor_t* a, const quantum_geometric_tensor_t* b) {
    if (!a || !b) return QGT_ERROR_INVALID_PARAMETER;
    return same_shape(a, b) ? QGT_SUCCESS : QGT_ERROR_DIMENSION_MISMATCH;
}

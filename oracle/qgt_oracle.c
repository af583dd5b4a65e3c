/*
 * qgt_oracle.c — CPU restatement of the reference's QGT hot path.   TEST INFRASTRUCTURE ONLY.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
 * load this file's library.  The product (quantum_geometric_tensor_b200/) never links or calls it.
 *
 * Parity pinning: the gate kernels and the g / F assembly are checked bit-for-bit (gate sweeps)
 * and to 1e-13 (assembly) against the UNMODIFIED reference sources compiled by oracle/Makefile
 * into oracle/_ref/libqgt_ref.so (tests/test_oracle.py), against the known-answer magnitudes of
 * the reference's tests/test_quantum_simulator_cpu.c (H, Bell, X, HZ, RX(pi), GHZ), and against
 * the independent dense-matrix QGT in oracle/dense_qgt.py.  For Q_mu_nu itself the reference holds
 * no golden value: "parity unpinned by reference tests" — pinned instead by the
 * reference's own diffgeo routines run here plus the analytic single-qubit closed form.
 *
 * All file:line citations are relative to the reference tree.
 */
#include <complex.h>
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include "../include/qgt_b200.h" /* POD circuit description only */

typedef double complex c128;

#define ORC_PI 3.14159265358979323846
#define ORC_SQRT2 1.41421356237309504880 /* same literal as hardware/quantum_simulator.c:36 */

/* ---- gate application: hardware/quantum_simulator.c:147-163 ------------------------------ */
static void orc_single(c128* a, int target, c128 g00, c128 g01, c128 g10, c128 g11, int n) {
    size_t dim = (size_t)1 << n, tm = (size_t)1 << target;
    for (size_t i = 0; i < dim; i++) {
        if ((i & tm) == 0) {
            size_t j = i | tm;
            c128 a0 = a[i], a1 = a[j];
            a[i] = g00 * a0 + g01 * a1;
            a[j] = g10 * a0 + g11 * a1;
        }
    }
}

/* hardware/quantum_simulator.c:166-185 */
static void orc_controlled(c128* a, int control, int target, c128 g00, c128 g01, c128 g10, c128 g11, int n) {
    size_t dim = (size_t)1 << n, cm = (size_t)1 << control, tm = (size_t)1 << target;
    for (size_t i = 0; i < dim; i++) {
        if ((i & cm) && (i & tm) == 0) {
            size_t j = i | tm;
            c128 a0 = a[i], a1 = a[j];
            a[i] = g00 * a0 + g01 * a1;
            a[j] = g10 * a0 + g11 * a1;
        }
    }
}

/* E_z of the MaxCut cost: algorithms/qaoa.c:258-289 (cut weight + vertex terms), in double */
static double orc_cost_energy(const qgt_b200_circuit* c, size_t z) {
    double e = 0.0;
    for (size_t k = 0; k < c->num_edges; k++) {
        int zi = (int)((z >> c->edges[k].i) & 1), zj = (int)((z >> c->edges[k].j) & 1);
        if (zi != zj) e += c->edges[k].weight;
    }
    if (c->vertex_weights)
        for (int q = 0; q < c->num_qubits; q++) e += c->vertex_weights[q] * (double)(1 - 2 * (int)((z >> q) & 1));
    return e;
}

/* Apply one gate with a resolved angle.  Tables: hardware/quantum_simulator.c:188-283 for the 13
 * kinds the reference simulator knows; the remaining kinds follow the standard definitions the
 * reference's enum names (core/quantum_base_types.h:32-72). */
static int orc_apply_gate(c128* a, int n, const qgt_b200_circuit* c, const qgt_b200_gate* g, double th) {
    int t = g->target, ctl = g->control;
    double cs = cos(th / 2.0), sn = sin(th / 2.0);
    switch (g->kind) {
    case QGT_B200_GATE_I: return 0;
    case QGT_B200_GATE_X: orc_single(a, t, 0, 1, 1, 0, n); return 0;
    case QGT_B200_GATE_Y: orc_single(a, t, 0, -I, I, 0, n); return 0;
    case QGT_B200_GATE_Z: orc_single(a, t, 1, 0, 0, -1, n); return 0;
    case QGT_B200_GATE_H: {
        c128 h = 1.0 / ORC_SQRT2;
        orc_single(a, t, h, h, h, -1.0 / ORC_SQRT2, n); return 0; }
    case QGT_B200_GATE_S: orc_single(a, t, 1, 0, 0, I, n); return 0;
    case QGT_B200_GATE_T: orc_single(a, t, 1, 0, 0, cexp(I * ORC_PI / 4.0), n); return 0;
    case QGT_B200_GATE_SDG: orc_single(a, t, 1, 0, 0, -I, n); return 0;
    case QGT_B200_GATE_TDG: orc_single(a, t, 1, 0, 0, cexp(-I * ORC_PI / 4.0), n); return 0;
    case QGT_B200_GATE_SX: orc_single(a, t, 0.5 + 0.5 * I, 0.5 - 0.5 * I, 0.5 - 0.5 * I, 0.5 + 0.5 * I, n); return 0;
    case QGT_B200_GATE_RX: orc_single(a, t, cs, -I * sn, -I * sn, cs, n); return 0;
    case QGT_B200_GATE_RY: orc_single(a, t, cs, -sn, sn, cs, n); return 0;
    case QGT_B200_GATE_RZ: orc_single(a, t, cexp(-I * th / 2.0), 0, 0, cexp(I * th / 2.0), n); return 0;
    case QGT_B200_GATE_U1:
    case QGT_B200_GATE_PHASE: orc_single(a, t, 1, 0, 0, cexp(I * th), n); return 0;
    case QGT_B200_GATE_CNOT: orc_controlled(a, ctl, t, 0, 1, 1, 0, n); return 0;
    case QGT_B200_GATE_CY: orc_controlled(a, ctl, t, 0, -I, I, 0, n); return 0;
    case QGT_B200_GATE_CZ: orc_controlled(a, ctl, t, 1, 0, 0, -1, n); return 0;
    case QGT_B200_GATE_CH: {
        c128 h = 1.0 / ORC_SQRT2;
        orc_controlled(a, ctl, t, h, h, h, -1.0 / ORC_SQRT2, n); return 0; }
    case QGT_B200_GATE_CRX: orc_controlled(a, ctl, t, cs, -I * sn, -I * sn, cs, n); return 0;
    case QGT_B200_GATE_CRY: orc_controlled(a, ctl, t, cs, -sn, sn, cs, n); return 0;
    case QGT_B200_GATE_CRZ: orc_controlled(a, ctl, t, cexp(-I * th / 2.0), 0, 0, cexp(I * th / 2.0), n); return 0;
    case QGT_B200_GATE_SWAP: /* three CNOTs, hardware/quantum_simulator.c:269-275 */
        orc_controlled(a, ctl, t, 0, 1, 1, 0, n);
        orc_controlled(a, t, ctl, 0, 1, 1, 0, n);
        orc_controlled(a, ctl, t, 0, 1, 1, 0, n);
        return 0;
    case QGT_B200_GATE_ZZ: { /* exp(-i th/2 Z_t Z_ctl) */
        size_t dim = (size_t)1 << n;
        c128 pe = cexp(-I * th / 2.0), po = cexp(I * th / 2.0);
        for (size_t i = 0; i < dim; i++) a[i] *= ((((i >> t) ^ (i >> ctl)) & 1) ? po : pe);
        return 0; }
    case QGT_B200_GATE_COST: { /* algorithms/qaoa.c:344-372: |z> -> exp(-i gamma E_z)|z> */
        size_t dim = (size_t)1 << n;
        for (size_t z = 0; z < dim; z++) {
            double ph = -th * orc_cost_energy(c, z);
            a[z] *= (cos(ph) + I * sin(ph));
        }
        return 0; }
    default: return -1;
    }
}

/* Generator of a parameterised gate applied to the state, times d(angle)/d(theta):
 *   d/dtheta exp(-i a P/2) = (-i/2) P exp(...)   with a = scale*theta + angle
 * (the reference's own derivative engine is non-functional, so the restatement
 * uses the exact generator).  Returns -1 for a kind that has no parameter. */
static int orc_apply_generator(c128* a, int n, const qgt_b200_circuit* c, const qgt_b200_gate* g) {
    size_t dim = (size_t)1 << n;
    int t = g->target, ctl = g->control;
    c128 f = -0.5 * I * g->scale;
    switch (g->kind) {
    case QGT_B200_GATE_RX: orc_single(a, t, 0, f, f, 0, n); return 0;
    case QGT_B200_GATE_RY: orc_single(a, t, 0, f * (-I), f * I, 0, n); return 0;
    case QGT_B200_GATE_RZ: orc_single(a, t, f, 0, 0, -f, n); return 0;
    case QGT_B200_GATE_U1:
    case QGT_B200_GATE_PHASE: orc_single(a, t, 0, 0, 0, I * g->scale, n); return 0;
    case QGT_B200_GATE_CRX:
    case QGT_B200_GATE_CRY:
    case QGT_B200_GATE_CRZ: {
        size_t cm = (size_t)1 << ctl;
        for (size_t i = 0; i < dim; i++) if (!(i & cm)) a[i] = 0;
        if (g->kind == QGT_B200_GATE_CRX) orc_controlled(a, ctl, t, 0, f, f, 0, n);
        else if (g->kind == QGT_B200_GATE_CRY) orc_controlled(a, ctl, t, 0, f * (-I), f * I, 0, n);
        else orc_controlled(a, ctl, t, f, 0, 0, -f, n);
        return 0; }
    case QGT_B200_GATE_ZZ:
        for (size_t i = 0; i < dim; i++) a[i] *= ((((i >> t) ^ (i >> ctl)) & 1) ? -f : f);
        return 0;
    case QGT_B200_GATE_COST:
        for (size_t z = 0; z < dim; z++) a[z] *= (-I * g->scale * orc_cost_energy(c, z));
        return 0;
    default: return -1;
    }
}

static double orc_angle(const qgt_b200_gate* g, const double* theta) {
    return g->param >= 0 ? g->scale * theta[g->param] + g->angle : g->angle;
}

/* init_simulator_state (hardware/quantum_simulator_cpu.c:98-101) / qaoa_prepare_initial_state */
void orc_init_state(double* amps, int n, int initial_state) {
    c128* a = (c128*)amps;
    size_t dim = (size_t)1 << n;
    if (initial_state == QGT_B200_INIT_PLUS) {
        double v = 1.0 / sqrt((double)dim);
        for (size_t i = 0; i < dim; i++) a[i] = v;
    } else {
        memset(a, 0, dim * sizeof(c128));
        a[0] = 1.0;
    }
}

/* sim_execute_circuit (hardware/quantum_simulator.c:499-533), noise off */
int orc_apply_circuit(double* amps, const qgt_b200_circuit* c, const double* theta) {
    c128* a = (c128*)amps;
    for (size_t k = 0; k < c->num_gates; k++)
        if (orc_apply_gate(a, c->num_qubits, c, &c->gates[k], orc_angle(&c->gates[k], theta))) return -1;
    return 0;
}

/* d_mu psi = sum over gates k carrying parameter mu of  U_{>k} G_k U_{<=k} |init>    */
int orc_derivative(double* out, const qgt_b200_circuit* c, const double* theta, int mu) {
    int n = c->num_qubits;
    size_t dim = (size_t)1 << n;
    c128* acc = (c128*)out;
    c128* w = (c128*)malloc(dim * sizeof(c128));
    if (!w) return -2;
    memset(acc, 0, dim * sizeof(c128));
    for (size_t k = 0; k < c->num_gates; k++) {
        if (c->gates[k].param != mu) continue;
        orc_init_state((double*)w, n, c->initial_state);
        for (size_t j = 0; j <= k; j++) orc_apply_gate(w, n, c, &c->gates[j], orc_angle(&c->gates[j], theta));
        if (orc_apply_generator(w, n, c, &c->gates[k])) { free(w); return -1; }
        for (size_t j = k + 1; j < c->num_gates; j++) orc_apply_gate(w, n, c, &c->gates[j], orc_angle(&c->gates[j], theta));
        for (size_t i = 0; i < dim; i++) acc[i] += w[i];
    }
    free(w);
    return 0;
}

/* diffgeo_compute_fubini_study (distributed/differential_geometry.c:2819-2862) and
 * diffgeo_compute_berry_curvature (:2864-2906) in one pass: the same triple loop, same
 * accumulation order (k inner), metric = Re Q, berry_core = Im Q (core convention,
 * core/quantum_geometric_curvature.c:201-247); diffgeo's F is -2*Im Q. */
void orc_qgt_from_columns(const double* psi_, const double* dpsi_, size_t dim, size_t P,
                          double* metric, double* berry, double* q_full) {
    const c128* psi = (const c128*)psi_;
    const c128* d = (const c128*)dpsi_;
    for (size_t i = 0; i < P; i++) {
        for (size_t j = 0; j < P; j++) {
            c128 inner = 0, vi = 0, vj = 0;
            for (size_t k = 0; k < dim; k++) {
                c128 dic = conj(d[i * dim + k]);
                inner += dic * d[j * dim + k];
                vi += dic * psi[k];
                vj += conj(psi[k]) * d[j * dim + k];
            }
            c128 q = inner - vi * vj;
            if (metric) metric[i * P + j] = creal(q);
            if (berry) berry[i * P + j] = cimag(q);
            if (q_full) { q_full[2 * (i * P + j)] = creal(q); q_full[2 * (i * P + j) + 1] = cimag(q); }
        }
    }
}

/* Full QGT of a circuit: psi, all P derivative columns, then the assembly above. */
int orc_qgt(const qgt_b200_circuit* c, const double* theta, double* metric, double* berry, double* q_full) {
    int n = c->num_qubits;
    size_t dim = (size_t)1 << n, P = (size_t)c->num_params;
    double* psi = (double*)malloc(dim * 16);
    double* J = (double*)malloc(P * dim * 16);
    if (!psi || !J) { free(psi); free(J); return -2; }
    orc_init_state(psi, n, c->initial_state);
    int rc = orc_apply_circuit(psi, c, theta);
    for (size_t mu = 0; mu < P && rc == 0; mu++) rc = orc_derivative(J + mu * dim * 2, c, theta, (int)mu);
    if (rc == 0) orc_qgt_from_columns(psi, J, dim, P, metric, berry, q_full);
    free(psi); free(J);
    return rc;
}

/* ---- natural gradient: core/quantum_geometric_gradient.c:2721-2964 ----------------------- */
/* The reference works in ComplexFloat and calls LAPACK cgesvd_ for the condition number; the
 * restatement keeps the control flow (adaptive lambda = 1e-6*sqrt(kappa) when kappa > threshold,
 * Tikhonov G + lambda I, inverse, G^-1 g) in double on the REAL symmetric metric, with kappa from a
 * Jacobi eigen-decomposition (singular values of a symmetric matrix = |eigenvalues|).
 * Third-party arithmetic (LAPACK, unpinned) => this function is "parity unpinned". */
static void orc_jacobi_eig(double* A, size_t n, double* w) {
    for (int sweep = 0; sweep < 100; sweep++) {
        double off = 0;
        for (size_t p = 0; p < n; p++) for (size_t q = p + 1; q < n; q++) off += A[p * n + q] * A[p * n + q];
        if (off < 1e-300) break;
        for (size_t p = 0; p < n; p++) for (size_t q = p + 1; q < n; q++) {
            double apq = A[p * n + q];
            if (fabs(apq) < 1e-300) continue;
            double tau = (A[q * n + q] - A[p * n + p]) / (2.0 * apq);
            double t = (tau >= 0 ? 1.0 : -1.0) / (fabs(tau) + sqrt(1.0 + tau * tau));
            double cc = 1.0 / sqrt(1.0 + t * t), ss = t * cc;
            for (size_t k = 0; k < n; k++) {
                double akp = A[k * n + p], akq = A[k * n + q];
                A[k * n + p] = cc * akp - ss * akq; A[k * n + q] = ss * akp + cc * akq;
            }
            for (size_t k = 0; k < n; k++) {
                double apk = A[p * n + k], aqk = A[q * n + k];
                A[p * n + k] = cc * apk - ss * aqk; A[q * n + k] = ss * apk + cc * aqk;
            }
        }
    }
    for (size_t i = 0; i < n; i++) w[i] = A[i * n + i];
}

int orc_natural_gradient(const double* metric, const double* grad, size_t n,
                         const qgt_b200_natgrad_config* cfg, double* out, double* lambda_used) {
    double lambda = cfg->regularization;
    if (cfg->adaptive) { /* :2898-2912 */
        double* A = (double*)malloc(n * n * sizeof(double));
        double* w = (double*)malloc(n * sizeof(double));
        memcpy(A, metric, n * n * sizeof(double));
        orc_jacobi_eig(A, n, w);
        double smax = fabs(w[0]), smin = fabs(w[0]);
        for (size_t i = 1; i < n; i++) { /* :2763-2767 */
            double s = fabs(w[i]);
            if (s > smax) smax = s;
            if (s < smin && s > 0) smin = s;
        }
        double kappa = smin > 1e-15 ? smax / smin : INFINITY; /* :2770-2774 */
        if (kappa > cfg->condition_threshold) {
            double al = 1e-6 * sqrt(kappa);
            if (al > lambda) lambda = al;
        }
        free(A); free(w);
    }
    if (lambda_used) *lambda_used = lambda;
    /* (G + lambda I) x = g by Gaussian elimination with partial pivoting (matrix_inverse + multiply, :2931-2959) */
    double* M = (double*)malloc(n * (n + 1) * sizeof(double));
    for (size_t i = 0; i < n; i++) {
        for (size_t j = 0; j < n; j++) M[i * (n + 1) + j] = metric[i * n + j] + (i == j ? lambda : 0.0);
        M[i * (n + 1) + n] = grad[i];
    }
    for (size_t col = 0; col < n; col++) {
        size_t piv = col;
        for (size_t r = col + 1; r < n; r++) if (fabs(M[r * (n + 1) + col]) > fabs(M[piv * (n + 1) + col])) piv = r;
        if (fabs(M[piv * (n + 1) + col]) < 1e-300) { free(M); return -55; }
        if (piv != col) for (size_t k = 0; k <= n; k++) { double tmp = M[col * (n + 1) + k]; M[col * (n + 1) + k] = M[piv * (n + 1) + k]; M[piv * (n + 1) + k] = tmp; }
        for (size_t r = col + 1; r < n; r++) {
            double f = M[r * (n + 1) + col] / M[col * (n + 1) + col];
            for (size_t k = col; k <= n; k++) M[r * (n + 1) + k] -= f * M[col * (n + 1) + k];
        }
    }
    for (size_t ii = n; ii-- > 0;) {
        double s = M[ii * (n + 1) + n];
        for (size_t k = ii + 1; k < n; k++) s -= M[ii * (n + 1) + k] * out[k];
        out[ii] = s / M[ii * (n + 1) + ii];
    }
    free(M);
    return 0;
}

/* <E> = sum_z |psi_z|^2 E_z (algorithms/qaoa.c:455-487) and its gradient 2 Re <d_mu psi| E |psi> */
int orc_expectation_gradient(const qgt_b200_circuit* c, const double* theta, double* energy, double* grad) {
    int n = c->num_qubits;
    size_t dim = (size_t)1 << n;
    c128* psi = (c128*)malloc(dim * sizeof(c128));
    c128* d = (c128*)malloc(dim * sizeof(c128));
    if (!psi || !d) { free(psi); free(d); return -2; }
    orc_init_state((double*)psi, n, c->initial_state);
    int rc = orc_apply_circuit((double*)psi, c, theta);
    double e = 0;
    for (size_t z = 0; z < dim; z++) e += (creal(psi[z]) * creal(psi[z]) + cimag(psi[z]) * cimag(psi[z])) * orc_cost_energy(c, z);
    if (energy) *energy = e;
    for (int mu = 0; mu < c->num_params && rc == 0 && grad; mu++) {
        rc = orc_derivative((double*)d, c, theta, mu);
        c128 s = 0;
        for (size_t z = 0; z < dim; z++) s += conj(d[z]) * psi[z] * orc_cost_energy(c, z);
        grad[mu] = 2.0 * creal(s);
    }
    free(psi); free(d);
    return rc;
}

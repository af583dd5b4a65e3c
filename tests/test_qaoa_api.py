"""The reference's QAOA driver API (algorithms/qaoa.h) served by the device library, against the unmodified reference
(oracle/_ref/libqaoa_ref.so = algorithms/qaoa.c) on the same graph and angles: state after qaoa_apply_circuit, expectation,
the reference's finite-shift gradient, sampling, the optimisation loop; plus the exact adjoint gradient against the oracle."""
import ctypes as C
import os

import numpy as np
import pytest

import qaoa_api as Q
from quantum_geometric_tensor_b200 import circuits as K

HAVE_REF = os.path.exists(Q.REF)


def problem(n, seed):
    rng = np.random.default_rng(seed)
    edges = [(i, j, float(rng.uniform(0.5, 1.5))) for (i, j, _) in K.random_regular3(n)] if n >= 4 and n % 2 == 0 else \
        [(q, (q + 1) % n, float(rng.uniform(0.5, 1.5))) for q in range(n)]
    vw = rng.uniform(-0.5, 0.5, n)
    return edges, vw


def shift_gradient(lib, s, gamma, beta, use_library):
    """[E(theta_k + pi/2) - E(theta_k - pi/2)] / 2 for every parameter: the formula qaoa_compute_gradient documents
    (qaoa.c:489-558).  The reference's own routine cannot be the comparison: qaoa_apply_circuit copies its arguments into
    state->gamma/beta, so the routine's `state->gamma[i] - SHIFT` reads the already shifted value and the "minus" circuit is
    the unshifted one; for the reference library the formula is evaluated from its
    qaoa_apply_circuit + qaoa_compute_expectation instead."""
    p = len(gamma)
    gg, bg = np.zeros(p), np.zeros(p)
    if use_library:
        assert lib.lib.qaoa_apply_circuit(s, Q.dp(gamma), Q.dp(beta)) == 0
        assert lib.lib.qaoa_compute_gradient(s, Q.dp(gg), Q.dp(bg)) == 0
        return gg, bg
    for which, out in ((0, gg), (1, bg)):
        for i in range(p):
            vals = []
            for sgn in (+1, -1):
                ga, be = gamma.copy(), beta.copy()
                (ga if which == 0 else be)[i] += sgn * np.pi / 2
                assert lib.lib.qaoa_apply_circuit(s, Q.dp(ga), Q.dp(be)) == 0
                e = C.c_double(0)
                assert lib.lib.qaoa_compute_expectation(s, C.byref(e)) == 0
                vals.append(e.value)
            out[i] = (vals[0] - vals[1]) / 2
    return gg, bg


def test_graph_helpers_without_gpu():
    lib = Q.Qaoa(Q.COMPAT)
    g = lib.graph(5, [(0, 1, 1.0), (1, 2, 2.0), (3, 4, 0.5)], [0.1, 0, 0, 0, -0.2])
    assert g.contents.num_edges == 3 and g.contents.edges[1].weight == 2.0 and g.contents.vertex_weights[4] == -0.2
    assert lib.lib.qaoa_add_edge(g, 2, 2, 1.0) == -1 and lib.lib.qaoa_add_edge(g, 0, 9, 1.0) == -1       # self-loop, range
    sol = (C.c_int * 5)(0, 1, 1, 0, 1)
    assert lib.lib.qaoa_evaluate_solution(g, sol) == 1.0 + 0.5
    assert lib.lib.qaoa_estimate_optimal_p(8, 12) == 4 + int(12 / 28 * 4)
    c = lib.lib.qaoa_default_config(3)
    assert c.p == 3 and c.max_iterations == 1000 and abs(c.learning_rate - 0.01) < 1e-15 and c.num_shots == 1024
    lib.lib.qaoa_destroy_graph(g)
    if HAVE_REF:
        ref = Q.Qaoa(Q.REF)
        assert ref.lib.qaoa_estimate_optimal_p(8, 12) == lib.lib.qaoa_estimate_optimal_p(8, 12)


@pytest.mark.gpu
@pytest.mark.skipif(not HAVE_REF, reason="oracle/_ref/libqaoa_ref.so not shipped")
@pytest.mark.parametrize("n,p,seed", [(4, 1, 1), (8, 2, 2), (12, 3, 3), (16, 2, 4)])
def test_circuit_expectation_gradient_match_reference(n, p, seed):
    edges, vw = problem(n, seed)
    rng = np.random.default_rng(seed)
    gamma, beta = rng.uniform(0, 2 * np.pi, p), rng.uniform(0, np.pi, p)
    out = {}
    for name, path in (("ours", Q.COMPAT), ("ref", Q.REF)):
        lib = Q.Qaoa(path)
        g = lib.graph(n, edges, vw)
        s = lib.init(g, p, gamma, beta)
        assert s
        assert lib.lib.qaoa_apply_circuit(s, Q.dp(gamma), Q.dp(beta)) == 0
        psi = lib.amplitudes(s)
        e = C.c_double(0)
        assert lib.lib.qaoa_compute_expectation(s, C.byref(e)) == 0
        gg, bg = shift_gradient(lib, s, gamma, beta, use_library=(name == "ours"))
        # layer-by-layer application gives the same state as the whole circuit (apply_circuit first: it is what stores
        # the angles in the state object)
        assert lib.lib.qaoa_apply_circuit(s, Q.dp(gamma), Q.dp(beta)) == 0
        assert lib.lib.qaoa_prepare_initial_state(s) == 0
        for l in range(p):
            assert lib.lib.qaoa_apply_layer(s, l) == 0
        psi2 = lib.amplitudes(s)
        out[name] = (psi, e.value, gg, bg, psi2)
        lib.lib.qaoa_destroy(s)
        lib.lib.qaoa_destroy_graph(g)
    o, r = out["ours"], out["ref"]
    assert np.abs(o[0] - r[0]).max() < 1e-5                       # ComplexFloat state
    assert abs(o[1] - r[1]) < 1e-5 * max(1.0, abs(r[1]))
    assert np.abs(o[2] - r[2]).max() < 2e-5 * max(1.0, np.abs(r[2]).max()) and np.abs(o[3] - r[3]).max() < 2e-5 * max(1.0, np.abs(r[3]).max())
    assert np.abs(o[4] - r[4]).max() < 1e-5 and np.abs(o[4] - o[0]).max() < 1e-6


@pytest.mark.gpu
def test_exact_gradient_is_the_oracles(oracle):
    n, p = 10, 2
    edges, vw = problem(n, 7)
    rng = np.random.default_rng(7)
    gamma, beta = rng.uniform(0, 2 * np.pi, p), rng.uniform(0, np.pi, p)
    lib = Q.Qaoa(Q.COMPAT)
    lib.lib.qgt_b200_qaoa_exact_gradient.argtypes = [C.POINTER(Q.State), Q._DP, Q._DP, Q._DP]
    g = lib.graph(n, edges, vw)
    s = lib.init(g, p, gamma, beta)
    e = C.c_double(0)
    gg, bg = np.zeros(p), np.zeros(p)
    assert lib.lib.qgt_b200_qaoa_exact_gradient(s, C.byref(e), Q.dp(gg), Q.dp(bg)) == 0
    c = K.qaoa_maxcut(n, p, edges)
    c.vertex_weights = [float(v) for v in vw]
    th = np.empty(2 * p); th[0::2] = gamma; th[1::2] = beta
    eo, go = oracle.expectation_gradient(c, th)
    assert abs(e.value - eo) < 1e-10 and np.abs(gg - go[0::2]).max() < 1e-10 and np.abs(bg - go[1::2]).max() < 1e-10
    lib.lib.qaoa_destroy(s)
    lib.lib.qaoa_destroy_graph(g)


@pytest.mark.gpu
def test_sampling_follows_the_distribution():
    n, p = 8, 2
    edges, vw = problem(n, 9)
    gamma, beta = np.array([0.4, 0.9]), np.array([0.3, 0.7])
    lib = Q.Qaoa(Q.COMPAT)
    g = lib.graph(n, edges, vw)
    s = lib.init(g, p, gamma, beta)
    assert lib.lib.qaoa_apply_circuit(s, Q.dp(gamma), Q.dp(beta)) == 0
    prob = np.abs(lib.amplitudes(s).astype(np.complex128)) ** 2
    shots = 20000
    arr = (C.POINTER(C.c_int) * shots)()
    assert lib.lib.qaoa_sample(s, arr, shots) == 0
    counts = np.zeros(1 << n)
    for k in range(shots):
        z = sum(arr[k][q] << q for q in range(n))
        counts[z] += 1
    libc = C.CDLL(None)
    for k in range(shots):
        libc.free(arr[k])
    # chi-square-like bound: every frequency within 5 sigma of its probability
    sigma = np.sqrt(prob * (1 - prob) / shots)
    assert np.all(np.abs(counts / shots - prob) < 5 * sigma + 1e-4)
    lib.lib.qaoa_destroy(s)
    lib.lib.qaoa_destroy_graph(g)


@pytest.mark.gpu
@pytest.mark.skipif(not HAVE_REF, reason="oracle/_ref/libqaoa_ref.so not shipped")
def test_optimisation_loop_against_reference_circuits():
    """qaoa_optimize's gradient-ascent loop (qaoa.c:595-628) replayed with the reference library's own circuits and
    expectations (and the documented shift gradient): the cost history matches step by step, and - unlike the reference,
    whose best_cost starts at +INFINITY - the best parameters and the most probable bit string are recorded."""
    n, p, iters, lr = 6, 1, 10, 0.05
    edges, vw = problem(n, 5)
    gamma0, beta0 = np.array([0.5]), np.array([0.4])
    lib = Q.Qaoa(Q.COMPAT)
    g = lib.graph(n, edges, np.zeros(n))
    s = lib.init(g, p, gamma0, beta0, max_iterations=iters, learning_rate=lr, tolerance=1e-12)
    r = lib.lib.qaoa_optimize(s)
    assert r
    ours = np.array([r.contents.cost_history[k] for k in range(r.contents.history_length)])
    best = r.contents.optimal_cost
    sol = (C.c_int * n)(*[r.contents.optimal_solution[q] for q in range(n)])
    og = r.contents.optimal_gamma[0]
    assert r.contents.num_iterations == iters
    # replay with the reference
    ref = Q.Qaoa(Q.REF)
    gr = ref.graph(n, edges, np.zeros(n))
    sr = ref.init(gr, p, gamma0, beta0)
    ga, be = gamma0.copy(), beta0.copy()
    hist, params = [], []
    for _ in range(iters):
        assert ref.lib.qaoa_apply_circuit(sr, Q.dp(ga), Q.dp(be)) == 0
        e = C.c_double(0)
        assert ref.lib.qaoa_compute_expectation(sr, C.byref(e)) == 0
        hist.append(e.value); params.append((ga.copy(), be.copy()))
        gg, bg = shift_gradient(ref, sr, ga, be, use_library=False)
        ga += lr * gg; be += lr * bg
    hist = np.array(hist)
    assert len(ours) == iters and np.abs(ours - hist).max() < 1e-4
    k = int(np.argmax(ours))
    assert abs(best - ours[k]) < 1e-12 and abs(og - params[k][0][0]) < 1e-4
    # the recorded bit string is the most probable one of the best circuit
    assert ref.lib.qaoa_apply_circuit(sr, Q.dp(params[k][0]), Q.dp(params[k][1])) == 0
    prob = np.abs(ref.amplitudes(sr)) ** 2
    z = sum(sol[q] << q for q in range(n))
    assert prob[z] > prob.max() - 1e-6
    lib.lib.qaoa_destroy_result(r); lib.lib.qaoa_destroy(s); lib.lib.qaoa_destroy_graph(g)
    ref.lib.qaoa_destroy(sr); ref.lib.qaoa_destroy_graph(gr)

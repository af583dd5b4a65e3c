"""CPU tests: the oracle restatement is pinned against the unmodified reference code (oracle/_ref),
the committed golden vectors generated from it, the reference tests' known answers, the independent
dense-matrix QGT and the analytic single-qubit closed form."""
import numpy as np
import pytest

from oracle import dense_qgt
from quantum_geometric_tensor_b200 import circuits as K

from helpers import golden_cases, known_answer_cases, load_golden, rel_err


@pytest.mark.parametrize("name", golden_cases())
def test_oracle_matches_golden_vectors(oracle, name):
    c, z = load_golden(name)
    th = z["theta"]
    psi = oracle.apply(c, th)
    n_keep = z["psi"].size
    assert np.array_equal(psi[:n_keep], z["psi"]), "gate sweeps must be bit-exact with the reference"
    assert np.sum(psi * np.arange(1, psi.size + 1)) == z["psi_checksum"][0]
    q = oracle.qgt(c, th)
    assert rel_err(q.real, z["metric"]) < 1e-13
    assert rel_err(-2 * q.imag, z["curvature"]) < 1e-13   # diffgeo returns F = -2 Im Q


@pytest.mark.parametrize("seed", range(5))
def test_oracle_bit_exact_with_reference_simulator(oracle, reference, seed):
    n = 3 + seed
    c = K.random_circuit(n, 50, seed, kinds=sorted(K.REFERENCE_KINDS))
    th = K.default_angles(max(1, c.num_params), seed)
    assert np.array_equal(oracle.apply(c, th), reference.apply(c, th))
    for mu in range(min(3, c.num_params)):
        assert np.array_equal(oracle.derivative(c, th, mu), reference.derivative(c, th, mu))


def test_oracle_assembly_matches_reference_diffgeo(oracle, reference):
    rng = np.random.default_rng(3)
    dim, P = 64, 7
    psi = rng.normal(size=dim) + 1j * rng.normal(size=dim)
    psi /= np.linalg.norm(psi)
    J = rng.normal(size=(P, dim)) + 1j * rng.normal(size=(P, dim))
    q = oracle.qgt_from_columns(psi, J)
    g, f = reference.fubini_berry(psi, J)
    assert np.array_equal(q.real, g)
    assert np.array_equal(-2.0 * q.imag, f)


@pytest.mark.parametrize("label,circ,expected", known_answer_cases(), ids=[c[0] for c in known_answer_cases()])
def test_oracle_reference_known_answers(oracle, label, circ, expected):
    psi = oracle.apply(circ, np.zeros(1))
    assert np.abs(np.abs(psi) - np.array(expected)).max() < 1e-10


@pytest.mark.parametrize("seed", range(4))
def test_oracle_vs_independent_dense_qgt(oracle, seed):
    kinds = [K.X, K.Y, K.Z, K.H, K.S, K.T, K.SDG, K.TDG, K.SX, K.RX, K.RY, K.RZ, K.PHASE, K.CNOT, K.CY, K.CZ, K.CH,
             K.SWAP, K.CRX, K.CRY, K.CRZ, K.ZZ]
    c = K.random_circuit(3 + seed, 35, 10 + seed, kinds=kinds, share_params=True)
    th = K.default_angles(max(1, c.num_params), seed)
    assert np.abs(oracle.qgt(c, th) - dense_qgt.qgt(c, th)).max() < 1e-12
    psi, _ = dense_qgt.state_and_jacobian(c, th)
    assert np.abs(oracle.apply(c, th) - psi).max() < 1e-13


def test_oracle_qaoa_vs_dense(oracle):
    c = K.qaoa_maxcut(6, 2)
    c.vertex_weights = [0.1, 0.2, -0.3, 0.0, 0.4, -0.1]
    th = K.default_angles(c.num_params, 5)
    assert np.abs(oracle.qgt(c, th) - dense_qgt.qgt(c, th)).max() < 1e-12


def test_analytic_single_qubit_closed_form(oracle):
    # RY(t1) then RZ(t2) on |0>: Q = [[1/4, (i/4) sin t1], [-(i/4) sin t1, sin^2 t1 / 4]]  
    t1, t2 = 0.7, 1.3
    c = K.Circuit(1)
    c.rot(K.RY, 0, 0)
    c.rot(K.RZ, 0, 1)
    q = oracle.qgt(c, np.array([t1, t2]))
    expect = np.array([[0.25, 0.25j * np.sin(t1)], [-0.25j * np.sin(t1), 0.25 * np.sin(t1) ** 2]])
    assert np.abs(q - expect).max() < 1e-15
    assert abs(q[0, 1].imag - 0.16105442) < 1e-8 and abs(q[1, 1].real - 0.10375411) < 1e-8


def test_product_state_metric_is_block_diagonal(oracle):
    # HEA without entanglers: parameters on different qubits are uncorrelated
    n = 4
    c = K.Circuit(n)
    for q in range(n):
        c.rot(K.RY, q, 2 * q)
        c.rot(K.RZ, q, 2 * q + 1)
    q_ = oracle.qgt(c, K.default_angles(2 * n))
    for a in range(2 * n):
        for b in range(2 * n):
            if a // 2 != b // 2:
                assert abs(q_[a, b]) < 1e-15


def test_derivative_matches_central_differences(oracle):
    c = K.hea_layers(4, 2)
    th = K.default_angles(c.num_params)
    h = 1e-5
    for mu in (0, 5, 11):
        tp, tm = th.copy(), th.copy()
        tp[mu] += h
        tm[mu] -= h
        fd = (oracle.apply(c, tp) - oracle.apply(c, tm)) / (2 * h)
        assert np.abs(fd - oracle.derivative(c, th, mu)).max() < 1e-9


def test_natural_gradient_restatement(oracle):
    rng = np.random.default_rng(0)
    A = rng.normal(size=(12, 12))
    G = A @ A.T / 12
    g = rng.normal(size=12)
    x, lam = oracle.natural_gradient(G, g)
    assert lam == 1e-4
    assert np.abs((G + lam * np.eye(12)) @ x - g).max() < 1e-10
    # ill-conditioned: adaptive lambda = 1e-6 sqrt(kappa)  (gradient.c:2898-2912)
    w = np.logspace(0, -12, 12)
    G2 = (A * 0 + np.linalg.qr(A)[0]) @ np.diag(w) @ np.linalg.qr(A)[0].T
    x2, lam2 = oracle.natural_gradient(G2, g)
    kappa = w.max() / w.min()
    assert abs(lam2 - 1e-6 * np.sqrt(kappa)) / lam2 < 1e-3

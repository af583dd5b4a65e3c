"""The ComplexFloat circuit path (quantum_circuit_execute, reference core/quantum_circuit_operations.c:1139-1197).
CPU: the committed goldens equal what the unmodified reference produces now, and an independent NumPy evaluation with the
path's phase conventions reproduces them.  GPU: the compat library's quantum_circuit_execute against the goldens and
against the live reference library, to the 1e-5 the ComplexFloat APIs are compared at."""
import glob
import os

import numpy as np
import pytest

import pathb

HERE = os.path.dirname(os.path.abspath(__file__))
GOLD = sorted(glob.glob(os.path.join(HERE, "golden_pathb", "*.npz")))
TOL = 1e-5


def dense(n, table):
    """|psi> by plain NumPy in complex128 with the conventions of this path: X = RX(pi), Y = RY(pi), Z = RZ(pi), S = RZ(pi/2)."""
    psi = np.zeros(1 << n, complex); psi[0] = 1

    def one(q, m):
        nonlocal psi
        t = psi.reshape(1 << (n - 1 - q), 2, 1 << q)
        psi = np.einsum("ab,xby->xay", m, t).reshape(-1)

    def rot(axis, th):
        c, s = np.cos(th / 2), np.sin(th / 2)
        return {0: np.array([[c, -1j * s], [-1j * s, c]]), 1: np.array([[c, -s], [s, c]], complex),
                2: np.array([[c - 1j * s, 0], [0, c + 1j * s]])}[axis]

    idx = np.arange(1 << n)
    for op, a, b, ang in table:
        op, a, b = int(op), int(a), int(b)
        if op == pathb.H: one(a, np.array([[1, 1], [1, -1]], complex) / np.sqrt(2))
        elif op in (pathb.X, pathb.Y, pathb.Z): one(a, rot(op - pathb.X, np.pi))
        elif op == pathb.PHASE: one(a, rot(2, ang))
        elif op in (pathb.RX, pathb.RY, pathb.RZ): one(a, rot(op - pathb.RX, ang))
        elif op == pathb.CNOT: psi = psi[np.where(idx >> a & 1, idx ^ (1 << b), idx)]
        elif op == pathb.CZ: psi = np.where((idx >> a & 1) & (idx >> b & 1), -psi, psi)
        else:
            sw = idx & ~((1 << a) | (1 << b)) | ((idx >> a & 1) << b) | ((idx >> b & 1) << a)
            psi = psi[sw]
    return psi


def test_goldens_present():
    assert len(GOLD) == 16


@pytest.mark.parametrize("path", GOLD, ids=[os.path.basename(p)[:-4] for p in GOLD])
def test_golden_matches_independent_numpy(path):
    z = np.load(path)
    ref = dense(int(z["n"]), z["table"])
    assert np.abs(ref - z["psi"]).max() < TOL          # float accumulation of up to 200 gates stays well inside 1e-5


@pytest.mark.skipif(not os.path.exists(pathb.REFB), reason="oracle/_ref/libqgt_refb.so not built")
@pytest.mark.parametrize("path", GOLD, ids=[os.path.basename(p)[:-4] for p in GOLD])
def test_golden_is_what_the_reference_produces(path):
    z = np.load(path)
    rc, psi, stats = pathb.PathB(pathb.REFB).run(int(z["n"]), z["table"])
    assert rc == 0
    assert np.array_equal(psi, z["psi"])                # bit-exact: same code, same machine arithmetic
    assert stats[0] == int(z["depth"]) and stats[1] == int(z["gate_count"])


def test_compat_builders_and_bookkeeping_without_gpu():
    """create / builders / depth / gate_count / validate are host code: same answers as the reference's on any machine."""
    lib = pathb.PathB(pathb.COMPAT)
    for path in GOLD:
        z = np.load(path)
        c = lib.build(int(z["n"]), z["table"])
        assert lib.lib.quantum_circuit_depth(c) == int(z["depth"])
        assert lib.lib.quantum_circuit_gate_count(c) == int(z["gate_count"])
        assert lib.lib.quantum_circuit_validate(c) == 0
        lib.lib.quantum_circuit_destroy(c)
    c = lib.lib.quantum_circuit_create(3)
    assert lib.lib.quantum_circuit_hadamard(c, 3) == -1          # QGT_ERROR_INVALID_ARGUMENT: qubit out of range
    assert lib.lib.quantum_circuit_rotation(c, 0, 0.3, 0) == -1  # PAULI_I is not a rotation axis (QGT_ERROR_INVALID_PARAMETER)
    assert lib.lib.quantum_circuit_cnot(c, 0, 5) == -1
    assert lib.lib.quantum_circuit_create(0) is None
    s = lib.lib.init_quantum_state(2)
    assert lib.lib.quantum_circuit_execute(c, s) == -19          # QGT_ERROR_INCOMPATIBLE: 3-qubit circuit, 2-qubit state
    assert lib.lib.quantum_circuit_execute(None, s) == -1
    amps = np.ctypeslib.as_array(s.contents.amplitudes, shape=(8,))
    assert amps[0] == 1.0 and not amps[1:].any() and s.contents.dimension == 4 and s.contents.is_normalized
    lib.lib.quantum_state_cleanup(s)
    lib.lib.quantum_circuit_destroy(c)


@pytest.mark.gpu
@pytest.mark.parametrize("path", GOLD, ids=[os.path.basename(p)[:-4] for p in GOLD])
def test_execute_matches_golden(path):
    z = np.load(path)
    rc, psi, stats = pathb.PathB(pathb.COMPAT).run(int(z["n"]), z["table"])
    assert rc == 0
    assert np.abs(psi - z["psi"]).max() < TOL
    assert stats[0] == int(z["depth"]) and stats[1] == int(z["gate_count"]) and stats[2] == 0


@pytest.mark.gpu
@pytest.mark.skipif(not os.path.exists(pathb.REFB), reason="oracle/_ref/libqgt_refb.so not shipped")
@pytest.mark.parametrize("n,ngates,seed", [(2, 25, 1), (5, 60, 2), (9, 120, 3), (12, 160, 4), (16, 220, 5), (18, 120, 6)])
def test_execute_matches_live_reference(n, ngates, seed):
    table = pathb.random_table(n, ngates, 100 + seed)
    rng = np.random.default_rng(seed)
    init = (rng.normal(size=1 << n) + 1j * rng.normal(size=1 << n)).astype(np.complex64)
    init /= np.linalg.norm(init)                            # a caller-supplied register, not |0...0>
    rc1, ours, _ = pathb.PathB(pathb.COMPAT).run(n, table, init)
    rc2, ref, _ = pathb.PathB(pathb.REFB).run(n, table, init)
    assert rc1 == 0 and rc2 == 0
    assert np.abs(ours - ref).max() < TOL


@pytest.mark.gpu
def test_measure_all_collapses_to_one_basis_state():
    lib = pathb.PathB(pathb.COMPAT)
    n = 5
    table = np.array([(pathb.H, q, 0, 0.0) for q in range(n)] + [(pathb.CNOT, 0, 3, 0.0)], dtype=np.float64)
    c = lib.build(n, table)
    s = lib.lib.init_quantum_state(n)
    assert lib.lib.quantum_circuit_execute(c, s) == 0
    import ctypes as C
    res = (C.c_size_t * n)()
    assert lib.lib.quantum_circuit_measure_all(c, s, res) == 0
    amps = np.ctypeslib.as_array(s.contents.amplitudes, shape=(2 << n,)).copy().view(np.complex64)
    k = sum(int(res[q]) << q for q in range(n))
    assert abs(abs(amps[k]) - 1.0) < 1e-5 and np.abs(np.delete(amps, k)).max() < 1e-6
    lib.lib.quantum_state_cleanup(s)
    lib.lib.quantum_circuit_destroy(c)

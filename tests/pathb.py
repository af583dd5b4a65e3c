"""ctypes view of the ComplexFloat circuit API (core/quantum_circuit_operations.h of the reference).  The same wrapper
drives the compat library (libqgt_b200_compat.so, the product) and oracle/_ref/libqgt_refb.so (the unmodified reference,
test infrastructure), so a test reads the same on both."""
import ctypes as C
import os

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
COMPAT = os.path.join(ROOT, "quantum_geometric_tensor_b200", "libqgt_b200_compat.so")
REFB = os.path.join(ROOT, "oracle", "_ref", "libqgt_refb.so")

# op codes of the circuit tables in tests/golden_pathb/*.npz: (op, qubit a, qubit b, angle)
H, X, Y, Z, PHASE, RX, RY, RZ, CNOT, CZ, SWAP = range(11)
NAMES = ["h", "x", "y", "z", "phase", "rx", "ry", "rz", "cnot", "cz", "swap"]


class QuantumState(C.Structure):
    _fields_ = [("num_qubits", C.c_size_t), ("amplitudes", C.POINTER(C.c_float)), ("workspace", C.c_void_p),
                ("dimension", C.c_size_t), ("is_normalized", C.c_bool)]


class PathB:
    def __init__(self, path):
        self.lib = L = C.CDLL(path, mode=os.RTLD_LOCAL if hasattr(os, "RTLD_LOCAL") else 0)
        L.init_quantum_state.restype = C.POINTER(QuantumState)
        L.init_quantum_state.argtypes = [C.c_size_t]
        L.quantum_state_cleanup.argtypes = [C.POINTER(QuantumState)]
        L.quantum_state_reset.argtypes = [C.POINTER(QuantumState)]
        L.quantum_circuit_create.restype = C.c_void_p
        L.quantum_circuit_create.argtypes = [C.c_size_t]
        L.quantum_circuit_destroy.argtypes = [C.c_void_p]
        for f in ("hadamard", "pauli_x", "pauli_y", "pauli_z"):
            getattr(L, "quantum_circuit_" + f).argtypes = [C.c_void_p, C.c_size_t]
        L.quantum_circuit_phase.argtypes = [C.c_void_p, C.c_size_t, C.c_double]
        L.quantum_circuit_rotation.argtypes = [C.c_void_p, C.c_size_t, C.c_double, C.c_int]
        for f in ("cnot", "cz", "swap"):
            getattr(L, "quantum_circuit_" + f).argtypes = [C.c_void_p, C.c_size_t, C.c_size_t]
        L.quantum_circuit_execute.argtypes = [C.c_void_p, C.POINTER(QuantumState)]
        L.quantum_circuit_measure_all.argtypes = [C.c_void_p, C.POINTER(QuantumState), C.POINTER(C.c_size_t)]
        L.quantum_circuit_depth.restype = C.c_size_t
        L.quantum_circuit_depth.argtypes = [C.c_void_p]
        L.quantum_circuit_gate_count.restype = C.c_size_t
        L.quantum_circuit_gate_count.argtypes = [C.c_void_p]
        L.quantum_circuit_validate.argtypes = [C.c_void_p]

    def build(self, n, table):
        L = self.lib
        c = L.quantum_circuit_create(n)
        assert c
        for op, a, b, ang in table:
            op, a, b = int(op), int(a), int(b)
            if op == H: rc = L.quantum_circuit_hadamard(c, a)
            elif op == X: rc = L.quantum_circuit_pauli_x(c, a)
            elif op == Y: rc = L.quantum_circuit_pauli_y(c, a)
            elif op == Z: rc = L.quantum_circuit_pauli_z(c, a)
            elif op == PHASE: rc = L.quantum_circuit_phase(c, a, float(ang))
            elif op in (RX, RY, RZ): rc = L.quantum_circuit_rotation(c, a, float(ang), 1 + op - RX)
            elif op == CNOT: rc = L.quantum_circuit_cnot(c, a, b)
            elif op == CZ: rc = L.quantum_circuit_cz(c, a, b)
            else: rc = L.quantum_circuit_swap(c, a, b)
            assert rc == 0, (NAMES[op], a, b, rc)
        return c

    def run(self, n, table, init=None):
        """Final ComplexFloat amplitudes (complex64 array) of the circuit on |0..0> (or on `init`)."""
        L = self.lib
        c = self.build(n, table)
        s = L.init_quantum_state(n)
        assert s
        dim = 1 << n
        buf = np.ctypeslib.as_array(s.contents.amplitudes, shape=(2 * dim,))
        if init is not None:
            buf[:] = np.asarray(init, dtype=np.complex64).view(np.float32)
        rc = L.quantum_circuit_execute(c, s)
        out = buf.copy().view(np.complex64)
        stats = (L.quantum_circuit_depth(c), L.quantum_circuit_gate_count(c), L.quantum_circuit_validate(c))
        L.quantum_state_cleanup(s)
        L.quantum_circuit_destroy(c)
        return rc, out, stats


def random_table(n, ngates, seed):
    """Gates of the path's switch (quantum_circuit_operations.c:1147-1190); phase gates carry pi/2, the one angle on which
    the reference's recorded-but-ignored parameter and RZ(angle) agree."""
    rng = np.random.default_rng(seed)
    rows = []
    for _ in range(ngates):
        op = int(rng.integers(0, 11 if n > 1 else 8))
        a = int(rng.integers(0, n))
        b = 0
        if op >= CNOT:
            b = int(rng.integers(0, n - 1))
            b += b >= a
        ang = float(rng.uniform(-np.pi, np.pi)) if op in (RX, RY, RZ) else (np.pi / 2 if op == PHASE else 0.0)
        rows.append((op, a, b, ang))
    return np.array(rows, dtype=np.float64)

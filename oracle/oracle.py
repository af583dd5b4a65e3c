"""ctypes loaders for the CPU checkers.   TEST INFRASTRUCTURE ONLY.

``Oracle``     -> oracle/_build/libqgt_oracle.so   the C restatement (oracle/qgt_oracle.c), kind "port"
``Reference``  -> oracle/_ref/libqgt_ref.so        the unmodified reference sources + ref_driver.c, kind "reference"

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs import this.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from typing import Optional, Tuple

import numpy as np

from quantum_geometric_tensor_b200.circuits import CCircuit, Circuit

_HERE = os.path.dirname(os.path.abspath(__file__))
_DP = C.POINTER(C.c_double)


def _ptr(a: Optional[np.ndarray]):
    return None if a is None else a.ctypes.data_as(_DP)


def build(quiet: bool = True) -> None:
    """(Re)build the restatement and, when /root/reference is present, the reference library."""
    subprocess.run(["make", "-C", _HERE], check=True,
                   stdout=subprocess.DEVNULL if quiet else None)


class NatGradCfg(C.Structure):
    _fields_ = [("regularization", C.c_double), ("condition_threshold", C.c_double),
                ("adaptive", C.c_int), ("pseudoinverse_fallback", C.c_int), ("singular_cutoff", C.c_double)]

    @classmethod
    def default(cls) -> "NatGradCfg":   # get_default_natural_gradient_config, gradient.c:2721
        return cls(1e-4, 1e8, 1, 1, 1e-10)


class Oracle:
    def __init__(self) -> None:
        path = os.path.join(_HERE, "_build", "libqgt_oracle.so")
        if not os.path.exists(path):
            build()
        self.lib = C.CDLL(path)
        L = self.lib
        L.orc_init_state.argtypes = [_DP, C.c_int, C.c_int]
        L.orc_apply_circuit.argtypes = [_DP, C.POINTER(CCircuit), _DP]
        L.orc_derivative.argtypes = [_DP, C.POINTER(CCircuit), _DP, C.c_int]
        L.orc_qgt_from_columns.argtypes = [_DP, _DP, C.c_size_t, C.c_size_t, _DP, _DP, _DP]
        L.orc_qgt.argtypes = [C.POINTER(CCircuit), _DP, _DP, _DP, _DP]
        L.orc_natural_gradient.argtypes = [_DP, _DP, C.c_size_t, C.POINTER(NatGradCfg), _DP, _DP]
        L.orc_expectation_gradient.argtypes = [C.POINTER(CCircuit), _DP, _DP, _DP]

    def init_state(self, n: int, initial_state: int = 0) -> np.ndarray:
        a = np.zeros(1 << n, dtype=np.complex128)
        self.lib.orc_init_state(_ptr(a), n, initial_state)
        return a

    def apply(self, circ: Circuit, theta: np.ndarray, state: Optional[np.ndarray] = None) -> np.ndarray:
        a = self.init_state(circ.num_qubits, circ.initial_state) if state is None else np.array(state, dtype=np.complex128)
        cc = circ.to_c()
        th = np.ascontiguousarray(theta, dtype=np.float64)
        rc = self.lib.orc_apply_circuit(_ptr(a), C.byref(cc), _ptr(th))
        if rc:
            raise RuntimeError(f"orc_apply_circuit -> {rc}")
        return a

    def derivative(self, circ: Circuit, theta: np.ndarray, mu: int) -> np.ndarray:
        out = np.zeros(1 << circ.num_qubits, dtype=np.complex128)
        cc = circ.to_c()
        th = np.ascontiguousarray(theta, dtype=np.float64)
        rc = self.lib.orc_derivative(_ptr(out), C.byref(cc), _ptr(th), mu)
        if rc:
            raise RuntimeError(f"orc_derivative -> {rc}")
        return out

    def qgt_from_columns(self, psi: np.ndarray, dpsi: np.ndarray) -> np.ndarray:
        P, dim = dpsi.shape
        q = np.zeros((P, P), dtype=np.complex128)
        psi = np.ascontiguousarray(psi, dtype=np.complex128)
        dpsi = np.ascontiguousarray(dpsi, dtype=np.complex128)
        self.lib.orc_qgt_from_columns(_ptr(psi), _ptr(dpsi), dim, P, None, None, _ptr(q))
        return q

    def qgt(self, circ: Circuit, theta: np.ndarray) -> np.ndarray:
        P = circ.num_params
        q = np.zeros((P, P), dtype=np.complex128)
        cc = circ.to_c()
        th = np.ascontiguousarray(theta, dtype=np.float64)
        rc = self.lib.orc_qgt(C.byref(cc), _ptr(th), None, None, _ptr(q))
        if rc:
            raise RuntimeError(f"orc_qgt -> {rc}")
        return q

    def natural_gradient(self, metric: np.ndarray, grad: np.ndarray, cfg: Optional[NatGradCfg] = None) -> Tuple[np.ndarray, float]:
        cfg = cfg or NatGradCfg.default()
        m = np.ascontiguousarray(metric, dtype=np.float64)
        g = np.ascontiguousarray(grad, dtype=np.float64)
        out = np.zeros_like(g)
        lam = C.c_double(0)
        rc = self.lib.orc_natural_gradient(_ptr(m), _ptr(g), g.size, C.byref(cfg), _ptr(out), C.byref(lam))
        if rc:
            raise RuntimeError(f"orc_natural_gradient -> {rc}")
        return out, lam.value

    @staticmethod
    def energy(circ: Circuit, psi: np.ndarray) -> float:
        """<psi|H|psi> for the diagonal observable of the circuit's edge list / vertex weights."""
        return float(((psi.real ** 2 + psi.imag ** 2) * cost_energy_table(circ)).sum())

    def expectation_gradient(self, circ: Circuit, theta: np.ndarray) -> Tuple[float, np.ndarray]:
        cc = circ.to_c()
        th = np.ascontiguousarray(theta, dtype=np.float64)
        e = C.c_double(0)
        g = np.zeros(circ.num_params)
        rc = self.lib.orc_expectation_gradient(C.byref(cc), _ptr(th), C.byref(e), _ptr(g))
        if rc:
            raise RuntimeError(f"orc_expectation_gradient -> {rc}")
        return e.value, g


def cost_energy_table(circ: Circuit) -> np.ndarray:
    """E_z of every basis state: cut weight + vertex terms (reference algorithms/qaoa.c:258-289), plain NumPy."""
    idx = np.arange(1 << circ.num_qubits, dtype=np.uint64)
    e = np.zeros(idx.size)
    for i, j, w in circ.edges:
        e += w * (((idx >> np.uint64(i)) ^ (idx >> np.uint64(j))) & np.uint64(1)).astype(np.float64)
    if circ.vertex_weights is not None:
        for q, v in enumerate(circ.vertex_weights):
            e += v * (1.0 - 2.0 * ((idx >> np.uint64(q)) & np.uint64(1)).astype(np.float64))
    return e


class Reference:
    """The unmodified reference code (sim_* and diffgeo_*), if oracle/_ref was built."""

    @staticmethod
    def available() -> bool:
        return os.path.exists(os.path.join(_HERE, "_ref", "libqgt_ref.so"))

    def __init__(self) -> None:
        self.lib = C.CDLL(os.path.join(_HERE, "_ref", "libqgt_ref.so"))
        L = self.lib
        L.ref_apply_circuit.argtypes = [_DP, C.POINTER(CCircuit), _DP]
        L.ref_derivative.argtypes = [_DP, C.POINTER(CCircuit), _DP, C.c_int]
        L.ref_fubini_berry.argtypes = [_DP, _DP, C.c_size_t, C.c_size_t, _DP, _DP]
        L.ref_qgt.argtypes = [C.POINTER(CCircuit), _DP, C.c_size_t, _DP, _DP]

    def apply(self, circ: Circuit, theta: np.ndarray, state: Optional[np.ndarray] = None) -> np.ndarray:
        n = circ.num_qubits
        if state is None:
            a = np.zeros(1 << n, dtype=np.complex128)
            a[0] = 1.0
        else:
            a = np.array(state, dtype=np.complex128)
        cc = circ.to_c()
        th = np.ascontiguousarray(theta, dtype=np.float64)
        rc = self.lib.ref_apply_circuit(_ptr(a), C.byref(cc), _ptr(th))
        if rc:
            raise RuntimeError(f"ref_apply_circuit -> {rc}")
        return a

    def derivative(self, circ: Circuit, theta: np.ndarray, mu: int) -> np.ndarray:
        out = np.zeros(1 << circ.num_qubits, dtype=np.complex128)
        cc = circ.to_c()
        th = np.ascontiguousarray(theta, dtype=np.float64)
        rc = self.lib.ref_derivative(_ptr(out), C.byref(cc), _ptr(th), mu)
        if rc:
            raise RuntimeError(f"ref_derivative -> {rc}")
        return out

    def fubini_berry(self, psi: np.ndarray, dpsi: np.ndarray) -> Tuple[np.ndarray, np.ndarray]:
        """(g, F) with F = -2 Im Q, exactly as diffgeo returns them."""
        P, dim = dpsi.shape
        g = np.zeros((P, P))
        f = np.zeros((P, P))
        psi = np.ascontiguousarray(psi, dtype=np.complex128)
        dpsi = np.ascontiguousarray(dpsi, dtype=np.complex128)
        rc = self.lib.ref_fubini_berry(_ptr(psi), _ptr(dpsi), dim, P, _ptr(g), _ptr(f))
        if rc:
            raise RuntimeError(f"ref_fubini_berry -> {rc}")
        return g, f

    def qgt(self, circ: Circuit, theta: np.ndarray, max_cols: int = 0) -> Tuple[np.ndarray, np.ndarray]:
        P = circ.num_params if not max_cols else min(max_cols, circ.num_params)
        g = np.zeros((P, P))
        f = np.zeros((P, P))
        cc = circ.to_c()
        th = np.ascontiguousarray(theta, dtype=np.float64)
        rc = self.lib.ref_qgt(C.byref(cc), _ptr(th), max_cols, _ptr(g), _ptr(f))
        if rc:
            raise RuntimeError(f"ref_qgt -> {rc}")
        return g, f

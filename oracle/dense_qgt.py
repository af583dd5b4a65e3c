"""Independent dense-matrix QGT (n <= ~10).   TEST INFRASTRUCTURE ONLY.

Shares no code with oracle/qgt_oracle.c or the reference: every gate is built as a full
2^n x 2^n matrix with Kronecker products, the derivative is the analytic product rule
d_mu U = sum_k U_{>k} (dU_k/dtheta_mu) U_{<k}, and Q is evaluated from its definition
Q = <d_mu psi|d_nu psi> - <d_mu psi|psi><psi|d_nu psi>   (north star's "independent dense-matrix QGT").
"""
from __future__ import annotations

import numpy as np

from quantum_geometric_tensor_b200 import circuits as K

_I2 = np.eye(2, dtype=complex)
_X = np.array([[0, 1], [1, 0]], dtype=complex)
_Y = np.array([[0, -1j], [1j, 0]], dtype=complex)
_Z = np.array([[1, 0], [0, -1]], dtype=complex)
_H = np.array([[1, 1], [1, -1]], dtype=complex) / np.sqrt(2.0)
_P0 = np.array([[1, 0], [0, 0]], dtype=complex)
_P1 = np.array([[0, 0], [0, 1]], dtype=complex)


def _embed(n: int, ops: dict) -> np.ndarray:
    """kron over qubits n-1 .. 0 (qubit 0 = least significant index bit)."""
    m = np.array([[1.0 + 0j]])
    for q in range(n - 1, -1, -1):
        m = np.kron(m, ops.get(q, _I2))
    return m


def _one(kind: int, a: float):
    """(U, dU/da) for single-qubit kinds."""
    c, s = np.cos(a / 2), np.sin(a / 2)
    if kind == K.I: return _I2, None
    if kind == K.X: return _X, None
    if kind == K.Y: return _Y, None
    if kind == K.Z: return _Z, None
    if kind == K.H: return _H, None
    if kind == K.S: return np.diag([1, 1j]), None
    if kind == K.T: return np.diag([1, np.exp(1j * np.pi / 4)]), None
    if kind == K.SDG: return np.diag([1, -1j]), None
    if kind == K.TDG: return np.diag([1, np.exp(-1j * np.pi / 4)]), None
    if kind == K.SX: return 0.5 * np.array([[1 + 1j, 1 - 1j], [1 - 1j, 1 + 1j]]), None
    if kind in (K.RX, K.CRX):
        return (np.array([[c, -1j * s], [-1j * s, c]]), 0.5 * np.array([[-s, -1j * c], [-1j * c, -s]]))
    if kind in (K.RY, K.CRY):
        return (np.array([[c, -s], [s, c]], dtype=complex), 0.5 * np.array([[-s, -c], [c, -s]], dtype=complex))
    if kind in (K.RZ, K.CRZ):
        return (np.diag([np.exp(-0.5j * a), np.exp(0.5j * a)]), np.diag([-0.5j * np.exp(-0.5j * a), 0.5j * np.exp(0.5j * a)]))
    if kind in (K.U1, K.PHASE):
        return np.diag([1, np.exp(1j * a)]), np.diag([0, 1j * np.exp(1j * a)])
    raise ValueError(kind)


def cost_diagonal(circ: K.Circuit) -> np.ndarray:
    n = circ.num_qubits
    z = np.arange(1 << n)
    e = np.zeros(1 << n)
    for (i, j, w) in circ.edges:
        e += w * (((z >> i) ^ (z >> j)) & 1)
    if circ.vertex_weights is not None:
        for q in range(n):
            e += circ.vertex_weights[q] * (1 - 2 * ((z >> q) & 1))
    return e


def gate_matrices(circ: K.Circuit, gate, theta):
    """(U, dU/dtheta_param or None) as dense 2^n x 2^n."""
    n = circ.num_qubits
    kind, t, ctl, p, angle, scale = gate
    a = scale * theta[p] + angle if p >= 0 else angle
    if kind in (K.CNOT, K.CY, K.CZ, K.CH, K.CRX, K.CRY, K.CRZ):
        base = {K.CNOT: K.X, K.CY: K.Y, K.CZ: K.Z, K.CH: K.H}.get(kind, kind)
        u, du = _one(base, a)
        U = _embed(n, {ctl: _P0}) + _embed(n, {ctl: _P1, t: u})
        dU = _embed(n, {ctl: _P1, t: du}) * scale if (du is not None and p >= 0) else None
        return U, dU
    if kind == K.SWAP:
        cn = lambda c_, t_: _embed(n, {c_: _P0}) + _embed(n, {c_: _P1, t_: _X})
        return cn(ctl, t) @ cn(t, ctl) @ cn(ctl, t), None
    if kind == K.ZZ:
        zz = np.real(np.diag(_embed(n, {t: _Z, ctl: _Z})))
        U = np.diag(np.exp(-0.5j * a * zz))
        return U, (np.diag(-0.5j * zz * scale) @ U if p >= 0 else None)
    if kind == K.COST:
        e = cost_diagonal(circ)
        U = np.diag(np.exp(-1j * a * e))
        return U, (np.diag(-1j * e * scale) @ U if p >= 0 else None)
    u, du = _one(kind, a)
    U = _embed(n, {t: u})
    dU = _embed(n, {t: du}) * scale if (du is not None and p >= 0) else None
    return U, dU


def initial_state(circ: K.Circuit) -> np.ndarray:
    dim = 1 << circ.num_qubits
    if circ.initial_state == K.INIT_PLUS:
        return np.full(dim, 1.0 / np.sqrt(dim), dtype=complex)
    v = np.zeros(dim, dtype=complex)
    v[0] = 1.0
    return v


def state_and_jacobian(circ: K.Circuit, theta):
    """psi and J[mu] = d psi / d theta_mu, by the product rule on dense matrices."""
    theta = np.asarray(theta, dtype=float)
    mats = [gate_matrices(circ, g, theta) for g in circ.gates]
    psi = initial_state(circ)
    prefixes = [psi]
    for U, _ in mats:
        prefixes.append(U @ prefixes[-1])
    psi = prefixes[-1]
    J = np.zeros((circ.num_params, psi.size), dtype=complex)
    for k, (U, dU) in enumerate(mats):
        if dU is None:
            continue
        v = dU @ prefixes[k]
        for U2, _ in mats[k + 1:]:
            v = U2 @ v
        J[circ.gates[k][3]] += v
    return psi, J


def qgt(circ: K.Circuit, theta) -> np.ndarray:
    psi, J = state_and_jacobian(circ, theta)
    gram = J.conj() @ J.T
    v = J.conj() @ psi          # <d_mu psi | psi>
    return gram - np.outer(v, v.conj())

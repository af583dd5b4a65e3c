"""Circuit descriptions shared by the CUDA library, the oracle and the tests.

ctypes mirrors of the POD types in ``include/qgt_b200.h`` plus the synthetic workload
generators of SURVEY.md §8(d):

* ``hea(n, P)``          hardware-efficient ansatz: layers of [RY on every qubit, RZ on every
                          qubit, CNOT ladder q -> q+1], truncated after the P-th rotation
                          (the RY-RZ-CNOT pattern of core/quantum_geometric_interface.c:579-617
                          written as a gate list for sim_add_gate, hardware/quantum_simulator.c:442).
* ``qaoa_maxcut(n, p)``  QAOA for MaxCut on a random 3-regular graph: |+>^n, then p layers of
                          exp(-i gamma E_z) and prod_q RX(2 beta) (algorithms/qaoa.c:344-449);
                          parameters ordered (gamma_1, beta_1, ..., gamma_p, beta_p) — our choice,
                          the reference keeps two arrays (qaoa.c:167-168).
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass, field
from typing import List, Optional, Sequence, Tuple

import numpy as np

# gate kinds (values follow gate_type_t, core/quantum_base_types.h:32-72)
I, X, Y, Z, H, S, T = 0, 1, 2, 3, 4, 5, 6
RX, RY, RZ = 7, 8, 9
CNOT, CY, CZ, SWAP = 10, 11, 12, 13
U1, PHASE = 15, 19
CRX, CRY, CRZ, CH = 22, 23, 24, 25
SDG, TDG, SX = 26, 27, 29
ZZ = 35
COST = 100

INIT_ZERO, INIT_PLUS = 0, 1

PARAMETRIC = {RX, RY, RZ, U1, PHASE, CRX, CRY, CRZ, ZZ, COST}
TWO_QUBIT = {CNOT, CY, CZ, SWAP, CRX, CRY, CRZ, CH, ZZ}
REFERENCE_KINDS = {I, X, Y, Z, H, S, T, RX, RY, RZ, CNOT, CZ, SWAP}  # quantum_simulator.c:188-283

SEED_ANGLES = 20240611
SEED_GRAPH = 1234


class CGate(C.Structure):
    _fields_ = [("kind", C.c_int32), ("target", C.c_int32), ("control", C.c_int32),
                ("param", C.c_int32), ("angle", C.c_double), ("scale", C.c_double)]


class CEdge(C.Structure):
    _fields_ = [("i", C.c_int32), ("j", C.c_int32), ("weight", C.c_double)]


class CCircuit(C.Structure):
    _fields_ = [("num_qubits", C.c_int32), ("num_params", C.c_int32),
                ("gates", C.POINTER(CGate)), ("num_gates", C.c_size_t),
                ("edges", C.POINTER(CEdge)), ("num_edges", C.c_size_t),
                ("vertex_weights", C.POINTER(C.c_double)), ("initial_state", C.c_int32)]


@dataclass
class Circuit:
    """A parameterised circuit; ``gates`` are tuples (kind, target, control, param, angle, scale)."""
    num_qubits: int
    gates: List[Tuple[int, int, int, int, float, float]] = field(default_factory=list)
    num_params: int = 0
    edges: List[Tuple[int, int, float]] = field(default_factory=list)
    vertex_weights: Optional[Sequence[float]] = None
    initial_state: int = INIT_ZERO
    name: str = ""

    # -- building -----------------------------------------------------------------------------
    def add(self, kind: int, target: int, control: int = -1, param: int = -1,
            angle: float = 0.0, scale: float = 1.0) -> "Circuit":
        self.gates.append((kind, target, control, param, float(angle), float(scale)))
        if param >= 0:
            self.num_params = max(self.num_params, param + 1)
        return self

    def rot(self, kind: int, target: int, param: int, scale: float = 1.0) -> "Circuit":
        return self.add(kind, target, -1, param, 0.0, scale)

    # -- C view ---------------------------------------------------------------------------------
    def to_c(self) -> CCircuit:
        """ctypes view; the returned struct keeps its arrays alive via ``_keep``."""
        ga = (CGate * max(1, len(self.gates)))()
        for k, (kind, t, c, p, a, s) in enumerate(self.gates):
            ga[k] = CGate(kind, t, c, p, a, s)
        cc = CCircuit()
        cc.num_qubits, cc.num_params = self.num_qubits, self.num_params
        cc.gates, cc.num_gates = ga, len(self.gates)
        keep = [ga]
        if self.edges:
            ea = (CEdge * len(self.edges))()
            for k, (i, j, w) in enumerate(self.edges):
                ea[k] = CEdge(i, j, w)
            cc.edges, cc.num_edges = ea, len(self.edges)
            keep.append(ea)
        else:
            cc.edges, cc.num_edges = None, 0
        if self.vertex_weights is not None:
            va = (C.c_double * self.num_qubits)(*[float(v) for v in self.vertex_weights])
            cc.vertex_weights = va
            keep.append(va)
        else:
            cc.vertex_weights = None
        cc.initial_state = self.initial_state
        cc._keep = keep
        return cc

    # -- bookkeeping used by bench.py (SURVEY.md §8d byte formulas) ---------------------------------
    def unfused_bytes(self) -> float:
        """Algorithmic bytes of one forward run with one pass per gate (32*D dense, 16*D controlled)."""
        d = float(1 << self.num_qubits)
        tot = 0.0
        for kind, *_ in self.gates:
            if kind == I:
                continue
            tot += (16.0 if kind in (CNOT, CY, CZ, CH, CRX, CRY, CRZ) else 48.0 if kind == SWAP else 32.0) * d
        return tot


def default_angles(num_params: int, seed: int = SEED_ANGLES) -> np.ndarray:
    return np.random.default_rng(seed).uniform(-np.pi, np.pi, num_params).astype(np.float64)


def hea(n: int, num_params: int, name: str = "") -> Circuit:
    """First ``num_params`` rotation slots of the RY/RZ + CNOT-ladder ansatz on n qubits."""
    c = Circuit(n, name=name or f"hea_n{n}_p{num_params}")
    p = 0
    while p < num_params:
        for kind in (RY, RZ):
            for q in range(n):
                if p >= num_params:
                    return c
                c.rot(kind, q, p)
                p += 1
        if p >= num_params:
            return c
        for q in range(n - 1):
            c.add(CNOT, q + 1, q)
    return c


def hea_layers(n: int, layers: int, name: str = "") -> Circuit:
    """Exactly ``layers`` full layers (P = 2 n layers), each ending with its CNOT ladder."""
    c = Circuit(n, name=name or f"hea_n{n}_l{layers}")
    p = 0
    for _ in range(layers):
        for kind in (RY, RZ):
            for q in range(n):
                c.rot(kind, q, p)
                p += 1
        for q in range(n - 1):
            c.add(CNOT, q + 1, q)
    return c


def random_regular3(n: int, seed: int = SEED_GRAPH) -> List[Tuple[int, int, float]]:
    """3-regular simple graph as the union of three random perfect matchings (n even)."""
    assert n % 2 == 0
    rng = np.random.default_rng(seed)
    for _ in range(10000):
        edges = set()
        ok = True
        for _m in range(3):
            perm = rng.permutation(n)
            for a, b in zip(perm[0::2], perm[1::2]):
                e = (int(min(a, b)), int(max(a, b)))
                if e in edges:
                    ok = False
                    break
                edges.add(e)
            if not ok:
                break
        if ok:
            return [(i, j, 1.0) for (i, j) in sorted(edges)]
    raise RuntimeError("no simple 3-regular graph found")


def qaoa_maxcut(n: int, p: int, edges: Optional[List[Tuple[int, int, float]]] = None, name: str = "") -> Circuit:
    c = Circuit(n, name=name or f"qaoa_n{n}_p{p}", initial_state=INIT_PLUS)
    c.edges = edges if edges is not None else random_regular3(n)
    for layer in range(p):
        c.add(COST, 0, -1, 2 * layer, 0.0, 1.0)
        for q in range(n):
            c.add(RX, q, -1, 2 * layer + 1, 0.0, 2.0)   # RX(2 beta), qaoa.c:386
    c.num_params = 2 * p
    return c


def random_circuit(n: int, num_gates: int, seed: int, kinds: Optional[Sequence[int]] = None,
                   share_params: bool = False) -> Circuit:
    """Random circuit over ``kinds`` (default: the reference simulator's 13) for parity tests."""
    rng = np.random.default_rng(seed)
    kinds = list(kinds) if kinds is not None else sorted(REFERENCE_KINDS)
    c = Circuit(n, name=f"rand_n{n}_g{num_gates}_s{seed}")
    p = 0
    for _ in range(num_gates):
        kind = int(rng.choice(kinds))
        t = int(rng.integers(n))
        ctl = -1
        if kind in TWO_QUBIT:
            if n < 2:
                continue
            ctl = int(rng.integers(n - 1))
            if ctl >= t:
                ctl += 1
        if kind in PARAMETRIC and kind != COST:
            if share_params and p > 0 and rng.random() < 0.3:
                c.add(kind, t, ctl, int(rng.integers(p)), float(rng.uniform(-1, 1)), float(rng.uniform(0.5, 2.0)))
            else:
                c.add(kind, t, ctl, p, 0.0, 1.0)
                p += 1
        else:
            c.add(kind, t, ctl, -1, float(rng.uniform(-np.pi, np.pi)), 1.0)
    c.num_params = p
    return c


# the BASELINE.json configurations
def config(name: str) -> Circuit:
    name = name.lower()
    if name == "c1":
        return hea_layers(12, 2, "C1 hea n=12 L=2 P=48")
    if name == "c2":
        return hea_layers(20, 4, "C2 hea n=20 L=4 P=160")
    if name == "c3":
        return hea(28, 256, "C3 hea n=28 P=256")
    if name == "c3s":   # C3's parameter count on fewer qubits (everything resident)
        return hea(24, 256, "C3s hea n=24 P=256")
    if name == "c4":
        return qaoa_maxcut(30, 8, name="C4 qaoa n=30 p=8")
    if name == "c4s":
        return qaoa_maxcut(24, 8, name="C4s qaoa n=24 p=8")
    if name == "c5":
        return hea(33, 256, "C5 hea n=33 P=256")
    if name == "t30":
        return hea(30, 256, "target hea n=30 P=256")
    raise KeyError(name)
